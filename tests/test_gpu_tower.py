"""GPU: head-tower epilogues against plain PyTorch references of the same ops (floating-point kernels: the
tolerance is stated in each test).  Reference semantics: ConvModule(conv, GN(32), ReLU) of every tower layer
(radet/models/dense_heads/atss_head.py:52-87,133-138) and relu(Scale(conv)) on the regression branch (:141-143,
radet_head.py:27-30)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from radet_b200 import functional as F  # noqa: E402
from radet_b200 import plugin as P  # noqa: E402

DEV = "cuda"


def _gn_relu_ref(x, w, b, groups, eps, dtype):
    x = x.detach().to(dtype).requires_grad_()
    w_ = None if w is None else w.detach().to(dtype).requires_grad_()
    b_ = None if b is None else b.detach().to(dtype).requires_grad_()
    y = torch.relu(torch.nn.functional.group_norm(x, groups, w_, b_, eps))
    return x, w_, b_, y


@pytest.mark.parametrize("shape,groups", [((2, 256, 60, 80), 32), ((3, 64, 7, 5), 32), ((1, 256, 120, 160), 32), ((2, 8, 6, 8), 4),
                                          ((2, 96, 15, 20), 32)])
@pytest.mark.parametrize("affine", [True, False])
def test_gn_relu_forward_backward_vs_torch(shape, groups, affine):
    g = torch.Generator(device=DEV).manual_seed(sum(shape) + groups)
    x = (torch.randn(shape, device=DEV, generator=g) * 1.7 + 3.0).requires_grad_()      # a mean well away from 0
    w = (torch.randn(shape[1], device=DEV, generator=g) * 0.5 + 1.0).requires_grad_() if affine else None
    b = (torch.randn(shape[1], device=DEV, generator=g) * 0.3).requires_grad_() if affine else None
    dy = torch.randn(shape, device=DEV, generator=g)
    y = F.gn_relu(x, w, b, groups, 1e-5)
    y.backward(dy)
    for dtype, rt, at in ((torch.float64, 2e-5, 2e-5), (torch.float32, 1e-4, 1e-4)):
        xr, wr, br, yr = _gn_relu_ref(x, w, b, groups, 1e-5, dtype)
        yr.backward(dy.to(dtype))
        # elements whose pre-activation is within rounding of 0 may sit on the other side of the ReLU: compare away from it
        z = torch.nn.functional.group_norm(xr.detach(), groups, None if wr is None else wr.detach(), None if br is None else br.detach(), 1e-5)
        safe = z.abs() > 1e-4
        assert safe.float().mean() > 0.999
        torch.testing.assert_close(y.double()[safe], yr.double()[safe], rtol=rt, atol=at)
        # dx: every element depends on the group sums, which include the unsafe elements' dy; their weight is ~1e-3 of the sum
        scale = float(xr.grad.abs().max())
        assert float((x.grad.double() - xr.grad.double())[safe].abs().max()) <= 5e-4 * scale
        if affine:
            torch.testing.assert_close(w.grad.double(), wr.grad.double(), rtol=2e-3, atol=2e-3 * float(wr.grad.abs().max()))
            torch.testing.assert_close(b.grad.double(), br.grad.double(), rtol=2e-3, atol=2e-3 * float(br.grad.abs().max()))


def test_gn_relu_is_deterministic_and_handles_empty_batch():
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn((2, 256, 30, 40), device=DEV, generator=g).requires_grad_()
    w = torch.randn(256, device=DEV, generator=g).requires_grad_()
    b = torch.randn(256, device=DEV, generator=g).requires_grad_()
    outs = []
    for _ in range(2):
        x.grad = w.grad = b.grad = None
        y = F.gn_relu(x, w, b, 32)
        y.sum().backward()
        outs.append((y.detach().clone(), x.grad.clone(), w.grad.clone(), b.grad.clone()))
    for a, c in zip(*outs):
        assert torch.equal(a, c)
    e = F.gn_relu(torch.zeros((0, 256, 4, 4), device=DEV), w, b, 32)
    assert e.shape == (0, 256, 4, 4)
    with pytest.raises(Exception):
        F.gn_relu(torch.zeros((1, 30, 4, 4), device=DEV), None, None, 32)


@pytest.mark.parametrize("n", [1, 4 * 60 * 80 * 2, 1000003])
def test_scale_relu_vs_torch(n):
    g = torch.Generator(device=DEV).manual_seed(n % 1000)
    x = torch.randn(n, device=DEV, generator=g).requires_grad_()
    s = torch.tensor(1.37, device=DEV).requires_grad_()
    dy = torch.randn(n, device=DEV, generator=g)
    y = F.scale_relu(x, s)
    y.backward(dy)
    xr = x.detach().clone().requires_grad_()
    sr = s.detach().clone().requires_grad_()
    yr = torch.relu(xr * sr)
    yr.backward(dy)
    assert torch.equal(y, yr)                                   # one multiply, one max: bit-exact
    assert torch.equal(x.grad, xr.grad)
    torch.testing.assert_close(s.grad, sr.grad, rtol=1e-4, atol=1e-4 * float(sr.grad.abs() + 1))   # fp64 partial sums vs torch's fp32 sum
    ref64 = float((dy.double() * x.detach().double() * (x.detach() * 1.37 > 0)).sum())
    assert abs(float(s.grad) - ref64) <= 1e-6 * max(1.0, abs(ref64))
    # negative scale flips the mask
    s2 = torch.tensor(-0.5, device=DEV)
    assert torch.equal(F.scale_relu(x.detach(), s2), torch.relu(x.detach() * s2))


def test_head_forward_uses_the_fused_epilogues_and_matches_the_torch_tower(monkeypatch):
    """plugin.RADetHead.forward on CUDA with the fused epilogues (GN+ReLU, Scale+ReLU) against the same module's torch path in
    fp64 on the CPU: outputs within 2e-4, parameter gradients as close as the plain torch CUDA tower gets (the cuDNN
    convolutions set the floor)."""
    import copy
    torch.manual_seed(3)
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)      # the convolutions are cuDNN: keep them fp32 for the comparison
    cfg = dict(type="RADetHead", num_classes=5, in_channels=64, feat_channels=64, stacked_convs=2,
               norm_cfg=dict(type="GN", num_groups=32, requires_grad=True))
    head = P.build_head(cfg)
    head.init_weights()
    for m in list(head.cls_convs) + list(head.reg_convs):          # non-trivial affine parameters
        torch.nn.init.normal_(m.gn.weight, 1.0, 0.2)
        torch.nn.init.normal_(m.gn.bias, 0.0, 0.2)
        torch.nn.init.normal_(m.conv.weight, std=0.05)
    torch.nn.init.normal_(head.atss_reg.weight, std=0.05)
    with torch.no_grad():
        head.scales[0].scale.fill_(1.3)
    ref = copy.deepcopy(head).double()                              # CPU, fp64: the torch path of the same module
    feats = [torch.randn(2, 64, 15, 20), torch.randn(2, 64, 8, 10)]
    out_r = ref([f.double() for f in feats])
    sum((t * t).sum() for grp in out_r for t in grp).backward()

    def run(fused):
        h = copy.deepcopy(head).to(DEV)
        calls = []
        if fused:
            real_gn, real_sr = F.gn_relu, F.scale_relu
            monkeypatch.setattr(F, "gn_relu", lambda *a, **k: (calls.append("gn"), real_gn(*a, **k))[1])
            monkeypatch.setattr(F, "scale_relu", lambda *a, **k: (calls.append("sr"), real_sr(*a, **k))[1])
        else:       # the plain torch CUDA tower: same ops as the reference's forward_single
            monkeypatch.setattr(F, "gn_relu", lambda x, w, b, g, eps=1e-5: torch.relu(torch.nn.functional.group_norm(x, g, w, b, eps)))
            monkeypatch.setattr(F, "scale_relu", lambda x, s: torch.relu(x * s))
        out = h([f.to(DEV) for f in feats])
        sum((t * t).sum() for grp in out for t in grp).backward()
        if fused:
            assert calls.count("gn") == 2 * 2 * 2 and calls.count("sr") == 2      # 2 towers x 2 layers x 2 levels; 2 levels
        errs = {}
        for (n, p), (_, q) in zip(h.named_parameters(), ref.named_parameters()):
            if p.grad is None:
                assert q.grad is None or float(q.grad.abs().max()) == 0, n
                continue
            errs[n] = float((p.grad.cpu().double() - q.grad).abs().max()) / float(q.grad.abs().max() + 1e-30)
        return out, errs

    out, errs = run(True)
    for grp, grp_r in zip(out, out_r):
        for a, b in zip(grp, grp_r):
            torch.testing.assert_close(a.cpu().double(), b.double(), rtol=2e-4, atol=2e-4)
    _, base = run(False)
    worst = {n: (e, base[n]) for n, e in errs.items() if e > max(2.0 * base[n], 1e-4)}
    assert not worst, worst
