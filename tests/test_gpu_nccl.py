"""2 GPUs, NCCL: the ONE collective north_star names -- the opt-in FCOS/ATSS-style reduce_mean of the loss normalisers
(core/utils/dist_utils.py:63-69; RADetHead(sync_num_pos=True)).  Every rank owns different images; with the sync on,
each rank's losses and gradients must equal the oracle evaluated with the all-reduced normalisers; with it off (the
reference behaviour of RADetHead, radet_head.py:254-259) they equal the rank-local oracle.  Skipped below 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from oracle import radet_oracle as orc
    from radet_b200 import functional as F
    from radet_b200 import sharding
    from radet_b200 import synthetic as syn

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        wl = syn.WORKLOADS["cfg1"]
        geom = F.Geometry()
        lo, hi = sharding.image_range(rank, world, wl.B)
        batch = syn.make_batch(wl, hi - lo, lo)                      # this rank's images
        a = [orc.assign_image_seeded(im.gt_bboxes, im.masks, im.H, im.W, im.seed) for im in batch]
        idx_l, w_l = [x[0] for x in a], [x[1] for x in a]
        ho = syn.make_head_outputs(wl, batch, idx_l, seed_base=100 + 7 * rank)
        T = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
        cls, bbox, iou = [T(m) for m in ho.cls], [T(m) for m in ho.bbox], [T(m) for m in ho.iou]
        counts = [im.gt_bboxes.shape[0] for im in batch]
        boxes, labels = T(np.concatenate([im.gt_bboxes for im in batch])), T(np.concatenate([im.gt_labels for im in batch]))
        idx, w = T(np.stack(idx_l)), T(np.stack(w_l))
        gtb, gtl = [im.gt_bboxes for im in batch], [im.gt_labels for im in batch]
        local = orc.head_loss(ho.cls, ho.bbox, ho.iou, gtb, gtl, idx_l, w_l, wl.C, wl.H, wl.W)
        # rank-local (reference behaviour)
        l0, g0 = F.loss_fwd_bwd(geom, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, F.LossConfig())
        # synced normalisers over NCCL
        l1, g1 = F.loss_fwd_bwd(geom, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, F.LossConfig(), sync_group=dist.group.WORLD)
        norm = torch.tensor([local["num_pos"], local["sum_wq"]], dtype=torch.float64, device=dev)
        dist.all_reduce(norm)
        norm /= world
        synced = orc.head_loss(ho.cls, ho.bbox, ho.iou, gtb, gtl, idx_l, w_l, wl.C, wl.H, wl.W, normalizers=tuple(norm.tolist()))
        res = {"rank": rank, "ok": True, "msg": ""}
        for tag, (lv, gr), ref in (("local", (l0, g0), local), ("synced", (l1, g1), synced)):
            lv = lv.cpu().numpy()
            for i, k in enumerate(("loss_cls", "loss_bbox", "loss_iou")):
                if not abs(lv[i] - ref[k]) <= 1e-5 * abs(ref[k]):
                    res["ok"], res["msg"] = False, f"{tag} {k}: {lv[i]} vs {ref[k]}"
            for mine, want in ((gr[0], ref["grad_cls"]), (gr[1], ref["grad_bbox"]), (gr[2], ref["grad_iou"])):
                for x, r in zip(mine, want):
                    if not np.allclose(x.cpu().numpy(), r, rtol=1e-4, atol=1e-6 * max(1e-30, float(np.abs(r).max()))):
                        res["ok"], res["msg"] = False, f"{tag} gradients differ"
        res["num_pos"] = (local["num_pos"], float(norm[0]))
        q.put(res)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_sync_num_pos_over_nccl():
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(world)), key=lambda r: r["rank"])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in res:
        assert r["ok"], r
    # the two ranks own different images: their local positive counts differ, the synced normaliser is their mean
    assert res[0]["num_pos"][1] == res[1]["num_pos"][1] == 0.5 * (res[0]["num_pos"][0] + res[1]["num_pos"][0])
