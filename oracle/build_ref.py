"""TEST INFRASTRUCTURE — recipe that compiles the reference's OWN native ops from where they lie.

    python oracle/build_ref.py        (build container only: needs /root/reference)

/root/reference/radet/ops/vote/vote_ext.cpp and ops/cluster/cluster_ext.cpp are pybind11/libtorch CPU extensions
(setup.py:138-145).  They are compiled unmodified into oracle/_ref/{vote_ext,cluster_ext}/*.so (git-ignored, not
gpurun-ignored: the binaries travel to the GPU box where /root/reference does not exist).  Sources are never copied.
They strengthen the restatement (oracle/vote_oracle.c is checked against them) and serve as the `reference` kind of
CPU baseline for the NMS stage in bench.py.
"""
import glob
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RADET_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
OPS = {"vote_ext": "radet/ops/vote/vote_ext.cpp", "cluster_ext": "radet/ops/cluster/cluster_ext.cpp"}


def build():
    from torch.utils.cpp_extension import load

    for name, rel in OPS.items():
        d = os.path.join(OUT, name)
        os.makedirs(d, exist_ok=True)
        load(name=name, sources=[os.path.join(REF, rel)], build_directory=d, extra_cflags=["-O2"], verbose=False)
    return OUT


def load_prebuilt(name):
    """Import oracle/_ref/<name>/<name>.so (prebuilt here; usable on the GPU box)."""
    import torch  # noqa: F401  (libtorch symbols)

    cands = glob.glob(os.path.join(OUT, name, f"{name}*.so"))
    if not cands:
        raise FileNotFoundError(f"oracle/_ref/{name} not built (run oracle/build_ref.py in the build container)")
    spec = importlib.util.spec_from_file_location(name, cands[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build())
    m = load_prebuilt("vote_ext")
    print([n for n in dir(m) if "nms" in n])
