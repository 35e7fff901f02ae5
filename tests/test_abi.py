"""CPU: the C-ABI library loads, exports every symbol include/radet_b200.h declares, and the ctypes mirrors of the
header structs have the C compiler's layout.  No compute calls (no GPU here)."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest

from radet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "radet_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(radet_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert set(names) <= set(_lib.EXPORTED_SYMBOLS), set(names) - set(_lib.EXPORTED_SYMBOLS)
    assert lib.radet_version().decode().startswith("radet_b200")


def test_ctypes_struct_layout_matches_the_header():
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "radet_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(radet_grid_t), sizeof(radet_maps_t), sizeof(radet_loss_cfg_t), sizeof(radet_detect_cfg_t),
         sizeof(radet_grad_maps_t));
  printf("%zu %zu %zu %zu\n", offsetof(radet_grid_t, stride), offsetof(radet_grid_t, range_lo), offsetof(radet_grid_t, anchor_scale),
         offsetof(radet_detect_cfg_t, iou_threshold));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), c, "-o", exe])   # header is plain C
        out = subprocess.check_output([exe]).decode().split()
    sizes = [int(x) for x in out]
    assert sizes[:5] == [ctypes.sizeof(_lib.Grid), ctypes.sizeof(_lib.Maps), ctypes.sizeof(_lib.LossCfg), ctypes.sizeof(_lib.DetectCfg),
                         ctypes.sizeof(_lib.Maps)]
    assert sizes[5:] == [_lib.Grid.stride.offset, _lib.Grid.range_lo.offset, _lib.Grid.anchor_scale.offset,
                         _lib.DetectCfg.iou_threshold.offset]


def test_host_side_argument_checks_without_gpu():
    """Argument errors are detected on the host before any launch (RADET_E_*), so they can be exercised without a GPU."""
    lib = _lib.load()
    g = _lib.make_grid([(60, 80), (30, 40)], [8, 16], [(-1, 64), (64, 1e8)])
    assert lib.radet_num_points(ctypes.byref(g)) == 60 * 80 + 30 * 40
    bad = _lib.Grid()
    assert lib.radet_num_points(ctypes.byref(bad)) == -1
    assert lib.radet_assign_workspace_bytes(ctypes.byref(g), 4) > 0
    assert lib.radet_loss_workspace_bytes(ctypes.byref(g), 4, 21) > 0
    # null pointers -> RADET_E_BADARG (-1), never a crash
    rc = lib.radet_get_targets(ctypes.byref(g), 2, 21, None, None, None, None, None, None, None, None, None, None)
    assert rc == -1
    rc = lib.radet_pack_masks(None, 3, 480, 640, 8, 60, 80, None, None, None)
    assert rc == -1
    with pytest.raises(_lib.RadetError):
        _lib.check(rc, "radet_pack_masks")


def test_product_has_no_cpu_path():
    import torch

    from radet_b200 import functional as F

    with pytest.raises(_lib.RadetError):
        F.pack_masks(torch.zeros((1, 60, 80), dtype=torch.uint8), 1, 60, 80)          # CPU tensor -> loud failure
    with pytest.raises(_lib.RadetError):
        F.tblr_encode(torch.zeros((1, 4)), torch.zeros((1, 4)), 0.125)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under radet_b200/ may import it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "radet_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dirpath, f)
