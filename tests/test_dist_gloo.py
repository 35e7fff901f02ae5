"""CPU, world_size 2 over gloo: the N>1 host logic.  The path shards images and exchanges nothing (reference parity);
the only collective is the OPT-IN reduce_mean of the two loss normalisers."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import radet_oracle as orc
from radet_b200 import sharding
from radet_b200 import synthetic as syn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) opt-in normaliser sync == mean over ranks (core/utils/dist_utils.py:63-69)
        norm = torch.tensor([10.0 * (rank + 1), 2.5 * (rank + 1)], dtype=torch.float64)
        sharding.reduce_mean_(norm)
        # (2) image sharding: every rank assigns its own images; the concatenation must equal the single-process run
        wl = syn.Workload("dist_160x128", 128, 160, 21, 2, 3, 5, 12)
        lo, hi = sharding.image_range(rank, world, wl.B)
        batch = syn.make_batch(wl, hi - lo, lo)
        out = [orc.assign_image_seeded(im.gt_bboxes, im.masks, im.H, im.W, im.seed)[:2] for im in batch]
        q.put((rank, norm.tolist(), [(i.tolist(), w.tolist()) for i, w in out]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_shard_images_and_reduce_mean():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, norm, _ in res:
        assert norm == [15.0, 3.75]
    wl = syn.Workload("dist_160x128", 128, 160, 21, 2, 3, 5, 12)
    single = [orc.assign_image_seeded(im.gt_bboxes, im.masks, im.H, im.W, im.seed)[:2] for im in syn.make_batch(wl, 4, 0)]
    sharded = [x for _, _, out in res for x in out]
    assert len(sharded) == 4
    for (i1, w1), (i2, w2) in zip(sharded, single):
        assert np.array_equal(np.asarray(i1), i2) and np.array_equal(np.asarray(w1, np.float32), w2)
