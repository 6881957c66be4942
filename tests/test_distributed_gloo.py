"""World-size-2 gloo test of the multi-GPU host logic (sample-range sharding + one reduce of the film).  On the CPU
box the per-rank renderer is the oracle (tests may use it); on the GPU box bench.py drives the same code with the
CUDA renderer and NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import craytracer_b200 as c
import oracle_lib as o
from craytracer_b200 import scenes
from craytracer_b200.distributed import render_sharded

SPP = 5
W, H = 48, 32


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hs = c.parse_scene(scenes.materials(width=W, height=H))
    orc = o.OracleScene(hs)

    def render_range(lo, hi, out):
        film, _ = orc.render(W, H, seed=1, sample_begin=lo, sample_end=hi, threads=2)
        out.copy_(torch.from_numpy(film))

    mean, (lo, hi) = render_sharded(render_range, SPP, (H, W, 3), "cpu")
    assert (lo, hi) == ((0, 2) if rank == 0 else (2, 5))
    if rank == 0:
        np.save(out_path, mean.numpy())
    else:
        assert mean is None
    dist.destroy_process_group()


def test_two_rank_sharded_render_equals_single_rank(tmp_path):
    out = str(tmp_path / "film.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    sharded = np.load(out)
    hs = c.parse_scene(scenes.materials(width=W, height=H))
    whole, _ = o.OracleScene(hs).render(W, H, seed=1, sample_begin=0, sample_end=SPP, threads=2)
    whole = whole / np.float32(SPP)
    assert np.abs(sharded - whole).max() <= 1e-5 * max(1.0, float(whole.max()))
