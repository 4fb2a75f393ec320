"""The N>1 path on CPU: row-interleaved pixel partition + one sum-reduce to rank 0,
exercised with world_size 2 over gloo.  Each rank renders only its rows with the oracle standing in for the
device kernels; the reduced framebuffer must equal the single-rank image bit-for-bit."""
import os
import socket
import sys
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from conftest import ROOT, load_scene


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import importlib
    from lysref import oracle
    par = importlib.import_module('msc-futhark-ray-tracer_b200.parallel')
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    oracle.set_threads(2)
    t, tm, m = load_scene('cornell')
    h, w = 20, 24
    full = oracle.State.init(t, tm, m, h, w).sample_n_frames(3)
    mine = np.zeros_like(full)
    rows = par.owned_rows(h, rank, world)
    mine[rows] = full[rows]                                       # what a rank's framebuffer holds: zeros elsewhere
    assert par.local_pixel_count(h, w, rank, world) == len(rows) * w
    buf = torch.from_numpy(mine.copy())
    par.reduce_framebuffer(buf, dst=0)
    if rank == 0:
        np.save(out_path, np.stack([buf.numpy(), full]))
    dist.barrier()
    dist.destroy_process_group()


def test_row_partition_reduce_world2(tmp_path):
    out = str(tmp_path / 'r.npy')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got, want = np.load(out)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_partition_arithmetic():
    import importlib
    par = importlib.import_module('msc-futhark-ray-tracer_b200.parallel')
    for h in (1, 7, 1080, 2160):
        for world in (1, 2, 3, 4, 8):
            rows = [par.owned_rows(h, r, world) for r in range(world)]
            assert sorted(np.concatenate(rows).tolist()) == list(range(h))
            assert max(len(r) for r in rows) - min(len(r) for r in rows) <= 1
