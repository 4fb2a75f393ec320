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


def _pass_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import importlib
    from lysref import oracle
    par = importlib.import_module('msc-futhark-ray-tracer_b200.parallel')
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    oracle.set_threads(2)
    t, tm, m = load_scene('cornell')
    h, w, total = 16, 20, 7
    ranges = par.pass_ranges(total, world)
    first, count = ranges[rank]
    base = oracle.State.init(t, tm, m, h, w)
    mine = base.advance_rng(first).sample_n_frames(count)          # this rank's running average of its pass range
    buf = torch.from_numpy(mine.copy())
    par.merge_pass_split(buf, count, [c for _, c in ranges], dst=0)
    if rank == 0:
        # what the merge must equal: the mean of the passes each rank's average holds (its first pass is dropped)
        kept = [k for f, c in ranges for k in range(f + (1 if c >= 2 else 0), f + c)]
        per_pass = np.stack([base.advance_rng(k).sample_n_frames(1).astype(np.float64) for k in kept])
        np.save(out_path, np.stack([buf.numpy().astype(np.float64), per_pass.mean(axis=0)]))
    dist.barrier()
    dist.destroy_process_group()


def test_pass_split_merge_world2(tmp_path):
    """Pass-split rendering (bench.py's N>1 mode, tools/bench_synthetic_multi.py): the weighted sum of the per-rank
    running averages equals the mean of the passes they hold, to f32 rounding (1e-4 relative is the north-star bar)."""
    out = str(tmp_path / 'p.npy')
    mp.spawn(_pass_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got, want = np.load(out)
    assert want.max() > 0
    assert np.allclose(got, want, rtol=1e-4, atol=1e-6 * want.max())


def test_pass_ranges_arithmetic():
    import importlib
    par = importlib.import_module('msc-futhark-ray-tracer_b200.parallel')
    for total in (1, 7, 16, 1024):
        for world in (1, 2, 3, 4, 8):
            r = par.pass_ranges(total, world)
            assert sum(c for _, c in r) == total and r[0][0] == 0
            assert all(r[i][0] + r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in r) - min(c for _, c in r) <= 1
    assert par.averaged_passes(1) == 1 and par.averaged_passes(2) == 1 and par.averaged_passes(128) == 127
