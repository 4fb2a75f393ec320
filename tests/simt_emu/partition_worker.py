#!/usr/bin/env python
"""One rank of the world-size-2 gloo test of the PRODUCT's partitioned rendering (tests/test_simt_emu.py): the library's own
kernels (emulator build, loaded through run_with_emu.py) render this rank's rows / this rank's passes, the framebuffer is
sum-reduced over gloo, and rank 0 compares with the oracle.  Environment: RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT, OUT."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
pkg = importlib.import_module('msc-futhark-ray-tracer_b200')
par = importlib.import_module('msc-futhark-ray-tracer_b200.parallel')
from lysref import oracle  # noqa: E402


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    dist.init_process_group('gloo', rank=rank, world_size=world)
    d = np.load(os.path.join(ROOT, 'tests', 'golden', 'scenes', 'cornell.npz'))
    h, w, passes = 21, 27, 5
    res = {}
    # rows: every rank samples its interleaved rows with the library's kernels; the reduced image is the single-rank image, bit for bit
    ctx = pkg.Context()
    ctx.set_partition(rank, world)
    s = pkg.State.init(ctx, d['tris'], d['tri_mats'], d['mats'], h, w)
    mine = s.sample_n_frames(passes)
    rows = par.owned_rows(h, rank, world)
    other = np.setdiff1d(np.arange(h), rows)
    res['rows_zero_elsewhere'] = bool(not mine[other].any()) and bool(mine[rows].any())
    buf = torch.from_numpy(mine.copy())
    par.reduce_framebuffer(buf, dst=0)
    s.free(); ctx.close()
    if rank == 0:
        want = oracle.State.init(d['tris'], d['tri_mats'], d['mats'], h, w).sample_n_frames(passes)
        res['rows_bit_exact'] = bool(np.array_equal(buf.numpy().view(np.uint32), want.view(np.uint32)))
    # passes: contiguous pass ranges, the weight folded into the library's last accumulate kernel, one reduce
    ctx = pkg.Context()
    s = pkg.State.init(ctx, d['tris'], d['tri_mats'], d['mats'], h, w)
    total = 7
    ranges = par.pass_ranges(total, world)
    first, count = ranges[rank]
    counts = [c for _, c in ranges]
    sr = s.advance_rng(first) if first else s
    hnd, ptr, shape, _ = sr.sample_n_frames_device(count, want_stats=False, weight=par.pass_weight(count, counts))
    img = sr.values_f32_3d(hnd, shape)
    sr.free_f32_3d(hnd)
    buf = torch.from_numpy(img.copy())
    par.merge_pass_split(buf, count, counts, dst=0, weighted=True)
    if rank == 0:
        base = oracle.State.init(d['tris'], d['tri_mats'], d['mats'], h, w)
        kept = [k for f, c in ranges for k in range(f + (1 if c >= 2 else 0), f + c)]
        want = np.stack([base.advance_rng(k).sample_n_frames(1).astype(np.float64) for k in kept]).mean(axis=0)
        got = buf.numpy().astype(np.float64)
        res['passes_close'] = bool(np.allclose(got, want, rtol=1e-4, atol=1e-6 * want.max())) and bool(want.max() > 0)
        import json
        json.dump(res, open(os.environ['OUT'], 'w'))
    else:
        assert res['rows_zero_elsewhere']
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
