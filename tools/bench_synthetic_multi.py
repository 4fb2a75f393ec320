#!/usr/bin/env python
"""BASELINE config 5 on N GPUs: the synthetic 1 003 244-triangle Cornell scene at 3840x2160, ONE frame of --passes
sample passes (BASELINE: 1024) split over the ranks, one NCCL sum-reduce of the framebuffer to rank 0 at frame end.

    python tools/bench_synthetic_multi.py --passes 1024                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_synthetic_multi.py --passes 1024

--mode passes (default): contiguous pass ranges per rank (parallel.render_frame), whole frame on every rank.
--mode rows: interleaved pixel rows, all passes on every rank, image bit-identical to the single-GPU one.
This is STRONG scaling (the frame is fixed); time = device time of the slowest rank including the reduce (CUDA events on
the context's stream, max over ranks).  Prints one JSON line on rank 0.  Not the driver's bench (that is bench.py)."""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('msc-futhark-ray-tracer_b200')
par = importlib.import_module('msc-futhark-ray-tracer_b200.parallel')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--passes', type=int, default=1024)
    ap.add_argument('--mode', default='passes', choices=['passes', 'rows'])
    ap.add_argument('--k', type=int, default=151, help='tessellation of every Cornell quad (151 -> 1 003 244 triangles)')
    ap.add_argument('--width', type=int, default=3840)
    ap.add_argument('--height', type=int, default=2160)
    ap.add_argument('--reps', type=int, default=2)
    ap.add_argument('--check-passes', type=int, default=0, help='rows mode: compare a frame of this many passes bit for bit with the unpartitioned render on rank 0')
    a = ap.parse_args()
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    d = np.load(os.path.join(ROOT, 'tests', 'golden', 'scenes', 'cornell.npz'))
    tris, tri_mats = pkg.scenes.synthetic_cornell(d['tris'], d['tri_mats'], a.k)
    ctx = pkg.Context(device=local)
    if a.mode == 'rows':
        ctx.set_partition(rank, world)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    state = pkg.State.init(ctx, tris, tri_mats, d['mats'], a.height, a.width)       # scene replicated: every rank builds the LBVH
    build_ms = state.bvh_rebuild_ms(3)

    def frame(passes):
        hnd, view = par.render_frame(ctx, state, passes, a.mode, rank, world, dev, stream)
        return hnd, view

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # warm-up: at least LYS_PIPELINE (8) passes on EVERY rank, so that all pipeline buffer sets exist before the timed frame
    # (a 4-pass warm-up left 4 of the 8 sets to be cudaMalloc'ed inside the first timed frame: ~0.1 s at 4K)
    hnd, _ = frame(min(a.passes, 16 * world) if a.mode == 'passes' else min(a.passes, 16))
    state.free_f32_3d(hnd)
    best = None
    for _ in range(a.reps):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        hnd, view = frame(a.passes)
        e1.record(stream)
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        mean = float(view.mean().item()) if rank == 0 else None
        state.free_f32_3d(hnd)
        if best is None or ms.item() < best[0]:
            best = (float(ms.item()), mean)
    # the one collective on its own: sum-reduce of a framebuffer-sized device buffer to rank 0
    reduce_ms = None
    if world > 1:
        buf = torch.zeros((a.height, a.width, 3), dtype=torch.float32, device=dev)
        with torch.cuda.stream(stream):
            par.reduce_framebuffer(buf, dst=0)
        sync()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            r0.record(stream)
            for _ in range(5):
                par.reduce_framebuffer(buf, dst=0)
            r1.record(stream)
        sync()
        t = torch.tensor([r0.elapsed_time(r1) / 5], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        reduce_ms = float(t.item())
    bit_exact = None
    if a.mode == 'rows' and a.check_passes > 0:
        hnd, view = frame(a.check_passes)
        sync()
        if rank == 0:
            got = view.cpu().numpy().copy()
        state.free_f32_3d(hnd)
        if rank == 0:
            full = pkg.Context(device=local)
            want = pkg.State.init(full, tris, tri_mats, d['mats'], a.height, a.width).sample_n_frames(a.check_passes)
            bit_exact = bool(np.array_equal(got.view(np.uint32), want.view(np.uint32)))
            full.close()
    if rank == 0:
        ms, mean = best
        print(json.dumps({'config': '5 synthetic cornell k=%d' % a.k, 'tris': int(len(tris)), 'res': '%dx%d' % (a.width, a.height),
                          'passes': a.passes, 'mode': a.mode, 'n_gpus': world, 'frame_ms': round(ms, 2),
                          'mpaths_s': round(a.width * a.height * a.passes / (ms * 1e-3) / 1e6, 1), 'scaling': 'strong',
                          'lbvh_build_ms': round(build_ms, 3), 'image_mean': mean,
                          'reduce_bytes': a.width * a.height * 12, 'reduce_ms': None if reduce_ms is None else round(reduce_ms, 3),
                          'rows_image_bit_exact_vs_single_gpu': bit_exact, 'check_passes': a.check_passes or None}), flush=True)
    sync()
    state.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
