#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/multi_gpu_check.py
Row-interleaved multi-GPU rendering over NCCL: every rank samples the grid rows r % N == rank,
one dist.reduce(SUM) of the framebuffer to rank 0, and rank 0 checks the result BIT-FOR-BIT against its own
unpartitioned render.  Also reports the strong-scaling throughput of this mode (same frame split over N GPUs)."""
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
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    res = {}
    for name, h, w, passes in (('spectrumsphere', 270, 480, 4), ('cornell', 1080, 1920, 16)):
        d = np.load(os.path.join(ROOT, 'tests', 'golden', 'scenes', name + '.npz'))
        ctx = pkg.Context(device=local)
        ctx.set_partition(rank, world)
        stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
        s = pkg.State.init(ctx, d['tris'], d['tri_mats'], d['mats'], h, w)

        def step():
            hnd, ptr, shape, _ = s.sample_n_frames_device(passes, want_stats=False)
            t = par.as_torch(ptr, shape, dev)
            with torch.cuda.stream(stream):
                par.reduce_framebuffer(t, dst=0)
            return hnd, t
        hnd, t = step()
        torch.cuda.synchronize(dev)
        got = t.cpu().numpy().copy() if rank == 0 else None
        s.free_f32_3d(hnd)
        # timing (strong scaling: the SAME frame over `world` GPUs)
        for _ in range(2):
            s.free_f32_3d(step()[0])
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        K = 10
        for _ in range(K):
            s.free_f32_3d(step()[0])
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            full_ctx = pkg.Context(device=local)
            want = pkg.State.init(full_ctx, d['tris'], d['tri_mats'], d['mats'], h, w).sample_n_frames(passes)
            res[name] = {'bit_exact_vs_single_gpu': bool(np.array_equal(got.view(np.uint32), want.view(np.uint32))),
                         'mpaths_s': h * w * passes * K / (float(ms.item()) * 1e-3) / 1e6, 'res': '%dx%d' % (w, h), 'passes': passes}
            full_ctx.close()
        s.free()
        ctx.close()
    if rank == 0:
        print(json.dumps({'n_gpus': world, 'mode': 'row partition (strong scaling)', **res}), flush=True)
        assert all(v['bit_exact_vs_single_gpu'] for v in res.values())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
