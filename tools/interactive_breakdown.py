#!/usr/bin/env python
"""Where a frame of the interactive loop goes (development aid): step, render, read-back timed separately, each followed by a
context sync, plus raw copy rates of the box.  python tools/interactive_breakdown.py [scene] [w] [h]"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('msc-futhark-ray-tracer_b200')
import ctypes as C


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'cornell'
    w = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
    h = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
    d = np.load(os.path.join(ROOT, 'tests', 'golden', 'scenes', name + '.npz'))
    res = {}
    with pkg.Context() as ctx:
        L = ctx._L
        s = pkg.State.init(ctx, d['tris'], d['tri_mats'], d['mats'], h, w).resize(h, w).key(pkg.KEY['m'])
        out = np.empty((h, w), np.int32)
        for _ in range(20):
            n = s.step(); s.free(); s = n; s.render(out)
        N = 200
        t0 = time.perf_counter()
        for _ in range(N):
            n = s.step(); s.free(); s = n; ctx.sync()
        res['step_ms'] = (time.perf_counter() - t0) / N * 1e3
        t0 = time.perf_counter()
        for _ in range(N):
            n = s.step(); s.free(); s = n
        ctx.sync()
        res['step_back_to_back_ms'] = (time.perf_counter() - t0) / N * 1e3
        arr = s._entry('futhark_entry_render', s._p); ctx.sync()
        t0 = time.perf_counter()
        for _ in range(N):
            ctx.check(L.futhark_values_i32_2d(ctx._ctx, arr, out.ctypes.data_as(C.c_void_p)), 'values')
        res['values_reused_buffer_ms'] = (time.perf_counter() - t0) / N * 1e3
        t0 = time.perf_counter()
        for _ in range(50):
            o2 = np.empty((h, w), np.int32)
            ctx.check(L.futhark_values_i32_2d(ctx._ctx, arr, o2.ctypes.data_as(C.c_void_p)), 'values')
        res['values_fresh_buffer_ms'] = (time.perf_counter() - t0) / 50 * 1e3
        L.futhark_free_i32_2d(ctx._ctx, arr)
        t0 = time.perf_counter()
        for _ in range(N):
            a = s._entry('futhark_entry_render', s._p); ctx.sync(); L.futhark_free_i32_2d(ctx._ctx, a)
        res['render_ms'] = (time.perf_counter() - t0) / N * 1e3
        t0 = time.perf_counter()
        for _ in range(N):
            n = s.step(); s.free(); s = n; s.render(out)
        res['frame_ms'] = (time.perf_counter() - t0) / N * 1e3
        src = np.empty((h, w), np.int32); src[:] = 1
        t0 = time.perf_counter()
        for _ in range(50):
            np.copyto(out, src)
        res['host_memcpy_frame_ms'] = (time.perf_counter() - t0) / 50 * 1e3
        s.free()
    try:
        import torch
        g = torch.empty(h * w, dtype=torch.int32, device='cuda')
        pin = torch.empty(h * w, dtype=torch.int32).pin_memory()
        pag = torch.empty(h * w, dtype=torch.int32)
        for nm, dst in (('pinned', pin), ('pageable', pag)):
            dst.copy_(g); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(50):
                dst.copy_(g)
            torch.cuda.synchronize()
            res['torch_d2h_%s_ms' % nm] = (time.perf_counter() - t0) / 50 * 1e3
    except Exception as e:      # noqa
        res['torch'] = repr(e)
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()}))


if __name__ == '__main__':
    main()
