#!/usr/bin/env python
"""The five BASELINE.json configs on one B200 next to the CPU oracle on a bounded sample (writes JSON lines).

Not the driver's bench (that is bench.py); this fills the per-config table in profiles/README.md."""
import importlib
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
pkg = importlib.import_module('msc-futhark-ray-tracer_b200')
from lysref import oracle  # noqa: E402


def scene(name):
    d = np.load(os.path.join(ROOT, 'tests', 'golden', 'scenes', name + '.npz'))
    return d['tris'], d['tri_mats'], d['mats']


def run(ctx, label, tris, tm, mats, h, w, passes, path_len=16, origin=(0.0, 0.8, 1.8), cpu_passes=1, reps=3):
    ctx.set_path_len(path_len)
    oracle.set_path_len(path_len)
    s = pkg.State.init(ctx, tris, tm, mats, h, w, origin=origin)
    hnd, _, _, _ = s.sample_n_frames_device(min(passes, 4))          # warm-up
    s.free_f32_3d(hnd)
    best = None
    for _ in range(reps):
        hnd, _, shape, st = s.sample_n_frames_device(passes)
        s.free_f32_3d(hnd)
        if best is None or st['device_ms'] < best['device_ms']:
            best = st
    build_ms = s.bvh_rebuild_ms(5)
    so = oracle.State.init(tris, tm, mats, h, w, origin=origin)
    oracle.set_threads(len(os.sched_getaffinity(0)))
    oracle.counters_reset()
    t0 = time.perf_counter()
    so.sample_n_frames(cpu_passes)
    dt = time.perf_counter() - t0
    c = oracle.counters()
    pp = {k: c[k] / max(c['paths'], 1) for k in c}
    gpu = h * w * passes / (best['device_ms'] * 1e-3) / 1e6
    cpu = h * w * cpu_passes / dt / 1e6
    b_path = 24 + 32 * pp['box_tests'] + 40 * pp['tri_tests'] + 112 * pp['vertices']
    rec = dict(config=label, tris=int(len(tris)), res='%dx%d' % (w, h), passes=passes, path_len=path_len, gpu_mpaths_s=round(gpu, 1),
               gpu_ms_per_pass=round(best['device_ms'] / passes, 3), cpu_mpaths_s=round(cpu, 2), cpu_cores=oracle.get_threads(),
               speedup=round(gpu / cpu, 1), vertices_per_path=round(best['vertices'] / best['paths'], 3),
               shadow_rays_per_path=round(best['shadow_rays'] / best['paths'], 3), box_tests_per_path=round(pp['box_tests'], 1),
               tri_tests_per_path=round(pp['tri_tests'], 1), b_path_bytes=round(b_path), algorithmic_gbs=round(b_path * gpu * 1e6 / 1e9, 1),
               lbvh_build_ms=round(build_ms, 3))
    print(json.dumps(rec), flush=True)
    s.free()
    ctx.set_path_len(16)
    oracle.set_path_len(16)
    return rec


def main():
    ctx = pkg.Context()
    out = []
    only = sys.argv[1:]                                   # optional: config labels to run, by prefix ("5", "2b", "metric")
    want = lambda label: not only or any(label.startswith(p) for p in only)
    c = scene('cornell')
    if want('1a'): out.append(run(ctx, '1a cornell 512x512 path_len 5', *c, 512, 512, 64, path_len=5, cpu_passes=4))
    if want('1b'): out.append(run(ctx, '1b cornell 512x512 path_len 16', *c, 512, 512, 64, cpu_passes=4))
    if want('metric'): out.append(run(ctx, 'metric cornell 1080p', *c, 1080, 1920, 16))
    m = scene('mirrorbox')
    if want('2a'): out.append(run(ctx, '2a mirrorbox 1080p default camera (outside the box)', *m, 1080, 1920, 64))
    if want('2b'): out.append(run(ctx, '2b mirrorbox 1080p camera inside (0,0.8,0.6)', *m, 1080, 1920, 64, origin=(0.0, 0.8, 0.6)))
    if want('3'): out.append(run(ctx, '3 spectrumsphere 1080p 256 passes', *scene('spectrumsphere'), 1080, 1920, 256, reps=1))
    if want('4'): out.append(run(ctx, '4 spectrumspherehigh 1080p', *scene('spectrumspherehigh'), 1080, 1920, 16))
    if want('5'):
        st, sm = pkg.scenes.synthetic_cornell(c[0], c[1], 151)
        out.append(run(ctx, '5 synthetic 1003244 tris 4K (16 of 1024 passes)', st, sm, c[2], 2160, 3840, 16, reps=2))
    for lab in only:                                      # "k38": every Cornell quad as a 38 x 38 grid (44 k^2 triangles) at 1080p: layout crossover measurements
        if lab[0] == 'k' and lab[1:].isdigit():
            st, sm = pkg.scenes.synthetic_cornell(c[0], c[1], int(lab[1:]))
            out.append(run(ctx, '%s synthetic %d tris 1080p' % (lab, len(st)), st, sm, c[2], 1080, 1920, 16))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    if not only:
        json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'configs.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
