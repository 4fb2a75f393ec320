#!/usr/bin/env python
"""A few 1080p CornellBox sample passes (profiling target for ncu)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('msc-futhark-ray-tracer_b200')
name = sys.argv[1] if len(sys.argv) > 1 else 'cornell'
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
d = np.load(os.path.join(ROOT, 'tests', 'golden', 'scenes', name + '.npz'))
ctx = pkg.Context()
kw = {'origin': (0.0, 0.8, 0.6)} if name == 'mirrorbox' else {}
H, W = (int(os.environ.get('LYS_H', 1080)), int(os.environ.get('LYS_W', 1920)))
s = pkg.State.init(ctx, d['tris'], d['tri_mats'], d['mats'], H, W, **kw)
if os.environ.get('LYS_WARM', '1') != '0':
    s.sample_n_frames_device(1)      # leaves the queue-length estimates that size the late grids / pick the fused tail
h, ptr, shape, st = s.sample_n_frames_device(passes)
ctx.sync()
print(name, shape, st)
if os.environ.get('LYS_DETAIL'):
    ctx.set_profiling(True); ctx.profile(reset=True)
    s.sample_n_frames_device(passes)
    d = ctx.profile_detail(); pr = ctx.profile()
    print('per-pass us  trace:', [round(x * 1000 / passes, 1) for x in d['trace']])
    print('per-pass us  shade:', [round(x * 1000 / passes, 1) for x in d['shade']])
    print({k: round(v[0] * 1000 / passes, 1) for k, v in pr.items()})
