#!/usr/bin/env python
"""A few sample passes (profiling target for ncu): tools/prof_pass.py [scene] [passes]; scene = a bundled scene name or
`synthetic` (BASELINE config 5: 1 003 244 triangles); LYS_H / LYS_W set the frame (default 1080 x 1920)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('msc-futhark-ray-tracer_b200')
name = sys.argv[1] if len(sys.argv) > 1 else 'cornell'
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
d = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'scenes', ('cornell' if name == 'synthetic' else name) + '.npz')))
if name == 'synthetic':
    d['tris'], d['tri_mats'] = pkg.scenes.synthetic_cornell(d['tris'], d['tri_mats'], int(os.environ.get('LYS_K', 151)))
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
