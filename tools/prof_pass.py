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
s = pkg.State.init(ctx, d['tris'], d['tri_mats'], d['mats'], 1080, 1920, **kw)
h, ptr, shape, st = s.sample_n_frames_device(passes)
ctx.sync()
print(name, shape, st)
