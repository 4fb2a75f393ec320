#!/usr/bin/env python
"""Builds the LBVH of the 1 003 244-triangle synthetic Cornell scene a few times (profiling target)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('msc-futhark-ray-tracer_b200')
d = np.load(os.path.join(ROOT, 'tests', 'golden', 'scenes', 'cornell.npz'))
k = int(sys.argv[1]) if len(sys.argv) > 1 else 151
st, sm = pkg.scenes.synthetic_cornell(d['tris'], d['tri_mats'], k)
ctx = pkg.Context()
s = pkg.State.init(ctx, st, sm, d['mats'], 64, 64)
print('tris', len(st), 'build ms (mean of 5):', s.bvh_rebuild_ms(5))
