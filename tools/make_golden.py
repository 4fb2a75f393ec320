#!/usr/bin/env python
"""Regenerates tests/golden/: scene arrays from the reference's bundled assets and oracle-derived vectors.

Run in the build container (needs /root/reference).  The GPU box has no /root/reference, so the loader OUTPUT
(tri [n,3,3] f32, tri_mats [n] u32, mats [m,28] f32) is committed as compressed .npz, not the OBJ/MTL sources.
The arrays come from the product's C++ loader (libljus.so) and are cross-checked against the independent
Python loader in tests/lysref/loader.py."""
import importlib
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
pkg = importlib.import_module('msc-futhark-ray-tracer_b200')
from lysref import loader, oracle  # noqa: E402

ASSETS = '/root/reference/assets'
SCENES = {'cornell': 'CornellBox-Original.obj', 'mirrorbox': 'MirrorBox.obj', 'spectrumsphere': 'SpectrumSphere.obj',
          'spectrumspherehigh': 'SpectrumSphereHigh.obj'}


def main():
    pkg.build()
    out = os.path.join(ROOT, 'tests', 'golden', 'scenes')
    os.makedirs(out, exist_ok=True)
    for name, fn in SCENES.items():
        t, tm, m = pkg.load_obj(os.path.join(ASSETS, fn))
        t2, tm2, m2 = loader.load_obj(os.path.join(ASSETS, fn))
        assert np.array_equal(t.view(np.uint32), t2.view(np.uint32)) and np.array_equal(tm, tm2) and np.array_equal(m.view(np.uint32), m2.view(np.uint32)), name
        np.savez_compressed(os.path.join(out, name + '.npz'), tris=t, tri_mats=tm, mats=m)
        print(name, t.shape, m.shape)
    # oracle-derived regression vectors (self-authored pins; the reference ships none)
    vec = {}
    for name in SCENES:
        d = np.load(os.path.join(out, name + '.npz'))
        s = oracle.State.init(d['tris'], d['tri_mats'], d['mats'], 48, 64)
        b = s.bvh()
        for k in ('bounds', 'morton', 'src_index', 'left', 'right', 'parent', 'node_aabb'):
            vec['%s_%s' % (name, k)] = b[k]
        pr = s.probe_primary()
        vec[name + '_first_hit_src'] = pr['src_tri']
        vec[name + '_first_hit_t'] = pr['t']
        vec[name + '_img3'] = s.sample_n_frames(3)
        vec[name + '_lights'] = s.light_indices()
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'oracle_vectors.npz'), **vec)
    print('wrote oracle_vectors.npz with', len(vec), 'arrays')


if __name__ == '__main__':
    main()
