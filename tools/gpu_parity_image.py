#!/usr/bin/env python
"""Accumulated image of the library against the oracle's at a BASELINE.json size, bit for bit (development / test tool).

    python tools/gpu_parity_image.py <scene> <h> <w> <passes>      scene = a bundled scene name or `synthetic` (k = 151: 1 003 244 triangles)

Kernel variants are selected by the LYS_* environment knobs (read once per process), which is why the variant tests run this
in a subprocess.  Prints one JSON line: img_bits, bad_pixels, oracle / library seconds."""
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


def load(name):
    d = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'scenes', ('cornell' if name == 'synthetic' else name) + '.npz')))
    if name == 'synthetic':
        d['tris'], d['tri_mats'] = pkg.scenes.synthetic_cornell(d['tris'], d['tri_mats'], 151)
    return d['tris'], d['tri_mats'], d['mats']


def compare(ctx, name, h, w, passes):
    t, tm, m = load(name)
    kw = {'origin': (0.0, 0.8, 0.6)} if name == 'mirrorbox' else {}      # the default camera is outside MirrorBox's front wall
    oracle.set_threads(len(os.sched_getaffinity(0)))
    sg = pkg.State.init(ctx, t, tm, m, h, w, **kw)
    t0 = time.time(); ig = sg.sample_n_frames(passes); tg = time.time() - t0
    so = oracle.State.init(t, tm, m, h, w, **kw)
    t0 = time.time(); io = so.sample_n_frames(passes); to = time.time() - t0
    bad = (io.view(np.uint32) != ig.view(np.uint32)).any(axis=2)
    sg.free()
    return {'scene': name, 'res': '%dx%d' % (w, h), 'passes': passes, 'img_bits': bool(not bad.any()), 'bad_pixels': int(bad.sum()),
            'nonzero': bool(ig.max() > 0), 'oracle_s': round(to, 2), 'library_s': round(tg, 2)}


if __name__ == '__main__':
    name, h, w, passes = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    with pkg.Context() as ctx:
        print(json.dumps(compare(ctx, name, h, w, passes)), flush=True)
