#!/usr/bin/env python
"""Throughput of N 1080p CornellBox passes for different path_len knobs (how much the sparse tail bounces cost)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('msc-futhark-ray-tracer_b200')
d = np.load(os.path.join(ROOT, 'tests', 'golden', 'scenes', 'cornell.npz'))
ctx = pkg.Context()
for pl in (16, 8, 5, 4, 3):
    ctx.set_path_len(pl)
    s = pkg.State.init(ctx, d['tris'], d['tri_mats'], d['mats'], 1080, 1920)
    h, _, _, _ = s.sample_n_frames_device(16); s.free_f32_3d(h)
    best = 1e9
    for _ in range(5):
        h, _, _, st = s.sample_n_frames_device(64); s.free_f32_3d(h)
        best = min(best, st['device_ms'])
    print('path_len', pl, 'Mpaths/s', round(1080 * 1920 * 64 / best / 1e3, 1), 'ms/pass', round(best / 64, 4), st['vertices'] / st['paths'])
