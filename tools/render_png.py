#!/usr/bin/env python
"""Renders the four bundled scenes on the GPU (futhark_entry_sample_n_frames) to small PNGs for a visual sanity check.
The images are the raw accumulated [0,1]-clamped framebuffer exactly as the reference's commented-out PNG path would
save it (demo-save/src/main.rs:43-48: clamp(0,1) * 255.99)."""
import importlib, os, sys
import numpy as np
from PIL import Image
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('msc-futhark-ray-tracer_b200')
out = os.path.join(ROOT, 'gpurun_out', 'renders')
os.makedirs(out, exist_ok=True)
ctx = pkg.Context()
for name, origin, passes in (('cornell', (0.0, 0.8, 1.8), 512), ('mirrorbox', (0.0, 0.8, 0.6), 512), ('spectrumsphere', (0.0, 0.8, 1.8), 1024),
                             ('spectrumspherehigh', (0.0, 0.8, 1.8), 512)):
    d = np.load(os.path.join(ROOT, 'tests', 'golden', 'scenes', name + '.npz'))
    s = pkg.State.init(ctx, d['tris'], d['tri_mats'], d['mats'], 360, 480, origin=origin)
    img = s.sample_n_frames(passes)
    Image.fromarray((np.clip(img, 0, 1) * 255.99).astype(np.uint8)).save(os.path.join(out, name + '.png'))
    print(name, img.shape, float(img.mean()))
# LIDAR distance view (cam_conf_id 2) of SpectrumSphere
d = np.load(os.path.join(ROOT, 'tests', 'golden', 'scenes', 'spectrumsphere.npz'))
s = pkg.State.init(ctx, d['tris'], d['tri_mats'], d['mats'], 360, 480, cam_conf_id=2)
img = s.key(pkg.KEY['m']).step().step().step().step().image()
Image.fromarray((np.clip(img, 0, 1) * 255.99).astype(np.uint8)).save(os.path.join(out, 'spectrumsphere_lidar_distance.png'))
