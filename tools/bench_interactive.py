#!/usr/bin/env python
"""Frame rate of the reference's interactive loop (demo-interactive/liblys.c:104-123) through the C ABI: per frame
futhark_entry_step (ONE sample pass + accumulate) + futhark_entry_render (ARGB pack) + futhark_values_i32_2d (blocking D2H of
the [h][w] i32 frame into host memory) + the two frees.  Nothing is pipelined across frames: a frame is displayed before the
next one starts, as in the SDL loop.  Wall-clock time (host calls and copies are part of a frame).  Prints one JSON line.

    python tools/bench_interactive.py [scene] [width] [height] [frames]

The only published interactive figure of the reference is a HUD screenshot (prism-dispersion.png: "FPS: 1" at 725x665 with 16
samples per frame on an unnamed GPU); this is the same loop on a B200.  Not the driver's bench (that is bench.py)."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('msc-futhark-ray-tracer_b200')


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'cornell'
    w = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
    h = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
    frames = int(sys.argv[4]) if len(sys.argv) > 4 else 300
    d = np.load(os.path.join(ROOT, 'tests', 'golden', 'scenes', name + '.npz'))
    with pkg.Context() as ctx:
        s = pkg.State.init(ctx, d['tris'], d['tri_mats'], d['mats'], h, w).resize(h, w).key(pkg.KEY['m'])   # accumulate on
        out = np.empty((h, w), np.int32)                        # the host's frame buffer (liblys.c:66 allocates it once per window size)
        for phase, n in (('warm-up', 20), ('timed', frames)):
            t0 = time.perf_counter()
            for _ in range(n):
                nxt = s.step()
                s.free()
                s = nxt
                s.render(out)                                    # render + values_i32_2d + free_i32_2d
            dt = time.perf_counter() - t0
        info = s.info()
        print(json.dumps({'loop': 'step + render + values_i32_2d per frame (liblys.c:104-123)', 'scene': name, 'res': '%dx%d' % (w, h),
                          'frames': frames, 'fps': round(frames / dt, 1), 'ms_per_frame': round(dt / frames * 1e3, 3),
                          'mpaths_s': round(w * h * frames / dt / 1e6, 1), 'accumulated_frames': int(info['n_frames']),
                          'd2h_bytes_per_frame': w * h * 4, 'argb_checksum': int((out.view(np.uint32) & 0xFFFFFF).sum())}), flush=True)
        s.free()


if __name__ == '__main__':
    main()
