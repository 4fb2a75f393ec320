#!/usr/bin/env python
"""Quick GPU-vs-oracle parity sweep (development tool; the formal version lives in tests/)."""
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


def beq(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.dtype == np.float32:
        return np.array_equal(a.view(np.uint32), b.view(np.uint32))
    return np.array_equal(a, b)


def main():
    res = {}
    ctx = pkg.Context()
    # math contract
    rng = np.random.default_rng(1)
    for fn, lo, hi in (('sin', 0, 6.3), ('cos', 0, 6.3), ('exp', -110, 89), ('log', 0, 4), ('pow5', 0, 1), ('acos', -1, 1), ('probit', 0, 1)):
        x = rng.uniform(lo, hi, 1 << (12 if os.environ.get('LYS_EMU_FAST_MATH_SWEEP') else 20)).astype(np.float32)   # the CPU emulator runs one fiber per element
        res['math_' + fn] = bool(beq(ctx.eval_math(fn, x), oracle.eval_math(fn, x)))
    scenes = sys.argv[1:] or ['cornell', 'mirrorbox', 'spectrumsphere', 'spectrumspherehigh']
    for name in scenes:
        d = np.load(os.path.join(ROOT, 'tests', 'golden', 'scenes', name + '.npz'))
        h, w = 96, 128
        kw = {'origin': (0.0, 0.8, 0.6)} if name == 'mirrorbox' else {}
        so = oracle.State.init(d['tris'], d['tri_mats'], d['mats'], h, w, **kw)
        sg = pkg.State.init(ctx, d['tris'], d['tri_mats'], d['mats'], h, w, **kw)
        bo, bg = so.bvh(), sg.bvh()
        r = {k: bool(beq(bo[k], bg[k])) for k in ('bounds', 'morton', 'src_index', 'left', 'right', 'parent', 'node_aabb', 'leaf_aabb')}
        if not r['node_aabb']:
            r['node_aabb_mismatch'] = int((bo['node_aabb'].view(np.uint32) != bg['node_aabb'].view(np.uint32)).any(axis=1).sum())
        r['lights'] = bool(beq(so.light_indices(), sg.light_indices()))
        po, pg = so.probe_primary(), sg.probe_primary()
        r['first_hit_leaf'] = bool(beq(po['leaf'], pg['leaf']))
        r['first_hit_leaf_mismatch'] = int((po['leaf'] != pg['leaf']).sum())
        r['first_hit_t'] = bool(beq(po['t'], pg['t']))
        qo, qg = so.probe_pass(), sg.probe_pass()
        r['pass_radiance_bits'] = bool(beq(qo['radiance'], qg['radiance']))
        bad = (qo['radiance'].view(np.uint32) != qg['radiance'].view(np.uint32)).any(axis=2)
        r['pass_radiance_bad_pixels'] = int(bad.sum())
        r['pass_distance_bits'] = bool(beq(qo['distance'], qg['distance']))
        r['pass_channel'] = bool(beq(qo['channel'], qg['channel']))
        t0 = time.time(); io = so.sample_n_frames(4); t_or = time.time() - t0
        t0 = time.time(); ig = sg.sample_n_frames(4); t_gpu = time.time() - t0
        r['img4_bits'] = bool(beq(io, ig))
        r['img4_maxabs'] = float(np.nanmax(np.abs(io - ig)))
        r['t_oracle_s'], r['t_gpu_s'] = t_or, t_gpu
        # step / accumulate / render through the interactive entry points
        so2, sg2 = so.key(0x6D), sg.key(0x6D)
        for _ in range(3):
            so2, sg2 = so2.step(), sg2.step()
        r['step3_img_bits'] = bool(beq(so2.image(), sg2.image()))
        r['render_bits'] = bool(beq(so2.render(), sg2.render()))
        r['build_ms'] = sg.bvh_rebuild_ms(3)
        res[name] = r
        print(name, json.dumps(r), flush=True)
    print(json.dumps({k: v for k, v in res.items() if k.startswith('math_')}))
    out = os.path.join(ROOT, 'gpurun_out')
    os.makedirs(out, exist_ok=True)
    json.dump(res, open(os.path.join(out, 'parity_quick.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
