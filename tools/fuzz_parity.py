#!/usr/bin/env python
"""Randomised parity of the library against the oracle (development / test tool; exit status 1 on any mismatch).

    python tools/fuzz_parity.py lbvh [seed] [scenes]    hostile geometry: inf / NaN / denormal / 1e30 / 1e-30 coordinates, exact
                                                        duplicates, flat and identical triangles, sizes around the warp, tile and
                                                        chunk boundaries -> every LBVH array, bit for bit (NaN payloads included)
    python tools/fuzz_parity.py soup [seed] [scenes]    lysref.objwriter.random_soup scenes (whole uber-BSDF parameter space, ten
                                                        lights) in a random camera preset and frame size -> per-vertex radiance,
                                                        distance, channel and two accumulated passes, bit for bit
    python tools/fuzz_parity.py keys [seed] [sessions]  random host sessions (key events, resizes, steps) -> state scalars, image, ARGB frame
Runs on the GPU by default; with LYS_LIBTRACER / LYS_ALLOW_EMULATOR=1 set by tests/test_simt_emu.py it runs the same CUDA sources
on the CPU SIMT emulator."""
import importlib
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
pkg = importlib.import_module('msc-futhark-ray-tracer_b200')
from lysref import oracle, objwriter  # noqa: E402


NAN_PAYLOAD_ONLY = [0]      # arrays that differed only in the sign / payload bits of NaNs (reported, not counted)


def same_bits(a, b):
    """Bit equality; for f32 a NaN equals any NaN.  The sign and payload of a NaN produced by an invalid operation
    (inf - inf for an all-infinite triangle) belong to the arithmetic unit, not to the algorithm: x86 SSE writes the
    default NaN 0xFFC00000, NVIDIA GPUs write 0x7FFFFFFF, and the reference itself would differ in the same way between
    its `c` and `cuda` backends.  Everything that is not a NaN must match bit for bit, and NaNs must sit at the same places."""
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype != np.float32:
        return np.array_equal(a, b)
    ua, ub = a.view(np.uint32), b.view(np.uint32)
    diff = ua != ub
    if not diff.any():
        return True
    both_nan = np.isnan(a) & np.isnan(b)
    if (diff & ~both_nan).any():
        return False
    if NAN_PAYLOAD_ONLY[0] == 0:
        i = np.flatnonzero(diff.reshape(-1))[0]
        print('note: NaN payloads differ (first: oracle 0x%08x, library 0x%08x); positions agree' % (ua.reshape(-1)[i], ub.reshape(-1)[i]), flush=True)
    NAN_PAYLOAD_ONLY[0] += 1
    return True


def hostile_triangles(rng, it):
    n = int(rng.choice([2, 3, 5, 31, 32, 33, 100, 257, 600, 1500]))
    kind = it % 10
    t = rng.random((n, 3, 3))
    if kind == 1:
        t = t * 1e-30
    elif kind == 2:
        t = t * 1e30
    elif kind == 3:
        t[rng.integers(0, n, max(1, n // 10))] = np.inf
    elif kind == 4:
        t[rng.integers(0, n, max(1, n // 10)), rng.integers(0, 3), rng.integers(0, 3)] = np.nan
    elif kind == 5:
        t = np.round(t * 4) / 4                                  # many exact duplicates, coplanar
    elif kind == 6:
        t[:, :, 0] = 0.5                                         # zero extent on x
    elif kind == 7:
        t = np.repeat(t[:1], n, axis=0)                          # all identical: the index tie-break decides the whole tree
    elif kind == 8:
        t = t - 0.5
        t[::2] *= -1e-3                                          # mixed signs and scales
    elif kind == 9:
        t = t * 1e-40                                            # denormals
    with np.errstate(all='ignore'):
        return np.ascontiguousarray(t, np.float32), kind


def fuzz_lbvh(ctx, rng, count):
    mats = np.load(os.path.join(ROOT, 'tests', 'golden', 'scenes', 'cornell.npz'))['mats']
    bad = 0
    for it in range(count):
        t, kind = hostile_triangles(rng, it)
        tm = np.zeros(len(t), np.uint32)
        so, sg = oracle.State.init(t, tm, mats, 4, 4), pkg.State.init(ctx, t, tm, mats, 4, 4)
        bo, bg = so.bvh(), sg.bvh()
        for key in ('bounds', 'morton', 'src_index', 'left', 'right', 'parent', 'node_aabb', 'leaf_aabb'):
            if not same_bits(bo[key], bg[key]):
                print('MISMATCH lbvh scene', it, 'kind', kind, 'n', len(t), key, flush=True)
                bad += 1
                break
        sg.free()
    return bad


def fuzz_soup(ctx, rng, count):
    bad = 0
    for it in range(count):
        seed = int(rng.integers(100, 1 << 30))
        scene = objwriter.random_soup(seed, n_tris=int(rng.integers(40, 400)))
        conf = int(rng.integers(0, 3))
        h, w = int(rng.integers(8, 48)), int(rng.integers(8, 64))
        origin = tuple(float(x) for x in rng.uniform([-0.6, 0.4, -0.6], [0.6, 1.6, 0.9]))
        kw = dict(cam_conf_id=conf, origin=origin, pitch=float(rng.uniform(-0.5, 0.5)), yaw=float(rng.uniform(-3, 3)), seed=int(rng.integers(0, 1000)))
        so, sg = oracle.State.init(*scene, h, w, **kw), pkg.State.init(ctx, *scene, h, w, **kw)
        qo, qg = so.probe_pass(), sg.probe_pass()
        ok = all(same_bits(qo[k], qg[k]) for k in ('radiance', 'distance', 'channel')) and same_bits(so.sample_n_frames(2), sg.sample_n_frames(2))
        if not ok:
            print('MISMATCH soup scene', it, 'seed', seed, 'conf', conf, 'frame', (h, w), flush=True)
            bad += 1
        sg.free()
    return bad


KEYS = [0x20, 0x31, 0x32, 0x61, 0x64, 0x69, 0x6B, 0x6C, 0x6D, 0x6E, 0x6F, 0x70, 0x73, 0x74, 0x77, 0x78, 0x7A,
        0x4000004F, 0x40000050, 0x40000051, 0x40000052, 0x71, 0x0]        # every key of lib.fut:120-185 plus two unbound ones


def fuzz_keys(ctx, rng, count):
    """random host sessions: key-down / key-up events, resizes and steps in any order (the state machine of lib.fut:108-185 and
    state.fut), then the state scalars, the accumulated image and the ARGB frame"""
    d = np.load(os.path.join(ROOT, 'tests', 'golden', 'scenes', 'cornell.npz'))
    scene = (d['tris'], d['tri_mats'], d['mats'])
    bad = 0
    for it in range(count):
        h, w = int(rng.integers(4, 24)), int(rng.integers(4, 32))
        kw = dict(cam_conf_id=int(rng.integers(0, 3)), seed=int(rng.integers(0, 100)))
        so, sg = oracle.State.init(*scene, h, w, **kw), pkg.State.init(ctx, *scene, h, w, **kw)
        trace = []
        errored = [False]
        free_old = [False]
        saved = None                        # an older pair of states, stepped again later (the library may have run ahead of it)

        def step_both():
            """one step on both sides; accumulating onto an image of another shape is a run-time size error in the reference
            (`img_new :> [m][n]vec3`, integrator.fut:184): the library must report it, the oracle (which aborts) is not asked"""
            nonlocal so, sg
            sc = so.scalars()
            _, _, gw, gh = so.dims()
            if sc['mode'] and sc['n_frames'] > 0 and so.image().shape[:2] != (gh, gw):
                try:
                    sg.step()
                    raise SystemExit('the library accepted an accumulate step onto an image of another shape: %r' % (trace,))
                except pkg.TracerError:
                    errored[0] = True
                    return False
            so = so.step()
            old, sg = sg, sg.step()
            if free_old[0] and not (saved is not None and old is saved[1]):
                old.free()                  # the interactive host's pattern (liblys.c:110): the stepped state is freed at once
            return True
        for _ in range(int(rng.integers(5, 40))):
            r = rng.random()
            if r < 0.55:
                key, e = int(rng.choice(KEYS)), int(rng.random() < 0.15)
                if key == 0x32 and so.scalars()['subsampling'] >= 6:
                    continue
                so, sg = so.key(key, e), sg.key(key, e)
                trace.append(('key', hex(key), e))
            elif r < 0.65:
                h, w = int(rng.integers(4, 24)), int(rng.integers(4, 32))
                so, sg = so.resize(h, w), sg.resize(h, w)
                trace.append(('resize', h, w))
            elif r < 0.80:
                if not step_both():
                    break
                trace.append(('step',))
            elif r < 0.93:                  # a stepping loop as the interactive host runs it: the library starts steps ahead
                free_old[0] = bool(rng.random() < 0.7)
                n = int(rng.integers(2, 7))
                ok = True
                for k in range(n):
                    ok = ok and step_both()
                    if ok and k == 1 and saved is None and not free_old[0] and rng.random() < 0.5:
                        saved = (so, sg)
                    if ok and rng.random() < 0.3:
                        sg.render()         # render + read-back between the steps
                free_old[0] = False
                trace.append(('steps', n))
                if not ok:
                    break
            elif saved is not None:         # back to an older state: what was computed ahead of the newer one is dropped
                so, sg = saved
                saved = None
                trace.append(('back',))
        else:
            step_both()
        if errored[0]:
            sg.free()
            continue
        a, b = so.scalars(), sg.info()
        cam_g = np.array([b['cam_pitch'], b['cam_yaw'], *b['cam_origin'], b['aperture'], b['focal_dist']], np.float32)
        ok = all(a[k] == b[k] for k in ('rng', 'n_frames', 'subsampling', 'render_mode', 'cam_conf_id')) and int(a['mode']) == int(b['mode'])
        ok = ok and same_bits(a['cam'], cam_g) and same_bits(a['ambience'], b['ambience'])
        ok = ok and same_bits(so.image(), sg.image()) and same_bits(so.render(), sg.render())
        if not ok:
            print('MISMATCH session', it, trace, flush=True)
            bad += 1
        sg.free()
    return bad


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else 'lbvh'
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    count = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    with pkg.Context() as ctx, np.errstate(all='ignore'):
        bad = {'lbvh': fuzz_lbvh, 'soup': fuzz_soup, 'keys': fuzz_keys}[what](ctx, rng, count)
    print('%s: %d scenes, %d mismatching (%d arrays equal up to NaN sign / payload)' % (what, count, bad, NAN_PAYLOAD_ONLY[0]), flush=True)
    sys.exit(1 if bad else 0)


if __name__ == '__main__':
    main()
