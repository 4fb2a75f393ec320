"""GPU parity: the sm_100a library, called through its C ABI, against the CPU oracle on identical inputs.
Bit-exact everywhere (integers, indices AND f32 results): the arithmetic contract makes that a meaningful bar."""
import importlib
import os
import sys
import numpy as np
import pytest
from conftest import SCENE_NAMES, bits_equal

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K = dict(SPACE=0x20, K1=0x31, K2=0x32, m=0x6D, t=0x74, p=0x70, w=0x77, UP=0x40000052, i=0x69)


# The default camera (0, 0.8, 1.8) is OUTSIDE MirrorBox's front wall and sees nothing; its tests look from inside.
ORIGIN = {'mirrorbox': (0.0, 0.8, 0.6)}


def both(orc, pkg, gpu, scene, h, w, name=None, **kw):
    t, tm, m = scene
    if name in ORIGIN:
        kw.setdefault('origin', ORIGIN[name])
    return orc.State.init(t, tm, m, h, w, **kw), pkg.State.init(gpu, t, tm, m, h, w, **kw)


def test_math_contract_bits(orc, gpu):
    rng = np.random.default_rng(1)
    for fn, lo, hi in (('sin', 0, 6.3), ('cos', 0, 6.3), ('exp', -110, 89), ('log', 0, 4), ('pow5', 0, 1), ('acos', -1, 1), ('probit', 0, 1)):
        x = rng.uniform(lo, hi, 1 << 20).astype(np.float32)
        x[:4] = [lo, hi, 0.0, 1.0]
        assert bits_equal(gpu.eval_math(fn, x), orc.eval_math(fn, x)), fn


@pytest.mark.parametrize('name', SCENE_NAMES)
def test_lbvh_bit_exact(orc, pkg, gpu, scenes, name):
    so, sg = both(orc, pkg, gpu, scenes[name], 8, 8)
    bo, bg = so.bvh(), sg.bvh()
    for k in ('bounds', 'morton', 'src_index', 'left', 'right', 'parent', 'node_aabb', 'leaf_aabb'):
        assert bits_equal(bo[k], bg[k]), k
    assert bits_equal(so.light_indices(), sg.light_indices())


@pytest.mark.parametrize('k', [2, 9, 40])
def test_lbvh_synthetic_bit_exact(orc, pkg, gpu, scenes, k):
    t, tm, m = scenes['cornell']
    st, sm = pkg.scenes.synthetic_cornell(t, tm, k)
    so, sg = orc.State.init(st, sm, m, 8, 8), pkg.State.init(gpu, st, sm, m, 8, 8)
    bo, bg = so.bvh(), sg.bvh()
    for key in ('bounds', 'morton', 'src_index', 'left', 'right', 'parent', 'node_aabb'):
        assert bits_equal(bo[key], bg[key]), key


def test_lbvh_edge_cases(orc, pkg, gpu, scenes):
    m = scenes['cornell'][2]
    rng = np.random.default_rng(7)
    cases = {
        'two': rng.random((2, 3, 3)),
        'duplicates': np.repeat(rng.random((3, 3, 3)), 50, axis=0),                  # equal Morton codes -> index tie-break
        'flat_z': np.concatenate([rng.random((300, 3, 2)), np.zeros((300, 3, 1))], axis=2),   # zero extent on z -> NaN axis
        'ragged_4097': rng.random((4097, 3, 3)) * 10 - 5,                            # one key past a sort tile
        'clustered': np.concatenate([rng.random((500, 3, 3)) * 1e-3, rng.random((500, 3, 3)) * 1e-3 + 100]),
        'sweep': np.cumsum(rng.random((3000, 1, 3)), axis=0) + rng.random((3000, 3, 3)),  # every triangle extends the bounds
    }
    for name, tri in cases.items():
        tri = np.ascontiguousarray(tri, np.float32)
        tm = np.zeros(len(tri), np.uint32)
        so, sg = orc.State.init(tri, tm, m, 4, 4), pkg.State.init(gpu, tri, tm, m, 4, 4)
        bo, bg = so.bvh(), sg.bvh()
        for key in ('bounds', 'morton', 'src_index', 'left', 'right', 'parent', 'node_aabb'):
            assert bits_equal(bo[key], bg[key]), (name, key)


def test_refit_converged_mode(orc, pkg, gpu, scenes):
    t, tm, m = scenes['spectrumsphere']
    orc.set_refit_mode(1)
    gpu.set_refit_mode(1)
    try:
        bo, bg = orc.State.init(t, tm, m, 4, 4).bvh(), pkg.State.init(gpu, t, tm, m, 4, 4).bvh()
    finally:
        orc.set_refit_mode(0)
        gpu.set_refit_mode(0)
    assert bits_equal(bo['node_aabb'], bg['node_aabb'])
    # converged boxes enclose their children
    for side in ('left', 'right'):
        c = bg[side]
        child = np.where(c[:, None] >= 0, bg['node_aabb'][np.maximum(c, 0)], bg['leaf_aabb'][np.maximum(~c, 0)])
        lo_p, hi_p = bg['node_aabb'][:, :3] - bg['node_aabb'][:, 3:], bg['node_aabb'][:, :3] + bg['node_aabb'][:, 3:]
        lo_c, hi_c = child[:, :3] - child[:, 3:], child[:, :3] + child[:, 3:]
        assert (lo_c >= lo_p - 1e-5).all() and (hi_c <= hi_p + 1e-5).all()


def test_refit_literal_sweeps_equal_worklists(pkg, gpu, scenes):
    """refit mode 2 (the reference's own Jacobi sweeps, also the overflow fallback) == mode 0 (crown worklists)."""
    t, tm, m = scenes['cornell']
    st, sm = pkg.scenes.synthetic_cornell(t, tm, 40)
    for tris, mats_ix, mm in ((scenes['spectrumspherehigh'][0], scenes['spectrumspherehigh'][1], scenes['spectrumspherehigh'][2]), (st, sm, m)):
        a = pkg.State.init(gpu, tris, mats_ix, mm, 4, 4).bvh()['node_aabb']
        gpu.set_refit_mode(2)
        try:
            b = pkg.State.init(gpu, tris, mats_ix, mm, 4, 4).bvh()['node_aabb']
        finally:
            gpu.set_refit_mode(0)
        assert bits_equal(a, b)


@pytest.mark.parametrize('name', SCENE_NAMES)
def test_first_hit_bit_exact(orc, pkg, gpu, scenes, name):
    so, sg = both(orc, pkg, gpu, scenes[name], 120, 160, name=name)
    po, pg = so.probe_primary(), sg.probe_primary()
    assert bits_equal(po['leaf'], pg['leaf']) and bits_equal(po['src_tri'], pg['src_tri']) and bits_equal(po['t'], pg['t'])
    assert (po['leaf'] >= 0).mean() > 0.5


@pytest.mark.parametrize('name', SCENE_NAMES)
@pytest.mark.parametrize('conf', [0, 1, 2])
def test_pass_radiance_bit_exact(orc, pkg, gpu, scenes, name, conf):
    so, sg = both(orc, pkg, gpu, scenes[name], 72, 96, name=name, cam_conf_id=conf)
    qo, qg = so.probe_pass(), sg.probe_pass()
    assert bits_equal(qo['channel'], qg['channel'])
    assert bits_equal(qo['distance'], qg['distance'])
    assert bits_equal(qo['radiance'], qg['radiance'])          # stronger than the 1e-4 relative bar of the north star
    assert qo['radiance'].max() > 0


@pytest.mark.parametrize('seed', [1, 2, 3])
def test_random_soup_scenes_bit_exact(orc, pkg, gpu, seed):
    """Fuzzed scenes (lysref.objwriter.random_soup): the whole uber-BSDF parameter space, ten light triangles (one of zero
    area), duplicated and degenerate triangles, in all three camera presets -- LBVH, first hits, per-vertex radiance /
    distance / channel of one pass, three accumulated passes and the LIDAR point cloud."""
    from lysref import objwriter
    scene = objwriter.random_soup(seed)
    for conf in (0, 1, 2):
        so, sg = both(orc, pkg, gpu, scene, 60, 84, cam_conf_id=conf, origin=(0.0, 1.0, 0.9))
        if conf == 0:
            bo, bg = so.bvh(), sg.bvh()
            for key in ('bounds', 'morton', 'src_index', 'left', 'right', 'parent', 'node_aabb', 'leaf_aabb'):
                assert bits_equal(bo[key], bg[key]), (seed, key)
            assert bits_equal(so.light_indices(), sg.light_indices()) and len(sg.light_indices()) == 10
            po, pg = so.probe_primary(), sg.probe_primary()
            assert bits_equal(po['leaf'], pg['leaf']) and bits_equal(po['t'], pg['t'])
        qo, qg = so.probe_pass(), sg.probe_pass()
        for key in ('radiance', 'distance', 'channel'):
            assert bits_equal(qo[key], qg[key]), (seed, conf, key)
        assert (qo['radiance'] > 0).any()
        assert bits_equal(so.sample_n_frames(3), sg.sample_n_frames(3)), (seed, conf)
    assert bits_equal(so.sample_points_n(3)[1], sg.sample_points_n(3)[1]), seed


def test_random_rays(orc, pkg, gpu, scenes):
    from test_oracle_scene import random_rays
    for name in ('cornell', 'spectrumspherehigh'):
        so, sg = both(orc, pkg, gpu, scenes[name], 4, 4)
        rays = random_rays(50000, seed=11)
        lo, to = so.closest_hits(rays)
        lg, tg = sg.trace_closest(rays)
        assert bits_equal(lo, lg) and bits_equal(to, tg)
        tmax = np.random.default_rng(2).uniform(0.0, 2.5, len(rays)).astype(np.float32)
        assert bits_equal(so.any_hits(rays, tmax), sg.trace_any(rays, tmax))


def test_material_probe_bits(orc, gpu, scenes):
    rng = np.random.default_rng(5)
    out_o = np.empty(9, np.float32)
    rows = np.concatenate([scenes['spectrumsphere'][2], scenes['mirrorbox'][2]])
    for it in range(300):
        row = rows[it % len(rows)].copy()
        if it % 3 == 0:
            row[12:16] = [rng.random(), rng.random(), 1 + rng.random(), rng.random()]
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        wo = rng.normal(size=3); wo /= np.linalg.norm(wo)
        wi = rng.normal(size=3); wi /= np.linalg.norm(wi)
        wl = float(rng.uniform(380, 700)); seed = int(rng.integers(1, 2 ** 31 - 2))
        a = [np.ascontiguousarray(v, np.float32) for v in (row, wo, wi, n)]
        orc.lib().orc_material_probe(a[0], wl, a[1], a[2], a[3], seed, out_o)
        out_g = gpu.material_probe(a[0], wl, a[1], a[2], a[3], seed)
        assert bits_equal(out_o, out_g), it


@pytest.mark.parametrize('name', ['cornell', 'mirrorbox', 'spectrumsphere'])
def test_entry_points_bit_exact(orc, pkg, gpu, scenes, name):
    so, sg = both(orc, pkg, gpu, scenes[name], 60, 80, name=name)
    io5 = so.sample_n_frames(5)
    assert bits_equal(io5, sg.sample_n_frames(5)) and io5.max() > 0
    so, sg = so.key(K['m']), sg.key(K['m'])
    for _ in range(4):
        so, sg = so.step(), sg.step()
    assert bits_equal(so.image(), sg.image()) and bits_equal(so.render(), sg.render())
    io, ig = so.scalars(), sg.info()
    assert io['rng'] == ig['rng'] and io['n_frames'] == ig['n_frames'] == 4
    # camera keys, subsampling, sky, aperture (thin lens -> sin/cos on the device), resize
    for seq in ([K['w'], K['UP']], [K['K2']], [K['p']], [K['i'], K['i']], [K['t']], [K['t'], K['t']]):
        a, b = so, sg
        for k in seq:
            a, b = a.key(k), b.key(k)
        a, b = a.step().step(), b.step().step()
        assert bits_equal(a.image(), b.image()), seq
        assert bits_equal(a.render(), b.render()), seq
    a, b = so.resize(33, 47).step(), sg.resize(33, 47).step()
    assert bits_equal(a.image(), b.image()) and a.image().shape == (33, 47, 3)


def test_stepping_loop_runs_ahead_bit_exact(orc, pkg, gpu, scenes):
    """The interactive host's loop (liblys.c:104-123: step, free the stepped state, render + read back, repeat).  Once the host
    steps the state the previous step returned, the library computes steps AHEAD on its pass-slot streams
    (futhark_entry_step); what the host sees must not depend on that: every frame equals the oracle's, also when the loop is
    interrupted by a key event (steps computed ahead are dropped), resumed, and when an older state is stepped again."""
    so, sg = both(orc, pkg, gpu, scenes['cornell'], 90, 120)
    so, sg = so.key(K['m']), sg.key(K['m'])
    frame = np.empty((90, 120), np.int32)
    keep = None
    for k in range(14):
        so_old, so = so, so.step()
        old, sg = sg, sg.step()
        if k == 5:
            keep = (so_old, old)                              # an older state the library has already run ahead of
        else:
            old.free()
        sg.render(frame)
        if k in (2, 9, 13):
            assert bits_equal(so.render(), frame), k
    assert bits_equal(so.image(), sg.image()) and so.scalars()['n_frames'] == sg.info()['n_frames'] == 14
    so, sg = so.key(K['w']), sg.key(K['w'])                   # camera moves: n_frames back to 0, ahead steps dropped
    for k in range(5):
        so = so.step()
        old, sg = sg, sg.step()
        old.free()
    assert bits_equal(so.image(), sg.image()) and bits_equal(so.render(), sg.render())
    so, sg = keep                                             # back to the state saved in the first loop
    for k in range(3):
        so, sg = so.step(), sg.step()
    assert bits_equal(so.image(), sg.image()) and so.scalars()['rng'] == sg.info()['rng']


def test_sample_points_bit_exact(orc, pkg, gpu, scenes):
    for name in ('cornell', 'spectrumsphere'):
        so, sg = both(orc, pkg, gpu, scenes[name], 48, 64, cam_conf_id=2)     # demo-save uses cam_conf_id 2 (wrapper.rs:50)
        (no, po), (ng, pg) = so.sample_points_n(4), sg.sample_points_n(4)
        assert bits_equal(po, pg) and no.scalars()['rng'] == ng.info()['rng']
        assert (po[..., 3] > 0).any()


def test_path_len_knob(orc, pkg, gpu, scenes):
    so, sg = both(orc, pkg, gpu, scenes['mirrorbox'], 40, 40, name='mirrorbox')
    orc.set_path_len(5)
    gpu.set_path_len(5)
    try:
        assert bits_equal(so.probe_pass()['radiance'], sg.probe_pass()['radiance'])
        assert bits_equal(so.sample_n_frames(3), sg.sample_n_frames(3))
    finally:
        orc.set_path_len(16)
        gpu.set_path_len(16)


def test_edge_configurations(orc, pkg, gpu, scenes):
    """Ragged / degenerate inputs: 1x1 and odd-sized frames, odd subsampling, a scene without lights (direct_radiance
    consumes no draws, direct.fut:116-117), sky ambience, path_len 1, huge and tiny coordinates."""
    t, tm, m = scenes['cornell']
    for h, w in ((1, 1), (1, 37), (33, 1), (67, 129)):
        so, sg = both(orc, pkg, gpu, (t, tm, m), h, w)
        assert bits_equal(so.sample_n_frames(3), sg.sample_n_frames(3)), (h, w)
    so, sg = both(orc, pkg, gpu, (t, tm, m), 45, 71)
    so, sg = so.key(0x32).key(0x32).key(0x6D), sg.key(0x32).key(0x32).key(0x6D)          # subsampling 3: grid 15 x 24
    so, sg = so.step().step(), sg.step().step()
    assert so.image().shape == (15, 24, 3) and bits_equal(so.image(), sg.image()) and bits_equal(so.render(), sg.render())
    dark = m.copy(); dark[:, 16:] = np.tile(np.array([-1, 0], np.float32), 6)             # no emissive material -> no lights
    so, sg = both(orc, pkg, gpu, (t, tm, dark), 40, 56)
    assert sg.info()['n_lights'] == 0
    so, sg = so.key(0x70), sg.key(0x70)                                                   # sky on: radiance only from misses
    qo, qg = so.probe_pass(), sg.probe_pass()
    assert bits_equal(qo['radiance'], qg['radiance']) and qo['radiance'].max() > 0
    orc.set_path_len(1); gpu.set_path_len(1)
    try:
        so, sg = both(orc, pkg, gpu, (t, tm, m), 40, 56)
        assert bits_equal(so.sample_n_frames(2), sg.sample_n_frames(2))
    finally:
        orc.set_path_len(16); gpu.set_path_len(16)
    for scale, off in ((1e4, 3e5), (1e-3, 0.0)):
        ts = (t * np.float32(scale) + np.float32(off)).astype(np.float32)
        org = tuple(np.float32(v) * np.float32(scale) + np.float32(off) for v in (0.0, 0.8, 1.8))
        so, sg = both(orc, pkg, gpu, (ts, tm, m), 36, 48, origin=org)
        bo, bg = so.bvh(), sg.bvh()
        for k in ('bounds', 'morton', 'src_index', 'left', 'right', 'node_aabb'):
            assert bits_equal(bo[k], bg[k]), (scale, k)
        assert bits_equal(so.probe_pass()['radiance'], sg.probe_pass()['radiance']), scale


def test_row_partition_sums_to_full_image(pkg, gpu, scenes):
    """Multi-GPU split emulated on one device: ranks 0..2 of a world of 3 each render their rows; the sum equals
    the single-GPU image bit-for-bit (every pixel is non-zero on exactly one rank)."""
    t, tm, m = scenes['spectrumsphere']
    full = pkg.State.init(gpu, t, tm, m, 50, 64).sample_n_frames(3)
    acc = np.zeros_like(full)
    for r in range(3):
        c = pkg.Context()
        c.set_partition(r, 3)
        part = pkg.State.init(c, t, tm, m, 50, 64).sample_n_frames(3)
        assert not part[np.arange(50) % 3 != r].any()
        acc += part
        c.close()
    assert bits_equal(acc, full)


def test_generated_surface_extras(pkg, gpu, scenes):
    """The rest of a `futhark cuda --library` header (tracer.h, last section): raw device access, profiling report."""
    L, c = gpu._L, gpu._ctx
    a = np.arange(2 * 3 * 5, dtype=np.float32).reshape(2, 3, 5)
    arr = L.futhark_new_f32_3d(c, a.ctypes.data, 2, 3, 5)
    dptr = L.futhark_values_raw_f32_3d(c, arr)
    assert dptr != 0
    cp = L.futhark_new_raw_f32_3d(c, dptr, 5 * 4, 1, 5, 5)               # 25 elements starting one row (20 bytes) in
    out = np.empty((1, 5, 5), np.float32)
    gpu.check(L.futhark_values_f32_3d(c, cp, out.ctypes.data), 'values')
    assert np.array_equal(out.reshape(-1), a.reshape(-1)[5:30])
    assert L.futhark_values_raw_f32_3d(c, cp) != dptr                      # new_raw copies, it does not alias
    L.futhark_free_f32_3d(c, cp); L.futhark_free_f32_3d(c, arr)
    u = np.arange(7, dtype=np.uint32)
    ua = L.futhark_new_u32_1d(c, u.ctypes.data, 7)
    ub = L.futhark_new_raw_u32_1d(c, L.futhark_values_raw_u32_1d(c, ua), 0, 7)
    back = np.empty(7, np.uint32)
    gpu.check(L.futhark_values_u32_1d(c, ub, back.ctypes.data), 'values')
    assert np.array_equal(back, u)
    L.futhark_free_u32_1d(c, ua); L.futhark_free_u32_1d(c, ub)
    t, tm, m = scenes['cornell']
    with pkg.Context(profiling=True) as pc:                                # futhark_context_config_set_profiling
        s = pkg.State.init(pc, t, tm, m, 48, 64)
        img = s.sample_n_frames(3)
        rep = pc.report()
        assert 'kernel launches' in rep and 'trace' in rep and 'shade' in rep and 'accumulate' in rep
        L.futhark_context_pause_profiling(pc._ctx)
        n0 = pc.profile(reset=False)['trace'][1]
        s.sample_n_frames(2)
        assert pc.profile(reset=False)['trace'][1] == n0                   # paused: nothing recorded
        L.futhark_context_unpause_profiling(pc._ctx)
        s.sample_n_frames(1)
        assert pc.profile(reset=False)['trace'][1] > n0
        s.free()
    assert bits_equal(img, pkg.State.init(gpu, t, tm, m, 48, 64).sample_n_frames(3))   # profiling does not change results


def test_error_behaviour(pkg, gpu, scenes):
    t, tm, m = scenes['cornell']
    with pytest.raises(pkg.TracerError, match='at least 2 triangles'):
        pkg.State.init(gpu, t[:1], tm[:1], m, 8, 8)
    with pytest.raises(pkg.TracerError, match='material index'):
        pkg.State.init(gpu, t, tm + 100, m, 8, 8)
    import ctypes as C
    L = gpu._L
    a = L.futhark_new_f32_3d(gpu._ctx, t.ctypes.data_as(C.c_void_p), len(t), 3, 3)
    bad = L.futhark_new_f32_2d(gpu._ctx, m.ctypes.data_as(C.c_void_p), 4, 56)       # wrong inner dimension
    tmv = L.futhark_new_u32_1d(gpu._ctx, tm.ctypes.data_as(C.c_void_p), len(tm))
    org = L.futhark_new_f32_1d(gpu._ctx, np.zeros(3, np.float32).ctypes.data_as(C.c_void_p), 3)
    out = C.c_void_p()
    rc = L.futhark_entry_init(gpu._ctx, C.byref(out), 0, 8, 8, 0, a, tmv, bad, 0.0, 0.0, org)
    assert rc != 0
    msg = C.string_at(L.futhark_context_get_error(gpu._ctx)).decode()
    assert '[m][28]' in msg
    # the context stays usable after an error
    assert pkg.State.init(gpu, t, tm, m, 8, 8).step().image().shape == (8, 8, 3)


def test_full_size_properties(orc, pkg, gpu, scenes):
    """BASELINE sizes, through size-independent properties: determinism, partition additivity, and the oracle on a sample of rows."""
    t, tm, m = scenes['cornell']
    s = pkg.State.init(gpu, t, tm, m, 1080, 1920)
    a, b = s.sample_n_frames(2), s.sample_n_frames(2)
    assert bits_equal(a, b) and np.isfinite(a).all() and a.max() > 0
    so = orc.State.init(t, tm, m, 1080, 1920)
    assert bits_equal(so.sample_n_frames(2), a)                       # 2 x 2M paths: a few seconds on the host


def test_million_triangle_bvh(orc, pkg, gpu, scenes):
    t, tm, m = scenes['cornell']
    st, sm = pkg.scenes.synthetic_cornell(t, tm, 151)
    sg = pkg.State.init(gpu, st, sm, m, 64, 64)
    bg = sg.bvh()
    assert np.all(np.diff(bg['morton'].astype(np.int64)) >= 0)
    assert np.array_equal(np.sort(bg['src_index']), np.arange(len(st)))
    same = bg['morton'][1:] == bg['morton'][:-1]
    assert np.all(bg['src_index'][1:][same] > bg['src_index'][:-1][same])
    so = orc.State.init(st, sm, m, 64, 64)
    bo = so.bvh()
    for key in ('bounds', 'morton', 'src_index', 'left', 'right', 'parent', 'node_aabb'):
        assert bits_equal(bo[key], bg[key]), key
    po, pg = so.probe_primary(), sg.probe_primary()
    assert bits_equal(po['src_tri'], pg['src_tri']) and bits_equal(po['t'], pg['t'])
    assert bits_equal(so.sample_n_frames(2), sg.sample_n_frames(2))


BASELINE_SIZES = [                       # BASELINE.json configs 2-5 at the resolutions they are quoted on
    ('mirrorbox', 1080, 1920, 3),        # camera inside the box: 9 vertices per path, every bounce launch carries a long queue
    ('spectrumsphere', 1080, 1920, 2),
    ('spectrumspherehigh', 1080, 1920, 2),
    ('synthetic', 2160, 3840, 2),        # 1 003 244 triangles at 4K: plain node array, one box stage, the large-scene grid sizing
]


@pytest.mark.parametrize('name,h,w,passes', BASELINE_SIZES, ids=[b[0] for b in BASELINE_SIZES])
def test_baseline_resolution_images_bit_exact(gpu, name, h, w, passes):
    """Queue lengths, the bounce the fused tail starts at, the hits-first order lists and the grid sizes all depend on the
    frame size: the accumulated image at the full BASELINE size against the oracle's, bit for bit (the second and later
    passes run with the queue-length estimates of the first)."""
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import gpu_parity_image
    r = gpu_parity_image.compare(gpu, name, h, w, passes)
    assert r['img_bits'] and r['nonzero'], r


@pytest.mark.parametrize('env,scene', [
    ({'LYS_TAIL_MAX': '100000000'}, 'mirrorbox'),      # deep specular paths through k_tail at 1080p (by default its queues never get short enough)
    ({'LYS_TAIL_MAX': '262144'}, 'synthetic'),         # k_tail on the large scene
], ids=['tail-mirrorbox-1080p', 'tail-synthetic-4k'])
def test_fused_tail_at_baseline_resolution(env, scene):
    import json
    import subprocess
    h, w = (2160, 3840) if scene == 'synthetic' else (1080, 1920)
    e = dict(os.environ)
    e.update(env)
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, 'tools', 'gpu_parity_image.py'), scene, str(h), str(w), '3'], env=e, text=True, timeout=900)
    r = json.loads(out.strip().splitlines()[-1])
    assert r['img_bits'] and r['nonzero'], (env, r)


def test_pass_split_merge_matches_single_gpu(pkg, gpu, scenes):
    """Pass-split multi-GPU frame on one device: two 'ranks' render passes [0, 9) and [9, 18) of one frame with their weights
    applied inside the last accumulate kernel (lys_sample_n_frames_weighted); their sum must agree with the single-GPU
    18-pass image within 1e-4 relative (north star: per-pass radiance tolerance; the running average is order dependent, so
    not bitwise).  Weight 1 must be the unweighted image bit for bit."""
    par = importlib.import_module('msc-futhark-ray-tracer_b200.parallel')
    t, tm, m = scenes['cornell']
    s = pkg.State.init(gpu, t, tm, m, 108, 192)
    total, world = 18, 2
    full = s.sample_n_frames(total)
    ranges = par.pass_ranges(total, world)
    counts = [c for _, c in ranges]
    acc = np.zeros_like(full, dtype=np.float64)
    for first, count in ranges:
        sr = s.advance_rng(first) if first else s
        hnd, _, shape, _ = sr.sample_n_frames_device(count, want_stats=False, weight=par.pass_weight(count, counts))
        acc += sr.values_f32_3d(hnd, shape)
        sr.free_f32_3d(hnd)
    # each rank drops its first pass (integrator.fut:183-186), so the two means cover 16 of the 17 passes the single image averages
    assert abs(float(acc.mean()) - float(full.mean())) / max(float(full.mean()), 1e-6) < 0.03      # two estimates of the same image: means agree (Monte Carlo noise averaged over the pixels)
    # exactness of the fold: weight w inside the kernel == the same image scaled afterwards
    hnd, _, shape, _ = s.sample_n_frames_device(5, want_stats=False, weight=0.25)
    scaled = s.values_f32_3d(hnd, shape); s.free_f32_3d(hnd)
    assert bits_equal(scaled, (np.float32(0.25) * s.sample_n_frames(5)).astype(np.float32))
    hnd, _, shape, _ = s.sample_n_frames_device(5, want_stats=False, weight=1.0)
    one = s.values_f32_3d(hnd, shape); s.free_f32_3d(hnd)
    assert bits_equal(one, s.sample_n_frames(5))
    # the same split evaluated by the oracle's arithmetic: mean of per-rank means == what the ranks produced, to 1e-4 relative
    ref = np.zeros_like(acc)
    for first, count in ranges:
        sr = s.advance_rng(first) if first else s
        ref += np.float64(par.pass_weight(count, counts)) * sr.sample_n_frames(count).astype(np.float64)
    assert np.allclose(acc, ref, rtol=1e-4, atol=1e-7)


def test_init_light_capacity_regrow(orc, pkg, gpu, scenes):
    """More emissive triangles than the first capacity of the light arrays (4096): init regrows them after its single
    read-back; light indices keep the input order (scene.fut:58-66)."""
    t, tm, m = scenes['cornell']
    st, sm = pkg.scenes.synthetic_cornell(t, tm, 40)                      # 35 200 triangles, 3 200 of them on the light quad
    mm = m.copy()
    emits = ((m[:, 16:28:2] >= 0) & (m[:, 17:28:2] > 0)).any(axis=1)        # nonzero_spectrum, scene.fut:59-60
    mm[:, 16:] = m[np.flatnonzero(emits)[0], 16:]                          # every material emits: all triangles are lights
    so, sg = orc.State.init(st, sm, mm, 8, 8), pkg.State.init(gpu, st, sm, mm, 8, 8)
    assert sg.info()['n_lights'] == len(st) > 4096
    assert bits_equal(so.light_indices(), sg.light_indices())
    assert bits_equal(so.sample_n_frames(2), sg.sample_n_frames(2))


@pytest.mark.parametrize('env', [
    {'LYS_OCT_ONE_COPY': '1'},         # one copy of the traversal records, select-based box test (LAY_SEL: what scenes above 64K nodes run)
    {'LYS_REFILL_MIN': '1'},           # lane refill of the closest-hit walk on the small scene too (octant copies; default from 1024 triangles)
    {'LYS_REFILL_MIN': '1', 'LYS_SHADE_ORDER': '0'},
    {'LYS_REFILL_MIN': '100000000'},   # no lane refill with octant copies (spectrumsphere: the batch loop)
    {'LYS_SHADE_ORDER': '0'},          # k_shade walks the queue in slot order instead of hits first
    {'LYS_FUSE_GENERATE': '0'},        # k_generate and k_trace(-1) as two launches (the per-class timing sequence)
    {'LYS_TAIL_MAX': '0'},             # no fused tail kernel: one launch per stage and bounce
    {'LYS_TAIL_MAX': '100000000'},     # fused tail from bounce 1 on (every queue is 'short')
    {'LYS_TAIL_MAX': '100000000', 'LYS_OCT_ONE_COPY': '1'},
    {'LYS_ADAPTIVE_GRIDS': '0'},       # every pass runs like the first pass of a frame (full grids, no tail)
], ids=lambda e: ','.join(f'{k}={v}' for k, v in e.items()))
def test_kernel_variants_bit_exact(env):
    """Every kernel variant the library selects by scene size or from the previous pass, forced by its environment knob (read once per process), gives the oracle's bits: the sweep of
    tools/gpu_parity_quick.py (BVH, first hits, per-vertex radiance, 4 accumulated passes, step/render) in a subprocess."""
    import json
    import subprocess
    import sys
    e = dict(os.environ)
    e.update(env)
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, 'tools', 'gpu_parity_quick.py'), 'cornell', 'spectrumsphere'],
                                  env=e, text=True, timeout=600)
    seen = 0
    for line in out.splitlines():
        name, _, js = line.partition(' ')
        if name in ('cornell', 'spectrumsphere'):
            r = json.loads(js)
            seen += 1
            for key in ('first_hit_leaf', 'first_hit_t', 'pass_radiance_bits', 'pass_distance_bits', 'pass_channel', 'img4_bits',
                        'step3_img_bits', 'render_bits'):
                assert r[key], (env, name, key)
    assert seen == 2
