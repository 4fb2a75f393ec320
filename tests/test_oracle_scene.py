"""Oracle behaviour on the bundled scenes: golden regression pins, traversal vs brute force, and the reference's
truncated-refit quirk."""
import numpy as np
import pytest
from conftest import SCENE_NAMES, bits_equal


@pytest.mark.parametrize('name', SCENE_NAMES)
def test_matches_committed_golden(orc, scenes, golden_vectors, name):
    t, tm, m = scenes[name]
    s = orc.State.init(t, tm, m, 48, 64)
    b = s.bvh()
    for k in ('bounds', 'morton', 'src_index', 'left', 'right', 'parent', 'node_aabb'):
        assert bits_equal(b[k], golden_vectors['%s_%s' % (name, k)]), k
    pr = s.probe_primary()
    assert bits_equal(pr['src_tri'], golden_vectors[name + '_first_hit_src'])
    assert bits_equal(pr['t'], golden_vectors[name + '_first_hit_t'])
    assert bits_equal(s.sample_n_frames(3), golden_vectors[name + '_img3'])
    assert bits_equal(s.light_indices(), golden_vectors[name + '_lights'])


@pytest.mark.parametrize('name', SCENE_NAMES)
def test_bvh_invariants(orc, scenes, name):
    t, tm, m = scenes[name]
    b = orc.State.init(t, tm, m, 8, 8).bvh()
    n = len(t)
    assert np.all(np.diff(b['morton'].astype(np.int64)) >= 0)                     # sorted
    assert np.array_equal(np.sort(b['src_index']), np.arange(n))                 # a permutation
    same = b['morton'][1:] == b['morton'][:-1]
    assert np.all(b['src_index'][1:][same] > b['src_index'][:-1][same])          # stable: ties keep input order
    assert (b['morton'] < (1 << 30)).all()
    # light triangles: every triangle whose material row has an emission knot with w >= 0 and x > 0, in input order
    em = m[:, 16:].reshape(-1, 6, 2)
    emissive = ((em[:, :, 0] >= 0) & (em[:, :, 1] > 0)).any(axis=1)
    assert np.array_equal(orc.State.init(t, tm, m, 8, 8).light_indices(), np.nonzero(emissive[tm])[0])


def random_rays(n, seed=0):
    rng = np.random.default_rng(seed)
    o = np.stack([rng.uniform(-0.9, 0.9, n), rng.uniform(0.1, 1.8, n), rng.uniform(-0.9, 0.9, n)], axis=1)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d], axis=1).astype(np.float32)


@pytest.mark.parametrize('name', SCENE_NAMES)
def test_converged_bvh_equals_brute_force(orc, scenes, name):
    """With converged boxes the left-first walk must find the same nearest t as testing every triangle
    (mk_fake_bvh semantics, bvh.fut:31-39); ties may pick another triangle with the same t."""
    t, tm, m = scenes[name]
    orc.set_refit_mode(1)
    try:
        s = orc.State.init(t, tm, m, 8, 8)
        rays = random_rays(4000)
        leaf, tt = s.closest_hits(rays)
        src_bf, t_bf = s.brute_force_hits(rays)
        assert np.array_equal(leaf >= 0, src_bf >= 0)
        hit = leaf >= 0
        assert np.array_equal(tt[hit], t_bf[hit])
    finally:
        orc.set_refit_mode(0)


def test_truncated_refit_numbers(orc, scenes):
    """bvh.fut:109 runs floor(log2 n)+2 Jacobi sweeps; SURVEY.md H1 measured how many node boxes that leaves
    unconverged on the bundled assets: 0 (Cornell), 0 (MirrorBox), 8 (SpectrumSphere), 16 (SpectrumSphereHigh)."""
    want = {'cornell': 0, 'mirrorbox': 0, 'spectrumsphere': 8, 'spectrumspherehigh': 16}
    for name, k in want.items():
        t, tm, m = scenes[name]
        a = orc.State.init(t, tm, m, 8, 8).bvh()['node_aabb']
        orc.set_refit_mode(1)
        try:
            c = orc.State.init(t, tm, m, 8, 8).bvh()['node_aabb']
        finally:
            orc.set_refit_mode(0)
        assert int((a.view(np.uint32) != c.view(np.uint32)).any(axis=1).sum()) == k, name


def test_shadow_rays_agree_with_closest(orc, scenes):
    t, tm, m = scenes['cornell']
    s = orc.State.init(t, tm, m, 8, 8)
    rays = random_rays(2000, seed=5)
    leaf, tt = s.closest_hits(rays)
    tmax = np.full(len(rays), 0.7, np.float32)
    anyh = s.any_hits(rays, tmax)
    assert np.array_equal(anyh == 1, (leaf >= 0) & (tt < 0.7))


def test_divergence_probes_are_consistent(orc, scenes):
    """The study probes used by tools/simt_model.py restate the same walk: the pattern's box / triangle visits equal the
    step counters, its last triangle hit is the closest hit, and the dumped bounce-0 rays are the camera rays."""
    t, tm, m = scenes['spectrumsphere']
    s = orc.State.init(t, tm, m, 24, 32)
    rays, n = s.probe_path_rays()
    prim = s.probe_primary(want_rays=True)
    assert n.min() >= 1 and n.max() <= 16
    assert np.array_equal(rays[:, :, 0].view(np.uint32), prim['rays'].view(np.uint32))
    hit0 = prim['leaf'] >= 0
    assert np.all((n > 1) <= hit0)                                     # a path continues only from a hit vertex
    live = np.nonzero(n.reshape(-1) > 1)[0]
    r1 = rays.reshape(-1, 16, 6)[live, 1]
    pat, ln = s.closest_hits_pattern(r1, 400)
    box, tri = s.closest_hits_steps(r1)
    assert np.array_equal(((pat == 0) | (pat == 1)).sum(axis=1), box)
    assert np.array_equal(((pat == 2) | (pat == 3)).sum(axis=1), tri)
    assert np.array_equal(ln, box + tri)
    leaf, _ = s.closest_hits(r1)
    assert np.array_equal((pat == 3).any(axis=1), leaf >= 0)
    assert np.all(pat[:, 0] <= 1)                                       # every walk starts with the root's box


def test_simt_model_tool_runs_and_orders_the_loop_shapes():
    """tools/simt_model.py (the lock-step cost model behind the traversal loop shape) on a small frame: a loop that lets a lane
    do a box visit and a triangle visit per iteration never needs more iterations than one visit per iteration, and the
    while-while shape needs the fewest iterations of all (it is the instruction count that makes it the worst)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('simt_model', os.path.join(root, 'tools', 'simt_model.py'))
    sm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sm)
    costs = {'B': 45, 'T': 60, 'loop': 6}
    seen = 0
    for b, r, pat, ln in sm.bounce_rays('cornell', 48, 64):
        if len(r) < 64:
            continue
        base = np.arange(len(r))
        it = {s: sm.schedule_cost(pat, ln, base, s, costs)[1] for s in ('X', 'BT', 'BBT', 'B*T')}
        assert it['BT'] <= it['X'] and it['BBT'] <= it['BT'] and it['B*T'] <= it['BBT'], it
        c_x, _ = sm.schedule_cost(pat, ln, base, 'X', costs)
        assert c_x > 0
        seen += 1
    assert seen >= 2
