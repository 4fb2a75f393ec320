"""CPU property tests of the two rewrites of hit_aabb (reference src/shapes.fut:114-135) that the traversal kernel relies on
(csrc/wavefront.cu: slab_test, slab_test_oct), in IEEE f32 with numpy (np.fmin / np.fmax drop NaN operands like fminf / fmaxf):

 1. the reference leaves after the first axis with tmax <= tmin; one test after the third axis gives the same boolean;
 2. picking near / far per axis at build time by the sign of 1/dir (the octant node arrays) gives the same t0 / t1 as the
    reference's swap;
 3. the escape links of the single-box layout (csrc/lbvh.cu: k_pack_records): the reference's parent-pointer walk
    (bvh.fut:126-142), a left-first walk with a stack and a walk that follows stored escape links visit the same nodes in
    the same order whatever the box tests and triangle tests answer.

Inputs include the cases that make the difference between a careful and a careless rewrite: zero and negative-zero direction
components (1/dir = +-inf, 0 * inf = NaN), origins exactly on slab planes, flat boxes, NaN directions, infinite tmax."""
import numpy as np

F = np.float32
FAR = F(1.0) + F(0.001)


def reference(o, inv, lo, hi, tmax):
    """hit_aabb as written: per axis swap by inv < 0, far slab * 1.001, fmax / fmin, early return."""
    n = len(tmax)
    tmin = np.zeros(n, F)
    tmx = tmax.copy()
    alive = np.ones(n, bool)
    for a in range(3):
        t0 = (lo[:, a] - o[:, a]) * inv[:, a]
        t1 = (hi[:, a] - o[:, a]) * inv[:, a]
        sw = inv[:, a] < 0
        t0, t1 = np.where(sw, t1, t0), np.where(sw, t0, t1)
        t1 = t1 * FAR
        ntmin, ntmax = np.fmax(t0, tmin), np.fmin(t1, tmx)
        tmin = np.where(alive, ntmin, tmin)              # a ray that has left keeps its state
        tmx = np.where(alive, ntmax, tmx)
        alive &= ~(tmx <= tmin)
    return alive


def branch_free(o, inv, lo, hi, tmax):
    tmin = np.zeros(len(tmax), F)
    tmx = tmax.copy()
    for a in range(3):
        t0 = (lo[:, a] - o[:, a]) * inv[:, a]
        t1 = (hi[:, a] - o[:, a]) * inv[:, a]
        sw = inv[:, a] < 0
        t0, t1 = np.where(sw, t1, t0), np.where(sw, t0, t1)
        t1 = t1 * FAR
        tmin, tmx = np.fmax(t0, tmin), np.fmin(t1, tmx)
    return ~(tmx <= tmin)


def octant(o, inv, lo, hi, tmax):
    """near / far chosen per axis by (inv < 0) before the arithmetic, as k_pack_records stores them per octant"""
    tmin = np.zeros(len(tmax), F)
    tmx = tmax.copy()
    for a in range(3):
        sw = inv[:, a] < 0
        near, far = np.where(sw, hi[:, a], lo[:, a]), np.where(sw, lo[:, a], hi[:, a])
        tmin = np.fmax((near - o[:, a]) * inv[:, a], tmin)
        tmx = np.fmin(((far - o[:, a]) * inv[:, a]) * FAR, tmx)
    return ~(tmx <= tmin)


def cases(seed, n):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-2, 2, (n, 3)).astype(F)
    h = np.abs(rng.normal(0, 0.7, (n, 3))).astype(F)
    h[rng.random((n, 3)) < 0.15] = 0                                  # flat boxes (axis-aligned walls)
    lo, hi = (c - h).astype(F), (c + h).astype(F)
    o = rng.uniform(-3, 3, (n, 3)).astype(F)
    on_plane = rng.random((n, 3)) < 0.1                               # origin exactly on a slab plane
    o = np.where(on_plane, np.where(rng.random((n, 3)) < 0.5, lo, hi), o).astype(F)
    d = rng.normal(0, 1, (n, 3)).astype(F)
    aimed = rng.random(n) < 0.6                                        # most rays point at their box, so that both outcomes are common
    d[aimed] = (c[aimed] + rng.normal(0, 0.3, (int(aimed.sum()), 3)).astype(F) * (h[aimed] + F(0.05)) - o[aimed]).astype(F)
    special = rng.random((n, 3))
    d[special < 0.08] = F(0.0)
    d[(special >= 0.08) & (special < 0.16)] = F(-0.0)
    d[(special >= 0.16) & (special < 0.17)] = np.nan
    d[(special >= 0.17) & (special < 0.18)] = F(1e-38)
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        inv = (F(1.0) / d).astype(F)
    tmax = np.where(rng.random(n) < 0.3, np.finfo(F).max, np.abs(rng.normal(0, 3, n))).astype(F)
    tmax[rng.random(n) < 0.01] = np.nan
    return o, inv, lo, hi, tmax


def test_one_comparison_after_the_third_axis_equals_the_early_returns():
    with np.errstate(invalid='ignore', over='ignore'):
        for seed in range(4):
            args = cases(seed, 500_000)
            ref, bf = reference(*args), branch_free(*args)
            assert np.array_equal(ref, bf), int((ref != bf).sum())
            assert 0.1 < ref.mean() < 0.9                                 # both outcomes are exercised


def test_octant_nodes_equal_the_swap():
    with np.errstate(invalid='ignore', over='ignore'):
        for seed in range(4):
            args = cases(100 + seed, 500_000)
            assert np.array_equal(branch_free(*args), octant(*args))


def _dot(a, b):
    return (a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]) + a[:, 2] * b[:, 2]          # vec3.dot, left to right


def _cross(a, b):
    return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2], a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], axis=1)


def test_early_rejections_of_the_triangle_test_do_not_change_it():
    """hit_triangle (shapes.fut:66-86) computes (t, u, v) = (1/det) * (n.s, m.e2, -(m.e1)) and then tests everything; the kernel
    (lys_device.cuh: tri_plane_test + tri_uv_test) rejects on the sign of n.s before the IEEE division and on t before u, v.
    Same boolean, same t, for random and degenerate inputs."""
    rng = np.random.default_rng(7)
    n = 2_000_000
    with np.errstate(invalid='ignore', over='ignore', divide='ignore'):
        a = rng.uniform(-1, 1, (n, 3)).astype(F)
        e1 = rng.normal(0, 0.5, (n, 3)).astype(F)
        e2 = rng.normal(0, 0.5, (n, 3)).astype(F)
        o = rng.uniform(-2, 2, (n, 3)).astype(F)
        tgt = (a + rng.uniform(-0.2, 0.7, (n, 1)).astype(F) * e1 + rng.uniform(-0.2, 0.7, (n, 1)).astype(F) * e2).astype(F)
        d = (tgt - o).astype(F)
        d = (d / np.sqrt(_dot(d, d))[:, None]).astype(F)
        flip = rng.random(n) < 0.3
        d[flip] = -d[flip]                                               # triangle behind the ray: t < 0
        on = rng.random(n) < 0.03
        o[on] = a[on]                                                    # origin on the vertex: n.s = 0
        d[rng.random(n) < 0.01] = np.nan
        tmax = np.where(rng.random(n) < 0.5, np.finfo(F).max, np.abs(rng.normal(0, 2, n))).astype(F)
        nn = _cross(e1, e2).astype(F)
        det = -(_dot(nn, d))
        s = (o - a).astype(F)
        # reference order of evaluation
        inv = (F(1.0) / det).astype(F)
        m = _cross(s, d).astype(F)
        t = inv * _dot(nn, s)
        u = inv * _dot(m, e2)
        v = inv * (-(_dot(m, e1)))
        approx_zero = (det > F(-0.00001)) & (det < F(0.00001))
        ref = ~approx_zero & (u >= 0) & (v >= 0) & (u + v <= 1) & (t < tmax) & (t > 0)
        # kernel order
        dn = _dot(nn, s)
        ok = ~approx_zero
        ok &= ((dn > 0) & (det > 0)) | ((dn < 0) & (det < 0))
        t2 = inv * dn
        ok &= (t2 < tmax) & (t2 > 0)
        ok &= (u >= 0) & (v >= 0) & (u + v <= 1)
    assert np.array_equal(ref, ok), int((ref != ok).sum())
    assert np.array_equal(t[ref].view(np.uint32), t2[ref].view(np.uint32))
    assert 0.1 < ref.mean() < 0.9


# ---- 3. escape links ------------------------------------------------------------------------------------------------
DONE = -(1 << 31)


def random_tree(rng, n_leaves):
    """random binary tree in the child encoding of the build: internal i -> i, leaf j -> ~j; node 0 is the root"""
    left, right, parent = [0] * (n_leaves - 1), [0] * (n_leaves - 1), [-1] * (n_leaves - 1)
    leaf_parent = [0] * n_leaves
    next_node, next_leaf = [1], [0]

    def build(i, k):                                   # node i covers k >= 2 leaves
        kl = int(rng.integers(1, k))
        for side, kk in ((0, kl), (1, k - kl)):
            if kk == 1:
                c = ~next_leaf[0]; leaf_parent[next_leaf[0]] = i; next_leaf[0] += 1
            else:
                c = next_node[0]; next_node[0] += 1; parent[c] = i
            (left if side == 0 else right)[i] = c
            if kk > 1:
                build(c, kk)
    build(0, n_leaves)
    return left, right, parent, leaf_parent


def thread_links(left, right, parent, leaf_parent):
    """k_pack_records: the right child of the nearest ancestor-or-self that is a left child, DONE on the right spine"""
    def link(self, p):
        while p >= 0 and right[p] == self:
            self, p = p, parent[p]
        return DONE if p < 0 else right[p]
    return [link(i, parent[i]) for i in range(len(left))], [link(~j, leaf_parent[j]) for j in range(len(leaf_parent))]


def walk_parent_pointers(left, right, parent, box_pass):
    """bvh.fut:126-142 as written: (current, prev); the box of `current` is tested when it is entered from above, the right
    child follows the left one, leaves are triangle-tested without a box test.  Returns the sequence of tests."""
    seq, current, prev = [], 0, None                    # prev: child pointer we came back from, or None = from above
    while current != -1:
        if prev is not None and prev == left[current]:
            ptr = right[current]
        elif not (prev is not None and prev == right[current]):
            seq.append(current)                         # hit_aabb of this node
            ptr = left[current] if box_pass[current] else None
        else:
            ptr = None
        if ptr is None:
            current, prev = parent[current], current
        elif ptr >= 0:
            current, prev = ptr, None
        else:
            seq.append(ptr)                             # hit_triangle of the leaf
            prev = ptr
    return seq


def walk_stack(left, right, box_pass):
    seq, stack, cur = [], [DONE], 0
    while cur != DONE:
        seq.append(cur)
        if cur >= 0 and box_pass[cur]:
            stack.append(right[cur]); cur = left[cur]
        else:
            cur = stack.pop()
    return seq


def walk_threaded(left, node_link, leaf_link, box_pass):
    seq, cur = [], 0
    while cur != DONE:
        seq.append(cur)
        cur = (left[cur] if box_pass[cur] else node_link[cur]) if cur >= 0 else leaf_link[~cur]
    return seq


def test_escape_links_visit_what_the_stack_and_the_parent_pointers_visit():
    rng = np.random.default_rng(77)
    for trial in range(300):
        n = int(rng.integers(2, 200))
        left, right, parent, leaf_parent = random_tree(rng, n)
        node_link, leaf_link = thread_links(left, right, parent, leaf_parent)
        for p in (0.0, 0.3, 0.7, 1.0):
            box_pass = (rng.random(n - 1) < p) if 0.0 < p < 1.0 else np.full(n - 1, p == 1.0)
            a = walk_stack(left, right, box_pass)
            assert a == walk_threaded(left, node_link, leaf_link, box_pass)
            assert a == walk_parent_pointers(left, right, parent, box_pass)
        if trial == 0:
            assert len(walk_threaded(left, node_link, leaf_link, np.ones(n - 1, bool))) == 2 * n - 1     # every node and leaf once
