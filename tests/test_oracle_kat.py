"""Known-answer tests pinning the oracle's integer pieces to independently written Python restatements of
the published algorithms (Karras 2012; cpprandom's LCG; the Morton bit-spreading of bvh.fut:52-73).
The reference ships no vectors of its own, so these are self-authored."""
import numpy as np
import pytest


def py_expand_bits(x):                    # slow, obviously-correct: insert two zeros after each of 10 bits
    r = 0
    for b in range(10):
        r |= ((x >> b) & 1) << (3 * b)
    return r


def test_expand_bits(orc):
    L = orc.lib()
    for x in list(range(0, 1024, 7)) + [1, 2, 3, 512, 1023]:
        assert L.orc_expand_bits(x) == py_expand_bits(x)
    assert L.orc_expand_bits(1023) == 0x09249249 and L.orc_expand_bits(1) == 1 and L.orc_expand_bits(2) == 8


def test_morton3d(orc):
    L = orc.lib()
    assert L.orc_morton3d(0.0, 0.0, 0.0) == 0
    assert L.orc_morton3d(1.0, 1.0, 1.0) == 0x3FFFFFFF                # min(v*1024, 1023) clamps
    assert L.orc_morton3d(0.5, 0.0, 0.0) == py_expand_bits(512) * 4
    assert L.orc_morton3d(0.0, 0.25, 0.0) == py_expand_bits(256) * 2
    assert L.orc_morton3d(0.0, 0.0, 0.999) == py_expand_bits(1022)
    assert L.orc_morton3d(float('nan'), 0.0, 0.0) == py_expand_bits(1023) * 4   # fmin semantics: NaN axis -> 1023
    assert L.orc_morton3d(-0.25, 0.0, 0.0) == 0                                   # negative -> 0
    rng = np.random.default_rng(0)
    for x, y, z in rng.random((200, 3)).astype(np.float32):
        q = [int(min(np.float32(v) * np.float32(1024), np.float32(1023))) for v in (x, y, z)]
        assert L.orc_morton3d(x, y, z) == py_expand_bits(q[0]) * 4 + py_expand_bits(q[1]) * 2 + py_expand_bits(q[2])


def py_lcg(s):
    return ((48271 * s) & 0xFFFFFFFF) % 2147483647


def py_hash(x):
    """cpprandom `hash (x: i32): i32` with Futhark's `>>` on i32 = ARITHMETIC shift (include/lys_pins.h,
    LYS_PIN_HASH_SHIFT_ARITHMETIC = 1), written on Python's unbounded signed integers."""
    def i32(v):
        v &= 0xFFFFFFFF
        return v - (1 << 32) if v & 0x80000000 else v
    x = i32(x)
    x = i32(((x >> 16) ^ x) * 0x45d9f3b)
    x = i32(((x >> 16) ^ x) * 0x45d9f3b)
    return ((x >> 16) ^ x) & 0xFFFFFFFF


def test_rng(orc):
    L = orc.lib()
    assert L.orc_rng_from_seed(0) == 263559660        # ((1>>16)^1) ^ (0 ^ 0x1555) = 0x1554; 48271*5460 mod (2^31-1)
    s = L.orc_rng_from_seed(0)
    for _ in range(1000):
        n = py_lcg(s)
        assert L.orc_rng_next(s) == n
        s = n
    for seed in (1, 42, -7, 2 ** 31 - 1):
        sp = ((1 >> 16) ^ 1) ^ ((seed & 0xFFFFFFFF) ^ 0x1555)
        assert L.orc_rng_from_seed(seed) == py_lcg(sp)
    for i in (0, 1, 2, 12345, 1920 * 1080 - 1, 2 ** 31 - 1):
        assert L.orc_hash(i) == py_hash(i)
    # uniform_real_distribution: lo + (f32(x)/2^31) * (hi - lo), x the NEW state
    import ctypes
    out = ctypes.c_uint32()
    v = L.orc_rng_uniform(5460, 0.0, 0.9999, ctypes.byref(out))
    assert out.value == 263559660
    assert np.float32(v) == np.float32(np.float32(263559660) / np.float32(2 ** 31)) * np.float32(0.9999)


def py_karras(keys):
    """Top-down restatement of Karras' radix tree over (key, index) with the duplicate tie-break, independent of
    the per-node binary searches in radix_tree.fut: node i covers a key range and splits at the highest differing bit."""
    n = len(keys)
    aug = [(int(k) << 32) | i for i, k in enumerate(keys)]     # 64-bit augmented keys are distinct and sorted

    def delta(i, j):
        if j < 0 or j >= n:
            return -1
        return 64 - (aug[i] ^ aug[j]).bit_length()

    left, right, parent = [0] * (n - 1), [0] * (n - 1), [-1] * (n - 1)
    for i in range(n - 1):
        d = 1 if delta(i, i + 1) > delta(i, i - 1) else -1
        dmin = delta(i, i - d)
        l = 0
        while delta(i, i + (l + 1) * d) > dmin:                  # linear scan instead of doubling + bisection
            l += 1
        j = i + l * d
        first, last = min(i, j), max(i, j)
        dn = delta(first, last)
        split = first
        while delta(first, split + 1) > dn:                      # linear scan for the split position
            split += 1
        if split == first:
            left[i] = ~split
        else:
            left[i] = split
            parent[split] = i
        if split + 1 == last:
            right[i] = ~(split + 1)
        else:
            right[i] = split + 1
            parent[split + 1] = i
    return np.array(left, np.int32), np.array(right, np.int32), np.array(parent, np.int32)


@pytest.mark.parametrize('keys', [
    [1, 2], [5, 5], [0, 0, 0, 0], [1, 2, 4, 8, 16, 32, 64, 128], [0, 1, 1, 1, 2, 3, 3, 7],
    [0x3FFFFFFF] * 5, list(range(17)), [0, 0, 1, 1, 2, 2, 0x20000000, 0x20000000, 0x3FFFFFFF]])
def test_karras_small(orc, keys):
    l, r, p = orc.radix_tree(np.array(keys, np.uint32))
    el, er, ep = py_karras(keys)
    assert np.array_equal(l, el) and np.array_equal(r, er) and np.array_equal(p, ep)


def test_karras_random(orc):
    rng = np.random.default_rng(3)
    for n in (3, 10, 100, 1000):
        keys = np.sort(rng.integers(0, 1 << 12, n).astype(np.uint32))       # many duplicates
        l, r, p = orc.radix_tree(keys)
        el, er, ep = py_karras(list(keys))
        assert np.array_equal(l, el) and np.array_equal(r, er) and np.array_equal(p, ep)
        # structure: root has no parent, every other node exactly one, every leaf referenced once
        assert p[0] == -1 and (p[1:] >= 0).all()
        leaves = np.concatenate([~l[l < 0], ~r[r < 0]])
        assert np.array_equal(np.sort(leaves), np.arange(n))


def test_spectrum_lookup(orc):
    L = orc.lib()
    s = np.array([380, 0.3, 450, 1.0, 540, 0.0, -1, 0, -1, 0, -1, 0], np.float32)          # SpectrumSphere.mtl bright-blue
    assert L.orc_spectrum_lookup(300.0, s) == np.float32(0.3)                                # below all knots -> nearest
    assert L.orc_spectrum_lookup(600.0, s) == np.float32(0.0)                                # above all knots
    assert L.orc_spectrum_lookup(450.0, s) == np.float32(1.0)
    assert abs(L.orc_spectrum_lookup(415.0, s) - 0.65) < 1e-6
    u = np.array([0, 5.0, -1, 0, -1, 0, -1, 0, -1, 0, -1, 0], np.float32)                   # uniform_spectrum 5
    assert L.orc_spectrum_lookup(1550.0, u) == np.float32(5.0)
    z = np.array([-1, 0] * 6, np.float32)
    assert L.orc_spectrum_lookup(500.0, z) == 0.0


def test_hit_triangle_against_an_independent_restatement(orc, scenes):
    """The oracle's closest hit over all triangles (orc_brute_force_hits: hit_triangle on every leaf, strictly smaller t wins)
    against hit_triangle written here from reference src/shapes.fut:66-86 and src/common.fut:35 in numpy f32 (operand order of
    athas/vector: dot = x*x' + y*y' + z*z' left to right; scale (1/a) v = (1/a) * each component):
    hit / miss, the winning source triangle and the bits of t."""
    F = np.float32
    t9, tm, m = scenes['spectrumsphere']
    tri = np.ascontiguousarray(t9, F).reshape(-1, 3, 3)
    s = orc.State.init(t9, tm, m, 48, 64)
    order = s.bvh()['src_index']                              # brute force walks the leaves in sorted order
    rays = s.probe_primary(want_rays=True)['rays'].reshape(-1, 6)
    rng = np.random.default_rng(3)
    extra = np.concatenate([rng.uniform(-1, 1, (2000, 3)) + (0, 0.8, 0), rng.normal(0, 1, (2000, 3))], axis=1).astype(F)
    extra[:, 3:] /= np.sqrt((extra[:, 3:].astype(np.float64) ** 2).sum(axis=1, keepdims=True)).astype(F)
    rays = np.ascontiguousarray(np.concatenate([rays, extra]), F)
    src_o, t_o = s.brute_force_hits(rays)

    def dot(a, b):
        return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]

    def cross(a, b):
        return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1], a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                         a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)

    A, B, Cc = tri[order, 0], tri[order, 1], tri[order, 2]
    e1, e2 = (B - A).astype(F), (Cc - A).astype(F)
    n = cross(e1, e2).astype(F)
    best_t = np.full(len(rays), np.inf, F)
    best = np.full(len(rays), -1, np.int64)
    o, d = rays[:, None, :3], rays[:, None, 3:]
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        a = -(dot(n[None], d))                                     # [rays][tris]
        sv = (o - A[None]).astype(F)
        mv = cross(sv, np.broadcast_to(d, sv.shape)).astype(F)
        inv = (F(1) / a).astype(F)
        t = inv * dot(n[None], sv)
        u = inv * dot(mv, e2[None])
        v = inv * (-(dot(mv, e1[None])))
        ok = ~((a > F(-0.00001)) & (a < F(0.00001))) & (u >= 0) & (v >= 0) & (u + v <= 1) & (t < np.finfo(F).max) & (t > 0)
    for k in range(len(order)):                                   # strictly smaller t wins, earlier leaf kept on ties
        better = ok[:, k] & ((best < 0) | (t[:, k] < best_t))
        best_t = np.where(better, t[:, k], best_t)
        best = np.where(better, k, best)
    src = np.where(best >= 0, order[np.maximum(best, 0)], -1)
    assert np.array_equal(src, src_o)
    hit = best >= 0
    assert 0.2 < hit.mean() < 0.98
    assert np.array_equal(best_t[hit].view(np.uint32), t_o[hit].view(np.uint32))


def test_stackless_walk_against_an_independent_restatement(orc, scenes):
    """closest_hit (reference src/bvh.fut:123-145: stackless, parent pointers, left first, strict t < tmax) and hit_aabb
    (src/shapes.fut:114-135) written here in scalar numpy f32 straight from the .fut text, run on the oracle's own tree and boxes
    (truncated refit included), against the oracle's walk: same leaf, same bits of t.  SpectrumSphere is the scene whose
    truncated boxes do not enclose all their triangles, so the order of the decisions matters."""
    F = np.float32
    t9, tm, m = scenes['spectrumsphere']
    tri = np.ascontiguousarray(t9, F).reshape(-1, 3, 3)
    s = orc.State.init(t9, tm, m, 16, 20)
    bv = s.bvh()
    left, right, parent, box, order = bv['left'], bv['right'], bv['parent'], bv['node_aabb'], bv['src_index']
    rays = s.probe_primary(want_rays=True)['rays'].reshape(-1, 6)
    rng = np.random.default_rng(5)
    extra = np.concatenate([rng.uniform(-0.8, 0.8, (120, 3)) + (0, 0.8, 0), rng.normal(0, 1, (120, 3))], axis=1).astype(F)
    rays = np.ascontiguousarray(np.concatenate([rays, extra]), F)
    leaf_o, t_o = s.closest_hits(rays)
    HIGHEST = np.finfo(F).max
    one, eps_far = F(1), F(0.001)

    def hit_aabb(tmax, o, d, c, h):
        mn, mx = c - h, c + h
        tmin = F(0)
        for a in range(3):
            with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
                inv = one / d[a]
                t0, t1 = (mn[a] - o[a]) * inv, (mx[a] - o[a]) * inv
                if inv < 0:
                    t0, t1 = t1, t0
                t1 = t1 * (one + eps_far)
            tmin, tmax = np.fmax(t0, tmin), np.fmin(t1, tmax)
            if tmax <= tmin:
                return False
        return True

    def dot(a, b):
        return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]

    def cross(a, b):
        return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], F)

    def hit_triangle(tmax, o, d, k):
        A, B, Cc = tri[order[k]]
        e1, e2 = B - A, Cc - A
        n = cross(e1, e2)
        a = -(dot(n, d))
        if a > F(-0.00001) and a < F(0.00001):
            return None
        sv = o - A
        mv = cross(sv, d)
        with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
            inv = one / a
            t, u, v = inv * dot(n, sv), inv * dot(mv, e2), inv * (-(dot(mv, e1)))
        if u >= 0 and v >= 0 and u + v <= 1 and t < tmax and t > 0:
            return t
        return None

    for r, lo, to in zip(rays, leaf_o, t_o):
        o, d = r[:3], r[3:]
        closest, tmax, current, prev = -1, HIGHEST, 0, -1                  # prev: a child pointer (internal i, leaf ~i); the root's "prev" is internal -1
        prev_is_root_marker = True
        steps = 0
        while current != -1:
            steps += 1
            assert steps < 10000
            l, rr = int(left[current]), int(right[current])
            came = None if prev_is_root_marker else prev
            if came is not None and came == l:
                child = rr
            elif (came is None or came != rr) and hit_aabb(tmax, o, d, box[current, :3], box[current, 3:]):
                child = l
            else:
                child = None
            prev_is_root_marker = False
            if child is None:
                prev, current = current, int(parent[current])
            elif child >= 0:
                prev, current = current, child
            else:
                t = hit_triangle(tmax, o, d, ~child)
                if t is not None:
                    closest, tmax = ~child, t
                prev = child
        assert closest == lo
        if closest >= 0:
            t = hit_triangle(HIGHEST, o, d, closest)                        # the final re-intersection with the outer tmax
            assert t is not None and np.float32(t).view(np.uint32) == to.view(np.uint32)
    assert (leaf_o >= 0).mean() > 0.2


@pytest.mark.parametrize('name', ['cornell', 'spectrumsphere', 'spectrumspherehigh'])
def test_build_against_an_independent_restatement(orc, scenes, name):
    """The whole BVH build of reference src/bvh.fut:86-121 restated here in numpy f32 from the .fut text -- triangle boxes
    (shapes.fut:96-110), the left fold of containing_aabb from the {0, -inf} neutral, Morton codes of the normalised centres, a
    stable sort by key, and `i32(log2 n) + 2` literal Jacobi sweeps from zero boxes over the oracle's tree topology -- against
    the oracle's bounds, sorted keys, permutation and (truncated) node boxes, bit for bit."""
    F = np.float32
    t9, tm, m = scenes[name]
    tri = np.ascontiguousarray(t9, F).reshape(-1, 3, 3)
    n = len(tri)
    bv = orc.State.init(t9, tm, m, 8, 8).bvh()

    def contain(c1, h1, c2, h2):
        mn = np.fmin(c1 - h1, c2 - h2)
        mx = np.fmax(c1 + h1, c2 + h2)
        c = F(0.5) * (mn + mx)
        return c.astype(F), (mx - c).astype(F)

    z = np.zeros((n, 3), F)
    cb, hb = contain(tri[:, 1], z, tri[:, 2], z)                     # containing_aabb b c
    c, h = contain(tri[:, 0], z, cb, hb)                             # containing_aabb a (...)
    bc, bh = np.zeros(3, F), np.full(3, -np.inf, F)
    for k in range(n):                                                # reduce on the c backend = sequential left fold
        bc, bh = contain(bc, bh, c[k], h[k])
    assert np.array_equal(np.concatenate([bc, bh]).view(np.uint32), bv['bounds'].view(np.uint32))
    with np.errstate(divide='ignore', invalid='ignore'):
        v = ((c - (bc - bh)) / (F(2) * bh)).astype(F)                 # normalise_position
    s = np.fmin(v * F(1024), F(1023))
    q = np.where(s > 0, s, 0).astype(np.uint32)                       # u32.f32 truncation

    def expand(x):
        x = (x * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
        x = (x * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
        x = (x * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
        x = (x * np.uint32(0x00000005)) & np.uint32(0x49249249)
        return x
    with np.errstate(over='ignore'):
        mort = expand(q[:, 0]) * np.uint32(4) + expand(q[:, 1]) * np.uint32(2) + expand(q[:, 2])
    perm = np.argsort(mort, kind='stable')
    assert np.array_equal(perm, bv['src_index'])
    assert np.array_equal(mort[perm], bv['morton'])
    # Jacobi refit on the oracle's topology (the topology itself is pinned by the Karras tests above)
    left, right = bv['left'], bv['right']
    lc, lh = c[perm], h[perm]
    assert np.array_equal(np.concatenate([lc, lh], axis=1).view(np.uint32), bv['leaf_aabb'].view(np.uint32))
    nc, nh = np.zeros((n - 1, 3), F), np.zeros((n - 1, 3), F)
    li, ri = np.where(left >= 0, left, 0), np.where(right >= 0, right, 0)
    ll, rl = np.where(left < 0, ~left, 0), np.where(right < 0, ~right, 0)
    depth = int(np.float32(np.log2(np.float32(n)))) + 2
    for _ in range(depth):
        c1 = np.where((left >= 0)[:, None], nc[li], lc[ll]); h1 = np.where((left >= 0)[:, None], nh[li], lh[ll])
        c2 = np.where((right >= 0)[:, None], nc[ri], lc[rl]); h2 = np.where((right >= 0)[:, None], nh[ri], lh[rl])
        nc, nh = contain(c1, h1, c2, h2)
    assert np.array_equal(np.concatenate([nc, nh], axis=1).view(np.uint32), bv['node_aabb'].view(np.uint32))


def test_material_against_an_independent_restatement(orc):
    """bsdf_f, bsdf_pdf and sample_dir (reference src/material.fut:60-410, rand.fut) written here in scalar numpy f32 straight
    from the .fut text, sharing with the oracle only what the reference does not define: the transcendental contract
    (include/lys_detmath.h through orc_eval_math), the LCG and spectrum_lookup (pinned by the tests above).  Bit-exact on
    dielectric, metallic, transparent and mixed materials, both hemispheres, all three pdf kinds."""
    import ctypes
    F = np.float32
    L = orc.lib()
    PI = F(np.pi)
    INV_PI = F(1.0) / PI
    fmax, fmin = np.fmax, np.fmin

    def m1(fn, x):
        return orc.eval_math(fn, np.array([x], F))[0]

    def dot(a, b):
        return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]

    def norm(v):
        return np.sqrt(dot(v, v))

    def normalise(v):
        return (F(1) / norm(v)) * v

    def cross(a, b):
        return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], F)

    def lerp(a, b, t):
        return a + (b - a) * t

    class Rng:
        def __init__(self, s):
            self.s = s

        def unit(self, lo=F(0), hi=F(0.9999)):
            out = ctypes.c_uint32()
            v = L.orc_rng_uniform(self.s, lo, hi, ctypes.byref(out))
            self.s = out.value
            return F(v)

    def same_hemisphere(w, u):
        return w[2] * u[2] > 0

    def sin2_theta(w):
        return fmax(F(0), F(1) - w[2] * w[2])

    def alpha_of(r):
        return F(1.62142) * fmax(F(0.004), r)

    def D(alpha, wh):
        with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
            t2 = sin2_theta(wh) / (wh[2] * wh[2])
            if np.isinf(t2):
                return F(0)
            return m1('exp', -t2 / (alpha * alpha)) / (PI * alpha * alpha * (wh[2] * wh[2]) * (wh[2] * wh[2]))

    def G(alpha, wo, wi):
        def lam(w):
            with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
                at = np.abs(np.sqrt(sin2_theta(w)) / w[2])
                if np.isinf(at):
                    return F(0)
                a = F(1) / (alpha * at)
            if a >= F(1.6):
                return F(0)
            return (F(1) - F(1.259) * a + F(0.396) * a * a) / (F(3.535) * a + F(2.181) * a * a)
        return F(1) / (F(1) + lam(wo) + lam(wi))

    def refl_bsdf(wo, wi, m):
        with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
            wh = normalise(wi + wo)
            a = alpha_of(m['roughness'])
            return (D(a, wh) * G(a, wo, wi)) / (F(4) * wo[2] * wi[2])

    def refl_pdf(wo, wi, m):
        if not same_hemisphere(wo, wi):
            return F(0)
        with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
            wh = normalise(wo + wi)
            return (D(alpha_of(m['roughness']), wh) * np.abs(wh[2])) / (F(4) * dot(wo, wh))

    def fresnel(wo, m):
        x = (F(1) - m['ref_ix']) / (F(1) + m['ref_ix'])
        r0 = x * x
        return r0 + (F(1) - r0) * m1('pow5', F(1) - wo[2])

    def diffuse_pdf(wo, wi):
        return wi[2] * INV_PI if same_hemisphere(wo, wi) else F(0)

    def refr_bsdf(m):
        return lerp(F(0), m['color'] * INV_PI, m['opacity'])

    def refr_pdf(wo, wi, m):
        return lerp(F(0), diffuse_pdf(wo, wi), m['opacity'])

    def uber_bsdf(wo, wi, m):
        refl = F(0) if wo[2] <= 0 else fresnel(wo, m)
        diel = lerp(refr_bsdf(m), refl_bsdf(wo, wi, m), refl)
        return lerp(diel, m['color'] * refl_bsdf(wo, wi, m), m['metalness'])

    def uber_pdf(wo, wi, m):
        if wo[2] <= 0:
            diel = refr_pdf(wo, wi, m)
        else:
            diel = lerp(refr_pdf(wo, wi, m), refl_pdf(wo, wi, m), fresnel(wo, m))
        return lerp(refl_pdf(wo, wi, m), diel, m['metalness'])            # operands as written in the reference (:360-361)

    def reflect(w, n):
        return F(-1) * w + (F(2) * dot(w, n)) * n

    def sample_reflection(wo, m, rng):
        u0, u1 = rng.unit(), rng.unit()
        ls = m1('log', F(1) - u0)
        if np.isinf(ls):
            wh, pdf_wh = np.zeros(3, F), F(0)
        else:
            a = alpha_of(m['roughness'])
            tan2 = -a * a * ls
            phi = u1 * F(2) * PI
            ct = F(1) / np.sqrt(F(1) + tan2)
            st = np.sqrt(fmax(F(0), F(1) - ct * ct))
            wh = np.array([st * m1('cos', phi), st * m1('sin', phi), ct], F)
            if not same_hemisphere(wo, wh):
                wh = -wh
            pdf_wh = D(a, wh) * np.abs(ct)
        wi = reflect(wo, wh)
        if not same_hemisphere(wo, wi):
            return np.zeros(3, F), F(0), 1, F(0)                          # null_sample: #impossible
        with np.errstate(divide='ignore', invalid='ignore'):
            kind, pdf = (2, pdf_wh / (F(4) * dot(wo, wh))) if pdf_wh > 0 else (1, F(0))
        return wi, refl_bsdf(wo, wi, m), kind, pdf

    def sample_refraction(wo, m, rng):
        p = rng.unit()
        if p < m['opacity']:
            theta = rng.unit(F(0), F(2) * PI)
            u = rng.unit()
            r = np.sqrt(u)
            d = r * np.array([m1('cos', theta), m1('sin', theta), F(0)], F)
            z = np.sqrt(fmax(F(0), F(1) - (d[0] * d[0] + d[1] * d[1])))
            return np.array([d[0], d[1], z], F), m['color'] * INV_PI, 2, z * INV_PI
        entering = wo[2] > 0
        n = np.array([0, 0, 1], F) if entering else np.array([-0.0, -0.0, -1.0], F)
        eta = F(1.0) / m['ref_ix'] if entering else m['ref_ix'] / F(1.0)
        ci = dot(n, wo)
        s2i = fmax(F(0), F(1) - ci * ci)
        s2t = eta * eta * s2i
        if s2t >= 1:
            wi = reflect(wo, n)
        else:
            wi = (-eta) * wo + (eta * ci - np.sqrt(F(1) - s2t)) * n
        with np.errstate(divide='ignore'):
            return wi, F(1) / np.abs(wi[2]), 0, F(0)

    def uber_sample(wo, m, rng):
        p = rng.unit()
        if p < m['metalness']:
            wi, b, k, pdf = sample_reflection(wo, m, rng)
            return wi, m['color'] * b, k, pdf
        if wo[2] <= 0:
            return sample_refraction(wo, m, rng)
        r = fresnel(wo, m)
        q = rng.unit()
        return sample_reflection(wo, m, rng) if q < r else sample_refraction(wo, m, rng)

    def onb(nrm):
        with np.errstate(divide='ignore', invalid='ignore'):
            if np.abs(nrm[0]) > np.abs(nrm[2]):
                b = normalise(np.array([-nrm[1], nrm[0], 0], F))
            else:
                b = normalise(np.array([0, -nrm[2], nrm[1]], F))
        return cross(b, nrm), b, nrm

    def to_local(o, w):
        return np.array([dot(w, o[0]), dot(w, o[1]), dot(w, o[2])], F)

    rs = np.random.default_rng(11)
    checked, kinds = 0, set()
    for trial in range(400):
        mat = np.zeros(28, F)
        mat[0:12:2] = -1
        mat[16:28:2] = -1
        mat[0:4] = (400, rs.uniform(0, 1), 700, rs.uniform(0, 1))          # two colour knots
        mat[12] = rs.choice([0.0, 0.001, 0.05, 0.3, 1.0])                  # roughness
        mat[13] = rs.choice([0.0, 0.0, 1.0, 0.4])                          # metalness
        mat[14] = rs.choice([1.0, 1.33, 1.5, 2.4])                         # ref_ix
        mat[15] = rs.choice([1.0, 1.0, 0.0, 0.5])                          # opacity
        wl = F(rs.uniform(380, 750))
        nrm = rs.normal(0, 1, 3).astype(F)
        nrm = (nrm / np.sqrt((nrm.astype(np.float64) ** 2).sum())).astype(F)
        wo = rs.normal(0, 1, 3).astype(F)
        wo = (wo / np.sqrt((wo.astype(np.float64) ** 2).sum())).astype(F)
        wi = rs.normal(0, 1, 3).astype(F)
        wi = (wi / np.sqrt((wi.astype(np.float64) ** 2).sum())).astype(F)
        seed = int(rs.integers(1, 2 ** 31 - 2))
        out = np.zeros(9, F)
        L.orc_material_probe(mat, wl, wo, wi, nrm, seed, out)
        m = dict(color=F(L.orc_spectrum_lookup(wl, np.ascontiguousarray(mat[:12]))), roughness=mat[12], metalness=mat[13],
                 ref_ix=mat[14] - (wl - F(589)) / F(10000), opacity=mat[15])
        o = onb(nrm)
        wo_l, wi_l = to_local(o, wo), to_local(o, wi)
        with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
            f, pdf = uber_bsdf(wo_l, wi_l, m), uber_pdf(wo_l, wi_l, m)
            rng = Rng(seed)
            swi, sb, sk, spdf = uber_sample(wo_l, m, rng)
            swi_w = (swi[0] * o[0] + swi[1] * o[1]) + swi[2] * o[2]
        got = np.array([f, pdf, swi_w[0], swi_w[1], swi_w[2], sb, F(sk), spdf], F)
        assert np.array_equal(got.view(np.uint32), out[:8].view(np.uint32)), (trial, mat[12:16], got, out[:8])
        assert rng.s == int(out[8:9].view(np.uint32)[0])
        kinds.add(sk)
        checked += 1
    assert checked == 400 and kinds == {0, 1, 2}
