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
    x &= 0xFFFFFFFF
    x = (((x >> 16) ^ x) * 0x45d9f3b) & 0xFFFFFFFF
    x = (((x >> 16) ^ x) * 0x45d9f3b) & 0xFFFFFFFF
    return (x >> 16) ^ x


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
