"""Writes (tris, tri_mats, mats) back to OBJ/MTL so that the loader reproduces the arrays exactly (test support)."""
import numpy as np


def _fmt(x):
    return repr(float(np.float32(x)))      # shortest decimal that round-trips through binary64 -> exact binary32 again


def write_obj(path_obj, tris, tri_mats, mats):
    path_mtl = path_obj[:-4] + '.mtl'
    with open(path_mtl, 'w') as f:
        for i, row in enumerate(mats):
            f.write('newmtl m%d\n' % i)
            for key, knots in (('Sp', row[0:12]), ('Em', row[16:28])):
                k = knots.reshape(6, 2)
                n = 6
                while n > 0 and k[n - 1, 0] == -1 and k[n - 1, 1] == 0:
                    n -= 1                  # trailing (-1, 0) pairs are what the loader pads with
                f.write('%s %s\n' % (key, ' '.join(_fmt(v) for v in k[:n].reshape(-1))))
            f.write('Pr %s\nPm %s\nNi %s\nTf %s\n\n' % tuple(_fmt(v) for v in row[12:16]))
    import os
    with open(path_obj, 'w') as f:
        f.write('mtllib %s\n' % os.path.basename(path_mtl))
        cur = None
        for t, m in zip(tris, tri_mats):
            for v in t:
                f.write('v %s %s %s\n' % tuple(_fmt(c) for c in v))
            if m != cur:
                f.write('usemtl m%d\n' % m)
                cur = m
            f.write('f -3 -2 -1\n')


def random_soup(seed, n_tris=240, n_mats=12, n_emissive=4):
    """A random closed-room scene for fuzzing: the Cornell-sized box [-1,1] x [0,2] x [-1,1] (12 wall triangles) filled with
    random triangles, random materials covering the whole uber-BSDF parameter space (rough / smooth metals, dielectrics with
    dispersion, partial opacity, knots with negative wavelengths = unused) and several emissive materials (more than the two
    light triangles of every bundled asset, so the light pick and its pdf matter).  Returns tris, tri_mats, mats."""
    rng = np.random.default_rng(seed)
    c = np.array([[-1, 0, -1], [1, 0, -1], [1, 2, -1], [-1, 2, -1], [-1, 0, 1], [1, 0, 1], [1, 2, 1], [-1, 2, 1]], np.float32)
    quads = [(0, 1, 2, 3), (5, 4, 7, 6), (4, 0, 3, 7), (1, 5, 6, 2), (3, 2, 6, 7), (4, 5, 1, 0)]
    walls = np.array([[c[a], c[b], c[d]] for a, b, d, _ in [(q[0], q[1], q[2], 0) for q in quads]] +
                     [[c[q[0]], c[q[2]], c[q[3]]] for q in quads], np.float32)
    centre = rng.uniform([-0.9, 0.1, -0.9], [0.9, 1.9, 0.9], (n_tris, 1, 3))
    soup = (centre + rng.normal(0, 0.12, (n_tris, 3, 3))).astype(np.float32)
    soup[:6] = soup[6:12]                                      # exact duplicates: equal Morton codes, t ties
    soup[12, 2] = soup[12, 1]                                  # a degenerate (zero-area) triangle
    tris = np.concatenate([walls, soup]).astype(np.float32)
    mats = np.zeros((n_mats, 28), np.float32)
    for i in range(n_mats):
        lam = np.sort(rng.uniform(380, 720, 6)).astype(np.float32)
        val = rng.uniform(0.05, 0.95, 6).astype(np.float32)
        used = int(rng.integers(1, 7))
        lam[used:] = -1.0
        mats[i, 0:12:2], mats[i, 1:12:2] = lam, val
        kind = i % 4
        mats[i, 12] = [rng.uniform(0.05, 1.0), 0.0, rng.uniform(0.0, 0.3), rng.uniform(0.2, 0.9)][kind]     # roughness
        mats[i, 13] = [0.0, 1.0, 0.0, rng.uniform(0.0, 1.0)][kind]                                          # metalness
        mats[i, 14] = [1.0, 1.0, rng.uniform(1.2, 2.0), rng.uniform(1.0, 1.6)][kind]                        # ref_ix
        mats[i, 15] = [1.0, 1.0, rng.uniform(0.0, 0.4), rng.uniform(0.3, 1.0)][kind]                        # opacity
        mats[i, 16:28:2], mats[i, 17:28:2] = -1.0, 0.0
        if i < n_emissive:
            mats[i, 16:20] = [400.0, rng.uniform(2, 20), 700.0, rng.uniform(2, 20)]
    tri_mats = rng.integers(n_emissive, n_mats, len(tris)).astype(np.uint32)
    lights = rng.choice(np.arange(12, len(tris)), 9, replace=False)
    tri_mats[lights] = rng.integers(0, n_emissive, 9)
    tri_mats[12 + 12] = 0                                      # the degenerate triangle is a light: area 0, pdf inf
    return np.ascontiguousarray(tris), tri_mats, mats
