"""Scene helpers: the BASELINE configs' inputs.

`synthetic_cornell(k)` is BASELINE config 5: every quad of the 22-quad Cornell box is
tessellated into a k x k grid of quads (2 triangles each) with the same materials; k = 151 gives 1 003 244
triangles.  Vertices are bilinear interpolants evaluated in float64 and rounded once to f32 (not a dyadic grid:
the scene-bounds fold is therefore exercised in its rounding-sensitive regime)."""
import numpy as np


def quads_of(tris, tri_mats):
    """Pairs of consecutive triangles (a,b,c),(a,c,d) as produced by fan triangulation -> quads [q,4,3]."""
    t = np.asarray(tris, np.float32).reshape(-1, 3, 3)
    assert len(t) % 2 == 0
    q = np.empty((len(t) // 2, 4, 3), np.float32)
    q[:, 0] = t[0::2, 0]
    q[:, 1] = t[0::2, 1]
    q[:, 2] = t[0::2, 2]
    q[:, 3] = t[1::2, 2]
    assert np.array_equal(t[1::2, 0], t[0::2, 0]) and np.array_equal(t[1::2, 1], t[0::2, 2])
    return q, np.asarray(tri_mats, np.uint32)[0::2]


def synthetic_cornell(cornell_tris, cornell_tri_mats, k):
    q, qm = quads_of(cornell_tris, cornell_tri_mats)
    q = q.astype(np.float64)
    s = np.arange(k + 1, dtype=np.float64) / k
    S, T = np.meshgrid(s, s, indexing='xy')                     # S varies along columns
    out_t, out_m = [], []
    for a, b, c, d in q:
        P = ((1 - S) * (1 - T))[..., None] * a + (S * (1 - T))[..., None] * b + (S * T)[..., None] * c + ((1 - S) * T)[..., None] * d
        P = P.astype(np.float32)
        p00, p10, p11, p01 = P[:-1, :-1], P[:-1, 1:], P[1:, 1:], P[1:, :-1]
        cell = np.stack([np.stack([p00, p10, p11], axis=2), np.stack([p00, p11, p01], axis=2)], axis=2)  # [k,k,2,3,3]
        out_t.append(cell.reshape(-1, 3, 3))
    tris = np.concatenate(out_t, axis=0)
    mats = np.repeat(qm, 2 * k * k)
    return np.ascontiguousarray(tris, np.float32), np.ascontiguousarray(mats, np.uint32)
