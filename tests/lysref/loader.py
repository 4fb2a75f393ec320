"""Independent Python restatement of the reference scene loader (ljus/src/lib.rs:41-105 on
tobj 0.1.12), used to cross-check the C++ loader and to regenerate tests/golden/scenes/*.npz."""
import os
import numpy as np


def _f32(tok):
    # decimal -> nearest binary64 -> nearest binary32; for the short decimals in OBJ/MTL files this
    # equals the directly rounded f32 (checked against the C++ loader's strtof in tests)
    return np.float32(float(tok))


def _load_mtl(path):
    mats, names = [], {}
    cur = None
    with open(path) as f:
        for line in f:
            w = line.split()
            if not w or w[0] == '#':
                continue
            key = w[0]
            if key == 'newmtl':
                cur = {'Kd': [0.0, 0.0, 0.0], 'Ni': 1.0, 'unknown': {}}
                names[w[1]] = len(mats)
                mats.append(cur)
            elif cur is None:
                continue
            elif key == 'Kd':
                cur['Kd'] = [float(x) for x in w[1:4]]
            elif key == 'Ni':
                cur['Ni'] = float(w[1])
            elif key in ('Ka', 'Ks', 'Ns', 'd', 'illum', 'map_Ka', 'map_Kd', 'map_Ks', 'map_Ns', 'map_d'):
                pass
            else:
                cur['unknown'][key] = line.strip()[len(key):].strip()
    return mats, names


def _spectrum(m, key, rgb):
    if key in m['unknown']:
        v = [float(x) for x in m['unknown'][key].split()]
        out = list(v[:12])
        while len(out) < 12:
            out.append(-1.0 if (len(out) - len(v)) % 2 == 0 else 0.0)
        return out
    return [610.0, rgb[0], 550.0, rgb[1], 460.0, rgb[2], -1.0, 0.0, -1.0, 0.0, -1.0, 0.0]


def load_obj(path):
    """-> (tris [n,3,3] f32, tri_mats [n] u32, mats [m,28] f32)"""
    pos, tris, tri_mats = [], [], []
    mats, names, cur = [], {}, None
    with open(path) as f:
        for line in f:
            w = line.split()
            if not w or w[0].startswith('#'):
                continue
            if w[0] == 'v':
                pos.append([float(x) for x in w[1:4]])
            elif w[0] == 'f':
                ix = []
                for tok in w[1:]:
                    v = int(tok.split('/')[0])
                    ix.append(len(pos) + v if v < 0 else v - 1)
                for k in range(1, len(ix) - 1):
                    if cur is None:
                        raise ValueError("Mesh doesn't have material")
                    tri_mats.append(cur)
                    tris.append([pos[ix[0]], pos[ix[k]], pos[ix[k + 1]]])
            elif w[0] == 'mtllib':
                mats, names = _load_mtl(os.path.join(os.path.dirname(path), w[1]))
            elif w[0] == 'usemtl':
                cur = names.get(w[1])
    rows = []
    for m in mats:
        u = m['unknown']
        ke = [float(x) for x in u['Ke'].split()] if 'Ke' in u else [0.0, 0.0, 0.0]
        row = _spectrum(m, 'Sp', m['Kd'])
        row += [float(u.get('Pr', 1.0)), float(u.get('Pm', 0.0)), m['Ni'], float(u.get('Tf', 1.0))]
        row += _spectrum(m, 'Em', ke)
        rows.append(row)
    return (np.asarray(tris, dtype=np.float64).astype(np.float32).reshape(-1, 3, 3),
            np.asarray(tri_mats, dtype=np.uint32),
            np.asarray(rows, dtype=np.float64).astype(np.float32).reshape(-1, 28))
