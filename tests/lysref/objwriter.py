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
