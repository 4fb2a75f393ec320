"""ctypes binding for oracle/liblys_oracle.so (the CPU restatement of the reference).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs."""
import ctypes as C
import os
import subprocess
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_SO = os.path.join(_ROOT, 'oracle', 'liblys_oracle.so')

f32p = np.ctypeslib.ndpointer(np.float32, flags='C_CONTIGUOUS')
u32p = np.ctypeslib.ndpointer(np.uint32, flags='C_CONTIGUOUS')
i32p = np.ctypeslib.ndpointer(np.int32, flags='C_CONTIGUOUS')


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ('paths', 'vertices', 'closest_rays', 'shadow_rays', 'node_visits',
                                           'box_tests', 'tri_tests', 'loop_iters', 'closest_box', 'closest_tri', 'shadow_box', 'shadow_tri')]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build():
    src = os.path.join(_ROOT, 'oracle', 'lys_oracle.cpp')
    if (not os.path.exists(_SO)) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)):
        subprocess.check_call(['make', '-C', os.path.join(_ROOT, 'oracle')], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp = C.c_void_p
    L.orc_init.restype = vp
    L.orc_init.argtypes = [C.c_int32, C.c_uint32, C.c_uint32, C.c_uint32, f32p, u32p, C.c_int64, f32p, C.c_int64,
                           C.c_float, C.c_float, f32p]
    for name, args in (('orc_resize', [C.c_uint32, C.c_uint32, vp]), ('orc_key', [C.c_int32, C.c_int32, vp]),
                       ('orc_step', [vp]), ('orc_sample_points_n', [vp, C.c_uint32, f32p])):
        getattr(L, name).restype = vp
        getattr(L, name).argtypes = args
    L.orc_advance_rng.restype = vp
    L.orc_advance_rng.argtypes = [vp, C.c_uint32]
    L.orc_render.argtypes = [vp, i32p]
    L.orc_sample_n_frames.argtypes = [vp, C.c_uint32, f32p]
    L.orc_free_state.argtypes = [vp]
    L.orc_state_dims.argtypes = [vp] + [C.POINTER(C.c_uint32)] * 4
    L.orc_state_image.argtypes = [vp, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.orc_state_scalars.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                    C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_uint32), f32p, f32p]
    L.orc_bvh_size.restype = C.c_int64
    L.orc_bvh_size.argtypes = [vp]
    L.orc_n_lights.restype = C.c_int64
    L.orc_n_lights.argtypes = [vp]
    L.orc_bvh_get.argtypes = [vp, f32p, u32p, i32p, i32p, i32p, i32p, f32p, f32p]
    L.orc_light_indices.argtypes = [vp, i32p]
    L.orc_expand_bits.restype = C.c_uint32
    L.orc_expand_bits.argtypes = [C.c_uint32]
    L.orc_morton3d.restype = C.c_uint32
    L.orc_morton3d.argtypes = [C.c_float] * 3
    L.orc_hash.restype = C.c_uint32
    L.orc_hash.argtypes = [C.c_int32]
    L.orc_rng_from_seed.restype = C.c_uint32
    L.orc_rng_from_seed.argtypes = [C.c_int32]
    L.orc_rng_next.restype = C.c_uint32
    L.orc_rng_next.argtypes = [C.c_uint32]
    L.orc_rng_uniform.restype = C.c_float
    L.orc_rng_uniform.argtypes = [C.c_uint32, C.c_float, C.c_float, C.POINTER(C.c_uint32)]
    L.orc_radix_tree.argtypes = [u32p, C.c_int64, i32p, i32p, i32p]
    L.orc_spectrum_lookup.restype = C.c_float
    L.orc_spectrum_lookup.argtypes = [C.c_float, f32p]
    L.orc_eval_math.argtypes = [C.c_int, f32p, f32p, C.c_int64]
    L.orc_material_probe.argtypes = [f32p, C.c_float, f32p, f32p, f32p, C.c_uint32, f32p]
    L.orc_probe_primary.argtypes = [vp, i32p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_probe_pass.argtypes = [vp, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_brute_force_hits.argtypes = [vp, f32p, C.c_int64, i32p, f32p]
    L.orc_closest_hits.argtypes = [vp, f32p, C.c_int64, i32p, f32p]
    L.orc_any_hits.argtypes = [vp, f32p, f32p, C.c_int64, i32p]
    L.orc_closest_hits_steps.argtypes = [vp, f32p, C.c_int64, i32p, i32p]
    L.orc_probe_path_rays.argtypes = [vp, C.c_void_p, i32p]
    L.orc_closest_hits_pattern.argtypes = [vp, f32p, C.c_int64, C.c_int32, C.c_void_p, i32p]
    L.orc_counters_get.argtypes = [C.POINTER(Counters)]
    _lib = L
    return L


MATH_FN = {'sin': 0, 'cos': 1, 'exp': 2, 'log': 3, 'pow5': 4, 'acos': 5, 'probit': 6}


def eval_math(fn, x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    lib().orc_eval_math(MATH_FN[fn], x, out, x.size)
    return out


def set_path_len(n):
    lib().orc_set_path_len(int(n))


def set_refit_mode(m):
    lib().orc_set_refit_mode(int(m))


def set_math_mode(m):
    lib().orc_set_math_mode(int(m))


def set_threads(n):
    lib().orc_set_threads(int(n))


def get_threads():
    return int(lib().orc_get_threads())


def counters_reset():
    lib().orc_counters_reset()


def counters():
    c = Counters()
    lib().orc_counters_get(C.byref(c))
    return c.as_dict()


def radix_tree(keys):
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    n = keys.size
    l = np.empty(n - 1, np.int32)
    r = np.empty(n - 1, np.int32)
    p = np.empty(n - 1, np.int32)
    lib().orc_radix_tree(keys, n, l, r, p)
    return l, r, p


class State:
    """Mirror of the reference's opaque `state` (state.fut:8-19) as evaluated by the oracle."""

    def __init__(self, ptr):
        if not ptr:
            raise RuntimeError('oracle returned a null state')
        self._p = ptr

    @classmethod
    def init(cls, tris, tri_mats, mats, h, w, seed=0, cam_conf_id=0, pitch=0.0, yaw=0.0, origin=(0.0, 0.8, 1.8)):
        tris = np.ascontiguousarray(tris, np.float32)
        tri_mats = np.ascontiguousarray(tri_mats, np.uint32)
        mats = np.ascontiguousarray(mats, np.float32)
        o = np.asarray(origin, np.float32)
        return cls(lib().orc_init(seed, h, w, cam_conf_id, tris.reshape(-1), tri_mats, tri_mats.size, mats.reshape(-1),
                                  mats.size // 28, pitch, yaw, o))

    def __del__(self):
        if getattr(self, '_p', None):
            lib().orc_free_state(self._p)
            self._p = None

    def step(self):
        return State(lib().orc_step(self._p))

    def key(self, key, e=0):
        return State(lib().orc_key(e, key, self._p))

    def resize(self, h, w):
        return State(lib().orc_resize(h, w, self._p))

    def advance_rng(self, k):
        return State(lib().orc_advance_rng(self._p, k))

    def dims(self):
        v = [C.c_uint32() for _ in range(4)]
        lib().orc_state_dims(self._p, *[C.byref(x) for x in v])
        return tuple(int(x.value) for x in v)  # w, h, grid_w, grid_h

    def render(self):
        w, h, _, _ = self.dims()
        out = np.empty((h, w), np.int32)
        lib().orc_render(self._p, out)
        return out

    def sample_n_frames(self, n):
        _, _, gw, gh = self.dims()
        out = np.empty((gh, gw, 3), np.float32)
        lib().orc_sample_n_frames(self._p, n, out.reshape(-1))
        return out

    def sample_points_n(self, spp):
        _, _, gw, gh = self.dims()
        out = np.empty((gh, gw, 4), np.float32)
        st = State(lib().orc_sample_points_n(self._p, spp, out.reshape(-1)))
        return st, out

    def image(self):
        ih, iw = C.c_uint32(), C.c_uint32()
        lib().orc_state_image(self._p, None, C.byref(ih), C.byref(iw))
        out = np.empty((ih.value, iw.value, 3), np.float32)
        lib().orc_state_image(self._p, out.ctypes.data_as(C.c_void_p), None, None)
        return out

    def scalars(self):
        rng, nf, sub, cid = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        mode, rm = C.c_int32(), C.c_int32()
        cam = np.empty(7, np.float32)
        amb = np.empty(12, np.float32)
        lib().orc_state_scalars(self._p, C.byref(rng), C.byref(nf), C.byref(sub), C.byref(mode), C.byref(rm),
                                C.byref(cid), cam, amb)
        return dict(rng=rng.value, n_frames=nf.value, subsampling=sub.value, mode=mode.value, render_mode=rm.value,
                    cam_conf_id=cid.value, cam=cam, ambience=amb)

    def bvh(self):
        n = int(lib().orc_bvh_size(self._p))
        d = dict(bounds=np.empty(6, np.float32), morton=np.empty(n, np.uint32), src_index=np.empty(n, np.int32),
                 left=np.empty(n - 1, np.int32), right=np.empty(n - 1, np.int32), parent=np.empty(n - 1, np.int32),
                 node_aabb=np.empty((n - 1, 6), np.float32), leaf_aabb=np.empty((n, 6), np.float32))
        lib().orc_bvh_get(self._p, d['bounds'], d['morton'], d['src_index'], d['left'], d['right'], d['parent'],
                          d['node_aabb'].reshape(-1), d['leaf_aabb'].reshape(-1))
        return d

    def light_indices(self):
        n = int(lib().orc_n_lights(self._p))
        out = np.empty(n, np.int32)
        if n:
            lib().orc_light_indices(self._p, out)
        return out

    def probe_primary(self, want_rays=False):
        _, _, gw, gh = self.dims()
        leaf = np.empty((gh, gw), np.int32)
        src = np.empty((gh, gw), np.int32)
        t = np.empty((gh, gw), np.float32)
        rays = np.empty((gh, gw, 6), np.float32) if want_rays else None
        wl = np.empty((gh, gw), np.float32) if want_rays else None
        vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        lib().orc_probe_primary(self._p, leaf.reshape(-1), vp(src), vp(t), vp(rays), vp(wl))
        return dict(leaf=leaf, src_tri=src, t=t, rays=rays, wavelen=wl)

    def probe_pass(self):
        _, _, gw, gh = self.dims()
        rad = np.empty((gh, gw, 16), np.float32)
        dist = np.empty((gh, gw, 16), np.float32)
        ch = np.empty((gh, gw), np.int32)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        lib().orc_probe_pass(self._p, vp(rad), vp(dist), vp(ch))
        return dict(radiance=rad, distance=dist, channel=ch)

    def brute_force_hits(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        src = np.empty(len(rays), np.int32)
        t = np.empty(len(rays), np.float32)
        lib().orc_brute_force_hits(self._p, rays.reshape(-1), len(rays), src, t)
        return src, t

    def closest_hits(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        leaf = np.empty(len(rays), np.int32)
        t = np.empty(len(rays), np.float32)
        lib().orc_closest_hits(self._p, rays.reshape(-1), len(rays), leaf, t)
        return leaf, t

    def closest_hits_steps(self, rays):
        """box tests and triangle tests of every ray's closest-hit walk"""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        box = np.empty(len(rays), np.int32)
        tri = np.empty(len(rays), np.int32)
        lib().orc_closest_hits_steps(self._p, rays.reshape(-1), len(rays), box, tri)
        return box, tri

    def closest_hits_pattern(self, rays, max_steps=256):
        """visit pattern of every ray's walk (0 box fail, 1 box pass, 2 triangle miss, 3 triangle hit, 255 padding) and its length"""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        pat = np.full((len(rays), max_steps), 255, np.uint8)
        ln = np.empty(len(rays), np.int32)
        lib().orc_closest_hits_pattern(self._p, rays.reshape(-1), len(rays), max_steps, pat.ctypes.data_as(C.c_void_p), ln)
        return pat, ln

    def probe_path_rays(self):
        """closest-hit rays of every path of the next pass: rays [gh][gw][16][6], count [gh][gw]"""
        _, _, gw, gh = self.dims()
        rays = np.zeros((gh, gw, 16, 6), np.float32)
        n = np.zeros((gh, gw), np.int32)
        lib().orc_probe_path_rays(self._p, rays.ctypes.data_as(C.c_void_p), n.reshape(-1))
        return rays, n

    def any_hits(self, rays, tmax):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        tmax = np.ascontiguousarray(tmax, np.float32)
        out = np.empty(len(rays), np.int32)
        lib().orc_any_hits(self._p, rays.reshape(-1), tmax, len(rays), out)
        return out
