"""ctypes host binding for libtracer.so -- the futhark_* C ABI (include/tracer.h) plus lys_* extensions.

Mirrors the reference hosts: `Context` ~ create_futhark_context (demo-interactive/liblys.c:166-209),
`State.init` ~ do_sdl's futhark_entry_init call (liblys.c:133-144) and Fut::init (demo-save/src/wrapper.rs:34-73),
`State.step/render/key/resize` ~ sdl_loop / handle_sdl_events / window_size_updated (liblys.c:30-123),
`State.sample_points_n` ~ Fut::sample_points_ (wrapper.rs:75-92), `State.sample_n_frames` ~ main.rs:37-41.
Errors surface as TracerError carrying futhark_context_get_error(), like FUT_CHECK (liblys.h:32-40)."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'libtracer.so')     # the one library this binding opens: no override, no fallback
_LJUS = os.path.join(_HERE, 'libljus.so')

# src/sdl.fut key codes used by lib.fut:120-185
KEY = dict(SPACE=0x20, K1=0x31, K2=0x32, a=0x61, d=0x64, i=0x69, k=0x6B, l=0x6C, m=0x6D, n=0x6E, o=0x6F, p=0x70,
           s=0x73, t=0x74, w=0x77, x=0x78, z=0x7A, RIGHT=0x4000004F, LEFT=0x40000050, DOWN=0x40000051, UP=0x40000052)


class TracerError(RuntimeError):
    pass


def lib_path():
    return _SO


def build(force=False):
    """Compile libtracer.so / libtracer.a / libljus.so in-tree with nvcc for sm_100a (no GPU needed)."""
    srcs = [os.path.join(_HERE, 'csrc', f) for f in os.listdir(os.path.join(_HERE, 'csrc'))]
    srcs += [os.path.join(_HERE, '..', 'include', f) for f in ('tracer.h', 'lys_ext.h', 'lys_detmath.h', 'lys_pins.h')]
    newest = max(os.path.getmtime(s) for s in srcs)
    stale = force or not all(os.path.exists(p) and os.path.getmtime(p) >= newest for p in (_SO, _LJUS, os.path.join(_HERE, 'libtracer.a')))
    if stale:
        subprocess.check_call(['make', '-C', _HERE, '-j4'], stdout=subprocess.DEVNULL)
    return _SO


class StateInfo(C.Structure):
    _fields_ = [('dim_w', C.c_uint32), ('dim_h', C.c_uint32), ('subsampling', C.c_uint32), ('rng', C.c_uint32),
                ('img_h', C.c_uint32), ('img_w', C.c_uint32), ('n_frames', C.c_uint32), ('cam_conf_id', C.c_uint32),
                ('mode', C.c_int32), ('render_mode', C.c_int32), ('cam_pitch', C.c_float), ('cam_yaw', C.c_float),
                ('cam_origin', C.c_float * 3), ('aperture', C.c_float), ('focal_dist', C.c_float),
                ('ambience', C.c_float * 12), ('n_tris', C.c_int64), ('n_mats', C.c_int64), ('n_lights', C.c_int64)]


class PassStats(C.Structure):
    _fields_ = [('paths', C.c_uint64), ('vertices', C.c_uint64), ('closest_rays', C.c_uint64), ('shadow_rays', C.c_uint64),
                ('launches', C.c_uint64), ('device_ms', C.c_float)]


_lib = None
vp = C.c_void_p


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise TracerError('libtracer.so is not built (run __graft_entry__.build() or make -C %s); there is no CPU fallback' % _HERE)
    L = C.CDLL(_SO)
    L.futhark_context_config_new.restype = vp
    L.futhark_context_config_free.argtypes = [vp]
    L.futhark_context_config_set_device.argtypes = [vp, C.c_char_p]
    L.futhark_context_new.restype = vp
    L.futhark_context_new.argtypes = [vp]
    L.futhark_context_free.argtypes = [vp]
    L.futhark_context_sync.argtypes = [vp]
    L.futhark_context_get_error.restype = vp
    L.futhark_context_get_error.argtypes = [vp]
    for nm, ct, rank in (('f32_1d', C.c_float, 1), ('f32_2d', C.c_float, 2), ('f32_3d', C.c_float, 3), ('u32_1d', C.c_uint32, 1), ('i32_2d', C.c_int32, 2)):
        f = getattr(L, 'futhark_new_' + nm)
        f.restype = vp
        f.argtypes = [vp, vp] + [C.c_int64] * rank
        getattr(L, 'futhark_free_' + nm).argtypes = [vp, vp]
        getattr(L, 'futhark_values_' + nm).argtypes = [vp, vp, vp]
        g = getattr(L, 'futhark_shape_' + nm)
        g.restype = C.POINTER(C.c_int64)
        g.argtypes = [vp, vp]
        r = getattr(L, 'futhark_new_raw_' + nm)                 # device-to-device copy from (CUdeviceptr, byte offset)
        r.restype = vp
        r.argtypes = [vp, C.c_uint64, C.c_int] + [C.c_int64] * rank
        v = getattr(L, 'futhark_values_raw_' + nm)              # the array's own device memory (CUdeviceptr)
        v.restype = C.c_uint64
        v.argtypes = [vp, vp]
    L.futhark_context_config_set_profiling.argtypes = [vp, C.c_int]
    L.futhark_context_pause_profiling.argtypes = [vp]
    L.futhark_context_unpause_profiling.argtypes = [vp]
    L.futhark_context_report.restype = vp
    L.futhark_context_report.argtypes = [vp]
    L.futhark_context_clear_caches.argtypes = [vp]
    L.futhark_free_opaque_state.argtypes = [vp, vp]
    L.futhark_entry_init.argtypes = [vp, C.POINTER(vp), C.c_int32, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, vp, C.c_float, C.c_float, vp]
    L.futhark_entry_resize.argtypes = [vp, C.POINTER(vp), C.c_uint32, C.c_uint32, vp]
    L.futhark_entry_step.argtypes = [vp, C.POINTER(vp), vp]
    L.futhark_entry_key.argtypes = [vp, C.POINTER(vp), C.c_int32, C.c_int32, vp]
    L.futhark_entry_render.argtypes = [vp, C.POINTER(vp), vp]
    L.futhark_entry_sample_n_frames.argtypes = [vp, C.POINTER(vp), vp, C.c_uint32]
    L.futhark_entry_sample_points_n.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), vp, C.c_uint32]
    # extensions
    L.lys_context_set_path_len.argtypes = [vp, C.c_int]
    L.lys_context_set_refit_mode.argtypes = [vp, C.c_int]
    L.lys_context_set_partition.argtypes = [vp, C.c_int, C.c_int]
    L.lys_context_set_profiling.argtypes = [vp, C.c_int]
    L.lys_context_profile_get.argtypes = [vp, vp, vp, C.c_int]
    L.lys_context_profile_detail.argtypes = [vp, vp]
    L.lys_state_advance_rng.argtypes = [vp, C.POINTER(vp), vp, C.c_uint32]
    L.lys_context_device.argtypes = [vp]
    L.lys_context_stream.restype = vp
    L.lys_context_stream.argtypes = [vp]
    L.lys_context_launch_count.restype = C.c_uint64
    L.lys_context_launch_count.argtypes = [vp]
    L.lys_device_ptr_f32_3d.restype = vp
    L.lys_device_ptr_f32_3d.argtypes = [vp, vp]
    L.lys_state_image_device_ptr.restype = vp
    L.lys_state_image_device_ptr.argtypes = [vp, vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.lys_new_f32_3d_from_device.restype = vp
    L.lys_new_f32_3d_from_device.argtypes = [vp, vp, C.c_int64, C.c_int64, C.c_int64]
    L.lys_new_u32_1d_from_device.restype = vp
    L.lys_new_u32_1d_from_device.argtypes = [vp, vp, C.c_int64]
    L.lys_state_info_get.argtypes = [vp, vp, C.POINTER(StateInfo)]
    L.lys_state_image.argtypes = [vp, vp, vp]
    L.lys_state_bvh_get.argtypes = [vp, vp] + [vp] * 9
    L.lys_state_light_indices.argtypes = [vp, vp, vp]
    L.lys_state_bvh_rebuild_timed.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_float)]
    L.lys_probe_primary.argtypes = [vp, vp, vp, vp, vp]
    L.lys_probe_pass.argtypes = [vp, vp, vp, vp, vp]
    L.lys_trace_closest.argtypes = [vp, vp, vp, C.c_int64, vp, vp]
    L.lys_trace_any.argtypes = [vp, vp, vp, vp, C.c_int64, vp]
    L.lys_eval_math.argtypes = [vp, C.c_int, vp, vp, C.c_int64]
    L.lys_material_probe.argtypes = [vp, vp, C.c_float, vp, vp, vp, C.c_uint32, vp]
    L.lys_sample_n_frames_stats.argtypes = [vp, C.POINTER(vp), vp, C.c_uint32, C.POINTER(PassStats)]
    L.lys_sample_n_frames_weighted.argtypes = [vp, C.POINTER(vp), vp, C.c_uint32, C.c_float, C.POINTER(PassStats)]
    _lib = L
    return L


def exported_symbols():
    """Names the shared library must export (checked against include/*.h by the CPU test-suite)."""
    return _load()


_libc = C.CDLL(None)
_libc.free.argtypes = [vp]


def _ptr(a):
    return a.ctypes.data_as(vp) if a is not None else None


def load_obj(path):
    """OBJ/MTL -> (tris [n,3,3] f32, tri_mats [n] u32, mats [m,28] f32) through libljus.so's load_obj_data
    (the C symbol demo-interactive/liblys.h:14-17 binds)."""
    if not os.path.exists(_LJUS):
        raise TracerError('libljus.so is not built')
    L = C.CDLL(_LJUS)
    L.load_obj_data.argtypes = [C.c_char_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.POINTER(C.c_float)),
                                C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.POINTER(C.c_float))]
    L.free_obj_data.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
    nt, nm = C.c_size_t(), C.c_size_t()
    pt, pm, pd = C.POINTER(C.c_float)(), C.POINTER(C.c_uint32)(), C.POINTER(C.c_float)()
    L.load_obj_data(os.fsencode(path), C.byref(nt), C.byref(nm), C.byref(pt), C.byref(pm), C.byref(pd))
    tris = np.ctypeslib.as_array(pt, shape=(nt.value * 9,)).copy().reshape(-1, 3, 3)
    tm = np.ctypeslib.as_array(pm, shape=(nt.value,)).copy()
    mats = np.ctypeslib.as_array(pd, shape=(nm.value,)).copy().reshape(-1, 28)
    L.free_obj_data(pt, pm, pd)
    return tris, tm, mats


class Context:
    """futhark_context_config_new + futhark_context_new (liblys.c:166-209)."""

    def __init__(self, device=None, profiling=False):
        L = _load()
        self._L = L
        cfg = L.futhark_context_config_new()
        if device is not None:
            # an int is a device ordinal ("#k", the generated backend's syntax); a str goes through as is ("#1 B200", "B200")
            L.futhark_context_config_set_device(cfg, ('#%d' % device if isinstance(device, int) else str(device)).encode())
        if profiling:
            L.futhark_context_config_set_profiling(cfg, 1)
        self._ctx = L.futhark_context_new(cfg)
        L.futhark_context_config_free(cfg)
        if not self._ctx:
            raise TracerError('futhark_context_new failed: no usable CUDA device (libtracer has no CPU path)')

    def close(self):
        if getattr(self, '_ctx', None):
            self._L.futhark_context_free(self._ctx)
            self._ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def check(self, rc, what=''):
        if rc != 0:
            p = self._L.futhark_context_get_error(self._ctx)
            msg = C.string_at(p).decode() if p else ''
            if p:
                _libc.free(p)
            raise TracerError('%s failed (%d): %s' % (what, rc, msg))

    def sync(self):
        self.check(self._L.futhark_context_sync(self._ctx), 'futhark_context_sync')

    def report(self):
        """futhark_context_report: launches, device time per kernel class (when profiling), pooled device memory."""
        p = self._L.futhark_context_report(self._ctx)
        txt = C.string_at(p).decode() if p else ''
        if p:
            _libc.free(p)
        return txt

    # knobs (lys_ext.h)
    def set_path_len(self, n):
        self.check(self._L.lys_context_set_path_len(self._ctx, int(n)), 'lys_context_set_path_len')

    def set_refit_mode(self, m):
        self.check(self._L.lys_context_set_refit_mode(self._ctx, int(m)), 'lys_context_set_refit_mode')

    def set_partition(self, rank, world):
        self.check(self._L.lys_context_set_partition(self._ctx, int(rank), int(world)), 'lys_context_set_partition')

    def set_profiling(self, on):
        # True / 1: per-class sequence (exclusive times); 2: the production sequence with events around its launches (shares)
        self.check(self._L.lys_context_set_profiling(self._ctx, 2 if on == 2 else int(bool(on))), 'lys_context_set_profiling')

    def profile(self, reset=True):
        """-> {class: (device ms, launches)} for generate / trace / shade / accumulate."""
        ms = np.zeros(5, np.float32)
        n = np.zeros(5, np.uint64)
        self.check(self._L.lys_context_profile_get(self._ctx, _ptr(ms), _ptr(n), int(reset)), 'lys_context_profile_get')
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(('generate', 'trace', 'shade', 'tail', 'accumulate'))}

    def profile_detail(self):
        ms = np.zeros(36, np.float32)
        self.check(self._L.lys_context_profile_detail(self._ctx, _ptr(ms)), 'lys_context_profile_detail')
        return {'trace': ms[:18].tolist(), 'shade': ms[18:].tolist()}

    @property
    def device(self):
        return int(self._L.lys_context_device(self._ctx))

    @property
    def stream(self):
        return int(self._L.lys_context_stream(self._ctx) or 0)

    @property
    def launches(self):
        return int(self._L.lys_context_launch_count(self._ctx))

    def eval_math(self, fn, x):
        fnid = {'sin': 0, 'cos': 1, 'exp': 2, 'log': 3, 'pow5': 4, 'acos': 5, 'probit': 6}[fn]
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty_like(x)
        self.check(self._L.lys_eval_math(self._ctx, fnid, _ptr(x), _ptr(out), x.size), 'lys_eval_math')
        return out

    def material_probe(self, mat28, wavelen, wo, wi, normal, rng):
        a = [np.ascontiguousarray(v, np.float32) for v in (mat28, wo, wi, normal)]
        out = np.empty(9, np.float32)
        self.check(self._L.lys_material_probe(self._ctx, _ptr(a[0]), wavelen, _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), rng, _ptr(out)), 'lys_material_probe')
        return out


class State:
    """The reference's opaque `state` (src/state.fut:8-19) living on the device."""

    def __init__(self, ctx, ptr):
        self.ctx, self._p = ctx, ptr

    @classmethod
    def init(cls, ctx, tris, tri_mats, mats, h, w, seed=0, cam_conf_id=0, pitch=0.0, yaw=0.0, origin=(0.0, 0.8, 1.8)):
        L = ctx._L
        tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 3, 3)
        tri_mats = np.ascontiguousarray(tri_mats, np.uint32)
        mats = np.ascontiguousarray(mats, np.float32).reshape(-1, 28)
        org = np.ascontiguousarray(origin, np.float32)
        a_t = L.futhark_new_f32_3d(ctx._ctx, _ptr(tris), tris.shape[0], 3, 3)
        a_m = L.futhark_new_u32_1d(ctx._ctx, _ptr(tri_mats), tri_mats.shape[0])
        a_d = L.futhark_new_f32_2d(ctx._ctx, _ptr(mats), mats.shape[0], 28)
        a_o = L.futhark_new_f32_1d(ctx._ctx, _ptr(org), 3)
        try:
            if not (a_t and a_m and a_d and a_o):
                ctx.check(1, 'futhark_new_*')
            out = vp()
            ctx.check(L.futhark_entry_init(ctx._ctx, C.byref(out), seed, h, w, cam_conf_id, a_t, a_m, a_d, pitch, yaw, a_o), 'futhark_entry_init')
        finally:
            for nm, a in (('f32_3d', a_t), ('u32_1d', a_m), ('f32_2d', a_d), ('f32_1d', a_o)):
                if a:
                    getattr(L, 'futhark_free_' + nm)(ctx._ctx, a)
        return cls(ctx, out.value)

    def free(self):
        if getattr(self, '_p', None) and self.ctx._ctx:
            self.ctx._L.futhark_free_opaque_state(self.ctx._ctx, self._p)
        self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def _entry(self, name, *args):
        out = vp()
        self.ctx.check(getattr(self.ctx._L, name)(self.ctx._ctx, C.byref(out), *args), name)
        return out.value

    def step(self):
        return State(self.ctx, self._entry('futhark_entry_step', self._p))

    def key(self, key, e=0):
        return State(self.ctx, self._entry('futhark_entry_key', e, key, self._p))

    def resize(self, h, w):
        return State(self.ctx, self._entry('futhark_entry_resize', h, w, self._p))

    def advance_rng(self, k):
        return State(self.ctx, self._entry('lys_state_advance_rng', self._p, k))

    def _take(self, arr, nm, dtype, out=None):
        L = self.ctx._L
        rank = int(nm[-2])
        shp = getattr(L, 'futhark_shape_' + nm)(self.ctx._ctx, arr)
        shape = tuple(int(shp[i]) for i in range(rank))
        if out is None:
            out = np.empty(shape, dtype)
        elif out.shape != shape or out.dtype != dtype or not out.flags['C_CONTIGUOUS']:
            getattr(L, 'futhark_free_' + nm)(self.ctx._ctx, arr)
            raise ValueError('out must be a C-contiguous %s array of shape %s' % (np.dtype(dtype).name, shape))
        try:
            self.ctx.check(getattr(L, 'futhark_values_' + nm)(self.ctx._ctx, arr, _ptr(out)), 'futhark_values_' + nm)
        finally:
            getattr(L, 'futhark_free_' + nm)(self.ctx._ctx, arr)
        return out

    def render(self, out=None):
        """futhark_entry_render + futhark_values_i32_2d + futhark_free_i32_2d (liblys.c:113-115).  `out`: a frame buffer to
        fill (the reference host reads every frame into the same SDL surface); default: a new array."""
        return self._take(self._entry('futhark_entry_render', self._p), 'i32_2d', np.int32, out)

    def sample_n_frames(self, n):
        return self._take(self._entry('futhark_entry_sample_n_frames', self._p, n), 'f32_3d', np.float32)

    def sample_n_frames_device(self, n, want_stats=True, weight=1.0):
        """-> (futhark_f32_3d handle, device pointer, shape, stats dict); caller frees with free_f32_3d.
        weight != 1: the image is multiplied by it inside the last accumulate kernel (pass-split multi-GPU frames)."""
        L = self.ctx._L
        out = vp()
        st = PassStats()
        self.ctx.check(L.lys_sample_n_frames_weighted(self.ctx._ctx, C.byref(out), self._p, n, float(weight), C.byref(st) if want_stats else None),
                       'lys_sample_n_frames_weighted')
        shp = L.futhark_shape_f32_3d(self.ctx._ctx, out.value)
        stats = {k: getattr(st, k) for k, _ in PassStats._fields_}
        return out.value, int(L.lys_device_ptr_f32_3d(self.ctx._ctx, out.value)), tuple(int(shp[i]) for i in range(3)), stats

    def free_f32_3d(self, handle):
        self.ctx._L.futhark_free_f32_3d(self.ctx._ctx, handle)

    def values_f32_3d(self, handle, shape):
        out = np.empty(shape, np.float32)
        self.ctx.check(self.ctx._L.futhark_values_f32_3d(self.ctx._ctx, handle, _ptr(out)), 'futhark_values_f32_3d')
        return out

    def sample_points_n(self, spp):
        L = self.ctx._L
        o0, o1 = vp(), vp()
        self.ctx.check(L.futhark_entry_sample_points_n(self.ctx._ctx, C.byref(o0), C.byref(o1), self._p, spp), 'futhark_entry_sample_points_n')
        return State(self.ctx, o0.value), self._take(o1.value, 'f32_3d', np.float32)

    # ---- introspection (lys_ext.h)
    def info(self):
        si = StateInfo()
        self.ctx.check(self.ctx._L.lys_state_info_get(self.ctx._ctx, self._p, C.byref(si)), 'lys_state_info_get')
        d = {k: getattr(si, k) for k, _ in StateInfo._fields_}
        d['cam_origin'] = np.array(list(si.cam_origin), np.float32)
        d['ambience'] = np.array(list(si.ambience), np.float32)
        return d

    def image(self):
        i = self.info()
        out = np.empty((i['img_h'], i['img_w'], 3), np.float32)
        self.ctx.check(self.ctx._L.lys_state_image(self.ctx._ctx, self._p, _ptr(out)), 'lys_state_image')
        return out

    def image_device_ptr(self):
        h, w = C.c_uint32(), C.c_uint32()
        p = self.ctx._L.lys_state_image_device_ptr(self.ctx._ctx, self._p, C.byref(h), C.byref(w))
        return int(p), (int(h.value), int(w.value), 3)

    def grid(self):
        i = self.info()
        s = i['subsampling']
        return (i['dim_h'] + s - 1) // s, (i['dim_w'] + s - 1) // s

    def bvh(self):
        n = self.info()['n_tris']
        d = dict(bounds=np.empty(6, np.float32), morton=np.empty(n, np.uint32), src_index=np.empty(n, np.int32),
                 left=np.empty(n - 1, np.int32), right=np.empty(n - 1, np.int32), parent=np.empty(n - 1, np.int32),
                 node_aabb=np.empty((n - 1, 6), np.float32), leaf_aabb=np.empty((n, 6), np.float32), height=np.empty(n - 1, np.int32))
        self.ctx.check(self.ctx._L.lys_state_bvh_get(self.ctx._ctx, self._p, *[_ptr(d[k]) for k in
                       ('bounds', 'morton', 'src_index', 'left', 'right', 'parent', 'node_aabb', 'leaf_aabb', 'height')]), 'lys_state_bvh_get')
        return d

    def light_indices(self):
        out = np.empty(self.info()['n_lights'], np.int32)
        self.ctx.check(self.ctx._L.lys_state_light_indices(self.ctx._ctx, self._p, _ptr(out)), 'lys_state_light_indices')
        return out

    def bvh_rebuild_ms(self, reps=5):
        ms = C.c_float()
        self.ctx.check(self.ctx._L.lys_state_bvh_rebuild_timed(self.ctx._ctx, self._p, reps, C.byref(ms)), 'lys_state_bvh_rebuild_timed')
        return float(ms.value)

    def probe_primary(self):
        gh, gw = self.grid()
        leaf, src, t = np.empty((gh, gw), np.int32), np.empty((gh, gw), np.int32), np.empty((gh, gw), np.float32)
        self.ctx.check(self.ctx._L.lys_probe_primary(self.ctx._ctx, self._p, _ptr(leaf), _ptr(src), _ptr(t)), 'lys_probe_primary')
        return dict(leaf=leaf, src_tri=src, t=t)

    def probe_pass(self):
        gh, gw = self.grid()
        rad, dist, ch = np.empty((gh, gw, 16), np.float32), np.empty((gh, gw, 16), np.float32), np.empty((gh, gw), np.int32)
        self.ctx.check(self.ctx._L.lys_probe_pass(self.ctx._ctx, self._p, _ptr(rad), _ptr(dist), _ptr(ch)), 'lys_probe_pass')
        return dict(radiance=rad, distance=dist, channel=ch)

    def trace_closest(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        leaf, t = np.empty(len(rays), np.int32), np.empty(len(rays), np.float32)
        self.ctx.check(self.ctx._L.lys_trace_closest(self.ctx._ctx, self._p, _ptr(rays), len(rays), _ptr(leaf), _ptr(t)), 'lys_trace_closest')
        return leaf, t

    def trace_any(self, rays, tmax):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        tmax = np.ascontiguousarray(tmax, np.float32)
        out = np.empty(len(rays), np.int32)
        self.ctx.check(self.ctx._L.lys_trace_any(self.ctx._ctx, self._p, _ptr(rays), _ptr(tmax), len(rays), _ptr(out)), 'lys_trace_any')
        return out
