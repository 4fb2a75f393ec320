"""An independent restatement of the reference's path tracer, used to pin the oracle.

path_trace (src/integrator.fut:27-76), direct_radiance / estimate_direct / sample_light / occluded (src/direct.fut),
diffuselight_incident_radiance (src/light.fut:19-30), get_lights (src/scene.fut:58-66), mkray_adjust_acne (src/shapes.fut:41-46),
closest_hit / any_hit (src/bvh.fut:123-167), hit_aabb / hit_triangle (src/shapes.fut) and the uber material (src/material.fut)
are written here in scalar numpy f32 straight from the .fut text.  Shared with the oracle: only what the reference itself does
not define -- the transcendental contract (include/lys_detmath.h via orc_eval_math), the cpprandom LCG and spectrum_lookup
(all three pinned by tests/test_oracle_kat.py) -- plus the oracle's own BVH arrays and camera rays as INPUTS (both pinned
separately: test_build_against_an_independent_restatement, test_stackless_walk_...).

The check: per-vertex radiance and cumulative distance of every path of one pass (orc_probe_pass), bit for bit."""
import ctypes
import numpy as np
import pytest

F = np.float32
PI = F(np.pi)
INV_PI = F(1.0) / PI
HIGHEST = np.finfo(F).max
INF = F(np.inf)
fmax, fmin = np.fmax, np.fmin


def dot(a, b):
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], F)


def norm(v):
    return np.sqrt(dot(v, v))


def normalise(v):
    return (F(1) / norm(v)) * v


def lerp(a, b, t):
    return a + (b - a) * t


def sgn(x):
    return F(-1) if x < 0 else (F(0) if x == 0 else F(1))


class Fut:
    """one instance per test: holds the oracle handle for the shared primitives"""

    def __init__(self, orc, tris, tri_mats, mats, bvh):
        self.orc, self.L = orc, orc.lib()
        self.tri = np.ascontiguousarray(tris, F).reshape(-1, 3, 3)
        self.tri_mats, self.mats = np.asarray(tri_mats), np.ascontiguousarray(mats, F).reshape(-1, 28)
        self.left, self.right, self.parent = bvh['left'], bvh['right'], bvh['parent']
        self.box, self.order = bvh['node_aabb'], bvh['src_index']
        # get_lights: emissive triangles in input order
        def emissive(row):
            return any(row[16 + 2 * k] >= 0 and row[17 + 2 * k] > 0 for k in range(6))
        self.light_ix = [i for i in range(len(self.tri)) if emissive(self.mats[self.tri_mats[i]])]
        self.lights = [dict(kind='diffuse', tri=self.tri[i], em=self.mats[self.tri_mats[i]][16:28], theta=F(0)) for i in self.light_ix]
        self.extra = []                                                  # the transmitter lights of the current pixel (camera.fut:112-122)

    # ---- shared primitives
    def m1(self, fn, x):
        return self.orc.eval_math(fn, np.array([x], F))[0]

    def lookup(self, wl, knots12):
        return F(self.L.orc_spectrum_lookup(wl, np.ascontiguousarray(knots12, F)))

    def lcg(self, s):
        return int(self.L.orc_rng_next(s))

    def uniform(self, s, lo=F(0), hi=F(0.9999)):
        out = ctypes.c_uint32()
        v = self.L.orc_rng_uniform(s, lo, hi, ctypes.byref(out))
        return out.value, F(v)

    # ---- shapes.fut
    def hit_aabb(self, tmax, o, d, c, h):
        mn, mx = c - h, c + h
        tmin = F(0)
        for a in range(3):
            inv = F(1) / d[a]
            t0, t1 = (mn[a] - o[a]) * inv, (mx[a] - o[a]) * inv
            if inv < 0:
                t0, t1 = t1, t0
            t1 = t1 * (F(1) + F(0.001))
            tmin, tmax = fmax(t0, tmin), fmin(t1, tmax)
            if tmax <= tmin:
                return False
        return True

    def hit_triangle(self, tmax, o, d, A, B, Cc):
        e1, e2 = B - A, Cc - A
        n = cross(e1, e2)
        a = -(dot(n, d))
        if a > F(-0.00001) and a < F(0.00001):
            return None
        sv = o - A
        mv = cross(sv, d)
        inv = F(1) / a
        t, u, v = inv * dot(n, sv), inv * dot(mv, e2), inv * (-(dot(mv, e1)))
        if u >= 0 and v >= 0 and u + v <= 1 and t < tmax and t > 0:
            return t, o + t * d, normalise(n)
        return None

    def mkray_adjust_acne(self, pos, n, wi):
        return pos + F(0.001) * (sgn(dot(wi, n)) * n), normalise(wi)

    # ---- bvh.fut: the stackless walk, closest and any
    def walk(self, tmax0, o, d, any_hit):
        closest, tmax, current, prev, first = -1, tmax0, 0, 0, True
        while current != -1:
            l, r = int(self.left[current]), int(self.right[current])
            came = None if first else prev
            first = False
            if came is not None and came == l:
                child = r
            elif (came is None or came != r) and self.hit_aabb(tmax, o, d, self.box[current, :3], self.box[current, 3:]):
                child = l
            else:
                child = None
            if child is None:
                prev, current = current, int(self.parent[current])
            elif child >= 0:
                prev, current = current, child
            else:
                h = self.hit_triangle(tmax, o, d, *self.tri[self.order[~child]])
                if h is not None:
                    if any_hit:
                        return True
                    closest, tmax = ~child, h[0]
                prev = child
        return False if any_hit else closest

    # ---- material.fut (local space)
    def D(self, alpha, wh):
        t2 = fmax(F(0), F(1) - wh[2] * wh[2]) / (wh[2] * wh[2])
        if np.isinf(t2):
            return F(0)
        return self.m1('exp', -t2 / (alpha * alpha)) / (PI * alpha * alpha * (wh[2] * wh[2]) * (wh[2] * wh[2]))

    def G(self, alpha, wo, wi):
        def lam(w):
            at = np.abs(np.sqrt(fmax(F(0), F(1) - w[2] * w[2])) / w[2])
            if np.isinf(at):
                return F(0)
            a = F(1) / (alpha * at)
            if a >= F(1.6):
                return F(0)
            return (F(1) - F(1.259) * a + F(0.396) * a * a) / (F(3.535) * a + F(2.181) * a * a)
        return F(1) / (F(1) + lam(wo) + lam(wi))

    @staticmethod
    def alpha_of(r):
        return F(1.62142) * fmax(F(0.004), r)

    def refl_bsdf(self, wo, wi, m):
        wh = normalise(wi + wo)
        a = self.alpha_of(m['roughness'])
        return (self.D(a, wh) * self.G(a, wo, wi)) / (F(4) * wo[2] * wi[2])

    def refl_pdf(self, wo, wi, m):
        if not wo[2] * wi[2] > 0:
            return F(0)
        wh = normalise(wo + wi)
        return (self.D(self.alpha_of(m['roughness']), wh) * np.abs(wh[2])) / (F(4) * dot(wo, wh))

    def fresnel(self, wo, m):
        x = (F(1) - m['ref_ix']) / (F(1) + m['ref_ix'])
        r0 = x * x
        return r0 + (F(1) - r0) * self.m1('pow5', F(1) - wo[2])

    def uber_bsdf(self, wo, wi, m):
        refr = lerp(F(0), m['color'] * INV_PI, m['opacity'])
        refl = F(0) if wo[2] <= 0 else self.fresnel(wo, m)
        rb = self.refl_bsdf(wo, wi, m)
        return lerp(lerp(refr, rb, refl), m['color'] * rb, m['metalness'])

    def uber_pdf(self, wo, wi, m):
        dp = wi[2] * INV_PI if wo[2] * wi[2] > 0 else F(0)
        refr = lerp(F(0), dp, m['opacity'])
        rp = self.refl_pdf(wo, wi, m)
        diel = refr if wo[2] <= 0 else lerp(refr, rp, self.fresnel(wo, m))
        return lerp(rp, diel, m['metalness'])

    def sample_reflection(self, wo, m, s):
        s, u0 = self.uniform(s)
        s, u1 = self.uniform(s)
        ls = self.m1('log', F(1) - u0)
        if np.isinf(ls):
            wh, pdf_wh = np.zeros(3, F), F(0)
        else:
            a = self.alpha_of(m['roughness'])
            tan2 = -a * a * ls
            phi = u1 * F(2) * PI
            ct = F(1) / np.sqrt(F(1) + tan2)
            st = np.sqrt(fmax(F(0), F(1) - ct * ct))
            wh = np.array([st * self.m1('cos', phi), st * self.m1('sin', phi), ct], F)
            if not wo[2] * wh[2] > 0:
                wh = -wh
            pdf_wh = self.D(a, wh) * np.abs(ct)
        wi = F(-1) * wo + (F(2) * dot(wo, wh)) * wh
        if not wo[2] * wi[2] > 0:
            return s, (np.zeros(3, F), F(0), 1, F(0))
        kind, pdf = (2, pdf_wh / (F(4) * dot(wo, wh))) if pdf_wh > 0 else (1, F(0))
        return s, (wi, self.refl_bsdf(wo, wi, m), kind, pdf)

    def sample_refraction(self, wo, m, s):
        s, p = self.uniform(s)
        if p < m['opacity']:
            s, theta = self.uniform(s, F(0), F(2) * PI)
            s, u = self.uniform(s)
            r = np.sqrt(u)
            dx, dy = r * self.m1('cos', theta), r * self.m1('sin', theta)
            z = np.sqrt(fmax(F(0), F(1) - (dx * dx + dy * dy)))
            return s, (np.array([dx, dy, z], F), m['color'] * INV_PI, 2, z * INV_PI)
        entering = wo[2] > 0
        n = np.array([0, 0, 1], F) if entering else np.array([-0.0, -0.0, -1.0], F)
        eta = F(1.0) / m['ref_ix'] if entering else m['ref_ix'] / F(1.0)
        ci = dot(n, wo)
        s2t = eta * eta * fmax(F(0), F(1) - ci * ci)
        if s2t >= 1:
            wi = F(-1) * wo + (F(2) * dot(wo, n)) * n
        else:
            wi = (-eta) * wo + (eta * ci - np.sqrt(F(1) - s2t)) * n
        return s, (wi, F(1) / np.abs(wi[2]), 0, F(0))

    def uber_sample(self, wo, m, s):
        s, p = self.uniform(s)
        if p < m['metalness']:
            s, (wi, b, k, pdf) = self.sample_reflection(wo, m, s)
            return s, (wi, m['color'] * b, k, pdf)
        if wo[2] <= 0:
            return self.sample_refraction(wo, m, s)
        r = self.fresnel(wo, m)
        s, q = self.uniform(s)
        return self.sample_reflection(wo, m, s) if q < r else self.sample_refraction(wo, m, s)

    @staticmethod
    def onb(nrm):
        if np.abs(nrm[0]) > np.abs(nrm[2]):
            b = normalise(np.array([-nrm[1], nrm[0], 0], F))
        else:
            b = normalise(np.array([0, -nrm[2], nrm[1]], F))
        return cross(b, nrm), b, nrm

    @staticmethod
    def to_local(o, w):
        return np.array([dot(w, o[0]), dot(w, o[1]), dot(w, o[2])], F)

    def mat_at(self, row, wl):
        return dict(color=self.lookup(wl, row[:12]), roughness=row[12], metalness=row[13], ref_ix=row[14] - (wl - F(589)) / F(10000),
                    opacity=row[15])

    def bsdf_f(self, wo, wi, n, m):
        o = self.onb(n)
        return self.uber_bsdf(self.to_local(o, wo), self.to_local(o, wi), m)

    def bsdf_pdf(self, wo, wi, n, m):
        o = self.onb(n)
        return self.uber_pdf(self.to_local(o, wo), self.to_local(o, wi), m)

    def sample_dir(self, wo, n, m, s):
        o = self.onb(n)
        s, (wi, b, k, pdf) = self.uber_sample(self.to_local(o, wo), m, s)
        return s, ((wi[0] * o[0] + wi[1] * o[1]) + wi[2] * o[2], b, k, pdf)

    # ---- light.fut / direct.fut
    def incident(self, light, hitp, lightp, wl):
        A, B, Cc = light['tri']
        v = lightp - hitp
        wi, d2 = normalise(v), dot(v, v)
        ln = normalise(cross(B - A, Cc - A))
        cl = dot(-wi, ln)
        E = self.lookup(wl, light['em'])
        if light['kind'] == 'diffuse':
            return fmax(F(0), E * cl / d2)
        return E / d2 if self.m1('acos', cl) <= light['theta'] else F(0)   # frustumlight_incident_radiance (light.fut:32-44)

    def occluded(self, pos, n, lightp):
        v = lightp - pos
        w = normalise(v)
        if dot(w, n) <= 0:
            return True
        o, d = self.mkray_adjust_acne(pos, n, w)
        return self.walk(norm(v) - F(0.01), o, d, True)

    def direct_radiance(self, s, wo, pos, n, m, wl):
        lights = self.lights + self.extra
        if not lights:
            return s, F(0)
        s = self.lcg(s)                                                  # random_select: one raw draw
        light = lights[s % len(lights)]
        A, B, Cc = light['tri']
        e1, e2 = B - A, Cc - A
        area = norm(cross(e1, e2)) / F(2)
        # sample_arealight peeks two draws (the advanced rng is dropped, direct.fut:38,42)
        s1, u = self.uniform(s)
        _, v = self.uniform(s1)
        su = np.sqrt(u)
        lu, lv = F(1) - su, v * su
        p = (A + lu * e1) + lv * e2
        wi = normalise(p - pos)
        in_rad = self.incident(light, pos, p, wl)
        pdf = F(1) / area
        if self.occluded(pos, n, p):
            in_rad = F(0)
        if pdf == 0 or in_rad == 0:
            L = F(0)
        else:
            f = self.bsdf_f(wo, wi, n, m) * np.abs(dot(wi, n))
            sp = self.bsdf_pdf(wo, wi, n, m)
            weight = F(1) * pdf / (F(1) * pdf + F(1) * sp)
            L = f * weight * in_rad / pdf
        s, (bwi, bsdf, kind, bpdf) = self.sample_dir(wo, n, m, s)
        ro, rd = self.mkray_adjust_acne(pos, n, bwi)
        lh = self.hit_triangle(HIGHEST, ro, rd, A, B, Cc)
        Bv = F(0)
        if lh is not None and not self.occluded(pos, n, lh[1]):
            in_rad = self.incident(light, pos, lh[1], wl)
            f = bsdf * np.abs(dot(bwi, n))
            if kind == 0:
                Bv = f * in_rad
            elif kind == 2:
                lp = F(1) / area
                weight = F(1) * bpdf / (F(1) * bpdf + F(1) * lp)
                Bv = f * in_rad * weight / bpdf
        light_pdf = F(1) / F(len(lights))
        return s, (L + Bv) / light_pdf

    def disk(self, p, normal, radius, libm):
        """shapes.fut:17-35: 8 sector triangles around p; per-frame trigonometry from the C library like the c backend"""
        a = F(2) * PI / F(8)
        c = cross(normal, np.array([0, 1, 0], F))
        right = np.array([1, 0, 0], F) if norm(c) == 0 else normalise(c)
        up = normalise(cross(right, normal))

        def vec(b):
            x = F(1) * F(libm.cosf(b)) - F(0) * F(libm.sinf(b))           # vec3.rot_z b (1, 0, 0)
            y = F(1) * F(libm.sinf(b)) + F(0) * F(libm.cosf(b))
            return x * right + y * up
        out = []
        for i in range(8):
            v0, v1 = vec(a * F(i)), vec(a * (F(i) + F(1)))
            out.append(np.stack([p, p + radius * v1, p + radius * v0]).astype(F))
        return out

    # ---- integrator.fut
    def path_trace(self, o, d, wl, s, ambience12, path_len=16):
        rad = np.zeros(16, F)
        dist = np.full(16, np.inf, F)
        amb = self.lookup(wl, ambience12)
        i, distance = 0, F(0)
        while i < path_len:
            leaf = self.walk(HIGHEST, o, d, False)
            if leaf < 0:
                dist[i], rad[i] = INF, amb
                break
            t, pos, n = self.hit_triangle(HIGHEST, o, d, *self.tri[self.order[leaf]])
            row = self.mats[self.tri_mats[self.order[leaf]]]
            m = self.mat_at(row, wl)
            s = self.lcg(s)                                              # advance_rng
            wo = -d
            s, direct = self.direct_radiance(s, wo, pos, n, m, wl)
            r = direct + (self.lookup(wl, row[16:28]) if i == 0 else F(0))
            distance = distance + t
            dist[i], rad[i] = distance, r
            s, (wi, bsdf, kind, pdf) = self.sample_dir(wo, n, m, s)
            pdf = F(0) if kind == 1 else (F(1) if kind == 0 else pdf)
            p_term = F(1) - bsdf * np.abs(dot(n, wi)) / pdf
            s, x = self.uniform(s)
            if pdf == 0 or x < p_term:
                break
            i += 1
            o, d = self.mkray_adjust_acne(pos, n, wi)
        return rad, dist


@pytest.mark.parametrize('name,h,w,origin', [('cornell', 36, 48, (0.0, 0.8, 1.8)), ('spectrumsphere', 30, 40, (0.0, 0.8, 1.8)),
                                             ('mirrorbox', 18, 24, (0.0, 0.8, 0.6)), ('spectrumspherehigh', 15, 20, (0.0, 0.8, 1.8))])
def test_path_tracer_against_an_independent_restatement(orc, scenes, name, h, w, origin):
    t9, tm, m = scenes[name]
    st = orc.State.init(t9, tm, m, h, w, origin=origin)
    fut = Fut(orc, t9, tm, m, st.bvh())
    sc = st.scalars()
    prim = st.probe_primary(want_rays=True)
    want = st.probe_pass()
    rays, wls = prim['rays'].reshape(-1, 6), prim['wavelen'].reshape(-1)
    assert fut.light_ix == st.light_indices().tolist()
    L = orc.lib()
    vertices = 0
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        for ix in range(h * w):
            s = sc['rng'] ^ int(L.orc_hash(ix))
            s = fut.lcg(s)                                               # channel draw (camera.fut:72-73)
            s = fut.lcg(s)                                               # wavelength draw (:77); the camera ray only peeks
            rad, dist = fut.path_trace(rays[ix, :3], rays[ix, 3:], wls[ix], s, sc['ambience'])
            assert np.array_equal(rad.view(np.uint32), want['radiance'].reshape(-1, 16)[ix].view(np.uint32)), (ix, rad, want['radiance'].reshape(-1, 16)[ix])
            assert np.array_equal(dist.view(np.uint32), want['distance'].reshape(-1, 16)[ix].view(np.uint32)), ix
            vertices += int(np.isfinite(dist).sum())
    assert vertices > h * w // 2                                          # the sample really exercises shading, not just misses


@pytest.mark.parametrize('pitch,yaw', [(0.0, 0.0), (0.1, 0.3), (-0.4, 2.5)])
def test_camera_against_an_independent_restatement(orc, scenes, pitch, yaw):
    """sample_camera_wavelength and sample_camera_ray (reference src/camera.fut:46-55,68-110, visual_conf of src/lib.fut:18-27)
    restated from the .fut text: channel pick, probit wavelength, pixel jitter and lens sample from PEEKED draws, pinhole ray.
    Per-frame scalars (sin / cos of yaw and pitch, tan of the half field of view) come from the C library like in the c backend."""
    libm = ctypes.CDLL('libm.so.6')
    for fn in ('sinf', 'cosf', 'tanf'):
        getattr(libm, fn).restype = ctypes.c_float
        getattr(libm, fn).argtypes = [ctypes.c_float]
    t9, tm, m = scenes['cornell']
    h, w = 24, 40
    origin = np.array([0.1, 0.7, 1.9], F)
    st = orc.State.init(t9, tm, m, h, w, pitch=pitch, yaw=yaw, origin=tuple(float(x) for x in origin))
    fut = Fut(orc, t9, tm, m, st.bvh())
    prim = st.probe_primary(want_rays=True)
    rng0 = st.scalars()['rng']
    L = orc.lib()
    sensor = [(F(455), F(22)), (F(535), F(32)), (F(610), F(26))]
    fov = F(80) * PI / F(180)                                             # from_deg (linalg.fut:52)
    cam_dir = normalise(np.array([libm.sinf(F(yaw)), libm.sinf(F(pitch)), -libm.cosf(F(yaw))], F))
    right = normalise(cross(cam_dir, np.array([0, 1, 0], F)))
    up = normalise(cross(right, cam_dir))
    ratio = F(w) / F(h)
    half_h = F(libm.tanf(fov / F(2.0)))
    half_w = ratio * half_h
    wv, focus = F(-1) * cam_dir, F(1)
    llc = ((origin - (half_w * focus) * right) - (half_h * focus) * up) - focus * wv
    horizontal, vertical = (F(2) * half_w * focus) * right, (F(2) * half_h * focus) * up
    for i in range(h):
        for j in range(w):
            ix = i * w + j
            s = rng0 ^ int(L.orc_hash(ix))
            s = fut.lcg(s)
            mu, sigma = sensor[s % 3]
            s, p = fut.uniform(s)
            wl = mu + sigma * fut.m1('probit', p)
            s1, ox = fut.uniform(s)                                        # peeked: the caller keeps s
            _, oy = fut.uniform(s1)
            x = (F(j) + F(1) * ox) / F(w)
            y = ((F(h) - F(i) - F(1.0)) + F(1) * oy) / F(h)
            s2, theta = fut.uniform(s, F(0), F(2) * PI)                    # the same draws again for the lens sample
            _, u = fut.uniform(s2)
            r = np.sqrt(u)
            lens = F(0) * (r * np.array([fut.m1('cos', theta), fut.m1('sin', theta), F(0)], F))
            o = origin + (lens[0] * right + lens[1] * up)
            d = normalise(((llc + x * horizontal) + y * vertical) - o)
            got = np.concatenate([o, d]).astype(F)
            assert np.array_equal(got.view(np.uint32), prim['rays'][i, j].view(np.uint32)), (i, j, got, prim['rays'][i, j])
            assert np.float32(wl).view(np.uint32) == prim['wavelen'][i, j].view(np.uint32), (i, j)


def test_accumulation_and_render_against_an_independent_restatement(orc, scenes):
    """visualize_pixels (#render_color, reference src/integrator.fut:133-170: sequential sum of intensity * channel vector over
    the 16 path entries, times the channel count), sample_n_frames / sample_frame_accum (src/lib.fut:67-74, integrator.fut:180-192:
    running average weighted with the OLD frame count, so the first frame is discarded) and render (src/lib.fut:187-196 with
    matte's argb.from_rgba: clamp, * 255, truncate), from the per-vertex radiance of three consecutive passes."""
    t9, tm, m = scenes['spectrumsphere']
    h, w = 20, 28
    st = orc.State.init(t9, tm, m, h, w)
    vis = [np.array([0, 0, 1], F), np.array([0, 1, 0], F), np.array([1, 0, 0], F)]     # visual_conf sensor (lib.fut:24-26)
    frames = []
    for k in range(3):
        pp = st.advance_rng(k).probe_pass()
        img = np.zeros((h, w, 3), F)
        for i in range(h):
            for j in range(w):
                acc = np.zeros(3, F)
                for v in range(16):
                    acc = acc + pp['radiance'][i, j, v] * vis[pp['channel'][i, j]]
                img[i, j] = F(3) * acc
        frames.append(img)
    img, nf = frames[0], 1
    while nf < 3:
        n = F(nf)
        img = (((n - F(1)) / n) * img + (F(1) / n) * frames[nf]).astype(F)
        nf += 1
    got = st.sample_n_frames(3)
    assert np.array_equal(img.view(np.uint32), got.view(np.uint32))
    # the interactive path: accumulate on (key m), three steps, then render
    s = st.key(0x6D)
    for _ in range(3):
        s = s.step()
    px = s.render()
    im = s.image()
    c = np.clip(im, 0, 1)
    c = np.where(np.isnan(im), 0, c)
    ch = (c * F(255)).astype(np.uint32)
    argb = ((np.uint32(255) << 24) | (ch[..., 0] << 16) | (ch[..., 1] << 8) | ch[..., 2]).astype(np.uint32)
    assert np.array_equal(argb.view(np.int32), px[:im.shape[0], :im.shape[1]])


def _libm():
    libm = ctypes.CDLL('libm.so.6')
    for fn in ('sinf', 'cosf', 'tanf', 'expf'):
        getattr(libm, fn).restype = ctypes.c_float
        getattr(libm, fn).argtypes = [ctypes.c_float]
    libm.powf.restype = ctypes.c_float
    libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
    return libm


@pytest.mark.parametrize('conf', [2, 1])
def test_transmitter_lights_against_an_independent_restatement(orc, scenes, conf):
    """The LIDAR (scanning frustum lights around every pixel's own ray, src/camera.fut:112-122, light.fut:32-44, lidar_conf of
    lib.fut:10-16) and flash (a disk of diffuse lights at the camera, visual_flash_conf lib.fut:29-33 with the normalised
    black-body spectrum of spectrum.fut:60-79) configurations: per-vertex radiance and distance of one pass."""
    libm = _libm()
    t9, tm, m = scenes['cornell']
    h, w = 14, 18
    st = orc.State.init(t9, tm, m, h, w, cam_conf_id=conf)
    fut = Fut(orc, t9, tm, m, st.bvh())
    sc = st.scalars()
    prim = st.probe_primary(want_rays=True)
    want = st.probe_pass()
    rays, wls = prim['rays'].reshape(-1, 6), prim['wavelen'].reshape(-1)
    origin = np.array([0.0, 0.8, 1.8], F)
    if conf == 2:
        em = np.array([0, 1500] + [-1, 0] * 5, F)                         # uniform_spectrum 1500
        theta = F(3) * PI / F(180.0)
    else:
        cc, hh, kb, T = F(299792458), F(6.62606957e-34), F(1.3806488e-23), F(5500)
        nm = [F(150), F(460), F(550), F(610), F(1000), F(2000)]           # blue / green / red_wavelen (spectrum.fut:8-10)
        ls = [x * F(1e-9) for x in nm]
        planck = [(F(2) * hh * cc * cc) / (F(libm.powf(l, F(5))) * (F(libm.expf((hh * cc) / (l * kb * T))) - F(1))) for l in ls]
        knots = np.array([v for l, pl in zip(ls, planck) for v in (l * F(1e9), pl)], F)
        lam_max = (F(2.8977721e-3) / T) * F(1e9)
        mx = fut.lookup(lam_max, knots)
        em = knots.copy()
        em[1::2] = (knots[1::2] / mx) * F(1000)
        cam_dir = normalise(np.array([libm.sinf(F(0)), libm.sinf(F(0)), -libm.cosf(F(0))], F))
        flash = [dict(kind='diffuse', tri=t, em=em, theta=F(0)) for t in fut.disk(origin, cam_dir, F(0.05), libm)]
    L = orc.lib()
    lit = 0
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        for ix in range(h * w):
            s = sc['rng'] ^ int(L.orc_hash(ix))
            s = fut.lcg(s)
            s = fut.lcg(s)
            if conf == 2:
                fut.extra = [dict(kind='frustum', tri=t, em=em, theta=theta) for t in fut.disk(origin, rays[ix, 3:], F(0.01), libm)]
            else:
                fut.extra = flash
            rad, dist = fut.path_trace(rays[ix, :3], rays[ix, 3:], wls[ix], s, sc['ambience'])
            assert np.array_equal(rad.view(np.uint32), want['radiance'].reshape(-1, 16)[ix].view(np.uint32)), (ix, rad, want['radiance'].reshape(-1, 16)[ix])
            assert np.array_equal(dist.view(np.uint32), want['distance'].reshape(-1, 16)[ix].view(np.uint32)), ix
            lit += int((rad > 0).any())
    assert lit > h * w // 8


@pytest.mark.parametrize('seed', [1, 2])
def test_path_tracer_on_random_soup_against_an_independent_restatement(orc, seed):
    """The same comparison on a fuzzed scene (lysref.objwriter.random_soup): rough and smooth metals, dispersive dielectrics,
    partial opacity, unused spectrum knots, ten light triangles (one of zero area), duplicated and degenerate triangles --
    the parts of material.fut / direct.fut the four bundled assets hardly reach."""
    from lysref import objwriter
    t9, tm, m = objwriter.random_soup(seed)
    h, w, origin = 16, 20, (0.0, 1.0, 0.9)
    st = orc.State.init(t9, tm, m, h, w, origin=origin)
    fut = Fut(orc, t9, tm, m, st.bvh())
    sc = st.scalars()
    prim = st.probe_primary(want_rays=True)
    want = st.probe_pass()
    rays, wls = prim['rays'].reshape(-1, 6), prim['wavelen'].reshape(-1)
    assert fut.light_ix == st.light_indices().tolist() and len(fut.light_ix) == 10
    L = orc.lib()
    vertices = lit = 0
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        for ix in range(h * w):
            s = sc['rng'] ^ int(L.orc_hash(ix))
            s = fut.lcg(s)
            s = fut.lcg(s)
            rad, dist = fut.path_trace(rays[ix, :3], rays[ix, 3:], wls[ix], s, sc['ambience'])
            wr, wd = want['radiance'].reshape(-1, 16)[ix], want['distance'].reshape(-1, 16)[ix]
            assert np.array_equal(rad.view(np.uint32), wr.view(np.uint32)), (ix, rad, wr)
            assert np.array_equal(dist.view(np.uint32), wd.view(np.uint32)), ix
            vertices += int(np.isfinite(dist).sum())
            lit += int((rad > 0).any())
    assert vertices > h * w and lit > h * w // 10
