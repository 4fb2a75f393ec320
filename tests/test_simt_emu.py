"""The product's CUDA sources (csrc/*.cu) compiled for the host and executed by the fiber SIMT emulator of tests/simt_emu
must give the oracle's bits: LBVH arrays, first hits, per-vertex radiance, accumulated images, ARGB frames (the sweep of
tools/gpu_parity_quick.py, run in a subprocess against tests/simt_emu/_build/libtracer_emu.so).  CPU only."""
import json
import os
import subprocess
import sys

import pytest
from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, 'tests', 'simt_emu'))
RUN_EMU = os.path.join(ROOT, 'tests', 'simt_emu', 'run_with_emu.py')      # the only place that points the binding at the emulator build


@pytest.fixture(scope='module')
def emu_lib():
    import emu_build
    return emu_build.build()


def sweep(emu_lib, scenes, env=None):
    e = dict(os.environ)
    e.update(env or {})
    e['LYS_EMU_LIB'] = emu_lib
    out = subprocess.check_output([sys.executable, RUN_EMU, os.path.join(ROOT, 'tools', 'gpu_parity_quick.py')] + scenes, env=e, text=True, timeout=1500)
    res = {}
    for line in out.splitlines():
        name, _, js = line.partition(' ')
        if name in scenes:
            res[name] = json.loads(js)
    assert sorted(res) == sorted(scenes), out
    return res


def check(res):
    for name, r in res.items():
        for key in ('bounds', 'morton', 'src_index', 'left', 'right', 'parent', 'node_aabb', 'leaf_aabb', 'lights', 'first_hit_leaf', 'first_hit_t',
                    'pass_radiance_bits', 'pass_distance_bits', 'pass_channel', 'img4_bits', 'step3_img_bits', 'render_bits'):
            assert r[key] is True, (name, key, r)


def test_emulated_library_equals_the_oracle(emu_lib):
    check(sweep(emu_lib, ['cornell', 'spectrumsphere'], {'LYS_EMU_FAST_MATH_SWEEP': '1'}))


@pytest.mark.parametrize('env', [
    {'LYS_OCT_ONE_COPY': '1'},                           # what scenes above 64K nodes run: one copy of the records, select-based box test
    {'LYS_REFILL_MIN': '1'},                             # lane refill of the closest-hit walk with octant copies (scenes from 1024 triangles run it)
    {'LYS_REFILL_MIN': '1', 'LYS_TAIL_MAX': '0', 'LYS_EMU_SCHEDULE': '55'},
    {'LYS_TAIL_MAX': '100000000'},                     # fused tail kernel from bounce 1 on
    {'LYS_TAIL_MAX': '0', 'LYS_SHADE_ORDER': '0', 'LYS_FUSE_GENERATE': '0', 'LYS_EMU_SMS': '32'},
    {'LYS_EMU_SCHEDULE': '1'},                         # CTAs, warps and lanes run in reverse order: results must not depend on the schedule
    {'LYS_EMU_SCHEDULE': '4242', 'LYS_TAIL_MAX': '100000000'},       # pseudo-random orders, redrawn per CTA (race / order-dependence probe)
    {'LYS_EMU_SCHEDULE': '977', 'LYS_OCT_ONE_COPY': '1', 'LYS_EMU_SMS': '8'},
])
def test_emulated_kernel_variants(emu_lib, env):
    e = dict(env)
    e['LYS_EMU_FAST_MATH_SWEEP'] = '1'
    check(sweep(emu_lib, ['cornell'], e))


def test_gpu_parity_suite_on_the_emulator(emu_lib):
    """tests/test_gpu_parity.py itself (the `-m gpu` parity tests: LBVH edge cases, fuzzed soup scenes, all camera presets,
    entry points, LIDAR points, row partition, error behaviour, raw device access ...) run in a subprocess against the
    emulated library, 1 003 244-triangle LBVH included.  Left out: the 1080p / 4K cases (a minute and more per frame on the
    emulator) and the variant sweep (covered above)."""
    e = dict(os.environ)
    e['LYS_EMU_LIB'] = emu_lib
    e['LYS_EMU_SCHEDULE'] = '20261018'                      # pseudo-random CTA / warp / lane order: also a race probe
    r = subprocess.run([sys.executable, RUN_EMU, '-m', 'pytest', os.path.join(ROOT, 'tests', 'test_gpu_parity.py'), '-m', 'gpu', '-q', '-x', '-p', 'no:cacheprovider',
                        '-k', 'not full_size and not kernel_variants and not baseline_resolution'], env=e, text=True, capture_output=True, timeout=1500, cwd=ROOT)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert ' passed' in tail and 'failed' not in tail and int(tail.split(' passed')[0].split()[-1]) >= 40, tail


def test_product_binding_has_no_library_override(emu_lib):
    """No CPU path reachable from the product: the Python binding ignores the environment and only opens the libtracer.so
    next to it (the emulator build is reached through tests/simt_emu/run_with_emu.py alone)."""
    e = dict(os.environ)
    e.update({'LYS_LIBTRACER': emu_lib, 'LYS_ALLOW_EMULATOR': '1', 'LYS_EMU_LIB': emu_lib})
    code = ("import importlib, os; p = importlib.import_module('msc-futhark-ray-tracer_b200'); "
            "assert os.path.basename(p.lib_path()) == 'libtracer.so' and os.path.dirname(p.lib_path()) == os.path.dirname(p.__file__), p.lib_path()")
    r = subprocess.run([sys.executable, '-c', code], env=e, text=True, capture_output=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    src = open(os.path.join(ROOT, 'msc-futhark-ray-tracer_b200', 'tracer.py')).read()
    assert 'environ' not in src and 'emu' not in src.lower()


@pytest.mark.parametrize('what,seed,count,env', [
    ('lbvh', 1, 40, {'LYS_EMU_SCHEDULE': '3'}),        # hostile geometry: inf / NaN / denormals / duplicates / identical triangles
    ('soup', 2, 10, {}),                               # random scenes, materials, camera presets, poses, frame sizes, seeds
    ('soup', 3, 8, {'LYS_OCT_ONE_COPY': '1', 'LYS_EMU_SCHEDULE': '7'}),
    ('soup', 5, 8, {'LYS_REFILL_MIN': '1', 'LYS_EMU_SCHEDULE': '11'}),      # lane refill of the closest-hit walks on small random scenes
    ('keys', 4, 12, {}),                               # random host sessions: key events, resizes, steps -> scalars, image, ARGB frame
], ids=lambda v: str(v) if not isinstance(v, dict) else ','.join(f'{k}={x}' for k, x in v.items()) or 'default')
def test_fuzzed_parity_on_the_emulator(emu_lib, what, seed, count, env):
    """tools/fuzz_parity.py (usable on the GPU as well) against the emulated library: zero mismatching scenes."""
    e = dict(os.environ)
    e.update(env)
    e.update({'LYS_EMU_LIB': emu_lib})
    r = subprocess.run([sys.executable, RUN_EMU, os.path.join(ROOT, 'tools', 'fuzz_parity.py'), what, str(seed), str(count)], env=e, text=True, capture_output=True, timeout=1500)
    assert r.returncode == 0 and '%d scenes, 0 mismatching' % count in r.stdout, r.stdout[-2000:] + r.stderr[-1000:]


def test_c_hosts_on_the_emulator(emu_lib, orc, scenes, tmp_path):
    """The C hosts (host/lys_save.c = the reference's demo-save, host/lys_headless.c = liblys.c's loop) linked against the
    emulated library: their whole flow -- OBJ/MTL loader, futhark_new_*, init, sample_points_n / step + render, file writers --
    on the CPU, against the oracle.  (The GPU versions of these checks are in tests/test_c_host.py.)"""
    import numpy as np
    import emu_build
    from conftest import bits_equal
    from lysref import objwriter
    exe = emu_build.build_hosts()
    t, tm, m = scenes['spectrumsphere']
    obj, pcd, ppm = str(tmp_path / 's.obj'), str(tmp_path / 'dump.pcd'), str(tmp_path / 'f.ppm')
    objwriter.write_obj(obj, t, tm, m)
    w, h, spp = 40, 30, 3
    out = subprocess.check_output([exe['lys_save'], '-o', obj, '-w', str(w), '-h', str(h), '-s', str(spp), '-p', pcd], text=True, timeout=600)
    assert 'points %d' % (w * h) in out
    lines = open(pcd).read().split('\n')
    assert lines[6] == 'WIDTH %d' % (w * h) and lines[10] == 'DATA ascii'
    got = np.array([[np.float32(x) for x in line.split(' ')] for line in lines[11:] if line], np.float32).reshape(h, w, 3)
    want = orc.State.init(t, tm, m, h, w, cam_conf_id=2).sample_points_n(spp)[1]
    assert bits_equal(got, np.ascontiguousarray(want[..., :3]))
    frames = 3
    out = subprocess.check_output([exe['lys_headless'], '-o', obj, '-w', str(w), '-h', str(h), '-n', str(frames), '-k', '109', '-p', ppm], text=True, timeout=600)
    assert 'frames %d' % frames in out
    raw = open(ppm, 'rb').read()
    hdr = ('P6\n%d %d\n255\n' % (w, h)).encode()
    img = np.frombuffer(raw[len(hdr):], np.uint8).reshape(h, w, 3)
    so = orc.State.init(t, tm, m, h, w).resize(h, w).key(109)
    for _ in range(frames):
        so = so.step()
    px = so.render().view(np.uint32)
    assert np.array_equal(img, np.stack([(px >> 16) & 255, (px >> 8) & 255, px & 255], axis=2).astype(np.uint8)) and img.max() > 0


@pytest.mark.skipif(not os.path.isdir('/root/reference/demo-interactive'), reason='the reference checkout only exists in the build container')
def test_unmodified_reference_host_runs_against_the_library(emu_lib, orc, scenes, tmp_path):
    """demo-interactive/liblys.c, compiled UNMODIFIED from the reference checkout, linked against the (emulated) library and a
    headless SDL stand-in (tests/sdl_stub), runs a scripted session -- its own init with an 800x600 window, a resize event, a
    key event, five frames -- and the last window contents equal the oracle's ARGB frame for the same session."""
    import numpy as np
    import emu_build
    from lysref import objwriter
    exe = emu_build.build_reference_host()
    assert 'emu' in exe
    t, tm, m = scenes['cornell']
    obj, ppm = str(tmp_path / 'c.obj'), str(tmp_path / 'w.ppm')
    objwriter.write_obj(obj, t, tm, m)
    e = dict(os.environ)
    e.update({'LYS_SDL_SCRIPT': '0:resize:64x48 1:key:32 5:quit', 'LYS_SDL_DUMP': ppm})
    out = subprocess.check_output([exe['emu'], '-o', obj], env=e, text=True, timeout=900)
    assert 'sdl_stub: 5 frames, 3 events, window 64x48' in out
    s = orc.State.init(t, tm, m, 600, 800).resize(600, 800).step().resize(48, 64).key(32)      # liblys.c:133-152, then the script
    for _ in range(4):
        s = s.step()
    px = s.render().view(np.uint32)
    want = np.stack([(px >> 16) & 255, (px >> 8) & 255, px & 255], axis=2).astype(np.uint8)
    raw = open(ppm, 'rb').read()
    hdr = b'P6\n64 48\n255\n'
    assert raw.startswith(hdr)
    assert np.array_equal(np.frombuffer(raw[len(hdr):], np.uint8).reshape(48, 64, 3), want) and want.max() > 0
    # the reference's own error path: accumulating (key m) onto an image of another size is a Futhark size error (integrator.fut:184)
    e['LYS_SDL_SCRIPT'] = '0:resize:32x24 0:key:109 3:quit'
    r = subprocess.run([exe['emu'], '-o', obj], env=e, text=True, capture_output=True, timeout=900)
    assert r.returncode != 0 and 'Futhark error' in r.stderr and 'shape does not match' in r.stderr


def test_partitioned_rendering_world2_gloo_runs_the_library_kernels(emu_lib, tmp_path):
    """SURVEY 8(e) with world_size 2 over gloo, THROUGH the library: each rank runs the product's kernels (this emulator build)
    on its interleaved rows (lys_context_set_partition) resp. its pass range (lys_state_advance_rng + the weighted last
    accumulate), one sum-reduce of the framebuffer, and rank 0 compares with the oracle: rows bit for bit, passes within 1e-4."""
    import json
    import socket
    sk = socket.socket(); sk.bind(('127.0.0.1', 0)); port = sk.getsockname()[1]; sk.close()
    out = str(tmp_path / 'res.json')
    procs = []
    for rank in range(2):
        e = dict(os.environ)
        e.update({'LYS_EMU_LIB': emu_lib, 'RANK': str(rank), 'WORLD_SIZE': '2', 'MASTER_ADDR': '127.0.0.1', 'MASTER_PORT': str(port), 'OUT': out,
                  'OMP_NUM_THREADS': '2'})
        procs.append(subprocess.Popen([sys.executable, RUN_EMU, os.path.join(ROOT, 'tests', 'simt_emu', 'partition_worker.py')], env=e,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = [p.communicate(timeout=900)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), '\n'.join(l[-1500:] for l in logs)
    res = json.load(open(out))
    assert res == {'rows_zero_elsewhere': True, 'rows_bit_exact': True, 'passes_close': True}, res
