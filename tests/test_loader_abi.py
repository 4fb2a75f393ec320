"""Host side without a GPU: the OBJ/MTL loader, the scene generator, and the C-ABI surface of libtracer."""
import ctypes
import os
import re
import subprocess
import numpy as np
import pytest
from conftest import ROOT, SCENE_NAMES, bits_equal

REF = '/root/reference'
ASSET = {'cornell': 'CornellBox-Original.obj', 'mirrorbox': 'MirrorBox.obj', 'spectrumsphere': 'SpectrumSphere.obj',
         'spectrumspherehigh': 'SpectrumSphereHigh.obj'}


def test_golden_scene_shapes(scenes):
    want = {'cornell': (44, 8), 'mirrorbox': (38, 9), 'spectrumsphere': (2188, 7), 'spectrumspherehigh': (8716, 7)}
    for n in SCENE_NAMES:
        t, tm, m = scenes[n]
        assert (len(t), len(m)) == want[n] and t.shape[1:] == (3, 3) and m.shape[1] == 28 and tm.max() < len(m)
    # Cornell light row: Kd 0.78, Pr 1, Pm 0, Ni 1, Tf default 1, Ke 27 22 14 (CornellBox-Original.mtl)
    row = scenes['cornell'][2][7]
    want_row = [610, 0.78, 550, 0.78, 460, 0.78, -1, 0, -1, 0, -1, 0, 1, 0, 1, 1, 610, 27, 550, 22, 460, 14, -1, 0, -1, 0, -1, 0]
    assert np.array_equal(row, np.array(want_row, np.float32))
    # SpectrumSphere glass: Sp 0 0, Pr 0, Pm 0, Tf 0, Ni 1.5; light: Em 800 30 801 0
    g = scenes['spectrumsphere'][2][1]
    assert list(g[:4]) == [0, 0, -1, 0] and list(g[12:16]) == [0, 0, 1.5, 0]
    assert list(scenes['spectrumsphere'][2][6][16:22]) == [800, 30, 801, 0, -1, 0]


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference assets only exist in the build container')
@pytest.mark.parametrize('name', SCENE_NAMES)
def test_cpp_loader_matches_python_loader_and_golden(pkg, scenes, name):
    from lysref import loader
    path = os.path.join(REF, 'assets', ASSET[name])
    a = pkg.load_obj(path)
    b = loader.load_obj(path)
    for x, y, z in zip(a, b, scenes[name]):
        assert bits_equal(x, y) and bits_equal(x, z)


def test_loader_on_handwritten_obj(pkg, tmp_path):
    (tmp_path / 'm.mtl').write_text('# c\nnewmtl red\n Kd 1 0 0\n Pr 0.25\nnewmtl lamp\n Em 500 2 600 3\n Ni 1.33\n Tf 0.5\n')
    (tmp_path / 's.obj').write_text('mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0 0 1\nusemtl red\nf 1 2 3 4\n'
                                    'usemtl lamp\nf -1 -4/1/1 -3/2\nf 1 2 3 4 5\n')
    t, tm, m = pkg.load_obj(str(tmp_path / 's.obj'))
    assert list(tm) == [0, 0, 1, 1, 1, 1]                                   # quad -> 2, tri -> 1, pentagon fan -> 3
    assert np.array_equal(t[0], [[0, 0, 0], [1, 0, 0], [1, 1, 0]]) and np.array_equal(t[1], [[0, 0, 0], [1, 1, 0], [0, 1, 0]])
    assert np.array_equal(t[2], [[0, 0, 1], [1, 0, 0], [1, 1, 0]])          # negative (relative) indices
    assert np.array_equal(t[5], [[0, 0, 0], [0, 1, 0], [0, 0, 1]])
    assert list(m[0][:6]) == [610, 1, 550, 0, 460, 0] and m[0][12] == 0.25 and m[0][14] == 1.0 and m[0][15] == 1.0
    assert list(m[1][16:22]) == [500, 2, 600, 3, -1, 0] and m[1][14] == np.float32(1.33) and m[1][15] == 0.5


def test_synthetic_scene(pkg, scenes):
    t, tm, _ = scenes['cornell']
    st, sm = pkg.scenes.synthetic_cornell(t, tm, 151)
    assert st.shape == (1003244, 3, 3) and sm.shape == (1003244,)          # BASELINE config 5
    s1, m1 = pkg.scenes.synthetic_cornell(t, tm, 1)
    assert bits_equal(s1, t) and np.array_equal(m1, tm)                    # k = 1 reproduces the Cornell box itself
    s3, _ = pkg.scenes.synthetic_cornell(t, tm, 3)
    assert np.allclose(s3.reshape(-1, 3).min(0), t.reshape(-1, 3).min(0)) and np.allclose(s3.reshape(-1, 3).max(0), t.reshape(-1, 3).max(0))


def declared_symbols():
    names = set()
    for h in ('tracer.h', 'lys_ext.h'):
        src = open(os.path.join(ROOT, 'include', h)).read()
        src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
        names |= set(re.findall(r'\b((?:futhark|lys)_[a-z0-9_]+)\s*\(', src))
    return sorted(names)


def test_library_exports_every_declared_symbol(pkg):
    """The .so loads without a GPU and exports everything include/*.h declares (no compute is called)."""
    lib = ctypes.CDLL(pkg.lib_path())
    syms = declared_symbols()
    assert len(syms) >= 55
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert os.path.exists(os.path.join(os.path.dirname(pkg.lib_path()), 'libtracer.a'))
    out = subprocess.check_output(['nm', '-g', '--defined-only', os.path.join(os.path.dirname(pkg.lib_path()), 'libtracer.a')]).decode()
    for s in ('futhark_entry_init', 'futhark_entry_step', 'futhark_entry_render', 'futhark_entry_sample_points_n', 'futhark_values_i32_2d'):
        assert re.search(r'\bT %s\b' % s, out), s


def test_no_gpu_means_loud_failure_not_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    with pytest.raises(pkg.TracerError):
        pkg.Context()


def test_product_does_not_reference_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'msc-futhark-ray-tracer_b200')):
        for f in files:
            if f.endswith(('.cu', '.cuh', '.h', '.cpp', '.py')) or f == 'Makefile':
                txt = open(os.path.join(dirpath, f)).read()
                assert 'lys_oracle' not in txt and 'lysref' not in txt and 'oracle/' not in txt, os.path.join(dirpath, f)
                assert 'simt_emu' not in txt and 'libtracer_emu' not in txt and 'LYS_SIMT_EMU' not in txt, os.path.join(dirpath, f)   # nor the CPU emulator of tests/


@pytest.mark.skipif(not os.path.isdir(REF), reason='needs the reference host sources')
def test_unchanged_reference_host_compiles_against_tracer_h():
    """demo-interactive/liblys.c is the ABI contract: it must compile, unmodified, against include/tracer.h."""
    cmd = ['gcc', '-fsyntax-only', '-std=c11', '-DLYS_BACKEND_cuda', '-I' + os.path.join(ROOT, 'include'),
           '-I' + os.path.join(REF, 'deps/SDL2/include'), '-I' + os.path.join(REF, 'demo-interactive'),
           os.path.join(REF, 'demo-interactive/liblys.c')]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
