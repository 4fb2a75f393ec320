"""The headless C host (host/lys_headless.c): the reference's liblys.c call sequence, linked against the STATIC
libtracer.a + libljus exactly as the reference links main-interactive (Makefile:48-49)."""
import os
import subprocess
import numpy as np
import pytest
from conftest import ROOT, bits_equal

HOST = os.path.join(ROOT, 'host', 'lys_headless')


def build_host(pkg):
    pkg.build()
    subprocess.check_call(['make', '-C', os.path.join(ROOT, 'host')], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert os.path.exists(HOST)


def test_c_host_links_statically_and_fails_loudly_without_gpu(pkg):
    build_host(pkg)
    undefined = subprocess.check_output(['nm', '-u', HOST]).decode()
    assert 'futhark_' not in undefined                      # every futhark_* symbol was resolved from libtracer.a
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([HOST, '-o', 'missing.obj'], capture_output=True, text=True)
        assert r.returncode != 0 and 'no CUDA device' in r.stderr


def test_obj_writer_roundtrip(pkg, scenes, tmp_path):
    from lysref import objwriter
    for name in ('cornell', 'spectrumsphere'):
        t, tm, m = scenes[name]
        p = str(tmp_path / (name + '.obj'))
        objwriter.write_obj(p, t, tm, m)
        t2, tm2, m2 = pkg.load_obj(p)
        assert bits_equal(t, t2) and bits_equal(tm, tm2) and bits_equal(m, m2)


@pytest.mark.gpu
def test_c_host_matches_python_host(pkg, gpu, scenes, tmp_path):
    from lysref import objwriter
    build_host(pkg)
    t, tm, m = scenes['spectrumsphere']
    obj, ppm = str(tmp_path / 's.obj'), str(tmp_path / 'o.ppm')
    objwriter.write_obj(obj, t, tm, m)
    w, h, frames = 160, 120, 4
    out = subprocess.check_output([HOST, '-o', obj, '-w', str(w), '-h', str(h), '-n', str(frames), '-k', '109', '-p', ppm]).decode()
    assert 'frames 4' in out
    raw = open(ppm, 'rb').read()
    hdr = ('P6\n%d %d\n255\n' % (w, h)).encode()
    img = np.frombuffer(raw[len(hdr):], np.uint8).reshape(h, w, 3)
    s = pkg.State.init(gpu, t, tm, m, h, w).resize(h, w).key(109)
    for _ in range(frames):
        s = s.step()
    px = s.render().view(np.uint32)
    want = np.stack([(px >> 16) & 255, (px >> 8) & 255, px & 255], axis=2).astype(np.uint8)
    assert np.array_equal(img, want) and img.max() > 0
