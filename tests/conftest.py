import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
SCENE_NAMES = ['cornell', 'mirrorbox', 'spectrumsphere', 'spectrumspherehigh']


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    if a.dtype == np.float32:
        return bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))
    return bool(np.array_equal(a, b))


def load_scene(name):
    d = np.load(os.path.join(GOLDEN, 'scenes', name + '.npz'))
    return d['tris'], d['tri_mats'], d['mats']


@pytest.fixture(scope='session')
def pkg():
    m = importlib.import_module('msc-futhark-ray-tracer_b200')
    m.build()
    return m


@pytest.fixture(scope='session')
def orc():
    from lysref import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope='session')
def scenes():
    return {n: load_scene(n) for n in SCENE_NAMES}


@pytest.fixture(scope='session')
def golden_vectors():
    return np.load(os.path.join(GOLDEN, 'oracle_vectors.npz'))


@pytest.fixture(scope='session')
def gpu(pkg):
    """A libtracer context on cuda:0.  Fails loudly (no skip, no fallback) when the GPU path is unavailable."""
    ctx = pkg.Context()
    yield ctx
    ctx.close()
