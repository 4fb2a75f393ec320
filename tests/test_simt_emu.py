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


@pytest.fixture(scope='module')
def emu_lib():
    import build as emu_build
    return emu_build.build()


def sweep(emu_lib, scenes, env=None):
    e = dict(os.environ)
    e.update(env or {})
    e['LYS_LIBTRACER'] = emu_lib
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, 'tools', 'gpu_parity_quick.py')] + scenes, env=e, text=True, timeout=1500)
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
    {'LYS_TRACE_OCT': '0', 'LYS_TRACE_NB': '1'},       # what scenes above 128K nodes run: select-based box test, one box stage
    {'LYS_TAIL_MAX': '100000000'},                     # fused tail kernel from bounce 1 on
    {'LYS_TRACE_MODE': '1'},                           # refill variant of the trace kernel
    {'LYS_SHADE_SPLIT': '2', 'LYS_FUSE_GENERATE': '0'},
], ids=lambda e: ','.join(f'{k}={v}' for k, v in e.items()))
def test_emulated_kernel_variants(emu_lib, env):
    e = dict(env)
    e['LYS_EMU_FAST_MATH_SWEEP'] = '1'
    check(sweep(emu_lib, ['cornell'], e))
