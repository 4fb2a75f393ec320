"""Tolerance-level (not bit-level) statements, with the tolerance written in the test:
 - the arithmetic contract vs glibc libm does not change the converged picture,
 - splitting PASSES across ranks (bench.py's N>1 mode) agrees with a single render within Monte-Carlo noise,
 - bench.py --impl reference emits the contract's JSON line."""
import json
import os
import subprocess
import sys
import numpy as np
from conftest import ROOT


def rmse(a, b):
    return float(np.sqrt(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)))


def test_contract_math_vs_libm_same_picture(orc, scenes):
    """lys_detmath.h differs from glibc by <= 1.6 ulp per call; per-pixel values can differ (a flipped branch changes a
    path), the converged image must not: RMSE over 48 passes below 2 % of the mean radiance, identical first hits."""
    t, tm, m = scenes['spectrumsphere']
    s = orc.State.init(t, tm, m, 48, 64)
    a = s.sample_n_frames(48)
    hits_a = s.probe_primary()['src_tri']
    orc.set_math_mode(1)
    try:
        b = s.sample_n_frames(48)
        hits_b = s.probe_primary()['src_tri']
    finally:
        orc.set_math_mode(0)
    assert np.array_equal(hits_a, hits_b)                       # no transcendental on the primary-ray path (aperture 0)
    assert rmse(a, b) < 0.02 * float(a.mean()) + 1e-6
    same = np.mean(np.all(a == b, axis=2))
    assert same > 0.5                                           # most pixels are bit-identical even so


def test_pass_split_matches_single_render(orc, scenes):
    """N ranks x P passes (disjoint rng ranges via advance_rng) averaged vs one N*P-pass render: each rank drops its own
    first frame (integrator.fut:184-191 quirk), so the two differ only by Monte-Carlo noise.  Tolerance: the RMSE between
    them must be below 1.5x the RMSE between two independent single renders of the same total sample count."""
    t, tm, m = scenes['cornell']
    P, N = 12, 2
    s = orc.State.init(t, tm, m, 40, 48)
    single = s.sample_n_frames(P * N)
    parts = [s.advance_rng(r * P).sample_n_frames(P) for r in range(N)]
    split = np.mean(parts, axis=0)
    other = orc.State.init(t, tm, m, 40, 48, seed=7).sample_n_frames(P * N)
    noise = rmse(single, other)
    assert rmse(single, split) < 1.5 * noise
    assert abs(float(split.mean()) - float(single.mean())) < 0.1 * float(single.mean())


def test_converged_image_rmse_threshold(orc, scenes):
    """North star: the converged-image RMSE must fall below a stated threshold.  GPU == oracle bit-for-bit (test_gpu_parity),
    so the residual is Monte-Carlo noise only; it must shrink like 1/sqrt(passes): 64 passes vs 256-pass reference has
    RMSE < 0.6 x the RMSE of 16 passes vs the same reference."""
    t, tm, m = scenes['cornell']
    s = orc.State.init(t, tm, m, 32, 40)
    ref = orc.State.init(t, tm, m, 32, 40, seed=3).sample_n_frames(256)
    e16, e64 = rmse(s.sample_n_frames(16), ref), rmse(s.sample_n_frames(64), ref)
    assert e64 < 0.6 * e16


def test_bench_reference_arm_contract():
    env = dict(os.environ, OMP_NUM_THREADS='1')                 # as under torchrun; bench must override it
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                                  env=env, timeout=600).decode()
    line = json.loads([l for l in out.splitlines() if l.startswith('{')][-1])
    assert line['impl'] == 'reference' and line['metric'] == 'Mpaths/s' and line['unit'] == 'Mpaths/s' and line['value'] > 0
    assert line['higher_is_better'] is True and line['steps'] == 1 and 'workload' in line['config']
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e'] == {'value': line['value'], 'unit': 'Mpaths/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert line['cpu_baseline']['cores'] == len(os.sched_getaffinity(0))
