"""The arithmetic contract (include/lys_detmath.h) against float64 libm: accuracy bound and special values."""
import numpy as np
import pytest


def ulp_err(got, ref64):
    ref32 = ref64.astype(np.float32)
    ulp = np.abs(np.nextafter(ref32, np.float32(np.inf)).astype(np.float64) - ref32.astype(np.float64))
    ulp = np.where(ulp == 0, 1e-45, ulp)
    return np.abs(got.astype(np.float64) - ref64) / ulp


CASES = [('sin', np.sin, 0.0, 2 * np.pi, 2.0), ('cos', np.cos, 0.0, 2 * np.pi, 2.0), ('exp', np.exp, -87.0, 88.0, 1.5),
         ('log', np.log, 1e-6, 50.0, 1.5), ('acos', np.arccos, -1.0, 1.0, 2.0), ('pow5', lambda x: x ** 5, 0.0, 1.0, 0.5)]


@pytest.mark.parametrize('fn,ref,lo,hi,bound', CASES)
def test_accuracy(orc, fn, ref, lo, hi, bound):
    x = np.linspace(lo, hi, 2_000_001).astype(np.float32)
    got = orc.eval_math(fn, x)
    want = ref(x.astype(np.float64))
    err = ulp_err(got, want)
    if fn in ('sin', 'cos'):                       # relative error near the zeros of sin/cos is not meaningful
        err = err[np.abs(want) > 1e-3]
    assert err.max() <= bound, (fn, err.max())     # bound in f32 ulps of the correctly rounded result


def test_probit_matches_scipy(orc):
    from scipy.stats import norm
    p = np.linspace(1e-6, 0.9999, 200001).astype(np.float32)
    got = orc.eval_math('probit', p).astype(np.float64)
    want = norm.ppf(p.astype(np.float64))
    assert np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))) < 1e-6


def test_special_values(orc):
    inf = np.float32(np.inf)
    assert orc.eval_math('log', [0.0])[0] == -inf and np.isnan(orc.eval_math('log', [-1.0])[0])
    assert orc.eval_math('exp', [-200.0])[0] == 0.0 and orc.eval_math('exp', [100.0])[0] == inf
    assert orc.eval_math('exp', [0.0])[0] == 1.0 and orc.eval_math('log', [1.0])[0] == 0.0
    assert orc.eval_math('probit', [0.0])[0] == -inf and orc.eval_math('probit', [0.5])[0] == 0.0
    assert np.isnan(orc.eval_math('acos', [1.5])[0]) and orc.eval_math('acos', [1.0])[0] == 0.0
    assert orc.eval_math('sin', [0.0])[0] == 0.0 and orc.eval_math('cos', [0.0])[0] == 1.0
    # gradual underflow is kept
    assert 0 < orc.eval_math('exp', [-100.0])[0] < 1e-40
