"""CPU check of the algorithm behind the device's fast HLLL norm (``nrm2_x87_dd2``): its numpy twin against
numpy long double (= x87 extended precision on x86-64), on random, zero-laden, wide-range and short-mantissa
4-vectors, and on the vectors the oracle's HLLL flux actually feeds to the norm."""
import numpy as np
import pytest

import cases
import nrm2_twin as tw

pytestmark = pytest.mark.skipif(np.finfo(np.longdouble).nmant != 63, reason="needs x87 long double")


def _check(X, min_ok):
    r, ok = tw.nrm2_dd(X)
    g = tw.nrm2_longdouble(X)
    assert ok.mean() >= min_ok, ok.mean()
    assert not np.any((r != g) & ok)


def test_random_vectors_match_x87():
    rng = np.random.default_rng(3)
    for spread in (0, 3, 20, 60, 120):
        n = 200000
        X = np.ldexp(1 + rng.random((n, 4)), rng.integers(-spread, spread + 1, (n, 4))) * rng.choice([-1.0, 1.0], (n, 4))
        X[rng.random((n, 4)) < 0.12] = 0
        _check(X, 0.9999)
    n = 200000   # one component at rounding-noise level (rotated frames), exponent gap >> 64 bits
    X = np.ldexp(1 + rng.random((n, 4)), rng.integers(-30, 0, (n, 4)))
    X[:, 2] *= 1e-17 * rng.random(n)
    _check(X, 0.9999)
    for bits in (16, 24, 30):   # differences of nearby doubles: few significant bits, exact ties are common
        X = rng.integers(-2**bits, 2**bits, (n, 4)).astype(float) * np.ldexp(1.0, rng.integers(-40, 0, (n, 4)))
        _check(X, 0.99)
    X = np.array([[3.0, 4.0, 0, 0], [0, 0, 0, 0], [1e-300, 0, 0, 0], [1e200, 1.0, 0, 0], [np.inf, 1, 0, 0], [2.0, 0, 0, 0]])
    r, ok = tw.nrm2_dd(X)
    assert list(ok) == [True, True, False, False, False, True] and r[0] == 5.0 and r[1] == 0.0 and r[5] == 2.0


def test_vectors_of_a_real_hlll_residual_match_x87():
    from oracle import muscl_oracle as mo

    seen = []
    orig = mo.nrm2_x87
    mo.nrm2_x87 = lambda x: (seen.append(np.array(x).reshape(-1, 4).copy()), orig(x))[1]
    try:
        prob = cases.build_oracle(cases.dmr_mesh(), 24, 24, cases.dmr_ic, flux="HLLL", recon="primitive", integrator="RK2", CFL=0.4)
        prob.run(0.0, 1e9, max_steps=3)
    finally:
        mo.nrm2_x87 = orig
    X = np.concatenate(seen)
    assert len(X) > 10000
    _check(X, 0.99)
