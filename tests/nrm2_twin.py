"""numpy twin of ``nrm2_x87_dd2`` (pyhype_b200/csrc/pyh_math.cuh): the double-double emulation of OpenBLAS
dnrm2's x87 sequence ``(double) sqrtl(((x0^2 + x1^2) + x2^2) + x3^2)`` with every operation rounded to a 64-bit
significand.  Same operations in the same order as the device code (fma replaced by Dekker's exact product), so
the algorithm -- including its "cannot decide" flag -- can be checked on the CPU against numpy long double,
which IS the x87 format on x86-64."""
import numpy as np

C_TIE = 7.401458596402802e-17  # (1 - 2^-18) / (3 * 2^52)


def _split(a):
    c = 134217729.0 * a
    hi = c - (c - a)
    return hi, a - hi


def _sq_exact(x):
    p = x * x
    h, l = _split(x)
    return p, ((h * h - p) + 2 * h * l) + l * l


def _hi_word(a):
    return (a.view(np.int64) >> 32).astype(np.int64)


def _rn64(h, l, bad, inexact):
    hh = _hi_word(h)
    w = (((hh & 0x7FF00000) - 0x00B00000) | 0x00080000) & 0xFFFFFFFF
    M = (w << 32).astype(np.int64).view(np.float64)  # 1.5 * 2^(exponent(h) - 11)
    with np.errstate(all="ignore"):
        r = (l + M) - M
        d = l - r
        lo32 = h.view(np.int64) & 0xFFFFFFFF
        bad |= (inexact & (np.abs(d) > np.abs(M) * C_TIE)) | ((((hh & 0xFFFFF) | lo32) == 0) & (l < 0))
    return r


def _efield(a):
    return ((a.view(np.int64) >> 52) & 0x7FF).astype(np.int64)


def nrm2_dd(X):
    """X: (n, 4) float64 -> (result, ok)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    n = len(X)
    bad = np.zeros(n, bool)
    mx = np.abs(X).max(axis=1)
    zero = mx == 0
    ex = np.frexp(mx)[1] - 1
    inr = (ex >= -127) & (ex < 127)
    P, L = [], []
    for k in range(4):
        p, e = _sq_exact(np.ascontiguousarray(X[:, k]))
        P.append(p)
        L.append(_rn64(p, e, bad, np.zeros(n, bool)))
    ah, al = P[0], L[0]
    for k in range(1, 4):
        s = ah + P[k]
        bb = s - ah
        t = (ah - (s - bb)) + (P[k] - bb)
        u = (t + al) + L[k]
        far = np.abs(_efield(ah) - _efield(P[k])) > 40
        ah = s + u
        al = u - (ah - s)
        al = _rn64(ah, al, bad, far)
    with np.errstate(all="ignore"):
        h = np.sqrt(ah)
        ph, pe = _sq_exact(h)
        rr = (ah - ph) - pe
        c = (rr + al) / (2 * h)
        vh = h + c
        vl = c - (vh - h)
        vl = _rn64(vh, vl, bad, np.ones(n, bool))
        res = vh + vl
    ok = zero | (inr & ~bad & (res == res))
    return np.where(zero, 0.0, res), ok


def nrm2_longdouble(X):
    xl = np.asarray(X).astype(np.longdouble)
    acc = xl[:, 0] * xl[:, 0]
    for k in range(1, 4):
        acc = acc + xl[:, k] * xl[:, k]
    return np.sqrt(acc).astype(np.float64)
