"""-m "not gpu": the device arithmetic headers (pyhype_b200/csrc/pyh_math.cuh, pyh_fastdiv.cuh) compiled for the
HOST by g++ (tests/host_twin/) and compared with the oracle bit for bit: every Riemann solver in both
reconstruction modes, the four limiters, the two emulations of OpenBLAS' x87 dnrm2, the branch-free division /
reciprocal / square-root sequences -- for the plain-operator policy (Ar<false>, the kernel's fallback) and for the
branch-free fast policy (Ar<true>, the kernel's hot path), with every scaling executed literally (PYH_FOLD_POW2=0), in the shipped build (folded scalings) and in the
PYH_LEAN_CHECKS build.
This is a check of the SOURCE the stage kernel inlines, on the CPU; the kernel itself is checked by the -m gpu
tests.  (The MUFU seed instructions are replaced by stand-ins of similar accuracy: the refinement sequences
converge to the correctly rounded result from any such seed, which is exactly what is verified here.)"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import muscl_oracle as mo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_twin", "twin.cpp")
SHIM = os.path.join(ROOT, "tests", "host_twin", "shim")
CSRC = os.path.join(ROOT, "pyhype_b200", "csrc")
OUT = os.path.join(ROOT, "build", "host_twin")
G = 1.4
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def _ptr(a, t=dp):
    return a.ctypes.data_as(t)


class Twin:
    def __init__(self, fold, lean=0, shortcut=0, cert=0):
        os.makedirs(OUT, exist_ok=True)
        lib = os.path.join(OUT, f"libpyh_twin_fold{fold}_lean{lean}_sc{shortcut}_hc{cert}.so")
        deps = [SRC, os.path.join(SHIM, "cuda_runtime.h"), os.path.join(CSRC, "pyh_math.cuh"), os.path.join(CSRC, "pyh_fastdiv.cuh")]
        if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
            gxx = shutil.which("g++")
            if gxx is None:
                pytest.skip("g++ not available")
            subprocess.run([gxx, "-O2", "-ffp-contract=off", "-std=c++20", "-pthread", "-shared", "-fPIC", f"-DPYH_FOLD_POW2={fold}", f"-DPYH_LEAN_CHECKS={lean}", f"-DPYH_UNIFORM_SHORTCUT={shortcut}", f"-DPYH_HARTEN_CERT={cert}",
                            "-I", SHIM, "-I", CSRC, "-o", lib, SRC], check=True)
        self.lib = C.CDLL(lib)
        self.lib.twin_flux_scale.restype = C.c_double
        self.fold, self.lean = fold, lean
        assert self.lib.twin_fold_pow2() == fold

    def riemann(self, flux, prim, fast, QL, QR):
        n = len(QL)
        F = np.empty((n, 4))
        ok = np.empty(n, dtype=np.int32)
        QL, QR = np.ascontiguousarray(QL), np.ascontiguousarray(QR)
        rc = self.lib.twin_riemann(flux, prim, fast, C.c_double(G), C.c_long(n), _ptr(QL), _ptr(QR), _ptr(F), _ptr(ok, ip))
        assert rc == 0
        return F, ok.astype(bool), self.lib.twin_flux_scale(flux, fast)

    def limiter4(self, lim, fast, dmx, dmn, davg):
        n = len(dmx)
        phi = np.empty(n)
        ok = np.empty(n, dtype=np.int32)
        davg = np.ascontiguousarray(davg)
        assert self.lib.twin_limiter4(lim, fast, C.c_long(n), _ptr(dmx), _ptr(dmn), _ptr(davg), _ptr(phi), _ptr(ok, ip)) == 0
        return phi, ok.astype(bool)

    def nrm2(self, mode, x):
        n = len(x)
        out = np.empty(n)
        ok = np.empty(n, dtype=np.int32)
        x = np.ascontiguousarray(x)
        self.lib.twin_nrm2(mode, C.c_long(n), _ptr(x), _ptr(out), _ptr(ok, ip))
        return out, ok.astype(bool)

    def arith(self, op, a, b):
        n = len(b)
        out = np.empty(n)
        ok = np.empty(n, dtype=np.int32)
        self.lib.twin_arith(op, C.c_long(n), _ptr(a), _ptr(b), _ptr(out), _ptr(ok, ip))
        return out, ok.astype(bool)

    def cons2prim(self, fast, U):
        n = len(U)
        W = np.empty((n, 4))
        ok = np.empty(n, dtype=np.int32)
        U = np.ascontiguousarray(U)
        self.lib.twin_cons2prim(fast, C.c_double(G), C.c_long(n), _ptr(U), _ptr(W), _ptr(ok, ip))
        return W, ok.astype(bool)


# (PYH_FOLD_POW2, PYH_LEAN_CHECKS, PYH_UNIFORM_SHORTCUT): the literal operation list, the shipped default, the opt-in builds
@pytest.fixture(scope="module", params=[(0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 0, 1), (0, 0, 1), (1, 0, 0, 1), (1, 1, 1, 1)],
                ids=["literal", "fold_pow2", "fold_pow2_lean", "fold_pow2_uniform", "literal_uniform", "harten_cert", "everything"])
def twin(request):
    return Twin(*request.param)


def face_states(n, seed):
    """Primitive left / right states in the face frame: independent pairs (strong jumps), near-equal pairs (smooth
    flow), identical pairs, sonic / stagnation points (Harten's correction, flux/base.py:119-146), supersonic pairs of
    both signs (the one-sided HLL branches) and noise-level transverse velocities."""
    rng = np.random.default_rng(seed)
    m = n // 6

    def rand(k):
        return np.stack((rng.uniform(0.05, 12.0, k), rng.uniform(-4.0, 4.0, k), rng.uniform(-4.0, 4.0, k), rng.uniform(0.02, 15.0, k)), axis=-1)

    a = rand(m), rand(m)
    b0 = rand(m)
    b = b0, b0 * (1.0 + 1e-3 * rng.standard_normal((m, 4)))
    c0 = rand(m)
    c = c0, c0.copy()
    d0 = rand(m)
    snd = np.sqrt(G * d0[:, 3] / d0[:, 0])
    d0[:, 1] = snd * rng.choice([-1.0, 1.0, 0.0], m) * (1.0 + 1e-4 * rng.standard_normal(m))
    d = d0, d0 * (1.0 + 1e-2 * rng.standard_normal((m, 4)))
    e0 = rand(m)
    e0[:, 1] = rng.choice([-1.0, 1.0], m) * rng.uniform(3.0, 9.0, m) * np.sqrt(G * e0[:, 3] / e0[:, 0])
    e = e0, e0 * (1.0 + 5e-2 * rng.standard_normal((m, 4)))
    f0 = rand(n - 5 * m)
    f0[:, 2] = 1e-17 * rng.standard_normal(len(f0))
    f0[: len(f0) // 2, 1] = 0.0
    f1 = f0 * (1.0 + 1e-6 * rng.standard_normal(f0.shape))
    f1[len(f0) // 4: len(f0) // 2, 2] = 0.0
    f = f0, f1
    WL = np.concatenate([p[0] for p in (a, b, c, d, e, f)])
    WR = np.concatenate([p[1] for p in (a, b, c, d, e, f)])
    for Wp in (WL, WR):
        Wp[:, 0] = np.abs(Wp[:, 0]) + 1e-3
        Wp[:, 3] = np.abs(Wp[:, 3]) + 1e-3
    return WL, WR


FLUX_ID = {"Roe": 0, "HLLE": 1, "HLLL": 2}


@pytest.mark.parametrize("flux", ["Roe", "HLLE", "HLLL"])
@pytest.mark.parametrize("prim", [0, 1], ids=["conservative", "primitive"])
def test_riemann_solvers_match_oracle(twin, flux, prim):
    WL, WR = face_states(60000, seed=11 + FLUX_ID[flux])
    if prim:
        QL, QR = WL, WR
    else:   # conservative reconstruction: the kernel converts the face states (fvm/base.py:283-303)
        QL, QR = mo.prim_to_cons(WL, G), mo.prim_to_cons(WR, G)
        WL, WR = mo.cons_to_prim(QL, G), mo.cons_to_prim(QR, G)
    with np.errstate(all="ignore"):
        ref = mo.FLUXES[flux](WL, WR, G)
    for fast in (0, 1):
        F, ok, scale = twin.riemann(FLUX_ID[flux], prim, fast, QL, QR)
        assert ok.mean() > 0.999, (fast, ok.mean())
        bad = np.nonzero(ok & np.any(F != scale * ref, axis=1))[0]
        assert len(bad) == 0, (flux, prim, fast, len(bad), QL[bad[:2]], QR[bad[:2]], F[bad[:2]], scale * ref[bad[:2]])


def test_fast_path_rejects_out_of_range_operands(twin):
    WL, WR = face_states(600, seed=3)
    QL, QR = mo.prim_to_cons(WL, G), mo.prim_to_cons(WR, G)
    QL[::3] *= 2.0**300     # what tests/test_gpu_parity.py's fallback case does to the whole state
    QR[::3] *= 2.0**300
    F, ok, _ = twin.riemann(0, 0, 1, QL, QR)
    assert not ok[::3].any() and ok[1::3].all() and ok[2::3].all()


@pytest.mark.parametrize("lim", ["Venkatakrishnan", "VanLeer", "VanAlbada", "BarthJespersen"])
def test_limiters_match_oracle(twin, lim):
    rng = np.random.default_rng(5)
    n = 80000
    q = rng.uniform(-2, 2, n)
    nb = q[:, None] + rng.standard_normal((n, 4)) * rng.choice([1e-12, 1e-6, 1e-2, 1.0], (n, 1))
    mx = np.maximum(q, nb.max(axis=1))
    mn = np.minimum(q, nb.min(axis=1))
    dmx, dmn = mx - q, mn - q
    term = rng.standard_normal((n, 4)) * rng.choice([0.0, 1e-20, 1e-9, 1e-3, 0.5], (n, 4))
    davg = (q[:, None] + term) - q[:, None]          # limiters/base.py:99-102: sub-ulp terms quantise to 0
    with np.errstate(all="ignore"):
        slope = np.where(davg > 0, dmx[:, None] / davg, np.where(davg < 0, dmn[:, None] / davg, 1.0))
        pf = mo.LIMITERS[lim](slope)
    ref = np.minimum(np.minimum(np.minimum(pf[:, 0], pf[:, 1]), pf[:, 2]), pf[:, 3])
    lid = ["Venkatakrishnan", "VanLeer", "VanAlbada", "BarthJespersen"].index(lim)
    for fast in (0, 1):
        phi, ok = twin.limiter4(lid, fast, dmx, dmn, davg)
        assert ok.mean() > 0.9
        assert np.array_equal(phi[ok], ref[ok]), (lim, fast, np.count_nonzero(phi[ok] != ref[ok]))
    assert np.count_nonzero(davg == 0.0) > 1000 and np.count_nonzero(ref > 1.0) > 0 if lim == "Venkatakrishnan" else True


def test_nrm2_emulations_match_x87(twin):
    rng = np.random.default_rng(9)
    n = 200000
    x = rng.standard_normal((n, 4)) * 10.0 ** rng.uniform(-12, 6, (n, 1))
    x[::7, rng.integers(0, 4)] = 0.0
    x[::11] = x[::11] * np.array([1.0, 1e-9, 1e-18, 0.0])
    x[::1000] = 0.0
    ref = mo.nrm2_x87(x)
    out, _ = twin.nrm2(0, x)
    assert np.array_equal(out, ref)
    out, ok = twin.nrm2(1, x)
    assert ok.mean() > 0.999
    assert np.array_equal(out[ok], ref[ok])


def test_division_reciprocal_sqrt_sequences_are_correctly_rounded(twin):
    rng = np.random.default_rng(2)
    n = 400000
    a = rng.standard_normal(n) * 10.0 ** rng.uniform(-30, 30, n)
    b = rng.standard_normal(n) * 10.0 ** rng.uniform(-30, 30, n)
    a[::50] = 0.0
    q, ok = twin.arith(0, a, b)
    assert ok.all() and np.array_equal(q, a / b)
    r, ok = twin.arith(1, a, b)
    assert ok.all() and np.array_equal(r, 1.0 / b)
    s, ok = twin.arith(2, a, np.abs(b))
    assert ok.all() and np.array_equal(s, np.sqrt(np.abs(b)))
    # mantissas next to a power of two and quotients next to rounding ties
    m = 1.0 + np.arange(1, 2001) * 2.0**-52
    bb = np.concatenate((m, 2.0 - m + 1.0, 3.0 * m, 1.0 / m))
    aa = np.concatenate((np.ones_like(m), m[::-1], m, 7.0 * m))
    q, ok = twin.arith(0, aa, bb)
    assert ok.all() and np.array_equal(q, aa / bb)
    s, ok = twin.arith(2, aa, bb)
    assert ok.all() and np.array_equal(s, np.sqrt(bb))
    # out-of-range operands are flagged, not mis-evaluated silently
    _, ok = twin.arith(0, np.array([1.0, 1e-300, 1.0]), np.array([1e300, 1.0, np.inf]))
    assert not ok.any()


def test_cons_to_prim_matches_oracle(twin):
    WL, _ = face_states(30000, seed=21)
    U = mo.prim_to_cons(WL, G)
    ref = mo.cons_to_prim(U, G)
    for fast in (0, 1):
        Wt, ok = twin.cons2prim(fast, U)
        assert ok.all() and np.array_equal(Wt, ref)


SCALES = [-1000, -700, -400, -260, -200, -130, -100, -40, 0, 40, 100, 130, 200, 260, 400, 700, 1000]


@pytest.mark.parametrize("flux", ["Roe", "HLLE", "HLLL"])
@pytest.mark.parametrize("prim", [0, 1], ids=["conservative", "primitive"])
def test_fast_path_is_sound_for_any_magnitude(twin, flux, prim):
    """`ok` must imply equality with the oracle whatever the magnitude of the operands: densities, pressures,
    momenta / velocities and whole states scaled by 2^-1000 .. 2^1000 (the range tests may reject as much as they like,
    but what they let through has to be right)."""
    WL0, WR0 = face_states(3000, seed=31)
    accepted = 0
    for what in ("rho", "p", "vel", "all"):
        for k in SCALES:
            WL, WR = WL0.copy(), WR0.copy()
            f = 2.0**k
            with np.errstate(all="ignore"):
                for Wp in (WL, WR):
                    if what in ("rho", "all"):
                        Wp[:, 0] *= f
                    if what in ("p", "all"):
                        Wp[:, 3] *= f
                    if what == "vel":
                        Wp[:, 1:3] *= f
                if prim:
                    QL, QR = WL, WR
                else:
                    QL, QR = mo.prim_to_cons(WL, G), mo.prim_to_cons(WR, G)
                    WL, WR = mo.cons_to_prim(QL, G), mo.cons_to_prim(QR, G)
                ref = mo.FLUXES[flux](WL, WR, G)
            F, ok, scale = twin.riemann(FLUX_ID[flux], prim, 1, QL, QR)
            with np.errstate(all="ignore"):
                want = scale * ref
                good = (F == want) | (np.isnan(F) & np.isnan(want))
                if twin.fold:
                    # the documented limits of PYH_FOLD_POW2 (pyh_math.cuh): an intermediate in the subnormal range may
                    # cost one unit of the subnormal grid, and an intermediate that overflows does so one factor of two
                    # earlier or later -- both far outside any realizable state
                    good |= (np.abs(F - want) <= 4 * 4.94e-324) | ~np.isfinite(F) | ~np.isfinite(want)
            bad = np.nonzero(ok & ~np.all(good, axis=1))[0]
            assert len(bad) == 0, (flux, prim, what, k, len(bad), QL[bad[:1]], QR[bad[:1]], F[bad[:1]], want[bad[:1]])
            if abs(k) <= 200:   # and inside any physically meaningful range the folded build is exact as well
                strict = (F == want) | (np.isnan(F) & np.isnan(want))
                assert np.all(strict[ok]), (flux, prim, what, k)
            accepted += int(ok.sum())
    assert accepted > 3000 * 4 * 3      # the moderate scalings are accepted


@pytest.mark.parametrize("lim", ["Venkatakrishnan", "VanLeer", "VanAlbada", "BarthJespersen"])
def test_limiter_fast_path_is_sound_for_any_magnitude(twin, lim):
    rng = np.random.default_rng(41)
    n = 4000
    lid = ["Venkatakrishnan", "VanLeer", "VanAlbada", "BarthJespersen"].index(lim)
    accepted = 0
    for kq in SCALES:
        for kt in (-600, -300, -120, -30, 0, 30, 120, 300, 600):
            with np.errstate(all="ignore"):
                q = rng.uniform(-2, 2, n) * 2.0**kq
                nb = q[:, None] * (1.0 + rng.standard_normal((n, 4)) * rng.choice([1e-12, 1e-6, 1e-2, 1.0], (n, 1)))
                mx = np.maximum(q, nb.max(axis=1))
                mn = np.minimum(q, nb.min(axis=1))
                dmx, dmn = mx - q, mn - q
                term = rng.standard_normal((n, 4)) * rng.choice([0.0, 1e-9, 1e-3, 0.5], (n, 4)) * np.ldexp(1.0, max(-1070, min(1020, kq + kt)))
                davg = (q[:, None] + term) - q[:, None]
                slope = np.where(davg > 0, dmx[:, None] / davg, np.where(davg < 0, dmn[:, None] / davg, 1.0))
                pf = mo.LIMITERS[lim](slope)
                ref = np.minimum(np.minimum(np.minimum(pf[:, 0], pf[:, 1]), pf[:, 2]), pf[:, 3])
            fin = np.isfinite(dmx) & np.isfinite(dmn) & np.all(np.isfinite(davg), axis=1)
            phi, ok = twin.limiter4(lid, 1, np.ascontiguousarray(dmx), np.ascontiguousarray(dmn), davg)
            good = (phi == ref) | (np.isnan(phi) & np.isnan(ref))
            bad = np.nonzero(ok & fin & ~good)[0]
            assert len(bad) == 0, (lim, kq, kt, len(bad), dmx[bad[:1]], dmn[bad[:1]], davg[bad[:1]], phi[bad[:1]], ref[bad[:1]])
            accepted += int((ok & fin).sum())
    assert accepted > n * 20


@pytest.mark.parametrize("lim", ["Venkatakrishnan", "VanLeer", "VanAlbada", "BarthJespersen"])
def test_limiter_as_the_kernel_evaluates_it_matches_numpy_including_nan(twin, lim):
    """What the stage kernel does -- the fast pass, and the plain-operator pass where the fast one declines -- against numpy for
    ANY finite operands, NaN results included: an overflowing slope turns the Venkatakrishnan / VanAlbada quotient into
    inf / inf, np.minimum.reduce (limiters/base.py:179-186) carries the NaN into phi, and the reference's run ends there
    (examples/shockbox, step 24).  A minimum that drops the NaN keeps running on a state the reference never had."""
    rng = np.random.default_rng(43)
    n = 4000
    lid = ["Venkatakrishnan", "VanLeer", "VanAlbada", "BarthJespersen"].index(lim)
    nans = 0
    for kq in SCALES:
        for kt in (-600, -300, -120, -30, 0, 30, 120, 300, 600):
            with np.errstate(all="ignore"):
                q = rng.uniform(-2, 2, n) * 2.0**kq
                nb = q[:, None] * (1.0 + rng.standard_normal((n, 4)) * rng.choice([1e-12, 1e-6, 1e-2, 1.0], (n, 1)))
                mx = np.maximum(q, nb.max(axis=1))
                mn = np.minimum(q, nb.min(axis=1))
                dmx, dmn = mx - q, mn - q
                term = rng.standard_normal((n, 4)) * rng.choice([0.0, 1e-9, 1e-3, 0.5], (n, 4)) * np.ldexp(1.0, max(-1070, min(1020, kq + kt)))
                if kt == 0 and kq <= -400:
                    # a variable that is (still) noise next to neighbours of order one, like the y momentum ahead of shockbox's fronts
                    nb = rng.uniform(-2, 2, (n, 4))
                    mx = np.maximum(q, nb.max(axis=1))
                    mn = np.minimum(q, nb.min(axis=1))
                    dmx, dmn = mx - q, mn - q
                    term = rng.standard_normal((n, 4)) * np.ldexp(1.0, kq + rng.integers(0, 40, (n, 4)))
                davg = (q[:, None] + term) - q[:, None]
                slope = np.where(davg > 0, dmx[:, None] / davg, np.where(davg < 0, dmn[:, None] / davg, 1.0))
                ref = np.minimum.reduce(tuple(mo.LIMITERS[lim](slope[:, f]) for f in range(4)))
            fin = np.isfinite(dmx) & np.isfinite(dmn) & np.all(np.isfinite(davg), axis=1)
            dmx, dmn = np.ascontiguousarray(dmx), np.ascontiguousarray(dmn)
            pf, okf = twin.limiter4(lid, 1, dmx, dmn, davg)
            ps, _ = twin.limiter4(lid, 0, dmx, dmn, davg)
            phi = np.where(okf, pf, ps)
            good = (phi == ref) | (np.isnan(phi) & np.isnan(ref))
            bad = np.nonzero(fin & ~good)[0]
            assert len(bad) == 0, (lim, kq, kt, len(bad), dmx[bad[:1]], dmn[bad[:1]], davg[bad[:1]], phi[bad[:1]], ref[bad[:1]])
            nans += int((fin & np.isnan(ref)).sum())
    if lim in ("Venkatakrishnan", "VanAlbada"):
        assert nans > 100      # the sweep does reach the overflow


@pytest.mark.parametrize("flux", ["Roe", "HLLE", "HLLL"])
@pytest.mark.parametrize("prim", [0, 1], ids=["conservative", "primitive"])
def test_riemann_solvers_on_unrealizable_face_states_match_numpy_including_nan(twin, flux, prim):
    """Face states a limited reconstruction can overshoot into -- negative pressure or density on one side or both -- as the
    kernel evaluates them (fast pass, plain-operator pass where that declines): the NaN of the square roots must come out where
    numpy's does (np.maximum.reduce / np.minimum.reduce of flux/HLLL.py:36-37 carry it), because that NaN is what stops the
    reference's run (Euler2D.py:144-152)."""
    WL, WR = face_states(12000, seed=57 + FLUX_ID[flux])
    rng = np.random.default_rng(58)
    n = len(WL)
    for Wp, lo in ((WL, 0), (WR, 1)):
        sel = rng.integers(0, 6, n)
        Wp[sel == lo, 3] *= -1.0                 # negative pressure on this side
        Wp[sel == 2 + lo, 0] *= -1.0             # negative density on this side
        Wp[sel == 4, 3] *= -rng.uniform(0.0, 1e-3, (sel == 4).sum())   # slightly negative pressure on both sides
    if prim:
        QL, QR = WL, WR
    else:
        with np.errstate(all="ignore"):
            QL, QR = mo.prim_to_cons(WL, G), mo.prim_to_cons(WR, G)
            WL, WR = mo.cons_to_prim(QL, G), mo.cons_to_prim(QR, G)
    with np.errstate(all="ignore"):
        ref = mo.FLUXES[flux](WL, WR, G)
    Ff, okf, scale = twin.riemann(FLUX_ID[flux], prim, 1, QL, QR)
    Fs, _, _ = twin.riemann(FLUX_ID[flux], prim, 0, QL, QR)
    F = np.where(okf[:, None].astype(bool), Ff, Fs)
    with np.errstate(all="ignore"):
        want = scale * ref
    good = (F == want) | (np.isnan(F) & np.isnan(want))
    bad = np.nonzero(~np.all(good, axis=1))[0]
    assert np.isnan(want).any(axis=1).mean() > 0.3
    assert len(bad) == 0, (flux, prim, len(bad), QL[bad[:2]], QR[bad[:2]], F[bad[:2]], want[bad[:2]])
