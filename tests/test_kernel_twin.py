"""-m "not gpu": the fused stage kernel itself (pyhype_b200/csrc/pyh_stage_march.cuh, plus the ghost / geometry /
layout kernels) compiled by g++ and EXECUTED on the CPU by a small thread-block emulator (tests/host_twin/), and
compared with the golden fixtures of the unmodified reference: ghost strips, Green-Gauss gradients, limiter, residual
and the state after one Euler update -- by value, max-abs-diff 0, for every fixture (all fluxes, limiters,
reconstruction modes, 1-3 quadrature points, Dirichlet / slip-wall / outflow edges, irregular block topology) and for
several thread-block shapes (strip boundaries inside a block, multiple row strips).  This checks the SOURCE of the
kernel -- indexing, ring buffers, barriers placement, edge handling, arithmetic -- before any GPU time is spent; what
nvcc makes of it on the device is the business of the -m gpu tests."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import golden_io
from pyhype_b200.mesh.quad_mesh import QuadMesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_twin", "kernel_twin.cpp")
SHIM = os.path.join(ROOT, "tests", "host_twin", "shim")
CSRC = os.path.join(ROOT, "pyhype_b200", "csrc")
OUT = os.path.join(ROOT, "build", "host_twin")
SIDES = ("E", "W", "N", "S")
FLUX = {"Roe": 0, "HLLE": 1, "HLLL": 2}
LIM = {"Venkatakrishnan": 0, "VanLeer": 1, "VanAlbada": 2, "BarthJespersen": 3}
BC = {None: 0, "Reflection": 1, "Slipwall": 2, "OutletDirichlet": 3}
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


# extra -D flags of the builds that are checked: the shipped default (folded power-of-two scalings, lean range checks, certified
# Harten test), the literal operation list of SURVEY.md section 8A with every check, and the opt-in uniform-flow shortcut
BUILDS = {"default": [],
          "literal": ["-DPYH_FOLD_POW2=0", "-DPYH_SKIP_UNIT_ROT=0", "-DPYH_LEAN_CHECKS=0", "-DPYH_HARTEN_CERT=0"],
          "full_checks": ["-DPYH_LEAN_CHECKS=0", "-DPYH_HARTEN_CERT=0"],
          "uniform_shortcut": ["-DPYH_UNIFORM_SHORTCUT=1"]}


def build(name):
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, f"libpyh_kernel_twin_{name}.so")
    deps = [SRC, os.path.join(SHIM, "cuda_runtime.h")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        gxx = shutil.which("g++")
        if gxx is None:
            pytest.skip("g++ not available")
        subprocess.run([gxx, "-O1", "-ffp-contract=off", "-std=c++20", "-pthread", "-shared", "-fPIC", *BUILDS[name], "-I", SHIM,
                        "-I", CSRC, "-o", so, SRC], check=True)
    return C.CDLL(so)


@pytest.fixture(scope="module", params=list(BUILDS))
def lib(request):
    return build(request.param)


@pytest.fixture(scope="module")
def lib_default():
    return build("default")


def marshal(fx):
    nx, ny, gids = fx.nx, fx.ny, fx.gids
    idx = {g: i for i, g in enumerate(gids)}
    sch = fx.scheme()
    mlen = max(nx, ny)
    nb = len(gids)
    arr = {k: [] for k in ("nodes_x", "nodes_y", "area", "cos_v", "sin_v", "cos_h", "sin_h")}
    nbr = np.full((nb, 4), -1, dtype=np.int32)
    bc = np.zeros((nb, 4), dtype=np.int32)
    cart = np.zeros(nb, dtype=np.int32)
    dirichlet = np.zeros((nb, 4, mlen, 4))
    U = np.empty((nb, ny, nx, 4))
    for g in gids:
        b = fx.blocks[g]
        m = QuadMesh(nx, ny, NE=b["NE"], NW=b["NW"], SE=b["SE"], SW=b["SW"])
        for k in arr:
            arr[k].append(np.ascontiguousarray(getattr(m, k), dtype=np.float64))
        cart[idx[g]] = int(m.is_cartesian)
        for s, side in enumerate(SIDES):
            n = b["Neighbor" + side]
            nbr[idx[g], s] = -1 if n is None else idx[n]
            v = b["BCType" + side]
            if isinstance(v, np.ndarray):
                bc[idx[g], s] = 4
                strip = np.asarray(v, dtype=np.float64).reshape(-1, 4)
                dirichlet[idx[g], s, : len(strip)] = strip
            else:
                bc[idx[g], s] = BC[v]
        U[idx[g]] = fx[f"U0_{g}"]
    A = {k: np.ascontiguousarray(np.stack(v)) for k, v in arr.items()}
    return idx, sch, nb, mlen, A, nbr, bc, cart, dirichlet, U


def run_stage(lib, fx, nt, tys, coef):
    nx, ny = fx.nx, fx.ny
    idx, sch, nb, mlen, A, nbr, bc, cart, dirichlet, U = marshal(fx)
    R = np.empty((nb, ny, nx, 4))
    Un = np.empty((nb, ny, nx, 4))
    G = np.empty((nb, 12, ny, nx))
    gh = np.zeros((nb, 4, mlen, 4))
    p = lambda a, t=dp: a.ctypes.data_as(t)
    rc = lib.twin_stage(FLUX[sch["flux"]], LIM[sch["limiter"]], int(sch["recon"] == "primitive"), int(sch["nqp"]), nx, ny, nb, nt, tys,
                        C.c_double(fx.meta["gamma"]), C.c_double(coef), p(A["nodes_x"]), p(A["nodes_y"]), p(A["area"]), p(A["cos_v"]),
                        p(A["sin_v"]), p(A["cos_h"]), p(A["sin_h"]), p(nbr, ip), p(bc, ip), p(cart, ip), p(dirichlet), p(U), p(R), p(Un),
                        p(G), p(gh))
    assert rc == 0
    return idx, U, R, Un, G, gh


def check(fx, idx, U, R, Un, G, gh, coef):
    nx, ny = fx.nx, fx.ny
    for g in fx.gids:
        i = idx[g]
        for s, side in enumerate(SIDES):
            ref = fx[f"ghost0_{g}_{side}"].reshape(-1, 4)
            assert np.array_equal(gh[i, s, : len(ref)], ref), (fx.name, g, side)
        for k, name in enumerate(("gx", "gy", "phi")):
            got = np.moveaxis(G[i, 4 * k: 4 * k + 4], 0, -1)
            assert np.array_equal(got, fx[f"{name}_{g}"]), (fx.name, g, name, np.abs(got - fx[f"{name}_{g}"]).max())
        assert np.array_equal(R[i], fx[f"R_{g}"]), (fx.name, g, np.abs(R[i] - fx[f"R_{g}"]).max())
        assert np.array_equal(Un[i], U[i] + coef * fx[f"R_{g}"]), (fx.name, g)      # explicit_runge_kutta.py:84-89


@pytest.mark.parametrize("name", golden_io.names())
def test_stage_kernel_source_matches_reference_fixture(lib, name):
    fx = golden_io.Fixture(name)
    coef = 0.37 * float(fx["dts"][0])
    out = run_stage(lib, fx, nt=128 if name.startswith(("em_roe", "dmr_hlll", "jet_hlll")) else 64, tys=64, coef=coef)
    check(fx, *out, coef)


@pytest.mark.parametrize("nt,tys", [(32, 5), (64, 4), (96, 7)])
@pytest.mark.parametrize("name", ["em_ragged_roe_rk4", "dmr_hlll_venkat_prim_rk2", "wedge_roe_cons_rk2", "jet_hlle_prim_rk2",
                                  "step_hlll_prim_rk2", "em_nqp3", "cart_roe_cons_rk4"])
def test_stage_kernel_source_with_strip_boundaries_inside_the_block(lib_default, name, nt, tys):
    lib = lib_default
    """28- / 60- / 92-column strips and 4- to 7-row strips: block-interior strip seams, ring lanes on real cells, ragged tails."""
    fx = golden_io.Fixture(name)
    coef = 0.37 * float(fx["dts"][0])
    out = run_stage(lib, fx, nt=nt, tys=tys, coef=coef)
    check(fx, *out, coef)


def run_steps(lib, fx, nt, tys):
    from pyhype_b200._lib import PYH_MAX_STAGES
    from pyhype_b200.time_marching import TABLEAUX

    nx, ny = fx.nx, fx.ny
    idx, sch, nb, mlen, A, nbr, bc, cart, dirichlet, U = marshal(fx)
    rows = TABLEAUX[sch["integrator"]]
    tab = np.zeros(PYH_MAX_STAGES * PYH_MAX_STAGES)
    for s_, row in enumerate(rows):
        for k, a in enumerate(row):
            tab[s_ * PYH_MAX_STAGES + k] = float(a)
    dts = np.ascontiguousarray(fx["dts"], dtype=np.float64)
    Uout = np.empty_like(U)
    p = lambda a, t=dp: a.ctypes.data_as(t)
    rc = lib.twin_steps(FLUX[sch["flux"]], LIM[sch["limiter"]], int(sch["recon"] == "primitive"), int(sch["nqp"]), nx, ny, nb, nt, tys,
                        C.c_double(fx.meta["gamma"]), len(rows), p(tab), len(dts), p(dts), p(A["nodes_x"]), p(A["nodes_y"]), p(A["area"]),
                        p(A["cos_v"]), p(A["sin_v"]), p(A["cos_h"]), p(A["sin_h"]), p(nbr, ip), p(bc, ip), p(cart, ip), p(dirichlet), p(U),
                        p(Uout))
    assert rc == 0
    return idx, Uout


STEP_SUBSET = ["em_roe_venkat_cons_rk4", "dmr_hlll_venkat_prim_rk2", "wedge_roe_cons_rk2", "jet_hlle_prim_rk2", "step_hlll_prim_rk2",
               "em_int_DormandPrince5", "em_int_ExplicitEuler1", "em_nqp2"]


@pytest.mark.parametrize("name", STEP_SUBSET)
def test_whole_time_steps_of_every_build_match_reference_fixture(lib, name):
    """The opt-in builds on a cross-section of the fixtures (all of them run on the default build below)."""
    fx = golden_io.Fixture(name)
    idx, Uout = run_steps(lib, fx, nt=32, tys=64)
    for g in fx.gids:
        assert np.array_equal(Uout[idx[g]], fx[f"U_{g}"]), (name, g, np.abs(Uout[idx[g]] - fx[f"U_{g}"]).max())


@pytest.mark.parametrize("name", [n for n in golden_io.names() if n not in STEP_SUBSET])
def test_whole_time_steps_of_the_kernel_source_match_reference_fixture(lib_default, name):
    lib = lib_default
    """Every stage of every step -- the product's plan logic (pyh_plan.cuh: buffer roles, running partial sums), the
    stage kernel and the ghost refresh in between -- with the reference's own dt sequence: the state after N steps is
    the reference's, bit for bit, for all ten tableaux."""
    fx = golden_io.Fixture(name)
    assert len(fx["dts"]) == fx.meta["steps"]
    idx, Uout = run_steps(lib, fx, nt=32, tys=64)   # two 28-column strips; other shapes are varied by the single-stage tests
    for g in fx.gids:
        assert np.array_equal(Uout[idx[g]], fx[f"U_{g}"]), (name, g, np.abs(Uout[idx[g]] - fx[f"U_{g}"]).max())


TSAN_SCRIPT = """
import sys, ctypes as C
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import numpy as np, golden_io, test_kernel_twin as T
lib = C.CDLL({so!r})
for name in {names!r}:
    fx = golden_io.Fixture(name)
    coef = 0.37 * float(fx["dts"][0])
    T.check(fx, *T.run_stage(lib, fx, nt=64, tys=7, coef=coef), coef)
    idx, Uout = T.run_steps(lib, fx, nt=32, tys=64)
    assert all(np.array_equal(Uout[idx[g]], fx[f"U_{{g}}"]) for g in fx.gids)
print("TWIN-RUN-COMPLETE")
"""


def test_shared_memory_protocol_of_the_stage_kernel_is_race_free_under_thread_sanitizer():
    """The emulator's threads are real threads and __syncthreads() is a real barrier, so ThreadSanitizer sees every
    shared-memory access of the kernel: the one-barrier-per-row protocol over the double-buffered rings has no
    unordered conflicting accesses (checked stricter than CUDA needs: lanes of one warp count as separate threads too).
    With the barrier disabled the same run produces race reports, so the detector does see the kernel's accesses."""
    import sys

    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    tsan = subprocess.run([gxx, "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(tsan) or not os.path.exists(tsan):
        pytest.skip("libtsan not available")
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, "libpyh_kernel_twin_tsan.so")
    deps = [SRC, os.path.join(SHIM, "cuda_runtime.h")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run([gxx, "-O1", "-g", "-fsanitize=thread", "-ffp-contract=off", "-std=c++20", "-pthread", "-shared", "-fPIC",
                        "-I", SHIM, "-I", CSRC, "-o", so, SRC], check=True)
    code = TSAN_SCRIPT.format(root=ROOT, tests=os.path.join(ROOT, "tests"), so=so,
                              names=["em_ragged_roe_rk4", "dmr_hlll_venkat_prim_rk2", "em_nqp3"])
    env = dict(os.environ, LD_PRELOAD=tsan, TSAN_OPTIONS="exitcode=0 halt_on_error=0")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    if "TWIN-RUN-COMPLETE" not in r.stdout and "ThreadSanitizer" not in r.stderr:
        pytest.skip("the interpreter does not run under a preloaded libtsan here: " + r.stderr[-300:])
    assert "TWIN-RUN-COMPLETE" in r.stdout, r.stderr[-2000:]
    assert "ThreadSanitizer: data race" not in r.stderr, r.stderr[:3000]


def run_loop(lib, fx, t0, t_final, max_steps, nt=32, tys=64):
    from pyhype_b200._lib import PYH_MAX_STAGES
    from pyhype_b200.time_marching import TABLEAUX

    nx, ny = fx.nx, fx.ny
    idx, sch, nb, mlen, A, nbr, bc, cart, dirichlet, U = marshal(fx)
    rows = TABLEAUX[sch["integrator"]]
    tab = np.zeros(PYH_MAX_STAGES * PYH_MAX_STAGES)
    for s_, row in enumerate(rows):
        for k, a in enumerate(row):
            tab[s_ * PYH_MAX_STAGES + k] = float(a)
    Uout = np.empty_like(U)
    dts = np.zeros(max_steps)
    t = C.c_double(0.0)
    nsteps, bad = C.c_int(0), C.c_int(0)
    p = lambda a, ty=dp: a.ctypes.data_as(ty)
    rc = lib.twin_run(FLUX[sch["flux"]], LIM[sch["limiter"]], int(sch["recon"] == "primitive"), int(sch["nqp"]), nx, ny, nb, nt, tys,
                      C.c_double(fx.meta["gamma"]), C.c_double(sch["CFL"]), len(rows), p(tab), C.c_double(t0), C.c_double(t_final),
                      max_steps, p(A["nodes_x"]), p(A["nodes_y"]), p(A["area"]), p(A["cos_v"]), p(A["sin_v"]), p(A["cos_h"]), p(A["sin_h"]),
                      p(nbr, ip), p(bc, ip), p(cart, ip), p(dirichlet), p(U), p(Uout), p(dts), C.byref(t), C.byref(nsteps), C.byref(bad))
    assert rc == 0
    return idx, Uout, dts[: nsteps.value], t.value, nsteps.value, bad.value


@pytest.mark.parametrize("name", ["em_roe_venkat_cons_rk4", "dmr_hlll_venkat_prim_rk2", "wedge_roe_cons_rk2", "jet_hlle_prim_rk2",
                                  "em_int_ExplicitEuler1", "em_int_DormandPrince5"])
def test_device_resident_time_loop_source_reproduces_dt_sequence_and_state(lib_default, name):
    """pyh_run's loop, kernel for kernel, on the CPU: k_dt (the CFL reduction with its warp shuffles, shared-memory stage
    and atomic minimum, quad_block.py:423-436), k_dt_finalize, the stages, k_step_end -- the reference's dt sequence and
    final state, bit for bit."""
    fx = golden_io.Fixture(name)
    n = fx.meta["steps"]
    idx, Uout, dts, t, nsteps, bad = run_loop(lib_default, fx, 0.0, 1e9, n)
    assert nsteps == n and not bad
    assert list(dts) == list(fx["dts"])
    for g in fx.gids:
        assert np.array_equal(Uout[idx[g]], fx[f"U_{g}"]), (name, g)


def test_device_resident_time_loop_source_clamps_dt_and_stops_at_t_final(lib_default):
    fx = golden_io.Fixture("em_roe_venkat_cons_rk4")
    ref = list(fx["dts"])
    t_final = ref[0] + ref[1] + 0.25 * ref[2]
    idx, Uout, dts, t, nsteps, bad = run_loop(lib_default, fx, 0.0, t_final, 16)
    assert nsteps == 3 and not bad and list(dts[:2]) == ref[:2]
    tt = 0.0 + ref[0]
    tt += ref[1]
    assert dts[2] == t_final - tt and t == tt + dts[2] and not (t < t_final)      # solvers/base.py:132-136


def test_device_resident_time_loop_source_flags_unrealizable_states(lib_default):
    fx = golden_io.Fixture("em_lim_VanLeer")
    g0 = fx.gids[0]
    U = fx.z[f"U0_{g0}"].copy()
    U[3, 4, 0] = -1.0        # negative density (states/conservative.py:161-165)
    fx.z = dict(fx.z)
    fx.z[f"U0_{g0}"] = U
    idx, Uout, dts, t, nsteps, bad = run_loop(lib_default, fx, 0.0, 1e9, 3)
    assert bad and nsteps == 0


@pytest.mark.parametrize("split", [1, 2, 3])
@pytest.mark.parametrize("name,nt,tys", [("em_roe_venkat_cons_rk4", 12, 4), ("jet_hlll_prim_rk2", 16, 4), ("wedge_hlll_prim_rk2", 10, 3)])
def test_edge_and_interior_launches_cover_every_cell_once(lib_default, monkeypatch, name, nt, tys, split):
    """A context with remote neighbours launches every stage as thin edge strips (first / last rows, first / last column strip)
    plus an interior launch (pyh_plan.cuh: plan_tiles) so that the strip exchange runs behind the interior.  The twin runs the
    same tile plan -- interior FIRST, edges last, i.e. not even in the product's order -- and must reproduce the reference's
    ghost strips, gradients, limiter, residual and updated state exactly as the single launch does."""
    monkeypatch.setenv("PYH_TWIN_SPLIT", str(split))
    fx = golden_io.Fixture(name)
    coef = 0.37 * float(fx["dts"][0])
    out = run_stage(lib_default, fx, nt=nt, tys=tys, coef=coef)
    check(fx, *out, coef)


def _tile_cover(lib, nx, ny, nt, tys, ns, ew):
    """cells each launch of plan_tiles writes, as the stage kernel maps them (pyh_stage_march.cuh: output columns
    bx * (nt - 4) .. + nt - 5 of strip bx, rows [row0 + by * rowstride, min(+ tys, row1)))"""
    out = (C.c_int * 27)()
    n = lib.twin_plan_tiles(nx, ny, nt, tys, int(ns), int(ew), out)
    cover = np.zeros((ny, nx), dtype=np.int32)
    edge = np.zeros((ny, nx), dtype=bool)
    for q in range(n):
        row0, rowstride, row1, xfirst, xstride, gx, gy, t_, is_edge = out[9 * q: 9 * q + 9]
        for by in range(gy):
            i0 = row0 + by * rowstride
            i1 = min(i0 + t_, row1)
            for x in range(gx):
                bx = xfirst + x * xstride
                j0, j1 = bx * (nt - 4), min(bx * (nt - 4) + nt - 4, nx)
                cover[i0:i1, j0:j1] += 1
                if is_edge:
                    edge[i0:i1, j0:j1] = True
    return n, cover, edge


@pytest.mark.parametrize("nx,ny,nt,tys", [(2048, 2048, 128, 64), (150, 150, 160, 3), (26, 22, 64, 4), (26, 22, 32, 4), (500, 500, 64, 16),
                                          (36, 6, 16, 4), (24, 8, 12, 3), (300, 9, 32, 2), (1080, 60, 128, 16), (7, 40, 64, 64)])
@pytest.mark.parametrize("ns,ew", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_tile_plan_covers_every_cell_once_and_edges_precede_the_exchange(lib_default, nx, ny, nt, tys, ns, ew):
    """Invariants of pyh_plan.cuh: plan_tiles for every shape the chooser can pick: (1) the launches of a stage write every cell
    exactly once; (2) every cell a remote neighbour needs -- first / last row with north / south neighbours on other ranks,
    first / last column with east / west ones -- is written by an EDGE launch, because the strip exchange is enqueued right
    behind the edge launches (round 2: the 8-rank run caught a plan that left the east / west columns of a block too narrow to
    split in the interior launch)."""
    n, cover, edge = _tile_cover(lib_default, nx, ny, nt, tys, ns, ew)
    assert 1 <= n <= 3
    assert (cover == 1).all(), (n, np.argwhere(cover != 1)[:5])
    if ns:
        assert edge[0, :].all() and edge[-1, :].all()
    if ew:
        assert edge[:, 0].all() and edge[:, -1].all()
    if not ns and not ew:
        assert n == 1 and edge.all()       # single rank: one launch, nothing to overlap


# ---- the three-kernel stage (pyh_stage_split.cuh) -------------------------------------------------------------------------------
# (256 OS threads per emulated thread block make this path slow on the CPU: a cross-section by default, the rest with
# PYH_TWIN_ALL=1 -- last full run: profiles/r02p_split_twin_full.txt; on the B200 every GPU test runs once per path)
SPLIT_STEPS = ["dmr_hlll_venkat_prim_rk2", "wedge_roe_cons_rk2", "jet_hlle_prim_rk2", "em_int_ExplicitEuler1", "em_lim_BarthJespersen",
               "em_ragged_roe_rk4", "cart_roe_cons_rk4"]
if os.environ.get("PYH_TWIN_ALL"):
    SPLIT_STEPS += ["em_roe_venkat_cons_rk4", "step_hlll_prim_rk2", "em_int_DormandPrince5", "em_lim_VanLeer", "em_lim_VanAlbada"]


@pytest.fixture
def split_path(monkeypatch, lib_default):
    monkeypatch.setenv("PYH_TWIN_SPLITPATH", "1")
    lib_default.twin_split_stages.restype = C.c_longlong
    return lib_default


@pytest.mark.parametrize("name", [n for n in SPLIT_STEPS if n in golden_io.names()])
def test_split_stage_kernels_whole_time_steps_match_reference_fixture(split_path, name):
    """k_split_recon -> k_split_flux -> k_split_update (what a context launches per stage when that is faster on its blocks than the fused
    kernel) through the product's plan logic, ghost push included: the state after N steps is the reference's, bit for bit --
    every flux, both reconstruction modes, the four limiters, Dirichlet / reflection / outflow edges, skewed and Cartesian
    blocks, a single-stage and a seven-stage tableau."""
    fx = golden_io.Fixture(name)
    before = split_path.twin_split_stages()
    idx, Uout = run_steps(split_path, fx, nt=32, tys=64)
    assert split_path.twin_split_stages() > before
    for g in fx.gids:
        assert np.array_equal(Uout[idx[g]], fx[f"U_{g}"]), (name, g, np.abs(Uout[idx[g]] - fx[f"U_{g}"]).max())


@pytest.mark.parametrize("name", ["cart_roe_cons_rk4", "em_int_ExplicitEuler1"] + (["em_roe_venkat_cons_rk4", "dmr_hlll_venkat_prim_rk2"] if os.environ.get("PYH_TWIN_ALL") else []))
def test_split_stage_kernels_in_the_device_resident_time_loop(split_path, name):
    """... and inside pyh_run's loop: the CFL minimum k_split_update reduces for the next step (warp shuffles + atomicMin) gives the
    reference's dt sequence."""
    fx = golden_io.Fixture(name)
    n = fx.meta["steps"]
    before = split_path.twin_split_stages()
    idx, Uout, dts, t, nsteps, bad = run_loop(split_path, fx, 0.0, 1e9, n)
    assert split_path.twin_split_stages() > before
    assert nsteps == n and not bad
    assert list(dts) == list(fx["dts"])
    for g in fx.gids:
        assert np.array_equal(Uout[idx[g]], fx[f"U_{g}"]), (name, g)


def test_split_stage_kernels_flag_unrealizable_states_and_nan(split_path):
    if not os.environ.get("PYH_TWIN_ALL"):
        pytest.skip("26 s of emulation: set PYH_TWIN_ALL=1 (the GPU suite replays shockbox through both stage paths)")
    import test_named_configs as N

    fp = N.Named("shockbox")            # the reference's own abort: NaN through the limiter in step 24
    n = fp.meta["aborts_in_step"]
    idx, Uout, dts, t, nsteps, bad = run_loop(split_path, fp, 0.0, fp.meta["t_final_nd"], n + 2, nt=64, tys=16)
    assert bad and nsteps == n and np.array_equal(dts[: n - 1], fp.dts[: n - 1])
    assert np.isnan(Uout[idx[fp.gids[0]]]).sum() == 16
