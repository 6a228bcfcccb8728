"""Drive the UNMODIFIED reference (/root/reference, read-only) as a parity oracle.

TEST INFRASTRUCTURE ONLY.  Works only in the build container (the reference tree does not
travel to the GPU box); it is used by ``oracle/make_golden.py`` to generate the fixtures under
``tests/golden/`` and by the ``-m "not gpu"`` tests that pin ``oracle/muscl_oracle.py`` against
the live reference when it is present.  Nothing under ``pyhype_b200/`` imports this module.

Recipe: stub ``mpi4py``/``matplotlib`` (absent in the image), shim the
removed ``np.int`` alias (pyhype/flux/eigen_system.py:251), then import ``pyhype`` from
``/root/reference`` and drive ``Euler2D`` step by step without calling ``MPI.Finalize``.
"""
from __future__ import annotations

import logging
import os
import sys

import numpy as np

REFERENCE_ROOT = os.environ.get("PYHYPE_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pyhype"))


def activate():
    """Put the stubs and the reference on sys.path; returns the imported ``pyhype`` package."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if not hasattr(np, "int"):
        np.int = int  # pyhype/flux/eigen_system.py:251 uses the removed alias
    if not hasattr(np, "float"):
        np.float = float  # annotation at pyhype/blocks/quad_block.py:423
    for p in (_STUBS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_pyhype_ref")
    import pyhype  # noqa: F401

    logging.getLogger().setLevel(logging.ERROR)
    return pyhype


def patch_hlle():
    """HLLE never runs in the reference (pyhype/flux/HLLE.py:25 passes ``config`` as ``fluid``;
    :41 subtracts State objects).  Two-edit patched variant; every HLLE
    result is labelled "patched oracle (2 edits)"."""
    activate()
    from pyhype.flux.HLLE import FluxHLLE
    from pyhype.states import ConservativeState
    from pyhype.states.primitive import RoePrimitiveState

    def compute_flux(self, WL, WR):
        Wroe = RoePrimitiveState(self.config.fluid, WL, WR)  # edit 1
        L_p, L_m = self.wavespeeds_x(WL)
        R_p, R_m = self.wavespeeds_x(WR)
        Lp, Lm = self.harten_correction_x(Wroe, WL, WR, L_p=L_p, L_m=L_m, R_p=R_p, R_m=R_m)
        L_plus = np.maximum.reduce((R_m, Lm))[:, :, None]
        L_minus = np.minimum.reduce((L_p, Lp))[:, :, None]
        UR = ConservativeState(fluid=WR.fluid, state=WR)
        UL = ConservativeState(fluid=WL.fluid, state=WL)
        FluxR = WR.F(U=UR)
        FluxL = WL.F(U=UL)
        Flux = (L_plus * FluxL - L_minus * FluxR + L_minus * L_plus * (UR.data - UL.data)) / (
            L_plus - L_minus
        )  # edit 2: .data
        Flux = np.where(L_minus >= 0, FluxL, Flux)
        Flux = np.where(L_plus <= 0, FluxR, Flux)
        return Flux

    FluxHLLE.compute_flux = compute_flux


class CallableIC:
    """InitialCondition adaptor: ``fn(x, y) -> (ny, nx, 4)`` dimensional primitive array."""

    def __init__(self, fn):
        self.fn = fn

    def apply_to_block(self, block):
        from pyhype.states.conservative import ConservativeState
        from pyhype.states.primitive import PrimitiveState

        W = np.ascontiguousarray(self.fn(block.mesh.x[:, :, 0], block.mesh.y[:, :, 0]), dtype=float)
        block.state.data = PrimitiveState(fluid=block.config.fluid, array=W).to_type(ConservativeState).data
        block.state.make_non_dimensional()


def make_config(**over):
    """SolverConfig of the reference with explosion_multi defaults (examples/explosion_multi/config.py)."""
    activate()
    from pyhype.fluids import Air
    from pyhype.solver_config import SolverConfig
    from pyhype.states import ConservativeState, PrimitiveState

    recon = over.pop("reconstruction_type", "conservative")
    if isinstance(recon, str):
        recon = {"conservative": ConservativeState, "primitive": PrimitiveState}[recon]
    kw = dict(
        fvm_type="MUSCL",
        fvm_spatial_order=2,
        fvm_num_quadrature_points=1,
        fvm_gradient_type="GreenGauss",
        fvm_flux_function_type="Roe",
        fvm_slope_limiter_type="Venkatakrishnan",
        time_integrator="RK4",
        initial_condition=None,
        interface_interpolation="arithmetic_average",
        reconstruction_type=recon,
        write_solution=False,
        CFL=0.7,
        t_final=0.07,
        realplot=False,
        profile=False,
        fluid=Air(a_inf=343.0, rho_inf=1.0),
        nx=40,
        ny=40,
        nghost=1,
        use_JIT=True,
    )
    kw.update(over)
    return SolverConfig(**kw)


class RefRun:
    """A live reference solver stepped manually (Euler2D.solve() minus MPI.Finalize)."""

    def __init__(self, config, mesh):
        activate()
        from pyhype.solvers import Euler2D

        self.solver = Euler2D(config=config, mesh_config=mesh)
        self.solver.apply_initial_condition()
        self.solver.apply_boundary_condition()
        self.dts = []

    @property
    def blocks(self):
        return list(self.solver.blocks)

    def step(self, n=1):
        s = self.solver
        for _ in range(n):
            if not (s.t < s.t_final):
                break
            dt = s.get_dt()
            s._update_solution_blocks(dt=dt)
            s._realizability_check()
            s.t += dt
            s.num_time_step += 1
            self.dts.append(float(dt))
        return self

    def states(self):
        return [b.state.data.copy() for b in self.blocks]

    def ghosts(self):
        return [{d: b.ghost[k].state.data.copy() for d, k in zip("EWNS", (1, -1, 2, -2))} for b in self.blocks]

    def residuals(self):
        """dUdt of every block at the current state (+ gradients, phi, face fluxes)."""
        out = []
        for b in self.blocks:
            R = b.dUdt().copy()
            rb = b.recon_block
            out.append(
                dict(
                    R=R,
                    gx=rb.grad.x.copy(),
                    gy=rb.grad.y.copy(),
                    phi=rb.fvm.limiter.phi.copy(),
                    FE=rb.fvm.Flux.E[0].copy(),
                    FW=rb.fvm.Flux.W[0].copy(),
                    FN=rb.fvm.Flux.N[0].copy(),
                    FS=rb.fvm.Flux.S[0].copy(),
                )
            )
            b.clear_cache()
            rb.clear_cache()
        return out
