"""-m gpu: the reference-facing Python API (SolverConfig, Euler2D(config, mesh).solve(), mesh
generators, InitialCondition, PrimitiveDirichletBC) driving the CUDA engine; written the way the
reference's example scripts are (examples/explosion_multi, examples/dmr, examples/supersonic_wedge)."""
import numpy as np
import pytest

import cases
from pyhype_b200.boundary_conditions.base import PrimitiveDirichletBC
from pyhype_b200.fluids import Air
from pyhype_b200.initial_conditions.base import InitialCondition
from pyhype_b200.mesh.rectangular import RectagularMeshGenerator
from pyhype_b200.solver_config import SolverConfig
from pyhype_b200.solvers import Euler2D
from pyhype_b200.states import ConservativeState, PrimitiveState

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("stage_path")]   # every test once per stage implementation (conftest.py)


class ExplosionInitialCondition(InitialCondition):
    """examples/explosion/initial_condition.py, verbatim usage of the public API"""

    def apply_to_block(self, block):
        left_state = PrimitiveState(
            fluid=block.config.fluid, array=np.array([4.6968, 0.0, 0.0, 404400.0]).reshape((1, 1, 4))
        ).to_type(ConservativeState)
        right_state = PrimitiveState(
            fluid=block.config.fluid, array=np.array([1.1742, 0.0, 0.0, 101100.0]).reshape((1, 1, 4))
        ).to_type(ConservativeState)
        _x_cond = np.logical_and(block.mesh.x >= 3, block.mesh.x <= 7)
        _y_cond = np.logical_and(block.mesh.y >= 3, block.mesh.y <= 7)
        block.state.data = np.where(np.logical_and(_x_cond, _y_cond), left_state.data, right_state.data)
        block.state.make_non_dimensional()


def em_config(**over):
    kw = dict(
        fvm_type="MUSCL", fvm_spatial_order=2, fvm_num_quadrature_points=1, fvm_gradient_type="GreenGauss",
        fvm_flux_function_type="Roe", fvm_slope_limiter_type="Venkatakrishnan", time_integrator="RK4",
        initial_condition=ExplosionInitialCondition(), interface_interpolation="arithmetic_average",
        reconstruction_type=ConservativeState, write_solution=False, CFL=0.7, t_final=0.002, realplot=False,
        profile=False, fluid=Air(a_inf=343.0, rho_inf=1.0), nx=30, ny=30, nghost=1, use_JIT=True,
    )
    kw.update(over)
    return SolverConfig(**kw)


def em_mesh():
    return RectagularMeshGenerator.generate(
        BCE=["Reflection"], BCW=["Reflection"], BCN=["Reflection"], BCS=["Reflection"],
        east=10.0, west=0.0, north=20.0, south=0.0, n_blocks_horizontal=2, n_blocks_vertical=4,
    )


def oracle_run(blocks, nx, ny, ic, t_final, **kw):
    prob = cases.build_oracle(blocks, nx, ny, ic, **kw)
    t, dts = prob.run(0.0, t_final)
    return prob, t, dts


def test_explosion_multi_solve_matches_oracle_to_t_final():
    config = em_config()
    sim = Euler2D(config=config, mesh_config=em_mesh())
    sim.solve()
    prob, t, dts = oracle_run(em_mesh().dict, 30, 30, cases.explosion_ic, 0.002 * 343.0)
    assert sim.num_time_step == len(dts) and sim.t == t
    for block in sim.blocks:
        ref = prob.blocks[block.global_block_num]
        assert np.array_equal(block.state.data, ref.U)
        assert np.array_equal(block.ghost.E.state.data, ref.ghost["E"])
        assert np.array_equal(block.mesh.x[:, :, 0], ref.geom.xc)


def test_manual_stepping_and_lazy_state_sync():
    config = em_config(nx=16, ny=16)
    sim = Euler2D(config=config, mesh_config=em_mesh())
    sim.apply_initial_condition()
    sim.apply_boundary_condition()
    prob = cases.build_oracle(em_mesh().dict, 16, 16, cases.explosion_ic)
    for _ in range(3):
        dt_ref = prob.get_dt(sim.t, sim.t_final)
        assert sim.get_dt() == dt_ref
        dt = sim.step()
        prob.step(dt_ref)
        assert dt == dt_ref
    for block in sim.blocks:
        assert np.array_equal(block.state.data, prob.blocks[block.global_block_num].U)
        assert np.array_equal(block.dUdt(), prob.residual(prob.blocks[block.global_block_num]))
    # host writes are picked up again
    blk0 = next(iter(sim.blocks))
    blk0.state.data[:, :, 3] *= 1.25
    prob.blocks[blk0.global_block_num].U[:, :, 3] *= 1.25
    sim.apply_boundary_condition()
    prob.apply_bc()
    dt = sim.step()
    prob.step(dt)
    for block in sim.blocks:
        assert np.array_equal(block.state.data, prob.blocks[block.global_block_num].U)


def test_write_solution_layout(tmp_path):
    config = em_config(nx=12, ny=12, t_final=0.005, write_solution=True, write_solution_mode="every_n_timesteps",
                       write_solution_name="explosion_multi", write_solution_base=str(tmp_path), write_every_n_timesteps=3)
    sim = Euler2D(config=config, mesh_config=em_mesh())
    sim.solve()
    prob = cases.build_oracle(em_mesh().dict, 12, 12, cases.explosion_ic)
    base = tmp_path / "explosion_multi"
    assert (base / "mesh" / "mesh_x_blk_0.npy").exists() and np.load(base / "mesh" / "mesh_x_blk_3.npy").shape == (12, 12, 1)
    t, n = 0.0, 0
    written = []
    while t < sim.t_final:
        dt = prob.get_dt(t, sim.t_final)
        prob.step(dt)
        if n % 3 == 0:
            written.append(n)
            for g, b in prob.blocks.items():
                got = np.load(base / str(n) / f"explosion_multi_blk_{g}.npy")
                assert np.array_equal(got, b.U), (n, g)
        t += dt
        n += 1
    assert n == sim.num_time_step and len(written) >= 2
    assert sorted(int(p.name) for p in base.iterdir() if p.name != "mesh") == written


def test_dmr_config_primitive_hlll_rk2():
    from pyhype_b200.mesh.base import QuadMeshGenerator

    class DMRInitialCondition(InitialCondition):
        def apply_to_block(self, block):
            left = PrimitiveState(fluid=block.config.fluid, array=np.array([8, 8.25, 0.0, 116.5]).reshape((1, 1, 4))).to_type(ConservativeState)
            right = PrimitiveState(fluid=block.config.fluid, array=np.array([1.4, 0.0, 0.0, 1.0]).reshape((1, 1, 4))).to_type(ConservativeState)
            block.state.data = np.where(block.mesh.x <= 0.95, left.data, right.data)
            block.state.make_non_dimensional()

    config = em_config(fvm_flux_function_type="HLLL", time_integrator="RK2", initial_condition=DMRInitialCondition(),
                       reconstruction_type=PrimitiveState, CFL=0.4, t_final=0.01, nx=24, ny=24)
    k, a, d = 1, 2 / np.sqrt(3), np.tan(30 * np.pi / 180)
    xs = [0, k, 2 * k, 3 * k, 4 * k]
    mesh = QuadMeshGenerator(
        nx_blk=4, ny_blk=1, BCE=["OutletDirichlet"], BCW=["OutletDirichlet"], BCN=["OutletDirichlet"],
        BCS=["OutletDirichlet", "Slipwall", "Slipwall", "Slipwall"], top_x=xs, bot_x=xs,
        top_y=[a, a, a + d, a + 2 * d, a + 3 * d], bot_y=[0, 0, d, 2 * d, 3 * d],
        left_x=[0, 0], right_x=[4 * k, 4 * k], left_y=[0, a], right_y=[3 * d, a + 3 * d],
    )
    sim = Euler2D(config=config, mesh_config=mesh)
    sim.solve()
    prob, t, dts = oracle_run(cases.dmr_mesh(), 24, 24, cases.dmr_ic, 0.01 * 343.0, flux="HLLL", recon="primitive",
                              integrator="RK2", CFL=0.4)
    assert sim.num_time_step == len(dts) > 3
    for block in sim.blocks:
        assert np.array_equal(block.state.data, prob.blocks[block.global_block_num].U)


def test_wedge_with_primitive_dirichlet_inlet():
    air = Air(a_inf=343.0, rho_inf=1.0)
    ny, nx = 14, 16
    inlet = PrimitiveState(fluid=air, array=np.tile(np.array([1.0, 2.0, 0.0, 1 / 1.4]).reshape(1, 1, 4), (ny, 1, 1)))
    bc = PrimitiveDirichletBC(primitive_state=inlet)
    blocks = cases.wedge_mesh(ny)
    blocks[0]["BCTypeW"] = bc

    class Flood(InitialCondition):
        def apply_to_block(self, block):
            st = PrimitiveState(fluid=block.config.fluid, array=np.array([1.0, 2.0, 0.0, 1 / 1.4]).reshape((1, 1, 4))).to_type(ConservativeState)
            block.state.data = st.data
            block.state.make_non_dimensional()

    config = em_config(fvm_flux_function_type="HLLL", time_integrator="RK2", initial_condition=Flood(),
                       reconstruction_type=PrimitiveState, CFL=0.3, t_final=0.15, nx=nx, ny=ny, fluid=air)
    sim = Euler2D(config=config, mesh_config=blocks)
    sim.solve()
    prob, t, dts = oracle_run(cases.wedge_mesh(ny), nx, ny, cases.wedge_ic, 0.15 * 343.0, flux="HLLL", recon="primitive",
                              integrator="RK2", CFL=0.3)
    assert sim.num_time_step == len(dts) > 5
    for block in sim.blocks:
        assert np.array_equal(block.state.data, prob.blocks[block.global_block_num].U)
    # the examples' (1, 1, 4) inlet state fails the shape check in the reference too
    bad = PrimitiveDirichletBC(primitive_state=PrimitiveState(fluid=air, array=np.array([1.0, 2.0, 0.0, 1 / 1.4]).reshape(1, 1, 4)))
    blocks2 = cases.wedge_mesh(ny)
    blocks2[0]["BCTypeW"] = bad
    with pytest.raises(ValueError, match="equal shape"):
        Euler2D(config=config, mesh_config=blocks2)


def test_unsupported_options_raise():
    with pytest.raises(ValueError):
        Euler2D(config=em_config(fvm_flux_function_type="AUSM"), mesh_config=em_mesh())
    with pytest.raises(ValueError):
        Euler2D(config=em_config(nghost=2), mesh_config=em_mesh())
    with pytest.raises(KeyError):   # quadratures.py:37: _QUADS has keys 1, 2, 3
        Euler2D(config=em_config(fvm_num_quadrature_points=4), mesh_config=em_mesh())
    with pytest.raises(KeyError):
        m = em_mesh().dict
        Euler2D(config=em_config(), mesh_config={k + 1: v for k, v in m.items()})  # 1-based ids, as in the shipped wedge example


def test_facade_two_quadrature_points():
    config = em_config(nx=16, ny=14, t_final=0.003, fvm_num_quadrature_points=2, initial_condition=ExplosionInitialCondition())
    sim = Euler2D(config=config, mesh_config=em_mesh())
    sim.solve()
    prob = cases.build_oracle(em_mesh().dict, 16, 14, cases.explosion_ic, nqp=2)
    t, dts = prob.run(0.0, 0.003 * 343.0)
    assert sim.num_time_step == len(dts) and sim.t == t
    for block in sim.blocks:
        assert np.array_equal(block.state.data, prob.blocks[block.global_block_num].U)


def test_builtin_flood_initial_condition_is_filled_on_the_device():
    """SupersonicFloodInitialCondition assigns a (1, 1, 4) state that the setter broadcasts
    (initial_conditions/supersonic_flood.py:50-59): here it becomes a fill kernel, with the same values."""
    from pyhype_b200.initial_conditions import SupersonicFloodInitialCondition

    def mesh():
        return RectagularMeshGenerator.generate(
            BCE=["OutletDirichlet"], BCW=["OutletDirichlet"], BCN=["OutletDirichlet"], BCS=["OutletDirichlet"],
            east=10.0, west=0.0, north=20.0, south=0.0, n_blocks_horizontal=2, n_blocks_vertical=4)

    fluid = Air(a_inf=343.0, rho_inf=1.0)
    ic = SupersonicFloodInitialCondition(fluid, rho=1.2, u=700.0, v=30.0, p=101325.0)
    sim = Euler2D(config=em_config(nx=14, ny=10, initial_condition=ic, fluid=fluid, t_final=0.002), mesh_config=mesh())
    sim.apply_initial_condition()
    blk = next(iter(sim.blocks))
    assert blk.state._uniform is not None            # nothing materialised, nothing uploaded yet
    uploads = []
    orig = sim._engine.upload
    sim._engine.upload = lambda gid, arr: (uploads.append(gid), orig(gid, arr))[1]
    sim.apply_boundary_condition()
    assert uploads == []
    W = np.array([1.2, 700.0, 30.0, 101325.0]).reshape(1, 1, 4)
    U = cases.prim_to_cons_nd(W)[0, 0]
    for b in sim.blocks:
        got = sim._engine.download(b.global_block_num)
        assert got.shape == (10, 14, 4) and np.all(got == got[0, 0]) and np.array_equal(got[0, 0], U)
        assert np.array_equal(b.state.data, got)      # host view materialises to the same numbers
    sim.solve()
    prob = cases.build_oracle(mesh().dict, 14, 10, lambda x, y: np.broadcast_to(U, x.shape + (4,)).copy())
    t, dts = prob.run(0.0, 0.002 * 343.0)
    assert sim.num_time_step == len(dts) and len(dts) >= 2
    for b in sim.blocks:
        assert np.array_equal(b.state.data, prob.blocks[b.global_block_num].U)
