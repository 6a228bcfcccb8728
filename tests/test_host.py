"""-m "not gpu": host-side logic of the facade (configuration, state containers, mesh generators,
tableaux, block distribution, RK partial-sum plan)."""
import numpy as np
import pytest

import cases
import golden_io
from pyhype_b200.distributed import distribute_blocks, exchange_plan
from pyhype_b200.fluids import Air
from pyhype_b200.mesh.quad_mesh import QuadMesh
from pyhype_b200.solver_config import SolverConfig
from pyhype_b200.states import ConservativeState, PrimitiveState
from pyhype_b200.time_marching import TABLEAUX, get_tableau


def test_solver_config_fields_and_defaults():
    air = Air(a_inf=343.0, rho_inf=1.0)
    c = SolverConfig(nx=10, ny=12, CFL=0.7, t_final=0.07, initial_condition=None, fvm_type="MUSCL",
                     time_integrator="RK4", fvm_gradient_type="GreenGauss", fvm_flux_function_type="Roe",
                     fvm_slope_limiter_type="Venkatakrishnan", fvm_spatial_order=2, fvm_num_quadrature_points=1, fluid=air)
    assert c.n == 120 and c.nghost == 1 and c.use_JIT is True and c.reconstruction_type is ConservativeState
    assert c.interface_interpolation == "arithmetic_average" and c.show_log_for_procs == [0]
    assert c.write_solution is False and c.write_every_n_timesteps == 40
    with pytest.raises(AttributeError):
        c.alpha = 0.5  # __slots__, like the reference (so Generic2/Generic3 cannot work there either)
    assert air.gamma() == 1.4 and air.one_over_gm1() == 1.0 / (1.4 - 1.0) and air.g_over_gm1() == 1.4 / (1.4 - 1.0)


def test_state_containers_convert_like_the_reference():
    air = Air(a_inf=343.0, rho_inf=1.0)
    W = PrimitiveState(fluid=air, array=np.array([4.6968, 3.0, -2.0, 404400.0]).reshape(1, 1, 4))
    U = W.to_type(ConservativeState)
    rho, u, v, p = 4.6968, 3.0, -2.0, 404400.0
    assert U.data[0, 0].tolist() == [rho, rho * u, rho * v, p / (1.4 - 1) + 0.5 * rho * (u * u + v * v)]
    back = U.to_type(PrimitiveState)
    assert back.data[0, 0, 0] == rho and back.data[0, 0, 1] == (rho * u) / rho
    U.make_non_dimensional()
    assert U.data[0, 0, 3] == (p / (1.4 - 1) + 0.5 * rho * (u * u + v * v)) / (1.0 * 343.0**2)
    with pytest.raises(ValueError):
        ConservativeState(fluid=air, array=np.zeros((3, 4)))
    with pytest.raises(TypeError):
        ConservativeState(fluid="air", shape=(1, 1, 4))
    big = ConservativeState(fluid=air, shape=(3, 2, 4))
    big.data = U.data  # (1,1,4) broadcasts into an existing (3,2,4) state (states/base.py:99-107)
    assert big.data.shape == (3, 2, 4) and np.all(big.data[2, 1] == U.data[0, 0])
    assert U.realizable() is True


def test_tableaux_and_factory_errors():
    assert TABLEAUX["RK4"][3] == [1 / 6, 1 / 3, 1 / 3, 1 / 6] and TABLEAUX["RK2"] == [[0.5], [0, 1]]
    with pytest.raises(ValueError, match="is not available"):
        get_tableau("LeapFrog")
    with pytest.raises(AttributeError):
        get_tableau("Generic2")


def test_block_distribution_rule():
    assert distribute_blocks(8, 3) == {0: 0, 1: 0, 2: 0, 3: 1, 4: 1, 5: 1, 6: 2, 7: 2}
    assert distribute_blocks(64, 8)[63] == 7 and distribute_blocks(4, 8) == {0: 0, 1: 1, 2: 2, 3: 3}


def test_exchange_plan_pairs_up_across_ranks():
    blocks = cases.em_mesh()  # 2 wide x 4 high
    owner = distribute_blocks(8, 4)
    plans = {}
    for r in range(4):
        mine = [g for g, o in owner.items() if o == r]
        slots, off = [], 0
        for g in sorted(mine):
            for s in ("E", "W", "N", "S"):
                nb = blocks[g]["Neighbor" + s]
                if nb is not None and owner[nb] != r and blocks[g]["BCType" + s] is None:
                    slots.append(dict(gid=g, side=s, nbr=nb, offset=off, length=40))
                    off += 40
        plans[r] = exchange_plan(slots, owner, r)
    for a in range(4):
        for b in range(4):
            sends = [m[1] for m in plans[a][0] if m[0] == b]
            recvs = [m[1] for m in plans[b][1] if m[0] == a]
            assert sends == recvs


def test_mesh_generators_match_reference_vertices():
    for name in ("em_roe_venkat_cons_rk4", "dmr_hlll_venkat_prim_rk2"):
        fx = golden_io.Fixture(name)
        mine = cases.em_mesh() if name.startswith("em") else cases.dmr_mesh()
        for g in fx.gids:
            for k in ("SW", "SE", "NW", "NE"):
                assert [float(v) for v in mine[g][k]] == fx.blocks[g][k]
            for s in ("E", "W", "N", "S"):
                assert mine[g]["Neighbor" + s] == fx.blocks[g]["Neighbor" + s]
                assert mine[g]["BCType" + s] == fx.blocks[g]["BCType" + s]
    em = cases.em_mesh()
    assert em[3]["SW"][0] == 4.999999999999999  # transfinite rounding: no block of explosion_multi is "cartesian"
    assert not any(QuadMesh(4, 4, NE=b["NE"], NW=b["NW"], SE=b["SE"], SW=b["SW"]).is_cartesian for b in em.values())


def test_quad_mesh_geometry_matches_reference():
    for name in ("em_roe_venkat_cons_rk4", "dmr_hlll_venkat_prim_rk2", "wedge_hlll_prim_rk2", "cart_roe_cons_rk4"):
        fx = golden_io.Fixture(name)
        for g in fx.gids:
            b = fx.blocks[g]
            m = QuadMesh(fx.nx, fx.ny, NE=b["NE"], NW=b["NW"], SE=b["SE"], SW=b["SW"])
            assert np.array_equal(m.x[:, :, 0], fx[f"xc_{g}"]) and np.array_equal(m.y[:, :, 0], fx[f"yc_{g}"])
            assert np.array_equal(m.area, fx[f"A_{g}"])
            assert np.array_equal(m.theta_v[:, 1:], fx[f"thetaE_{g}"]) and np.array_equal(m.theta_h[1:], fx[f"thetaN_{g}"])
    assert golden_io.Fixture("cart_roe_cons_rk4").blocks[0]["NE"] == [10.0, 20.0]


def test_facade_validation_without_gpu():
    from pyhype_b200.solvers.euler2d import _validate

    air = Air(a_inf=343.0, rho_inf=1.0)

    def cfg(**over):
        kw = dict(nx=8, ny=8, CFL=0.7, t_final=0.07, initial_condition=None, fvm_type="MUSCL", time_integrator="RK4",
                  fvm_gradient_type="GreenGauss", fvm_flux_function_type="Roe", fvm_slope_limiter_type="Venkatakrishnan",
                  fvm_spatial_order=2, fvm_num_quadrature_points=1, fluid=air)
        kw.update(over)
        return SolverConfig(**kw)

    _validate(cfg())
    for bad in (dict(fvm_type="WENO"), dict(fvm_spatial_order=1), dict(nghost=2), dict(fvm_gradient_type="LeastSquares"),
                dict(fvm_flux_function_type="AUSM"), dict(fvm_slope_limiter_type="Minmod"),
                dict(interface_interpolation="harmonic")):
        with pytest.raises(ValueError):
            _validate(cfg(**bad))


def test_async_solution_writer_orders_backpressures_and_reports_errors(tmp_path):
    """The every_n_timesteps writer (solvers/euler2d.py) with a stand-in engine: dumps land in submission
    order with the data of their own submission, a third dump waits for a free buffer set, and an I/O
    error surfaces on the solver thread."""
    import threading

    from pyhype_b200.solvers.euler2d import _AsyncSolutionWriter

    class FakeEngine:
        def __init__(self):
            self.value = 0.0
            self.synced = threading.Event()

        def pinned_state_buffer(self):
            return np.zeros((3, 2, 4))

        def download_async(self, gid, out):
            out[...] = self.value + gid   # "device state" at submission time

        def downloads_sync(self):
            self.synced.set()

    eng = FakeEngine()
    w = _AsyncSolutionWriter(eng, [0, 1], depth=2)
    for step in range(5):
        eng.value = 10.0 * step
        w.submit({g: str(tmp_path / f"s{step}_b{g}.npy") for g in (0, 1)})
    w.close()
    assert eng.synced.is_set()
    for step in range(5):
        for g in (0, 1):
            assert np.all(np.load(tmp_path / f"s{step}_b{g}.npy") == 10.0 * step + g)

    w = _AsyncSolutionWriter(eng, [0], depth=2)
    w.submit({0: str(tmp_path / "no_such_dir" / "x.npy")})
    with pytest.raises(OSError):
        w.close()
