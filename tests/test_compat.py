"""-m "not gpu": the ``pyhype`` -> ``pyhype_b200`` import alias (pyhype_b200/compat.py).  Runs in a subprocess so
that the alias can never meet the real reference package, which other tests import from /root/reference."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def run(code, *paths):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join((ROOT,) + paths))
    r = subprocess.run([sys.executable, "-W", "ignore", "-c", textwrap.dedent(code)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_alias_resolves_reference_import_paths_to_this_package():
    out = run("""
        import pyhype_b200.compat as c
        c.install(); c.install()                      # idempotent
        from pyhype.solvers import Euler2D
        from pyhype.solvers.Euler2D import Euler2D as E2
        from pyhype.solver_config import SolverConfig
        from pyhype.states import PrimitiveState, ConservativeState
        from pyhype.states.primitive import PrimitiveState as P2
        from pyhype.states.conservative import ConservativeState as C2
        from pyhype.fluids import Air
        from pyhype.mesh.base import QuadMeshGenerator
        from pyhype.mesh.rectangular import RectagularMeshGenerator
        from pyhype.initial_conditions.base import InitialCondition
        from pyhype.initial_conditions.supersonic_flood import SupersonicFloodInitialCondition
        from pyhype.boundary_conditions.bc import PrimitiveDirichletBC
        from pyhype.boundary_conditions.base import PrimitiveDirichletBC as B2
        import pyhype, pyhype_b200, pyhype_b200.states.primitive as real
        assert pyhype is pyhype_b200 and P2 is real.PrimitiveState is PrimitiveState and C2 is ConservativeState
        assert Euler2D is E2 and Euler2D.__module__.startswith("pyhype_b200.") and B2 is PrimitiveDirichletBC
        for missing in ("pyhype.utils.visualizer", "pyhype.solvers.base"):   # off the hot path / broken in the reference too
            try:
                __import__(missing)
            except ImportError:
                pass
            else:
                raise AssertionError(missing)
        c.uninstall()
        import sys
        assert "pyhype" not in sys.modules and "pyhype.states" not in sys.modules
        print("alias ok")
    """)
    assert "alias ok" in out


def test_alias_refuses_to_shadow_an_imported_package():
    out = run("""
        import sys, types
        sys.modules["pyhype"] = types.ModuleType("pyhype")      # stands for the real reference package
        import pyhype_b200.compat as c
        try:
            c.install()
        except RuntimeError as e:
            print("refused:", e)
    """)
    assert "refused:" in out


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "examples")), reason="reference checkout not present")
def test_reference_example_configs_and_meshes_build_unchanged_through_the_alias():
    """The config.py / mesh.py / initial_condition.py modules of the reference's own examples, imported as they are."""
    out = run("""
        import pyhype_b200.compat as c
        c.install()
        from examples.explosion_multi.config import config as ec
        from examples.explosion_multi.mesh import mesh as em
        from examples.dmr.config import config as dc
        from examples.dmr.mesh import mesh_gen as dm
        from examples.jet.config import config as jc
        from examples.jet.mesh import mesh as jm
        from examples.supersonic_step.config import config as sc
        from examples.supersonic_step.mesh import step_ten_block
        from examples.implosion.config import config as ic
        from examples.implosion.mesh import mesh as im
        from examples.explosion.config import config as xc
        from examples.explosion.mesh import mesh_dict as xm
        for cfg in (ec, dc, jc, sc, ic, xc):
            assert type(cfg).__module__ == "pyhype_b200.solver_config", type(cfg).__module__
            assert hasattr(cfg.initial_condition, "apply_to_block")
        assert (ec.nx, ec.fvm_flux_function_type, ec.time_integrator) == (150, "Roe", "RK4") and len(em.dict) == 8
        assert (dc.fvm_flux_function_type, dc.reconstruction_type.__name__) == ("HLLL", "PrimitiveState") and len(dm.dict) == 4
        assert (jc.nx, jc.ny) == (1080, 60) and len(jm.dict) == 9
        assert type(jm.dict[4]["BCTypeW"]).__name__ == "PrimitiveDirichletBC" and jm.dict[3]["BCTypeW"] == "Slipwall"
        blocks = step_ten_block()
        assert len(blocks) == 10 and blocks[1]["NeighborE"] == 6 and type(blocks[0]["BCTypeW"]).__name__ == "PrimitiveDirichletBC"
        assert list(im) == [0] and list(xm) == [0]
        print("examples ok")
    """, REF)
    assert "examples ok" in out
