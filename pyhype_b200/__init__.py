"""pyhype_b200 -- B200-native engine for pyHype's second-order MUSCL residual + explicit-RK time
march, behind pyHype's own Python API (``SolverConfig``, ``Euler2D(config, mesh).solve()``, the
mesh generators, ``InitialCondition``, ``PrimitiveDirichletBC``).

Importing the package never touches the GPU; the first engine call loads
``pyhype_b200/lib/libpyh_b200.so`` (hand-written sm_100a CUDA behind ``include/pyh_b200.h``) and
raises if it is missing -- there is no CPU fallback.
"""
from . import solvers  # noqa: F401

__version__ = "0.1.0"
