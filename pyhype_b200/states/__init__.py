"""Host-side state containers used by initial conditions, boundary-condition descriptions and
result inspection.  They mirror the public surface of ``pyhype.states`` (constructor keywords,
``data``, ``to_type``, ``from_state``, ``make_non_dimensional``, the variable properties and the
thermodynamic helpers a user IC may call); the per-step arithmetic the reference does with these
objects (pyhype/states/*.py operator overloads on the hot path) runs on the GPU instead.
"""
from .base import RealizabilityException, State
from .conservative import ConservativeState
from .primitive import PrimitiveState

__all__ = ["State", "PrimitiveState", "ConservativeState", "RealizabilityException"]
