"""pyhype/initial_conditions/supersonic_flood.py:33-59"""
import numpy as np

from ..states import ConservativeState, PrimitiveState
from .base import InitialCondition


class SupersonicFloodInitialCondition(InitialCondition):
    def __init__(self, fluid, rho: float, u: float, v: float, p: float):
        if rho <= 0 or p <= 0:
            raise ValueError(f"Unrealizable density (rho={rho}) or pressure (p={p}).")
        self._rho, self._u, self._v, self._p = rho, u, v, p
        a = np.sqrt(fluid.gamma() * p / rho)
        velocity = np.hypot(u, v)
        mach_number = velocity / a
        if mach_number < 1.0:
            raise ValueError(
                "The given set of conditions do not produce a supersonic flow:\n"
                f"Speed of Sound = {a}, Total Velocity = {velocity}, Mach Number = {mach_number}."
            )

    def apply_to_block(self, block):
        state = PrimitiveState(
            fluid=block.config.fluid,
            array=np.array([self._rho, self._u, self._v, self._p]).reshape((1, 1, 4)),
        ).to_type(ConservativeState)
        block.state.data = state.data
        block.state.make_non_dimensional()
