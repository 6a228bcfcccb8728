"""Two-state initial conditions of the reference's examples as ONE class evaluated on the device.

examples/explosion/initial_condition.py:35-60, examples/explosion_multi, examples/implosion and examples/dmr/
initial_condition.py:36-59 all build a left and a right ``PrimitiveState``, convert both to ``ConservativeState`` and assign
``np.where(condition(block.mesh.x, block.mesh.y), left, right)`` followed by ``make_non_dimensional()``.  Written with
numpy on ``block.state.data`` (as a user of the reference would) that costs a (ny, nx, 4) host array and an upload per
block -- 1 GiB per GPU at the weak-scaling size.  ``BoxInitialCondition`` states the same fill as data; ``block.state.fill_box``
keeps the 2 x 4 numbers and lets ``pyh_fill_box`` evaluate the condition on the device's bit-identical centroids."""
import numpy as np

from ..states import ConservativeState, PrimitiveState
from .base import InitialCondition


class BoxInitialCondition(InitialCondition):
    """``inside`` / ``outside``: dimensional primitive states (rho, u, v, p); a cell gets ``inside`` when its centroid satisfies
    x0 <= x <= x1 and y0 <= y <= y1 (closed box, bounds default to -inf / +inf)."""

    def __init__(self, inside, outside, x0=-np.inf, x1=np.inf, y0=-np.inf, y1=np.inf):
        self.inside, self.outside = tuple(inside), tuple(outside)
        self.box = (x0, x1, y0, y1)

    def apply_to_block(self, block):
        fluid = block.config.fluid
        left = PrimitiveState(fluid=fluid, array=np.array(self.inside, dtype=float).reshape((1, 1, 4))).to_type(ConservativeState)
        right = PrimitiveState(fluid=fluid, array=np.array(self.outside, dtype=float).reshape((1, 1, 4))).to_type(ConservativeState)
        block.state.fill_box(*self.box, left.data, right.data)
        block.state.make_non_dimensional()
