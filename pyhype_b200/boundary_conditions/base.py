"""Boundary-condition descriptions. Mirrors pyhype/boundary_conditions/base.py:26-44.

On the GPU path a ``PrimitiveDirichletBC`` is not *called* per stage; its (non-dimensionalised)
primitive inlet state is handed to the engine once (``pyh_block_desc.dirichlet_prim``), which fills
the ghost strip and the ghost-side Riemann state from it every stage.  ``__call__`` is kept so
host code written against the reference keeps working."""
from __future__ import annotations

from abc import ABC, abstractmethod

from ..states import PrimitiveState


class BoundaryCondition(ABC):
    def __call__(self, state, *args, **kwargs):
        self._apply_boundary_condition(state, *args, **kwargs)

    @abstractmethod
    def _apply_boundary_condition(self, state, *args, **kwargs):
        raise NotImplementedError


class PrimitiveDirichletBC(BoundaryCondition):
    def __init__(self, primitive_state: PrimitiveState):
        if not isinstance(primitive_state, PrimitiveState):
            raise TypeError("primitive_array must be a PrimitiveState.")
        super().__init__()
        self._primitive_state = primitive_state
        self._primitive_state.make_non_dimensional()

    @property
    def primitive_state(self):
        return self._primitive_state

    def _apply_boundary_condition(self, state, *args, **kwargs):
        state.from_state(self._primitive_state)
