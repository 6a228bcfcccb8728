"""The reference defines PrimitiveDirichletBC twice (boundary_conditions/base.py:35-44 and bc.py:26-35);
both import paths are kept."""
from .base import BoundaryCondition, PrimitiveDirichletBC

__all__ = ["BoundaryCondition", "PrimitiveDirichletBC"]
