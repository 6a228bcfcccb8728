from .base import BoundaryCondition, PrimitiveDirichletBC

__all__ = ["BoundaryCondition", "PrimitiveDirichletBC"]
