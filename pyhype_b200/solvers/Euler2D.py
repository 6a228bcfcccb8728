"""Import path kept from the reference (``from pyhype.solvers.Euler2D import Euler2D``)."""
from .euler2d import Euler2D

__all__ = ["Euler2D"]
