from .euler2d import Euler2D

__all__ = ["Euler2D"]
