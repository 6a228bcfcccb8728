from .air import Air
from .base import Fluid

__all__ = ["Air", "Fluid"]
