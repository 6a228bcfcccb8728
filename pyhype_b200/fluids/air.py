"""pyhype/fluids/air.py:24-28"""
from .base import Fluid


class Air(Fluid):
    _molecular_mass = 28.97

    def gamma(self, *args, temperature: float = None, **kwargs) -> float:
        return 1.4
