"""Fluid constants (host-only). Mirrors pyhype/fluids/base.py:27-64."""
from __future__ import annotations

from abc import ABC, abstractmethod
from collections import namedtuple

import numpy as np

_FarField = namedtuple("far_field", ["a", "rho"])


class Fluid(ABC):
    gas_constant = 8314.0
    _molecular_mass = -np.inf

    def __init__(self, a_inf: float = 1.0, rho_inf: float = 1.0):
        self._far_field = _FarField(a_inf, rho_inf)
        self._R = self.gas_constant / self.molecular_mass

    @property
    def far_field(self):
        return self._far_field

    @property
    def R(self):
        return self._R

    @property
    def molecular_mass(self):
        return self._molecular_mass

    @abstractmethod
    def gamma(self, *args, temperature: float = None, **kwargs) -> float:
        raise NotImplementedError

    def g_over_gm1(self, *args, temperature: float = None, **kwargs) -> float:
        g = self.gamma(*args, temperature, **kwargs)
        return g / (g - 1.0)

    def one_over_gm1(self, *args, temperature: float = None, **kwargs) -> float:
        g = self.gamma(*args, temperature, **kwargs)
        return 1.0 / (g - 1.0)
