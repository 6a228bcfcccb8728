"""Conservative state [rho, rho*u, rho*v, e] (host container). Mirrors pyhype/states/conservative.py:32-165."""
from __future__ import annotations

import numpy as np

from .base import State


class ConservativeState(State):
    RHO_IDX, RHOU_IDX, RHOV_IDX, E_IDX = 0, 1, 2, 3

    rho = property(lambda s: s.data[:, :, 0], lambda s, x: s.data.__setitem__((slice(None), slice(None), 0), x))
    rhou = property(lambda s: s.data[:, :, 1], lambda s, x: s.data.__setitem__((slice(None), slice(None), 1), x))
    rhov = property(lambda s: s.data[:, :, 2], lambda s, x: s.data.__setitem__((slice(None), slice(None), 2), x))
    e = property(lambda s: s.data[:, :, 3], lambda s, x: s.data.__setitem__((slice(None), slice(None), 3), x))

    @property
    def u(self):
        return self.rhou / self.rho

    @property
    def v(self):
        return self.rhov / self.rho

    def Ek(self):
        return 0.5 * (self.u * self.u + self.v * self.v)

    def ek(self):
        return self.rho * self.Ek()

    @property
    def p(self):
        return (self.fluid.gamma() - 1) * (self.e - self.ek())

    def h(self):
        _ek = self.ek()
        return self.fluid.gamma() * (self.e - _ek) + _ek

    def H(self):
        return self.h() / self.rho

    def a(self):
        return np.sqrt(self.fluid.gamma() * self.p / self.rho)

    def V(self):
        return np.sqrt(self.u**2 + self.v**2)

    def Ma(self):
        return self.V() / self.a()

    def _as_array_of(self, target_type):
        from .primitive import PrimitiveState

        if target_type is ConservativeState or issubclass(target_type, ConservativeState):
            return self.data.copy()
        if target_type is PrimitiveState or issubclass(target_type, PrimitiveState):
            # ConservativeConverter.to_primitive (states/converter/concrete_defs.py:85-100)
            return np.dstack((self.rho.copy(), self.u, self.v, (self.fluid.gamma() - 1) * (self.e - self.ek())))
        raise TypeError(f"cannot convert ConservativeState to {target_type}")

    def realizability_conditions(self):
        return dict(rho_good=self.rho > 0, energy_good=self.e > 0)
