"""Primitive state [rho, u, v, p] (host container). Mirrors pyhype/states/primitive.py:34-244."""
from __future__ import annotations

import numpy as np

from .base import State


class PrimitiveState(State):
    RHO_IDX, U_IDX, V_IDX, P_IDX = 0, 1, 2, 3

    rho = property(lambda s: s.data[:, :, 0], lambda s, x: s.data.__setitem__((slice(None), slice(None), 0), x))
    u = property(lambda s: s.data[:, :, 1], lambda s, x: s.data.__setitem__((slice(None), slice(None), 1), x))
    v = property(lambda s: s.data[:, :, 2], lambda s, x: s.data.__setitem__((slice(None), slice(None), 2), x))
    p = property(lambda s: s.data[:, :, 3], lambda s, x: s.data.__setitem__((slice(None), slice(None), 3), x))

    def ek(self):
        return 0.5 * self.rho * (self.u * self.u + self.v * self.v)  # ek_JIT, primitive.py:93-104

    def Ek(self):
        return 0.5 * (self.u * self.u + self.v * self.v)

    def H(self):
        return self.fluid.g_over_gm1() * self.p / self.rho + self.Ek()

    def a(self):
        return np.sqrt(self.fluid.gamma() * self.p / self.rho)

    def e(self):
        return self.fluid.one_over_gm1() * self.p + self.ek()

    def V(self):
        return np.sqrt(self.u * self.u + self.v * self.v)

    def Ma(self):
        return self.V() / self.a()

    def _as_array_of(self, target_type):
        from .conservative import ConservativeState

        if target_type is PrimitiveState or issubclass(target_type, PrimitiveState):
            return self.data.copy()
        if target_type is ConservativeState or issubclass(target_type, ConservativeState):
            # PrimitiveConverter.to_conservative (states/converter/concrete_defs.py:127-141)
            return np.dstack((
                self.rho.copy(), self.rho * self.u, self.rho * self.v,
                self.p / (self.fluid.gamma() - 1) + self.ek(),
            ))
        raise TypeError(f"cannot convert PrimitiveState to {target_type}")

    def realizability_conditions(self):
        return dict(rho_good=self.rho > 0, pressure_good=self.p > 0)
