"""State base class (host container). Mirrors pyhype/states/base.py:37-313."""
from __future__ import annotations

import numpy as np

from ..fluids.base import Fluid

_COMPAT = (int, float, np.ndarray)


class RealizabilityException(Exception):
    pass


class State:
    def __init__(self, fluid, state=None, array=None, shape=None, fill=None):
        if not isinstance(fluid, Fluid):
            raise TypeError("fluid must be of type Fluid.")
        self.fluid = fluid
        self._data = None
        self.cache = {}
        if state is not None:
            self._data = np.zeros(state.shape)
            self.from_state(state)
        elif array is not None:
            self.from_array(array)
        elif shape is not None:
            self.data = np.full(shape=shape, fill_value=fill if fill is not None else 0.0)
        else:
            raise ValueError("State constructor must recieve either a state, array or a shape.")

    # -- data ---------------------------------------------------------------------------------------
    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, data):
        self.from_array(data)

    def from_array(self, array):
        if not isinstance(array, np.ndarray):
            raise TypeError(f"Input array must be a Numpy array, but it is a {type(array)}.")
        if array.ndim != 3 or array.shape[-1] != 4:
            raise ValueError("Array must have 3 dims and a depth of 4.")
        if self._data is None or self._data.shape == array.shape:
            self._data = array
        else:
            self._data[:, :, :] = array  # broadcast (states/base.py:99-107)
        self.clear_cache()

    def make_non_dimensional(self):
        ff = self.fluid.far_field
        self.data[:, :, 0] /= ff.rho
        self.data[:, :, 1] /= ff.rho * ff.a
        self.data[:, :, 2] /= ff.rho * ff.a
        self.data[:, :, 3] /= ff.rho * ff.a**2

    @property
    def shape(self):
        return self.data.shape

    def clear_cache(self):
        self.cache.clear()

    def reshape(self, shape):
        self._data = np.reshape(self.data, shape)

    def transpose(self, axes):
        self._data = self.data.transpose(axes)

    def __getitem__(self, index):
        return type(self)(fluid=self.fluid, array=self.data[index].copy())

    # -- conversions (pyhype/states/converter) ----------------------------------------------------------
    def _as_array_of(self, target_type):
        raise NotImplementedError

    def from_state(self, state):
        if not self.shape == state.shape:
            raise ValueError(
                f"States must have equal shape, but state has {self.shape} and from_state has {state.shape}"
            )
        self.data = state._as_array_of(type(self))

    def to_type(self, to_type):
        return to_type(fluid=self.fluid, array=self._as_array_of(to_type))

    # -- arithmetic (host convenience; pyhype/states/base.py:160-245) --------------------------------------
    def _binary(self, other, op, reverse=False):
        if isinstance(other, type(self)):
            o = other.data
        elif isinstance(other, _COMPAT):
            o = other
        else:
            raise TypeError(f"Other must be of type {self.__class__.__name__} or {_COMPAT}")
        return type(self)(fluid=self.fluid, array=op(o, self.data) if reverse else op(self.data, o))

    def __add__(self, other):
        return self._binary(other, np.add)

    __radd__ = __add__

    def __sub__(self, other):
        return self._binary(other, np.subtract)

    def __rsub__(self, other):
        return self._binary(other, np.subtract, reverse=True)

    def __mul__(self, other):
        return self._binary(other, np.multiply)

    __rmul__ = __mul__

    def __truediv__(self, other):
        return self._binary(other, np.divide)

    def __rtruediv__(self, other):
        return self._binary(other, np.divide, reverse=True)

    # -- realizability (pyhype/states/base.py:270-287) ----------------------------------------------------
    def realizability_conditions(self):
        raise NotImplementedError

    def realizable(self):
        conditions = self.realizability_conditions()
        if all(np.all(c) for c in conditions.values()):
            return True
        bad = {n: np.where(np.bitwise_not(g)) for n, g in conditions.items() if not np.all(g)}
        return RealizabilityException(
            f"{self.__class__.__name__} has unrealizable values in the following conditions: {list(bad.values())}"
        )
