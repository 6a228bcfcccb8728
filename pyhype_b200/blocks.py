"""Host-side views of the solution blocks.

The reference's ``QuadBlock`` (pyhype/blocks/quad_block.py:243-574) owns the state, four ghost
blocks, a reconstruction block and ten FVM objects; here the block that user code sees (initial
conditions, output, inspection) is a thin view: geometry on the host, the state mirrored lazily
from the GPU.  All per-step work lives behind the C ABI (``pyhype_b200.engine.Engine``).
"""
from __future__ import annotations

import numpy as np

from .mesh.quad_mesh import QuadMesh
from .states import ConservativeState

SIDES = ("E", "W", "N", "S")


class DeviceBackedState(ConservativeState):
    """``block.state``: a ConservativeState whose ``data`` is synchronised with the device copy
    on demand -- downloaded when the device is newer, re-uploaded before the next device call
    whenever host code may have written to it (a getter hands out a writable array)."""

    def __init__(self, fluid, shape, sync):
        self._sync = sync
        self._device_newer = False
        self._host_touched = True
        self._uniform = None   # pending (4,) state assigned as a (1, 1, 4) array: filled on the device, no upload
        super().__init__(fluid=fluid, shape=shape)

    def _materialize_uniform(self):
        if self._uniform is not None:
            self._data[:, :, :] = self._uniform.reshape(1, 1, 4)   # the reference's broadcast (states/base.py:99-107)
            self._uniform = None

    @property
    def data(self):
        if self._device_newer:
            self._data = self._sync.download()
            self._device_newer = False
            self._uniform = None
        self._materialize_uniform()
        self._host_touched = True
        return self._data

    @data.setter
    def data(self, array):
        self._device_newer = False
        self._host_touched = True
        self.from_array(array)

    def from_array(self, array):
        if not isinstance(array, np.ndarray):
            raise TypeError(f"Input array must be a Numpy array, but it is a {type(array)}.")
        if array.ndim != 3 or array.shape[-1] != 4:
            raise ValueError("Array must have 3 dims and a depth of 4.")
        self._uniform = None
        if self._data is None or self._data.shape == array.shape:
            self._data = array
        elif array.shape == (1, 1, 4):
            # built-in flood initial conditions (initial_conditions/supersonic_flood.py:50-59): keep the 4 numbers,
            # broadcast lazily on the host and by a fill kernel on the device
            self._uniform = np.array(array, dtype=np.float64).reshape(4)
        else:
            self._data[:, :, :] = array
        self.cache.clear()

    def make_non_dimensional(self):
        if self._uniform is not None and not self._device_newer:
            ff = self.fluid.far_field   # same divisions, on the 4 numbers instead of every cell (states/base.py:93-97)
            self._uniform[0] /= ff.rho
            self._uniform[1] /= ff.rho * ff.a
            self._uniform[2] /= ff.rho * ff.a
            self._uniform[3] /= ff.rho * ff.a**2
            return
        super().make_non_dimensional()

    def push_if_touched(self):
        if self._host_touched and not self._device_newer:
            if self._uniform is not None:
                self._sync.fill_uniform(self._uniform)
            else:
                self._sync.upload(np.ascontiguousarray(self._data, dtype=np.float64))
        self._host_touched = False

    def mark_device_newer(self):
        self._device_newer = True
        self._host_touched = False


class _Sync:
    def __init__(self, engine, gid):
        self.engine, self.gid = engine, gid

    def upload(self, arr):
        self.engine.upload(self.gid, arr)

    def download(self):
        return self.engine.download(self.gid)

    def fill_uniform(self, state):
        self.engine.fill_uniform(self.gid, state)


class _GhostView:
    """``block.ghost.E.state.data`` etc.: conservative ghost strips fetched from the device."""

    class _Strip:
        def __init__(self, block, side):
            self._block, self._side = block, side

        @property
        def state(self):
            return self

        @property
        def data(self):
            self._block._solver._flush_host_states()
            self._block._solver._wait_halo()   # an overlapped remote exchange may still be in flight
            return self._block._engine.download_ghost(self._block.global_block_num, self._side)

    def __init__(self, block):
        for s in SIDES:
            setattr(self, s, self._Strip(block, s))


class BlockInfo:
    def __init__(self, blk):
        self.nBLK = blk["nBLK"]
        self.neighbors = {s: blk["Neighbor" + s] for s in SIDES}
        self.bc = {s: blk["BCType" + s] for s in SIDES}


class QuadBlock:
    def __init__(self, config, blk_input, solver):
        self.config = config
        self.block_data = blk_input
        self.info = BlockInfo(blk_input)
        self.global_block_num = blk_input["nBLK"]
        self.mesh = QuadMesh(config.nx, config.ny, NE=blk_input["NE"], NW=blk_input["NW"], SE=blk_input["SE"],
                             SW=blk_input["SW"], nghost=config.nghost)
        self.is_cartesian = self.mesh.is_cartesian
        self.neighbors = self.info.neighbors
        self._solver = solver
        self._engine = None
        self.state = None
        self.ghost = _GhostView(self)

    def _attach(self, engine):
        self._engine = engine
        self.state = DeviceBackedState(self.config.fluid, (self.mesh.ny, self.mesh.nx, 4), _Sync(engine, self.global_block_num))

    @property
    def reconstruction_type(self):
        return self.config.reconstruction_type

    def dUdt(self):
        """Residual of this block at the current state (test hook; pyhype/blocks/quad_block.py:512-521)."""
        self._solver._flush_host_states()
        return self._engine.residual(self.global_block_num)

    def realizable(self):
        return self.state.realizable()
