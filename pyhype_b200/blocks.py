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
    whenever host code may have written to it (a getter hands out a writable array).

    Contract (differs from the reference, where ``state.data`` IS the solution): every READ of ``.data`` hands out the
    current host array and marks it as possibly modified, so it is uploaded again before the next device call; an array
    obtained earlier and kept across a device call (``a = block.state.data; solver.step(); a[...] = 0``) is a stale copy
    -- the write is not seen by the device and the next read of ``.data`` returns a fresh download.  Re-read ``.data``
    after every solver call, or assign through the setter (``block.state.data = array``), which always takes effect."""

    def __init__(self, fluid, shape, sync):
        self._sync = sync
        self._device_newer = False
        self._host_touched = True
        self._uniform = None   # pending (4,) state assigned as a (1, 1, 4) array: filled on the device, no upload
        self._boxes = []       # pending two-state fills (fill_box) on top of it, in order; also evaluated on the device
        super().__init__(fluid=fluid, shape=shape)

    def _materialize_uniform(self):
        if self._uniform is not None:
            self._data[:, :, :] = self._uniform.reshape(1, 1, 4)   # the reference's broadcast (states/base.py:99-107)
            self._uniform = None
        if self._boxes:
            x, y = self._sync.centroids()
            for x0, x1, y0, y1, inside, outside in self._boxes:   # the reference's np.where over the centroids
                cond = np.logical_and(np.logical_and(x >= x0, x <= x1), np.logical_and(y >= y0, y <= y1))
                self._data[:, :, :] = np.where(cond, inside.reshape(1, 1, 4), self._data if outside is None else outside.reshape(1, 1, 4))
            self._boxes = []

    def fill_box(self, x0, x1, y0, y1, inside, outside=None):
        """Two-state fill over the cell centroids, the pattern of the reference's explosion / implosion / dmr initial
        conditions (``np.where(cond(block.mesh.x, block.mesh.y), left_state.data, right_state.data)``,
        examples/explosion_multi/initial_condition.py:53-59): cells with x0 <= x <= x1 and y0 <= y <= y1 get ``inside``
        (a 4-vector or (1, 1, 4) array), the others ``outside`` (None: keep).  Kept as 2 x 4 numbers and evaluated by a fill
        kernel on the device's bit-identical centroids -- no (ny, nx, 4) array is built or uploaded unless host code reads
        ``data``.  ``make_non_dimensional`` acts on the pending numbers."""
        if self._device_newer:
            self.data  # pull the device copy first: the fill applies on top of it
        inside = np.array(inside, dtype=np.float64).reshape(4)
        outside = None if outside is None else np.array(outside, dtype=np.float64).reshape(4)
        if outside is not None:
            self._uniform, self._boxes = None, []   # the whole block is redefined
        self._boxes.append((float(x0), float(x1), float(y0), float(y1), inside, outside))
        self._host_touched = True
        self.cache.clear()

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

    def _pending_defines_block(self):
        """True when the pending fills alone determine every cell (uniform flood, or a box fill with an outside state)."""
        return self._uniform is not None or (bool(self._boxes) and self._boxes[0][5] is not None)

    def make_non_dimensional(self):
        if self._pending_defines_block() and not self._device_newer:
            ff = self.fluid.far_field   # same divisions, on the 4 numbers instead of every cell (states/base.py:93-97)
            vecs = ([self._uniform] if self._uniform is not None else []) + [v for b in self._boxes for v in (b[4], b[5]) if v is not None]
            for v in vecs:
                v[0] /= ff.rho
                v[1] /= ff.rho * ff.a
                v[2] /= ff.rho * ff.a
                v[3] /= ff.rho * ff.a**2
            return
        self.data   # materialise pending fills on the host, then the reference's elementwise division
        super().make_non_dimensional()

    def push_if_touched(self):
        if self._host_touched and not self._device_newer:
            if self._pending_defines_block():
                if self._uniform is not None:
                    self._sync.fill_uniform(self._uniform)
                for x0, x1, y0, y1, inside, outside in self._boxes:
                    self._sync.fill_box(x0, x1, y0, y1, inside, outside)
                # the host copy is now stale; it is re-created (download) on the next read of .data
                self._uniform, self._boxes = None, []
                self._device_newer = True
            else:
                self._materialize_uniform()
                self._sync.upload(np.ascontiguousarray(self._data, dtype=np.float64))
        self._host_touched = False

    def mark_device_newer(self):
        self._device_newer = True
        self._host_touched = False


class _Sync:
    def __init__(self, engine, gid, mesh=None):
        self.engine, self.gid, self.mesh = engine, gid, mesh

    def upload(self, arr):
        self.engine.upload(self.gid, arr)

    def download(self):
        return self.engine.download(self.gid)

    def fill_uniform(self, state):
        self.engine.fill_uniform(self.gid, state)

    def fill_box(self, x0, x1, y0, y1, inside, outside):
        self.engine.fill_box(self.gid, x0, x1, y0, y1, inside, outside)

    def centroids(self):
        return self.mesh.x, self.mesh.y


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
        self.state = DeviceBackedState(self.config.fluid, (self.mesh.ny, self.mesh.nx, 4), _Sync(engine, self.global_block_num, self.mesh))

    @property
    def reconstruction_type(self):
        return self.config.reconstruction_type

    def dUdt(self):
        """Residual of this block at the current state (test hook; pyhype/blocks/quad_block.py:512-521)."""
        self._solver._flush_host_states()
        return self._engine.residual(self.global_block_num)

    def realizable(self):
        return self.state.realizable()
