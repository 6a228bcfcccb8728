"""Thin Python owner of one ``pyh_ctx`` (one GPU).  Host-side plumbing only: every number is
produced by the CUDA library behind ``include/pyh_b200.h``."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import PYH_MAX_STAGES, PyhBlockDesc, PyhConfig, c_double_p, check

FLUX_IDS = {"Roe": 0, "HLLE": 1, "HLLL": 2}
LIMITER_IDS = {"Venkatakrishnan": 0, "VanLeer": 1, "VanAlbada": 2, "BarthJespersen": 3}
RECON_IDS = {"conservative": 0, "primitive": 1}
BC_IDS = {None: 0, "Reflection": 1, "Slipwall": 2, "OutletDirichlet": 3}
BC_DIRICHLET = 4
SIDES = ("E", "W", "N", "S")  # SidePropertyDict order (pyhype/utils/utils.py:229-237)


def _dp(a: np.ndarray):
    return a.ctypes.data_as(c_double_p)


def _c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


class Engine:
    def __init__(self, nx, ny, flux, limiter, recon, tableau, gamma, cfl, device=0, num_quadrature_points=1):
        self.lib = _lib.load()
        if flux not in FLUX_IDS:
            raise ValueError("Flux function type not specified.")
        if limiter not in LIMITER_IDS:
            raise ValueError("Slope limiter type not specified.")
        cfg = PyhConfig()
        cfg.abi_version = _lib.PYH_ABI_VERSION
        cfg.device = int(device)
        cfg.nx, cfg.ny = int(nx), int(ny)
        cfg.flux = FLUX_IDS[flux]
        cfg.limiter = LIMITER_IDS[limiter]
        cfg.recon = RECON_IDS[recon]
        cfg.num_quadrature_points = int(num_quadrature_points)
        cfg.num_stages = len(tableau)
        if not 1 <= len(tableau) <= PYH_MAX_STAGES:
            raise ValueError(f"Butcher tableau must have 1..{PYH_MAX_STAGES} stages")
        for s, row in enumerate(tableau):
            for k, a in enumerate(row):
                cfg.tableau[s * PYH_MAX_STAGES + k] = float(a)
        cfg.gamma = float(gamma)
        cfg.cfl = float(cfl)
        self.nx, self.ny, self.num_stages = int(nx), int(ny), len(tableau)
        self.device = int(device)
        self._ctx = C.c_void_p()
        check(self.lib.pyh_create(C.byref(cfg), C.byref(self._ctx)))
        self.gids = []
        self._pinned = []

    # -- construction ---------------------------------------------------------------------------
    def add_block(self, gid, mesh, neighbors, bcs, local_gids=None, is_cartesian=None):
        """``mesh``: pyhype_b200.mesh.quad_mesh.QuadMesh; ``neighbors``/``bcs``: dicts keyed E,W,N,S.
        A bc value is None, a string, or an (edge_len, 4) non-dimensional primitive array."""
        d = PyhBlockDesc()
        d.gid = int(gid)
        d.is_cartesian = int(mesh.is_cartesian if is_cartesian is None else is_cartesian)
        keep = []
        for s, side in enumerate(SIDES):
            nb = neighbors.get(side)
            d.neighbor[s] = -1 if nb is None else int(nb)
            d.neighbor_is_local[s] = int(nb is not None and (local_gids is None or nb in local_gids))
            bc = bcs.get(side)
            if isinstance(bc, np.ndarray):
                n = self.ny if side in ("E", "W") else self.nx
                arr = _c(bc).reshape(-1, 4)
                if arr.shape[0] != n:
                    raise ValueError(
                        f"States must have equal shape, but the ghost strip has {n} cells and the inlet state {arr.shape[0]}"
                    )
                keep.append(arr)
                d.bc[s] = BC_DIRICHLET
                d.dirichlet_prim[s] = _dp(arr)
            elif bc in BC_IDS:
                d.bc[s] = BC_IDS[bc]
            else:
                raise ValueError("Boundary Condition type " + str(bc) + " has not been specialized.")
        arrs = dict(
            nodes_x=_c(mesh.nodes_x), nodes_y=_c(mesh.nodes_y), area=_c(mesh.area),
            cos_v=_c(mesh.cos_v), sin_v=_c(mesh.sin_v), cos_h=_c(mesh.cos_h), sin_h=_c(mesh.sin_h),
        )
        assert arrs["nodes_x"].shape == (self.ny + 1, self.nx + 1)
        assert arrs["area"].shape == (self.ny, self.nx)
        assert arrs["cos_v"].shape == (self.ny, self.nx + 1) and arrs["cos_h"].shape == (self.ny + 1, self.nx)
        for k, a in arrs.items():
            setattr(d, k, _dp(a))
        check(self.lib.pyh_add_block(self._ctx, C.byref(d)))
        self.gids.append(int(gid))

    def finalize(self):
        check(self.lib.pyh_finalize(self._ctx))

    def close(self):
        if self._ctx:
            self.lib.pyh_destroy(self._ctx)
            self._ctx = C.c_void_p()
            for p in self._pinned:
                self.lib.pyh_host_free(p)
            self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state ----------------------------------------------------------------------------------------
    def upload(self, gid, aos):
        a = _c(aos)
        if a.shape != (self.ny, self.nx, 4):
            raise ValueError(f"state must have shape {(self.ny, self.nx, 4)}, got {a.shape}")
        check(self.lib.pyh_upload_state(self._ctx, int(gid), _dp(a)))

    def fill_uniform(self, gid, state):
        """Set every interior cell of block ``gid`` to the conservative 4-vector ``state`` (no upload)."""
        v = _c(np.asarray(state, dtype=np.float64).reshape(4))
        check(self.lib.pyh_fill_uniform(self._ctx, int(gid), _dp(v)))

    def fill_box(self, gid, x0, x1, y0, y1, inside, outside=None):
        """Cells of block ``gid`` whose centroid lies in the closed box [x0, x1] x [y0, y1] become the conservative 4-vector
        ``inside``, the others ``outside`` (None: left as they are).  No upload: evaluated on the device's centroids."""
        vi = _c(np.asarray(inside, dtype=np.float64).reshape(4))
        vo = None if outside is None else _c(np.asarray(outside, dtype=np.float64).reshape(4))
        check(self.lib.pyh_fill_box(self._ctx, int(gid), float(x0), float(x1), float(y0), float(y1), _dp(vi),
                                    None if vo is None else _dp(vo)))

    def download(self, gid, out=None):
        """Conservative state of block ``gid`` as (ny, nx, 4); ``out`` may be a caller-owned
        (e.g. pinned) C-contiguous float64 array to receive it without an extra copy."""
        if out is None:
            out = np.empty((self.ny, self.nx, 4))
        elif out.shape != (self.ny, self.nx, 4) or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array of shape (ny, nx, 4)")
        check(self.lib.pyh_download_state(self._ctx, int(gid), _dp(out)))
        return out

    # -- asynchronous streaming (copy streams overlap the compute stream) ----------------------------
    def _check_stream_buf(self, a):
        if a.shape != (self.ny, self.nx, 4) or a.dtype != np.float64 or not a.flags.c_contiguous:
            raise ValueError("buffer must be a C-contiguous float64 array of shape (ny, nx, 4)")

    def upload_async(self, gid, aos):
        """Enqueue the H2D copy of a (page-locked) host state; the buffer must stay untouched until
        ``transfers_sync``.  Takes effect at the next ``commit_uploads``."""
        self._check_stream_buf(aos)
        check(self.lib.pyh_upload_state_async(self._ctx, int(gid), _dp(aos)))

    def commit_uploads(self):
        check(self.lib.pyh_commit_uploads(self._ctx))

    def download_async(self, gid, out):
        """Enqueue conversion + D2H copy of the current solution of block ``gid`` into ``out``
        (page-locked); valid after ``transfers_sync``."""
        self._check_stream_buf(out)
        check(self.lib.pyh_download_state_async(self._ctx, int(gid), _dp(out)))

    def transfers_sync(self):
        check(self.lib.pyh_transfers_sync(self._ctx))

    def downloads_sync(self):
        """Wait for the D2H copies enqueued so far only (safe from a writer thread)."""
        check(self.lib.pyh_downloads_sync(self._ctx))

    def pinned_state_buffer(self):
        """A page-locked (ny, nx, 4) float64 array owned by the engine (freed in ``close``)."""
        nbytes = self.ny * self.nx * 4 * 8
        p = C.c_void_p()
        check(self.lib.pyh_host_alloc(nbytes, C.byref(p)))
        self._pinned.append(p)
        buf = (C.c_double * (self.ny * self.nx * 4)).from_address(p.value)
        return np.frombuffer(buf, dtype=np.float64).reshape(self.ny, self.nx, 4)

    def download_ghost(self, gid, side):
        s = SIDES.index(side)
        shape = (self.ny, 1, 4) if side in ("E", "W") else (1, self.nx, 4)
        out = np.empty(shape)
        check(self.lib.pyh_download_ghost(self._ctx, int(gid), s, _dp(out)))
        return out

    # -- hot path --------------------------------------------------------------------------------------
    def apply_bc(self):
        check(self.lib.pyh_apply_bc(self._ctx))

    def get_dt(self, t, t_final):
        dt = C.c_double()
        check(self.lib.pyh_get_dt(self._ctx, float(t), float(t_final), C.byref(dt)))
        return dt.value

    def local_dt(self, dev_ptr=None):
        """CFL * min on the device (global after comm_init); None keeps it in the context's own scratch double."""
        check(self.lib.pyh_local_dt(self._ctx, C.c_void_p(dev_ptr) if dev_ptr else None))

    def step(self, dt):
        check(self.lib.pyh_step(self._ctx, float(dt)))

    def step_begin(self, dt):
        check(self.lib.pyh_step_begin(self._ctx, float(dt)))

    def step_begin_dev(self, dev_ptr=None):
        check(self.lib.pyh_step_begin_dev(self._ctx, C.c_void_p(dev_ptr) if dev_ptr else None))

    def stage(self, s):
        check(self.lib.pyh_stage(self._ctx, int(s)))

    def run(self, t, t_final, max_steps=-1, poll_every=64, record_dts=0):
        """Device-resident time loop; returns (t, steps_done, unrealizable, dts)."""
        tt = C.c_double(float(t))
        steps = C.c_int64(0)
        bad = C.c_int32(0)
        dts = np.zeros(max(int(record_dts), 1))
        check(
            self.lib.pyh_run(
                self._ctx, C.byref(tt), float(t_final), int(max_steps), int(poll_every), C.byref(steps),
                C.byref(bad), _dp(dts), int(record_dts),
            )
        )
        n = min(steps.value, int(record_dts))
        return tt.value, steps.value, bool(bad.value), dts[:n].copy()

    def realizable(self):
        ok = C.c_int32(0)
        check(self.lib.pyh_realizable(self._ctx, C.byref(ok)))
        return bool(ok.value)

    # -- multi-rank transport owned by the library (NCCL bound inside the C layer) ------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """128-byte NCCL id (rank 0 creates it; the host hands it to every rank through any side channel)."""
        buf = C.create_string_buffer(128)
        check(_lib.load().pyh_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, rank, world, unique_id: bytes, owner):
        """Collective.  ``owner``: {global block id: rank} or a sequence indexed by block id.  Afterwards apply_bc /
        step / run / get_dt / realizable are collective calls and exchange the remote ghost strips themselves."""
        n = len(owner)
        table = (C.c_int32 * n)(*[int(owner[g]) for g in range(n)])
        idbuf = C.create_string_buffer(bytes(unique_id), 128)
        check(self.lib.pyh_comm_init(self._ctx, int(rank), int(world), idbuf, table, n))
        self.comm_world = int(world)

    def comm_info(self):
        r, w, n = C.c_int32(), C.c_int32(), C.c_int32()
        d = C.c_int64()
        check(self.lib.pyh_comm_info(self._ctx, C.byref(r), C.byref(w), C.byref(n), C.byref(d)))
        return dict(rank=r.value, world=w.value, messages=n.value, doubles_per_exchange=d.value)

    # -- remote halo (host-driven transport: the pack / unpack halves of the exchange) ----------------------------
    def halo_slots(self):
        n = C.c_int64()
        nd = C.c_int64()
        check(self.lib.pyh_halo_count(self._ctx, C.byref(n), C.byref(nd)))
        out = []
        for s in range(n.value):
            gid, side, nbr = C.c_int32(), C.c_int32(), C.c_int32()
            off, ln = C.c_int64(), C.c_int64()
            check(self.lib.pyh_halo_slot(self._ctx, s, C.byref(gid), C.byref(side), C.byref(nbr), C.byref(off), C.byref(ln)))
            out.append(dict(gid=gid.value, side=SIDES[side.value], nbr=nbr.value, offset=off.value, length=ln.value))
        return out, nd.value

    def pack_halo(self, dev_ptr):
        check(self.lib.pyh_pack_halo(self._ctx, C.c_void_p(dev_ptr)))

    def unpack_halo(self, dev_ptr):
        check(self.lib.pyh_unpack_halo(self._ctx, C.c_void_p(dev_ptr)))

    # -- test hooks / counters ---------------------------------------------------------------------------
    def residual(self, gid):
        out = np.empty((self.ny, self.nx, 4))
        check(self.lib.pyh_residual(self._ctx, int(gid), _dp(out)))
        return out

    def debug_fetch(self, gid, what):
        out = np.empty((self.ny, self.nx, 4))
        check(self.lib.pyh_debug_fetch(self._ctx, int(gid), {"gx": 0, "gy": 1, "phi": 2}[what], _dp(out)))
        return out

    def march_shape(self):
        """(lanes per thread block, rows per strip) of the stage kernel"""
        a, b = C.c_int32(), C.c_int32()
        check(self.lib.pyh_march_shape(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def stage_path(self):
        """'fused' (row-marching stage kernel) or 'split' (recon / flux / update kernels of small problems)"""
        a = C.c_int32()
        check(self.lib.pyh_stage_path(self._ctx, C.byref(a), None))
        return "split" if a.value else "fused"

    def stage_path_tuning(self):
        """ms per stage launch the first run() measured for (fused, split); (0, 0) when the path was not chosen by measurement"""
        a, ms = C.c_int32(), (C.c_double * 2)()
        check(self.lib.pyh_stage_path(self._ctx, C.byref(a), ms))
        return ms[0], ms[1]

    def launch_count(self):
        n = C.c_int64()
        check(self.lib.pyh_launch_count(self._ctx, C.byref(n)))
        return n.value

    def stream(self):
        s = C.c_uint64()
        check(self.lib.pyh_stream(self._ctx, C.byref(s)))
        return s.value

    def sync(self):
        check(self.lib.pyh_sync(self._ctx))
