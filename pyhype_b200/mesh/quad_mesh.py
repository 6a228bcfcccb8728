"""Per-block mesh geometry on the host (one-time setup; the GPU hot path takes it as input).

Mirrors ``pyhype.mesh.quad_mesh.QuadMesh`` (pyhype/mesh/quad_mesh.py:40-184): the numpy
expressions are evaluated in the reference's order so that nodes, centroids, areas and face
angles are bit-identical.  Unlike the reference, face data are kept once per *face*
(vertical faces (ny, nx+1), horizontal faces (ny+1, nx)) instead of four per-cell copies: the E
face of cell j and the W face of cell j+1 are built from the same two nodes with the same
formula (mesh/base.py:57-78), so they carry the same bits.  Everything that needs libm
(arctan, arccos, cos, sin) is computed here; the exactly rounded rest (lengths, midpoints,
centroid offsets) is derived on the device by ``k_geometry``.
"""
from __future__ import annotations

import hashlib
import os

import numpy as np

# Host geometry cache: arctan / arccos / cos / sin over (ny, nx) arrays cost ~1.5 s per
# 2048 x 2048 block and stay on the host because libm's results are the contract (they are not reproducible bit for bit on
# the device).  With PYH_GEOM_CACHE=<directory> every QuadMesh stores its arrays under a key of (nx, ny, the four vertices,
# numpy's version) and later constructions with the same key load them instead (memory-mapped npy files, bit-identical by
# construction since they ARE the earlier results).  Off by default: a cache directory is the user's decision.
_CACHED = ("nodes_x", "nodes_y", "x", "y", "theta_v", "theta_h", "cos_v", "sin_v", "cos_h", "sin_h", "area")


def _cache_dir():
    return os.environ.get("PYH_GEOM_CACHE") or None


class GridLocation:
    def __init__(self, x=None, y=None):
        self.x = x
        self.y = y


class _Vertices:
    def __init__(self, NE, NW, SE, SW):
        self.NE, self.NW, self.SE, self.SW = NE, NW, SE, SW


class QuadMesh:
    def __init__(self, nx, ny, NE, NW, SE, SW, nghost=1):
        self.nx, self.ny, self.nghost = int(nx), int(ny), nghost
        self.shape = (self.ny, self.nx)
        self.vertices = _Vertices(NE=NE, NW=NW, SE=SE, SW=SW)
        self.from_cache = False
        if not self._load_cached():
            self._create()
            self._store_cached()
        v = self.vertices
        self.nodes = GridLocation(self.nodes_x[:, :, None], self.nodes_y[:, :, None])
        self.A = self.area[:, :, None]
        # BaseBlockGhost._is_cartesian (blocks/quad_block.py:96-113): exact corner comparison
        self.is_cartesian = bool(
            (v.NE[1] == v.NW[1]) and (v.SE[1] == v.SW[1]) and (v.SE[0] == v.NE[0]) and (v.SW[0] == v.NW[0])
        )

    def _cache_key(self):
        v = self.vertices
        corners = np.array([v.NE[0], v.NE[1], v.NW[0], v.NW[1], v.SE[0], v.SE[1], v.SW[0], v.SW[1]], dtype=np.float64)
        h = hashlib.sha256(corners.tobytes() + f"|{self.nx}|{self.ny}|numpy {np.__version__}".encode())
        return h.hexdigest()[:32]

    def _load_cached(self):
        d = _cache_dir()
        if not d:
            return False
        path = os.path.join(d, self._cache_key())
        if not os.path.exists(os.path.join(path, "complete")):
            return False
        try:
            for name in _CACHED:
                arr = np.load(os.path.join(path, name + ".npy"), mmap_mode="r")
                setattr(self, name, arr[:, :, None] if name in ("x", "y") else arr)
        except Exception:
            return False
        self.from_cache = True
        return True

    def _store_cached(self):
        d = _cache_dir()
        if not d:
            return
        path = os.path.join(d, self._cache_key())
        try:
            os.makedirs(path, exist_ok=True)
            for name in _CACHED:
                arr = getattr(self, name)
                np.save(os.path.join(path, name + ".npy"), arr[:, :, 0] if name in ("x", "y") else arr)
            open(os.path.join(path, "complete"), "w").close()
        except OSError:
            pass   # a read-only or full cache directory must not stop the run

    def _create(self):
        nx, ny, v = self.nx, self.ny, self.vertices
        # block edges, then one linspace per node row (quad_mesh.py:79-99)
        east_x = np.linspace(v.SE[0], v.NE[0], ny + 1)
        east_y = np.linspace(v.SE[1], v.NE[1], ny + 1)
        west_x = np.linspace(v.SW[0], v.NW[0], ny + 1)
        west_y = np.linspace(v.SW[1], v.NW[1], ny + 1)
        xn = np.empty((ny + 1, nx + 1))
        yn = np.empty((ny + 1, nx + 1))
        for r in range(ny + 1):
            xn[r] = np.linspace(west_x[r], east_x[r], nx + 1)
            yn[r] = np.linspace(west_y[r], east_y[r], nx + 1)
        self.nodes_x, self.nodes_y = xn, yn
        ne = (xn[1:, 1:], yn[1:, 1:])
        nw = (xn[1:, :-1], yn[1:, :-1])
        se = (xn[:-1, 1:], yn[:-1, 1:])
        sw = (xn[:-1, :-1], yn[:-1, :-1])
        # centroids (quad_mesh.py:172-184); kept (ny, nx, 1) like the reference's mesh.x / mesh.y
        self.x = (0.25 * (ne[0] + nw[0] + se[0] + sw[0]))[:, :, None]
        self.y = (0.25 * (ne[1] + nw[1] + se[1] + sw[1]))[:, :, None]
        # face angles (mesh/base.py:66-74).  vertical face J of row i: high = node (i+1, J), low = node (i, J)
        hx, hy, lx, ly = xn[1:, :], yn[1:, :], xn[:-1, :], yn[:-1, :]
        theta_v = np.arctan((lx - hx) / (hy - ly))
        # horizontal face I of column j: high = node (I, j), low = node (I, j+1)
        hx, hy, lx, ly = xn[:, :-1], yn[:, :-1], xn[:, 1:], yn[:, 1:]
        theta_h = np.pi / 2 - np.arctan((hy - ly) / (lx - hx))
        self.theta_v, self.theta_h = theta_v, theta_h
        self.cos_v, self.sin_v = np.cos(theta_v), np.sin(theta_v)
        self.cos_h, self.sin_h = np.cos(theta_h), np.sin(theta_h)
        # area by Bretschneider's formula (quad_mesh.py:137-170, side lengths as in _face_length :258-268)
        def side(p, q):
            dx = p[0] - q[0]
            dy = q[1] - p[1]
            return np.sqrt(dx * dx + dy * dy)

        s1, s3 = side(sw, nw), side(se, ne)
        s2, s4 = side(nw, ne), side(sw, se)
        d2 = (nw[0] - se[0]) ** 2 + (nw[1] - se[1]) ** 2
        a1 = np.arccos((s1**2 + s4**2 - d2) / (2 * s1 * s4))
        a2 = np.arccos((s2**2 + s3**2 - d2) / (2 * s2 * s3))
        s = 0.5 * (s1 + s2 + s3 + s4)
        p1 = (s - s1) * (s - s2) * (s - s3) * (s - s4)
        p2 = s1 * s2 * s3 * s4
        self.area = np.sqrt(p1 - 0.5 * p2 * (1 + np.cos(a1 + a2)))

    # reference-style accessors (pyhype/mesh/quad_mesh.py:371-400)
    def get_NE_vertices(self):
        return self.nodes.x[1:, 1:, :], self.nodes.y[1:, 1:, :]

    def get_NW_vertices(self):
        return self.nodes.x[1:, :-1, :], self.nodes.y[1:, :-1, :]

    def get_SE_vertices(self):
        return self.nodes.x[:-1, 1:, :], self.nodes.y[:-1, 1:, :]

    def get_SW_vertices(self):
        return self.nodes.x[:-1, :-1, :], self.nodes.y[:-1, :-1, :]
