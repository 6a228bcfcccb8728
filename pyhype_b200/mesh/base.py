"""Multi-block mesh description generators (host-only setup).

Mirror of ``pyhype.mesh.base`` (pyhype/mesh/base.py:88-226): ``QuadMeshGenerator`` places the
block corner vertices by transfinite interpolation of the four boundary curves and emits the
``{block_id: {nBLK, NW, NE, SW, SE, Neighbor*, BCType*}}`` dictionary consumed by
``Euler2D``.  The interpolation is evaluated in the reference's operation order because the
exact vertex bits decide ``is_cartesian`` and every derived geometry value.
"""
from __future__ import annotations

import numpy as np


class MeshGenerator:
    def __init__(self):
        self.dict = {}


class QuadMeshGenerator(MeshGenerator):
    def __init__(
        self, nx_blk, ny_blk, BCE, BCW, BCN, BCS, BCNE=None, BCNW=None, BCSE=None, BCSW=None,
        NE=None, NW=None, SE=None, SW=None, left_x=None, left_y=None, right_x=None, right_y=None,
        top_x=None, top_y=None, bot_x=None, bot_y=None, blk_num_offset=0,
    ):
        super().__init__()
        self.nx, self.ny = nx_blk + 1, ny_blk + 1  # vertex counts, as in the reference
        self._blk_num_offset = blk_num_offset
        self.BCE = [BCE[0]] * ny_blk if len(BCE) == 1 else BCE
        self.BCW = [BCW[0]] * ny_blk if len(BCW) == 1 else BCW
        self.BCN = [BCN[0]] * nx_blk if len(BCN) == 1 else BCN
        self.BCS = [BCS[0]] * nx_blk if len(BCS) == 1 else BCS
        self.BCNE = self.BCNW = self.BCSE = self.BCSW = None

        nvx, nvy = self.nx, self.ny
        x, y = np.meshgrid(np.linspace(0, 1, nvx), np.linspace(0, 1, nvy))
        x[0, :] = np.linspace(SW[0], SE[0], nvx) if bot_x is None else bot_x
        y[0, :] = np.linspace(SW[1], SE[1], nvx) if bot_y is None else bot_y
        x[-1, :] = np.linspace(NW[0], NE[0], nvx) if top_x is None else top_x
        y[-1, :] = np.linspace(NW[1], NE[1], nvx) if top_y is None else top_y
        x[:, 0] = np.linspace(SW[0], NW[0], nvy) if left_x is None else left_x
        y[:, 0] = np.linspace(SW[1], NW[1], nvy) if left_y is None else left_y
        x[:, -1] = np.linspace(SE[0], NE[0], nvy) if right_x is None else right_x
        y[:, -1] = np.linspace(SE[1], NE[1], nvy) if right_y is None else right_y
        self.x, self.y = self._fill_interior(x, y)
        self.dict = self._create_block_descriptions()

    def _fill_interior(self, x, y):
        """Transfinite interpolation of the interior vertices (pyhype/mesh/base.py:88-121)."""
        nvx, nvy = self.nx, self.ny
        up_j, up_i = np.meshgrid(
            np.linspace(1 / nvx, (nvx - 1) / nvx, nvx - 2), np.linspace(1 / nvy, (nvy - 1) / nvy, nvy - 2)
        )
        dn_j, dn_i = np.meshgrid(
            np.linspace((nvx - 1) / nvx, 1 / nvx, nvx - 2), np.linspace((nvy - 1) / nvy, 1 / nvy, nvy - 2)
        )

        def blend(c):
            return (
                dn_i * c[0, 1:-1] + up_i * c[-1, 1:-1] + dn_j * c[1:-1, 0, None] + up_j * c[1:-1, -1, None]
                - dn_i * dn_j * c[0, 0] - dn_i * up_j * c[0, -1] - up_i * dn_j * c[-1, 0] - up_i * up_j * c[-1, -1]
            )

        x[1:-1, 1:-1] = blend(x)
        y[1:-1, 1:-1] = blend(y)
        return x, y

    def _create_block_descriptions(self):
        """Block id = nx_blk * i + j from the SW corner (pyhype/mesh/base.py:191-226)."""
        nby, nbx = self.ny - 1, self.nx - 1
        out = {}
        for i in range(nby):
            for j in range(nbx):
                num = nbx * i + j + self._blk_num_offset
                east_edge, west_edge = j == nbx - 1, j == 0
                north_edge, south_edge = i == nby - 1, i == 0
                out[num] = {
                    "nBLK": num,
                    "NW": [self.x[i + 1, j], self.y[i + 1, j]],
                    "NE": [self.x[i + 1, j + 1], self.y[i + 1, j + 1]],
                    "SW": [self.x[i, j], self.y[i, j]],
                    "SE": [self.x[i, j + 1], self.y[i, j + 1]],
                    "NeighborE": None if east_edge else num + 1,
                    "NeighborW": None if west_edge else num - 1,
                    "NeighborN": num + nbx if num + nbx < nby * nbx else None,
                    "NeighborS": num - nbx if num - nbx >= 0 else None,
                    "NeighborNE": None, "NeighborNW": None, "NeighborSE": None, "NeighborSW": None,
                    "BCTypeE": self.BCE[i] if east_edge else None,
                    "BCTypeW": self.BCW[i] if west_edge else None,
                    "BCTypeN": self.BCN[j] if north_edge else None,
                    "BCTypeS": self.BCS[j] if south_edge else None,
                    "BCTypeNE": self.BCE[i] if east_edge else None,
                    "BCTypeNW": self.BCW[i] if west_edge else None,
                    "BCTypeSE": self.BCN[j] if north_edge else None,
                    "BCTypeSW": self.BCS[j] if south_edge else None,
                }
        return out
