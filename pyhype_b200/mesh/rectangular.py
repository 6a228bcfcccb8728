"""Rectangular multi-block domains (mirror of pyhype/mesh/rectangular.py:23-62)."""
from __future__ import annotations

import numpy as np

from .base import QuadMeshGenerator


class RectagularMeshGenerator:  # (sic) the reference's public name
    @staticmethod
    def generate(BCE, BCW, BCN, BCS, east, west, north, south, n_blocks_horizontal, n_blocks_vertical):
        if east <= west:
            raise ValueError(f"East value {east} must be larger than west {west}")
        if north <= south:
            raise ValueError(f"North value {north} must be larger than south {south}")
        ones_h = np.ones(n_blocks_horizontal + 1)
        ones_v = np.ones(n_blocks_vertical + 1)
        xs = np.linspace(west, east, n_blocks_horizontal + 1)
        ys = np.linspace(south, north, n_blocks_vertical + 1)
        return QuadMeshGenerator(
            nx_blk=n_blocks_horizontal, ny_blk=n_blocks_vertical, BCE=BCE, BCW=BCW, BCN=BCN, BCS=BCS,
            top_x=xs, bot_x=xs, top_y=north * ones_h, bot_y=south * ones_h,
            left_x=west * ones_v, right_x=east * ones_v, left_y=ys, right_y=ys,
        )
