"""Loader for the fixtures written by oracle/make_golden.py (outputs of the unmodified reference)."""
from __future__ import annotations

import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SIDES = ("E", "W", "N", "S")


def names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


class Fixture:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.meta = json.loads(str(self.z["meta"]))
        self.name = name
        self.gids = self.meta["gids"]
        self.nx, self.ny = self.meta["nx"], self.meta["ny"]
        self.blocks = {}
        for g in self.gids:
            mb = dict(self.meta["blocks"][str(g)])
            mb["nBLK"] = g
            for s in SIDES:
                if mb["BCType" + s] == "@dirichlet":
                    mb["BCType" + s] = self.z[f"dirichlet_{g}_{s}"]
            self.blocks[g] = mb

    def scheme(self):
        m = self.meta
        return dict(flux=m["flux"], limiter=m["limiter"], recon=m["recon"], integrator=m["integrator"], CFL=m["CFL"],
                    nqp=m.get("nqp", 1))

    def __getitem__(self, key):
        return self.z[key]
