"""-m "not gpu": host geometry cache (PYH_GEOM_CACHE, pyhype_b200/mesh/quad_mesh.py): a second construction with the same key loads
the stored arrays and is bit-identical to a fresh build; another block (other vertices / size) misses; default is off."""
import os
import time

import numpy as np

import cases
from pyhype_b200.mesh import quad_mesh
from pyhype_b200.mesh.quad_mesh import QuadMesh

NAMES = quad_mesh._CACHED


def _mesh(b, nx, ny):
    return QuadMesh(nx, ny, NE=b["NE"], NW=b["NW"], SE=b["SE"], SW=b["SW"])


def test_cache_round_trip_is_bit_identical(tmp_path, monkeypatch):
    blocks = cases.dmr_mesh()          # skewed blocks: every array is non-trivial
    monkeypatch.delenv("PYH_GEOM_CACHE", raising=False)
    fresh = _mesh(blocks[2], 37, 23)
    assert not fresh.from_cache
    monkeypatch.setenv("PYH_GEOM_CACHE", str(tmp_path))
    first = _mesh(blocks[2], 37, 23)
    second = _mesh(blocks[2], 37, 23)
    assert not first.from_cache and second.from_cache
    for name in NAMES:
        assert np.array_equal(np.asarray(getattr(second, name)), np.asarray(getattr(fresh, name))), name
    assert second.is_cartesian == fresh.is_cartesian
    assert np.array_equal(second.A, fresh.A) and np.array_equal(second.nodes.x, fresh.nodes.x)
    assert not _mesh(blocks[1], 37, 23).from_cache      # other vertices
    assert not _mesh(blocks[2], 36, 23).from_cache      # other size
    assert len(os.listdir(tmp_path)) == 3


def test_unwritable_cache_directory_is_ignored(monkeypatch):
    monkeypatch.setenv("PYH_GEOM_CACHE", "/proc/definitely/not/writable")
    m = _mesh(cases.em_mesh()[0], 8, 6)
    assert not m.from_cache and m.area.shape == (6, 8)
