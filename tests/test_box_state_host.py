"""-m "not gpu": the lazy two-state fill of DeviceBackedState (pyhype_b200/blocks.py) on the host side -- pending fills are kept
as numbers, non-dimensionalised as numbers, pushed to the (here: recording stand-in) engine as fill calls instead of an upload,
and materialise into exactly the reference's np.where array when host code reads ``data``."""
import numpy as np

import cases
from pyhype_b200.blocks import DeviceBackedState
from pyhype_b200.fluids import Air
from pyhype_b200.mesh.quad_mesh import QuadMesh


class RecordingSync:
    def __init__(self, mesh):
        self.mesh, self.calls = mesh, []

    def centroids(self):
        return self.mesh.x, self.mesh.y

    def upload(self, arr):
        self.calls.append(("upload", arr.copy()))

    def fill_uniform(self, v):
        self.calls.append(("uniform", np.array(v)))

    def fill_box(self, *a):
        self.calls.append(("box",) + a)

    def download(self):
        raise AssertionError("no download expected")


def _state():
    b = cases.em_mesh()[0]
    mesh = QuadMesh(12, 10, NE=b["NE"], NW=b["NW"], SE=b["SE"], SW=b["SW"])
    air = Air(a_inf=343.0, rho_inf=1.0)
    sync = RecordingSync(mesh)
    return DeviceBackedState(air, (10, 12, 4), sync), sync, mesh


def _dimensional(W):
    g = cases.GAMMA
    rho, u, v, p = W
    return np.array([rho, rho * u, rho * v, p / (g - 1) + 0.5 * rho * (u * u + v * v)])


def test_pending_box_is_pushed_as_fill_calls_not_as_upload():
    st, sync, mesh = _state()
    UL, UR = _dimensional((4.6968, 0.0, 0.0, 404400.0)), _dimensional((1.1742, 0.0, 0.0, 101100.0))
    st.fill_box(3, 7, 3, 7, UL.reshape(1, 1, 4), UR.reshape(1, 1, 4))
    st.make_non_dimensional()
    st.push_if_touched()
    assert [c[0] for c in sync.calls] == ["box"]
    _, x0, x1, y0, y1, inside, outside = sync.calls[0]
    ref = cases.explosion_ic(mesh.x[:, :, 0], mesh.y[:, :, 0])
    assert (x0, x1, y0, y1) == (3.0, 7.0, 3.0, 7.0)
    assert set(map(tuple, ref.reshape(-1, 4))) <= {tuple(inside), tuple(outside)}   # the same 2 x 4 doubles as the numpy IC


def test_pending_box_materialises_to_the_numpy_fill_on_read():
    st, sync, mesh = _state()
    UL, UR = _dimensional((4.6968, 0.0, 0.0, 404400.0)), _dimensional((1.1742, 0.0, 0.0, 101100.0))
    st.fill_box(3, 7, 3, 7, UL, UR)
    st.make_non_dimensional()
    assert np.array_equal(st.data, cases.explosion_ic(mesh.x[:, :, 0], mesh.y[:, :, 0]))
    st.push_if_touched()                      # host code has seen (and may have written) the array: plain upload
    assert [c[0] for c in sync.calls] == ["upload"]


def test_box_on_top_of_host_data_is_applied_on_the_host():
    st, sync, mesh = _state()
    base = np.random.default_rng(0).random((10, 12, 4)) + 1.0
    st.data = base.copy()
    mark = np.array([9.0, 8.0, 7.0, 6.0])
    st.fill_box(-np.inf, 5.0, -np.inf, np.inf, mark, None)
    st.push_if_touched()
    assert [c[0] for c in sync.calls] == ["upload"]
    x = mesh.x[:, :, 0]
    assert np.array_equal(sync.calls[0][1], np.where((x <= 5.0)[..., None], mark, base))


def test_stale_alias_contract_is_what_the_docstring_says():
    """an array kept across a device call is a stale copy; re-reading .data downloads, the setter always wins."""
    st, sync, mesh = _state()

    class Dev(RecordingSync):
        def download(self):
            self.calls.append(("download",))
            return np.full((10, 12, 4), 7.0)

    dev = Dev(mesh)
    st._sync = dev
    st.data = np.ones((10, 12, 4))
    alias = st.data
    st.push_if_touched()
    st.mark_device_newer()                 # a solver call happened
    alias[...] = 0.0                       # stale write: not tracked
    st.push_if_touched()
    assert [c[0] for c in dev.calls] == ["upload"]
    assert np.all(st.data == 7.0)          # fresh download
    st.data = np.full((10, 12, 4), 3.0)    # the setter always takes effect
    st.push_if_touched()
    assert dev.calls[-1][0] == "upload" and np.all(dev.calls[-1][1] == 3.0)
