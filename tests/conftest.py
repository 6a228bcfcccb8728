import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the live reference tree (/root/reference, build container only)")


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(params=["fused", "split"])
def stage_path(request, monkeypatch):
    """Run a GPU test once per stage implementation: the fused row-marching kernel (pyh_stage_march.cuh) and the three-kernel
    stage (pyh_stage_split.cuh).  pyh_finalize reads PYH_SPLIT; without it a context takes whichever path its first run()
    measures as faster on its blocks.  (Two / three quadrature points always run fused.)"""
    monkeypatch.setenv("PYH_SPLIT", "1" if request.param == "split" else "0")
    return request.param
