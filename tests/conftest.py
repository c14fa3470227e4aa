import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bih-gpu-raytracer_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no GPU in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def scenes():
    from bihrt import scenes as S
    return S


@pytest.fixture()
def renderer():
    """A fresh context on cuda:0; fails loudly if libbihrt.so is missing (no fallback)."""
    import bihrt
    r = bihrt.Renderer(device=0)
    yield r
    r.close()


def assert_view_equals_oracle(view, ob):
    """Bit-exact comparison of bihrt_export_reference_view with the oracle's arrays."""
    assert view["n"] == ob.n and view["nu"] == ob.nu
    np.testing.assert_array_equal(view["morton_codes"], ob.codes)
    np.testing.assert_array_equal(view["tris_indexes"], ob.tris_idx)
    np.testing.assert_array_equal(view["unique_morton_codes"], ob.umc)
    np.testing.assert_array_equal(view["duplicates_cnts"], ob.cnt)
    np.testing.assert_array_equal(view["first_idxs"], ob.first)
    np.testing.assert_array_equal(view["children"], ob.children)
    np.testing.assert_array_equal(view["is_leaf"], ob.is_leaf)
    np.testing.assert_array_equal(view["axis"], ob.axis)
    np.testing.assert_array_equal(view["parent"], ob.parent)
    np.testing.assert_array_equal(view["leaf_parents"], ob.leaf_parents)
    # clip planes: same floats (== so that +0/-0, which no ray can tell apart, compare equal)
    np.testing.assert_array_equal(view["clip_planes"], ob.clip)
    np.testing.assert_array_equal(view["scene_lo"], ob.scene_lo)
    np.testing.assert_array_equal(view["scene_hi"], ob.scene_hi)
