"""The C-ABI library loads and exports every symbol include/bihrt.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bihrt.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bihrt_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("bihrt_create", "bihrt_scene_load_triangles", "bihrt_scene_load_obj", "bihrt_build",
                 "bihrt_export_reference_view", "bihrt_trace", "bihrt_render", "bihrt_framebuffer",
                 "bihrt_framebuffer_read"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    import bihrt
    lib = bihrt.load_library()
    syms = declared_symbols()
    assert sorted(bihrt.ABI_SYMBOLS) == syms
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.bihrt_version() == 100


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, never route through the oracle."""
    import torch
    import bihrt
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(bihrt.BihrtError):
        bihrt.Renderer(device=0)


def test_product_never_imports_oracle():
    for top in (os.path.join(ROOT, "bih-gpu-raytracer_b200"), os.path.join(ROOT, "tools")):      # the package and the dev tools
      for dirpath, _, files in os.walk(top):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                src = open(os.path.join(dirpath, fn), errors="ignore").read()
                code = "\n".join(l for l in src.splitlines() if not l.strip().startswith(("//", "#", "*", "/*", '"""')))
                assert "import oracle" not in code and "from oracle" not in code and "libbih_oracle" not in code, fn
