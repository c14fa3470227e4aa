"""Pins the CPU oracle -- and this repo's CUDA path -- against the REFERENCE'S OWN kernels run on the GPU:
oracle/_ref/libref_harness.so is R/src/CUDAKernels.cu (+ Camera.cu, Ray.cu, GPUArrayManager.cpp) compiled
unmodified for sm_100a by oracle/Makefile in the dev container; it travels to the GPU box prebuilt."""
import numpy as np
import pytest

from oracle import ref_harness

pytestmark = pytest.mark.gpu


def test_reference_kernel_library_travelled_to_this_box():
    """The pin must not vanish silently: oracle/_ref/libref_harness.so is built by __graft_entry__.build() where the
    reference tree exists and travels with the snapshot (git-ignored, NOT gpurun-ignored).  On a GPU box its absence
    is a FAILURE, not a skip."""
    assert ref_harness.available(), ("oracle/_ref/libref_harness.so is missing on this GPU box: run __graft_entry__.build() in the "
                                     "dev container (needs /root/reference) before shipping the snapshot")


ID_MISMATCH_MAX = 1e-4      # north star: ids exact except documented ties
T_REL_TOL = 1e-5            # nvcc contracts the reference's glm expressions to FMAs; the oracle does not


@pytest.mark.parametrize("name", ["dodecahedron", "cornell", "sphere187", "atrium", "soup", "sphere361"])
def test_oracle_and_library_match_reference_kernels(renderer, scenes, oracle, name):
    assert ref_harness.available(), "oracle/_ref/libref_harness.so missing (see test_reference_kernel_library_travelled_to_this_box)"
    tri, cam, w, h = {
        "dodecahedron": lambda: (scenes.dodecahedron(), scenes.pinhole_camera(aspect=1.0), 128, 128),
        "cornell": lambda: (scenes.cornell_box(), scenes.cornell_camera(), 256, 256),
        "sphere187": lambda: (scenes.displaced_sphere(187), scenes.pinhole_camera(), 640, 360),
        "sphere361": lambda: (scenes.displaced_sphere(361), scenes.pinhole_camera(), 320, 180),
        "atrium": lambda: (scenes.atrium(0.3), scenes.atrium_camera(), 320, 180),
        "soup": lambda: (scenes.random_soup(50000), scenes.pinhole_camera(), 320, 180),
    }[name]()
    ob = oracle.Bih(tri)
    ref = ref_harness.RefScene(ob)
    assert ref.build() == ob.nu
    rv = ref.export()
    # --- build: the reference's BuildTree + FindClipPlanes vs the oracle, bit for bit
    for k, o in (("morton_codes", ob.codes), ("tris_indexes", ob.tris_idx), ("unique_morton_codes", ob.umc),
                 ("duplicates_cnts", ob.cnt), ("first_idxs", ob.first), ("children", ob.children),
                 ("is_leaf", ob.is_leaf), ("axis", ob.axis), ("parent", ob.parent), ("leaf_parents", ob.leaf_parents),
                 ("clip_planes", ob.clip)):
        np.testing.assert_array_equal(rv[k], o, err_msg=k)
    # --- and vs this repo's library
    renderer.load_models(tri).build()
    v = renderer.reference_view()
    for k in ("morton_codes", "tris_indexes", "children", "is_leaf", "axis", "parent", "leaf_parents", "clip_planes"):
        np.testing.assert_array_equal(v[k], rv[k], err_msg=k)
    # --- trace: the reference's TraverseTree vs the oracle's literal restatement vs the library
    rays = oracle.camera_rays(cam, w, h)
    t_ref, s_ref, _ = ref.trace(rays)
    t_orc, s_orc, _ = ob.trace(rays, "ref")
    t_lib, s_lib, _ = renderer.trace(rays)
    for (t, s, what) in ((t_orc, s_orc, "oracle"), (t_lib, s_lib, "library")):
        bad = s != s_ref
        assert bad.sum() <= ID_MISMATCH_MAX * len(rays), "%s: %d of %d ids differ from the reference kernels" % (what, bad.sum(), len(rays))
        # (ids differ only where a ray grazes a triangle edge: nvcc contracts the reference's glm
        # expressions into FMAs, the oracle evaluates them uncontracted, so the u/v edge tests can fall
        # on different sides and the ray continues to the next surface -- the documented tie class)
        hit = ~bad & (s_ref >= 0)
        np.testing.assert_allclose(t[hit], t_ref[hit], rtol=T_REL_TOL, atol=0, err_msg=what)
    ref.close()
