"""Quality mode: 63-bit Morton keys + capped leaves.  NOT a parity path -- the tree differs from the
reference's by design -- so the check is the only one a different tree must still pass: every ray's closest hit equals
brute force over all triangles (oracle 'brute': same Moller-Trumbore arithmetic, so t is bit-exact; the primitive id may
differ only where two triangles give exactly the same t, <= 1e-4 of the rays)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ID_MISMATCH_MAX = 1e-4


def quality(renderer, cap=4):
    renderer.set_option("morton_bits", 63)
    renderer.set_option("leaf_cap", cap)
    return renderer


def check_vs_brute(renderer, ob, rays):
    t, s, p, cnt = renderer.trace(rays, counted=True)
    t0, s0, p0 = ob.trace(rays, "brute")
    np.testing.assert_array_equal(t, t0)
    bad = p != p0
    assert bad.sum() <= max(1, ID_MISMATCH_MAX * len(rays)), "%d of %d primitive ids differ from brute force" % (bad.sum(), len(rays))
    assert np.array_equal(s >= 0, s0 >= 0)
    t2, s2, p2 = renderer.trace(rays)                       # uninstrumented kernel: same answer
    np.testing.assert_array_equal(t2, t)
    np.testing.assert_array_equal(p2, p)
    np.testing.assert_array_equal(s2, s)
    assert cnt["max_stack"] < 96
    return cnt


CASES = {
    "dodecahedron": lambda S: (S.dodecahedron(), S.pinhole_camera(aspect=1.0), 96, 96),
    "cornell": lambda S: (S.cornell_box(), S.cornell_camera(), 192, 192),
    "sphere64": lambda S: (S.displaced_sphere(64), S.pinhole_camera(), 320, 180),
    "sphere187": lambda S: (S.displaced_sphere(187), S.pinhole_camera(), 160, 90),
    "atrium": lambda S: (S.atrium(0.05), S.atrium_camera(), 160, 90),
    "soup": lambda S: (S.random_soup(20000), S.pinhole_camera(), 160, 90),
}


@pytest.mark.parametrize("cap", [1, 4, 16])
@pytest.mark.parametrize("name", list(CASES))
def test_quality_mode_hits_equal_brute_force(renderer, scenes, oracle, name, cap):
    tri, cam, w, h = CASES[name](scenes)
    ob = oracle.Bih(tri)
    rays = oracle.camera_rays(cam, w, h)
    quality(renderer, cap).load_models(tri).build()
    info = renderer.build_info()
    assert info["n"] == len(tri) and info["sort_passes"] == 8
    assert info["nu"] >= (len(tri) + cap - 1) // cap            # every leaf holds at most cap triangles
    check_vs_brute(renderer, ob, rays)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 9, 33, 257, 2049])
def test_quality_mode_ragged_sizes_and_duplicates(renderer, scenes, oracle, n):
    """Tiny scenes around the leaf cap, and identical triangles (equal 63-bit keys: ties are broken by position)."""
    base = scenes.displaced_sphere(40)
    tri = np.ascontiguousarray(base[:: max(1, len(base) // n)][:n])
    cam = scenes.pinhole_camera(aspect=1.0)
    rays = oracle.camera_rays(cam, 48, 48)
    for t9 in (tri, np.ascontiguousarray(np.repeat(tri[:max(1, n // 4)], 4, axis=0)[:n])):
        ob = oracle.Bih(t9)
        quality(renderer, 4).load_models(t9).build()
        check_vs_brute(renderer, ob, rays)


def test_quality_mode_many_identical_triangles(renderer, scenes, oracle):
    one = np.array([[-0.5, -0.5, 0.2, 0.5, -0.5, 0.2, 0.0, 0.6, 0.2]], np.float32)
    tri = np.ascontiguousarray(np.concatenate([np.repeat(one, 1000, axis=0), scenes.displaced_sphere(16)]))
    ob = oracle.Bih(tri)
    rays = oracle.camera_rays(scenes.pinhole_camera(aspect=1.0), 64, 64)
    quality(renderer, 4).load_models(tri).build()
    cnt = check_vs_brute(renderer, ob, rays)
    assert cnt["max_stack"] >= 1


def test_quality_mode_frame_and_switching_back(renderer, scenes, oracle):
    """render() in quality mode = the brute-force frame; switching the option back gives the parity tree again; the
    parity-only entry points refuse a quality tree."""
    import bihrt
    tri = scenes.displaced_sphere(96)
    cam = scenes.pinhole_camera(aspect=96 / 54)
    ob = oracle.Bih(tri)
    quality(renderer).load_models(tri).build()
    fb = renderer.render(cam, 96, 54, spp=4, jitter=True).framebuffer()
    rays = oracle.camera_rays(cam, 96, 54, spp=4, jitter=True)
    _, s0, _ = ob.trace(rays, "brute")
    np.testing.assert_array_equal(fb.ravel(), oracle.pack_framebuffer(s0, 96, 54, 4))
    with pytest.raises(bihrt.BihrtError):
        renderer.reference_view()
    with pytest.raises(bihrt.BihrtError):
        renderer.refit()
    # blob replication keeps the kind
    import torch
    nbytes = renderer.bih_blob_bytes()
    blob = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    renderer.bih_export(blob, nbytes); renderer.sync()
    r2 = bihrt.Renderer(0)
    r2.bih_import(blob, nbytes)
    fb2 = r2.render(cam, 96, 54, spp=4, jitter=True).framebuffer()
    np.testing.assert_array_equal(fb2, fb)
    r2.close()
    # a replica adopted in place by a context that was not told the kind is refused, not traced
    r3 = bihrt.Renderer(0)
    ptr, nb = r3.bih_region(len(tri))
    src_ptr, _ = renderer.bih_region(len(tri))
    t_src = torch.as_tensor(bihrt.multi._CudaView(src_ptr, (nb,), "|u1"), device="cuda")
    t_dst = torch.as_tensor(bihrt.multi._CudaView(ptr, (nb,), "|u1"), device="cuda")
    renderer.sync()
    t_dst.copy_(t_src); torch.cuda.synchronize()
    r3.bih_adopt(len(tri))
    with pytest.raises(bihrt.BihrtError):
        r3.render(cam, 96, 54, spp=4, jitter=True).framebuffer()
    r3.set_option("morton_bits", 63)
    r3.bih_adopt(len(tri))
    np.testing.assert_array_equal(r3.render(cam, 96, 54, spp=4, jitter=True).framebuffer(), fb)
    r3.close()
    # back to the reference's tree
    renderer.set_option("morton_bits", 30)
    renderer.build()
    v = renderer.reference_view()
    np.testing.assert_array_equal(v["children"], ob.children)
    np.testing.assert_array_equal(v["clip_planes"], ob.clip)


def test_quality_mode_one_million_triangles(renderer, scenes, oracle):
    """Size-independent property at 1 M triangles: quality-mode hits == the parity path's hits wherever both are unambiguous
    (both are exact closest hits of the same triangles), and t agrees bit for bit on every ray."""
    tri = scenes.displaced_sphere(scenes.SPHERE_NSEG["1m"])
    cam = scenes.pinhole_camera()
    rays = oracle.camera_rays(cam, 960, 540)
    renderer.load_models(tri).build()
    t0, s0, p0 = renderer.trace(rays)
    quality(renderer).build()
    t1, s1, p1, cnt = renderer.trace(rays, counted=True)
    np.testing.assert_array_equal(t1, t0)
    assert (p1 != p0).sum() <= ID_MISMATCH_MAX * len(rays)
    info = renderer.build_info()
    assert info["nu"] >= len(tri) // 4
