"""Device-side secondary-ray generation: the stage after the path.  The generator has no
reference counterpart (Color() is a stub, R/src/CUDAKernels.cu:370-389); what is checked is that it does what
include/bihrt.h says and that tracing ITS rays matches the oracle bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _expected_origins(tri, cam, w, h, spp, jitter, oracle, t, prim, src):
    rays = oracle.camera_rays(cam, w, h, spp=spp, jitter=jitter, seed=1984)[src]
    P = rays[:, :3] + t[src, None] * rays[:, 3:]
    v = tri[prim[src]].reshape(-1, 3, 3)
    n = np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]).astype(np.float64)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    return P + 1e-3 * n, n


@pytest.mark.parametrize("scene", ["sphere", "atrium"])
def test_secondary_rays(renderer, scenes, oracle, scene):
    if scene == "sphere":
        tri, cam, w, h = scenes.displaced_sphere(96), scenes.pinhole_camera(aspect=160 / 90), 160, 90
    else:
        tri, cam, w, h = scenes.atrium(0.1), scenes.atrium_camera(160 / 90), 160, 90
    spp, jitter = 2, True
    ob = oracle.Bih(tri)
    renderer.load_models(tri).build()
    t, slot, prim = renderer.render_hits(cam, w, h, spp=spp, jitter=jitter)
    hits = np.nonzero(slot >= 0)[0]
    light = (0.1, 0.8, -0.2)
    for kind in ("shadow", "diffuse"):
        rays_d, src_d = renderer.secondary_rays(cam, w, h, spp=spp, kind=kind, light=light, jitter=jitter)
        renderer.sync()
        rays, src = rays_d.cpu().numpy(), src_d.cpu().numpy()
        # compacted, in sample order, one ray per hit sample
        np.testing.assert_array_equal(src, hits)
        P, n = _expected_origins(tri, cam, w, h, spp, jitter, oracle, t, prim, src)
        np.testing.assert_allclose(rays[:, :3], P, rtol=1e-4, atol=2e-5)
        if kind == "shadow":
            np.testing.assert_allclose(rays[:, 3:], np.asarray(light)[None, :] - P, rtol=1e-4, atol=2e-5)
        else:
            d = rays[:, 3:].astype(np.float64)
            np.testing.assert_allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-4)
            assert np.all((d * n).sum(1) > -1e-4)                      # upper hemisphere about the normal
            cos = (d * n).sum(1)
            assert 0.55 < cos.mean() < 0.78                            # cosine-weighted: E[cos] = 2/3
            again, _ = renderer.secondary_rays(cam, w, h, spp=spp, kind=kind, light=light, jitter=jitter)
            np.testing.assert_array_equal(again.cpu().numpy(), rays)   # counter-based: deterministic
        # the rays it emits, traced on the device, match the oracle's literal reference traversal exactly
        tg, sg, pg = renderer.trace(rays_d)
        renderer.sync()
        t0, s0, p0 = ob.trace(rays, "ref")
        np.testing.assert_array_equal(sg.cpu().numpy(), s0)
        np.testing.assert_array_equal(tg.cpu().numpy(), t0)
        np.testing.assert_array_equal(pg.cpu().numpy(), p0)
        # occlusion query: some hit before tmax <=> the closest hit is before tmax (the light sits at t = 1 on shadow rays)
        for tmax in (1.0, 0.25, 3.0):
            blk_d = renderer.trace_any(rays_d, tmax=tmax)
            renderer.sync()                                   # device outputs are asynchronous on the context's stream
            blk = blk_d.cpu().numpy()
            np.testing.assert_array_equal(blk >= 0, (s0 >= 0) & (t0 < np.float32(tmax)))
            assert blk.max() < len(tri)
        np.testing.assert_array_equal(renderer.trace_any(rays, tmax=1.0) >= 0, (s0 >= 0) & (t0 < 1.0))      # host buffers


@pytest.mark.parametrize("n_rays", [5000, 70001, 1 << 20])
def test_sorted_ray_lists_give_the_same_results_in_list_order(renderer, scenes, oracle, n_rays):
    """Option trace_sort_rays: the list is traced through a permutation (origin cell + direction octant, 3 onesweep passes) but
    every result lands in the ray's own slot: outputs equal the unsorted trace's -- and the oracle's -- bit for bit,
    for closest hits, occlusion queries, counters (same rays, same per-ray work), device and host buffers."""
    import torch
    tri = scenes.atrium(0.2)
    ob = oracle.Bih(tri)
    renderer.load_models(tri).build()
    rng = np.random.default_rng(n_rays)
    # incoherent rays: origins all over the hall (some outside the scene box), random directions, a few axis-parallel
    o = rng.uniform([-2.2, -1.1, -1.1], [2.2, 1.1, 1.1], (n_rays, 3))
    d = rng.normal(size=(n_rays, 3))
    d[::97, 0] = 0.0
    rays = np.ascontiguousarray(np.concatenate([o, d], 1), dtype=np.float32)
    t0, s0, p0, c0 = renderer.trace(rays, counted=True)
    renderer.set_option("trace_sort_rays", 1)
    t1, s1, p1, c1 = renderer.trace(rays, counted=True)
    t2, s2, p2 = renderer.trace(rays)
    d_rays = torch.from_numpy(rays).cuda()
    t3, s3, p3 = renderer.trace(d_rays)
    blk = renderer.trace_any(d_rays, tmax=0.7)
    renderer.set_option("trace_sort_rays", 0)
    blk0 = renderer.trace_any(d_rays, tmax=0.7)
    for (t, s, p) in ((t1, s1, p1), (t2, s2, p2), (t3.cpu().numpy(), s3.cpu().numpy(), p3.cpu().numpy())):
        np.testing.assert_array_equal(s, s0)
        np.testing.assert_array_equal(t, t0)
        np.testing.assert_array_equal(p, p0)
    assert c1["nodes"] == c0["nodes"] and c1["tris"] == c0["tris"]
    np.testing.assert_array_equal(blk.cpu().numpy() >= 0, blk0.cpu().numpy() >= 0)
    if n_rays <= 70001:
        tr, sr, pr = ob.trace(rays, "ref")
        np.testing.assert_array_equal(s0, sr)
        np.testing.assert_array_equal(t0, tr)
