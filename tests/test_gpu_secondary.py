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
