"""Contracts around the hot path that a wrong answer would not reveal: stream ordering with torch, the build
watchdog, reuse of the tile schedule across launch geometries, the exported framebuffer, work-counter limits."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_default_context_is_ordered_with_torch(scenes, oracle):
    """Renderer() runs on its own stream; device tensors in / out are ordered with torch's current stream by
    events (bihrt_get_stream + ExternalStream), and stream handle 0 means CUDA's legacy default stream."""
    import torch
    import bihrt
    tri = scenes.displaced_sphere(187)
    cam = scenes.pinhole_camera()
    rays = oracle.camera_rays(cam, 640, 360)
    r_np = bihrt.Renderer(0)
    t0, s0, p0 = r_np.load_models(tri).build().trace(rays)
    r_np.close()
    for which in ("own", "legacy0", "torch_side_stream"):
        r = bihrt.Renderer(0)
        side = torch.cuda.Stream()
        if which == "legacy0":
            r.set_stream(torch.cuda.current_stream().cuda_stream)      # 0 on torch's default stream
            assert r.stream_handle() == 0
        elif which == "torch_side_stream":
            r.set_stream(side.cuda_stream)
        else:
            assert r.stream_handle() not in (0, side.cuda_stream)
        for rep in range(3):
            # inputs produced on torch's current stream right before the call, outputs consumed right after
            d_tri = torch.from_numpy(tri).cuda() * 1.0
            r.load_models(d_tri).build()
            d_rays = torch.from_numpy(rays).cuda() + 0.0
            t, s, p = r.trace(d_rays)
            s_sum = (s.long() + 1).sum()                   # torch kernel on torch's stream, immediately
            np.testing.assert_array_equal(s.cpu().numpy(), s0)
            np.testing.assert_array_equal(t.cpu().numpy(), t0)
            np.testing.assert_array_equal(p.cpu().numpy(), p0)
            assert int(s_sum) == int((s0.astype(np.int64) + 1).sum())
            b = r.trace_any(d_rays, tmax=1e30)
            np.testing.assert_array_equal(b.cpu().numpy() >= 0, s0 >= 0)
        r.set_stream(None)                                  # back to the private stream
        assert r.stream_handle() not in (0, side.cuda_stream)
        r.close()


def test_tripped_build_watchdog_is_reported_and_nothing_is_traced(renderer, scenes, oracle):
    """a build whose device watchdog tripped must not be traced silently."""
    import bihrt
    tri = scenes.displaced_sphere(64)
    cam = scenes.pinhole_camera()
    rays = oracle.camera_rays(cam, 64, 36)
    renderer.load_models(tri).build()
    t0, s0, _ = renderer.trace(rays)
    for graph in (1, 0):
        renderer.set_option("build_graph", graph)
        renderer.set_option("debug_trip_watchdog", 2)
        renderer.build()                                    # asynchronous: the trip is known once the build has run
        with pytest.raises(bihrt.BihrtError) as e:
            renderer.sync()
        assert e.value.code == -6 and "watchdog" in str(e.value)
        for call in (lambda: renderer.trace(rays), lambda: renderer.render(cam, 64, 36), lambda: renderer.trace_any(rays)):
            with pytest.raises(bihrt.BihrtError) as e:
                call()
            assert e.value.code == -6
        # a trace enqueued BEFORE the host could know (straight after the build) writes nothing instead of garbage
        renderer.set_option("debug_trip_watchdog", 0)
        renderer.build(); renderer.sync()
        renderer.set_option("debug_trip_watchdog", 2)
        renderer.build()
        s = np.full(len(rays), -7, np.int32)
        with pytest.raises(bihrt.BihrtError):
            renderer.trace(rays, slot=s)
        assert np.all(s == -7)                              # refused up front, or traced nothing and the copy-back was skipped
        renderer.set_option("debug_trip_watchdog", 0)
        renderer.build(); renderer.sync()
        t1, s1, _ = renderer.trace(rays)
        np.testing.assert_array_equal(s1, s0)
        np.testing.assert_array_equal(t1, t0)


def test_tile_schedule_is_not_reused_across_geometries(scenes):
    """1920xH at 1 spp and 1408xH at 3 spp collided in the old packed signature; a permutation of the wrong
    size skips tiles.  Alternate many geometries on one context and compare every frame with a fresh context's."""
    import bihrt
    tri = scenes.displaced_sphere(187)
    r = bihrt.Renderer(0)
    r.load_models(tri).build()
    r.set_option("trace_tile_order", 2)
    geoms = [(1920, 96, 1), (1408, 96, 3), (1920, 96, 1), (1408, 96, 3), (640, 360, 2), (1408, 96, 3), (352, 200, 1), (1920, 96, 1),
             (704, 100, 4), (1408, 96, 3)]
    fresh = {}
    for (w, h, spp) in geoms:
        cam = scenes.pinhole_camera(aspect=w / h)
        fb = r.render(cam, w, h, spp=spp, jitter=spp > 1).framebuffer().copy()
        if (w, h, spp) not in fresh:
            q = bihrt.Renderer(0)
            q.load_models(tri).build()
            q.set_option("trace_tile_order", 0)
            fresh[(w, h, spp)] = q.render(cam, w, h, spp=spp, jitter=spp > 1).framebuffer().copy()
            q.close()
        np.testing.assert_array_equal(fb, fresh[(w, h, spp)], err_msg=str((w, h, spp)))
    r.close()


def test_exported_framebuffer_does_not_move(renderer, scenes):
    import bihrt
    renderer.load_models(scenes.displaced_sphere(32)).build()
    cam = scenes.pinhole_camera()
    renderer.framebuffer_ipc_export(64, 36)
    p0 = renderer.framebuffer_ptr()[0]
    renderer.render(cam, 64, 36)                            # fits: fine
    with pytest.raises(bihrt.BihrtError) as e:
        renderer.render(cam, 640, 360)                      # would reallocate under the peers' feet
    assert e.value.code == -4
    assert renderer.framebuffer_ptr()[0] == p0
    renderer.framebuffer_ipc_unexport()
    renderer.render(cam, 640, 360)
    assert renderer.framebuffer().shape == (360, 640)


def test_work_counter_limits(renderer, scenes, oracle):
    """padded items = tiles x 1024 x lane groups must stay below 2^32 - 2^24: the lane groups shrink for very
    large frames, and the frame is still the oracle's."""
    tri = scenes.displaced_sphere(32)
    renderer.load_models(tri).build()
    # 16384 x 16384 x 32 spp would be 2^33 items with 32 lane groups; 8 x 8 px windows are enough to see it is refused or correct
    cam = scenes.pinhole_camera(aspect=1.0)
    w = h = 12288
    try:
        renderer.render(cam, w, h, spp=32, jitter=True)
    except Exception as ex:                                  # only an explicit refusal is acceptable
        assert "work counter" in str(ex)
        return
    fb = renderer.framebuffer()
    ob = oracle.Bih(tri)
    for (x0, y0) in ((6144 - 16, 6144 - 16), (3000, 6100), (0, 0)):
        rays = oracle.camera_rays_window(cam, w, h, x0, y0, 32, 32, spp=32, jitter=True)
        _, s, _ = ob.trace(rays, "ref")
        np.testing.assert_array_equal(fb[y0:y0 + 32, x0:x0 + 32], oracle.pack_framebuffer(s, 32, 32, 32).reshape(32, 32))
