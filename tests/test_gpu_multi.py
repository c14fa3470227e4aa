"""Two GPUs in one process (no NCCL): BIH blob replicated to the second device, frame split by unit interleave /
samples / tiles, shards summed and resolved -- bit-identical to the single-GPU image.
Skipped on a one-GPU box; the N = 2/4/8 torchrun path is exercised by bench.py, which reports the same check."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_two_gpus_render_the_same_image(scenes):
    import torch
    import bihrt
    from bihrt import multi
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    tri = scenes.displaced_sphere(128)
    cam = scenes.pinhole_camera(aspect=320 / 200)
    w, h, spp = 320, 200, 8
    r0, r1 = bihrt.Renderer(0), bihrt.Renderer(1)
    r0.load_models(tri).build()
    full = r0.render(cam, w, h, spp=spp, jitter=True).framebuffer().copy()
    # replicate: export on device 0, copy across, import on device 1
    nbytes = r0.bih_blob_bytes()
    blob0 = torch.empty(nbytes, dtype=torch.uint8, device="cuda:0")
    r0.bih_export(blob0, nbytes); r0.sync()
    blob1 = blob0.to("cuda:1")
    torch.cuda.synchronize(1)
    r1.bih_import(blob1, nbytes); r1.sync()
    for mode in ("interleave", "samples", "tiles"):
        parts = []
        for rank, r in ((0, r0), (1, r1)):
            if mode == "interleave":
                r.render_interleaved(cam, w, h, spp, rank, 2, jitter=True)
            elif mode == "samples":
                s0, s1 = multi.sample_range(spp, rank, 2)
                r.render_samples(cam, w, h, spp, s0, s1, jitter=True)
            else:
                r.render(cam, w, h, spp=spp, jitter=True, shard=(rank, 2))
            parts.append(r.framebuffer().astype(np.int64))
        total = (parts[0] + parts[1]).astype(np.uint32)
        if mode == "tiles":
            np.testing.assert_array_equal(total, full)
        else:
            r0.render_samples(cam, w, h, spp, 0, spp, jitter=True); r0.sync()      # any frame of the right size
            torch.cuda.synchronize(0)
            multi.framebuffer_tensor(r0).copy_(torch.from_numpy(total.astype(np.int32)).to("cuda:0"))
            torch.cuda.synchronize(0)
            np.testing.assert_array_equal(r0.framebuffer_resolve(spp).framebuffer(), full)
    # in-place replication: the blob region of device 0 copied straight into device 1's region, then adopted
    r2 = bihrt.Renderer(1)
    p0, nb0 = r0.bih_region(len(tri))
    p2, nb2 = r2.bih_region(len(tri))
    assert nb0 == nb2 == 64 + (64 + 48) * len(tri)           # header | 64-byte node slots | 48-byte triangle records
    src = torch.as_tensor(multi._CudaView(p0, (nb0,), "|u1"), device="cuda:0")
    dst = torch.as_tensor(multi._CudaView(p2, (nb2,), "|u1"), device="cuda:1")
    r0.sync(); dst.copy_(src); torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    r2.bih_adopt(len(tri))
    np.testing.assert_array_equal(r2.render(cam, w, h, spp=spp, jitter=True).framebuffer(), full)
    r2.close()
    # fused gather: device 1's trace kernel stores its pixels straight into device 0's framebuffer (peer access)
    r0.render(cam, w, h, spp=1); r0.sync()
    ptr0 = r0.framebuffer_ptr()[0]
    multi.framebuffer_tensor(r0).zero_()
    torch.cuda.synchronize(0)
    r0.render_interleaved_to(cam, w, h, spp, 0, 2, None, jitter=True)
    r1.render_interleaved_to(cam, w, h, spp, 1, 2, ptr0, jitter=True)
    r0.sync(); r1.sync()
    np.testing.assert_array_equal(r0.framebuffer(), full)
    r0.close(); r1.close()


@pytest.mark.parametrize("spp,count", [(8, 2), (16, 8), (4, 4), (1, 2), (3, 4), (6, 8)])
def test_interleaved_to_one_framebuffer_is_the_full_frame(scenes, spp, count):
    """The fused-gather partition on ONE device: `count` launches, each owning every count-th unit, write final
    colours into the same framebuffer without clearing it; together they must reproduce bihrt_render exactly."""
    import torch
    import bihrt
    from bihrt import multi
    tri = scenes.displaced_sphere(96)
    cam = scenes.pinhole_camera(aspect=333 / 190)
    w, h = 333, 190                        # ragged: edge tiles with padding pixels
    r = bihrt.Renderer(0)
    r.load_models(tri).build()
    full = r.render(cam, w, h, spp=spp, jitter=True).framebuffer().copy()
    multi.framebuffer_tensor(r).fill_(0x55555555)
    torch.cuda.synchronize()
    for k in range(count):
        r.render_interleaved_to(cam, w, h, spp, k, count, None, jitter=True)
    np.testing.assert_array_equal(r.framebuffer(), full)
    r.close()


@pytest.mark.parametrize("quality", [False, True])
def test_create_multi_nccl_group_renders_the_same_image(scenes, quality):
    """bihrt_create_multi: N contexts + one NCCL communicator inside the library.  Context 0 builds,
    bihrt_multi_broadcast replicates the blob with one ncclBroadcast, bihrt_multi_render = the single-GPU frame, bit for bit,
    for several frames with a rebuild in between (stream ordering between the devices)."""
    import torch
    import bihrt
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n = min(torch.cuda.device_count(), 4)
    n = 2 if n == 3 else n
    w, h, spp = 330, 200, 8
    cam = scenes.pinhole_camera(aspect=w / h)
    single = bihrt.Renderer(0)
    m = bihrt.MultiRenderer(n)
    assert m.nccl_version() > 0
    if quality:
        single.set_option("morton_bits", 63)
        m.ctx[0].set_option("morton_bits", 63)
    for frame, nseg in enumerate((96, 128, 96)):
        tri = scenes.displaced_sphere(nseg, phase=0.3 * frame)
        want = single.load_models(tri).build().render(cam, w, h, spp=spp, seed=7 + frame, jitter=True).framebuffer().copy()
        m.ctx[0].load_models(tri).build()
        m.broadcast().render(cam, w, h, spp=spp, seed=7 + frame, jitter=True)
        np.testing.assert_array_equal(m.framebuffer(), want)
    m.sync()
    m.close()
    single.close()
