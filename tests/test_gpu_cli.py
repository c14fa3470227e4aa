"""The C host program (bih-gpu-raytracer_b200/host/bihrt_cli.c): the reference's main / LoadModels / frame loop
on top of the C ABI, checked against the Python mirror pixel for pixel."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "bih-gpu-raytracer_b200", "bihrt_cli")


def read_ppm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"P6"
        w, h = map(int, f.readline().split())
        assert int(f.readline()) == 255
        return np.frombuffer(f.read(), np.uint8).reshape(h, w, 3)


@pytest.mark.skipif(not os.path.exists(CLI), reason="bihrt_cli not built")
def test_cli_renders_like_the_library(renderer, scenes, tmp_path):
    tri = (scenes.displaced_sphere(48) * np.float32(0.8) + np.tile(np.float32([2.4, 0.0, 0.0]), 3)).astype(np.float32)
    raw = tmp_path / "mesh.tri9"
    tri.tofile(raw)
    out = tmp_path / "frame.ppm"
    w, h, spp = 160, 120, 2
    p = subprocess.run([CLI, str(raw), "-w", str(w), "-h", str(h), "-s", str(spp), "-o", str(out)],
                       capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "%d triangles" % len(tri) in p.stdout
    img = read_ppm(out)
    cam = scenes.reference_camera(aspect=w / h)                      # Camera((2,0,-2), W/H), R/src/Renderer.cpp:99
    fb = renderer.load_models(tri).build().render(cam, w, h, spp=spp, seed=1984, jitter=True).framebuffer()
    exp = np.stack([fb & 255, (fb >> 8) & 255, (fb >> 16) & 255], -1).astype(np.uint8)[::-1]   # row 0 = bottom
    np.testing.assert_array_equal(img, exp)
    assert (img[..., 2] == 0).any() and (img[..., 2] == 40).any()    # both hits (yellow) and misses in the frame


@pytest.mark.skipif(not os.path.exists(CLI), reason="bihrt_cli not built")
def test_cli_two_gpus_render_the_same_frame(scenes, tmp_path):
    """-g 2: BIH peer-copied to the second device, both trace kernels store into the first device's framebuffer."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    tri = (scenes.displaced_sphere(64) * np.float32(0.8) + np.tile(np.float32([2.4, 0.0, 0.0]), 3)).astype(np.float32)
    raw = tmp_path / "mesh.tri9"
    tri.tofile(raw)
    imgs = []
    for k, extra in enumerate((["-g", "1"], ["-g", "2"], ["-g", "2", "-p"])):      # one GPU; bihrt_create_multi (NCCL); peer copies
        out = tmp_path / ("frame%d.ppm" % k)
        p = subprocess.run([CLI, str(raw), "-w", "330", "-h", "200", "-s", "4", "-f", "2"] + extra + ["-o", str(out)],
                           capture_output=True, text=True, timeout=180)
        assert p.returncode == 0, p.stdout + p.stderr
        if extra == ["-g", "2"]:
            assert "bihrt_create_multi: 2 contexts" in p.stdout, p.stdout + p.stderr
        imgs.append(read_ppm(out))
    np.testing.assert_array_equal(imgs[0], imgs[1])
    np.testing.assert_array_equal(imgs[0], imgs[2])


def test_cli_reports_errors():
    p = subprocess.run([CLI, "/nonexistent.obj"], capture_output=True, text=True, timeout=60)
    assert p.returncode == 1 and "cannot open" in p.stderr
