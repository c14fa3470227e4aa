import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/bih-gpu-raytracer_b200")
import bihrt
from bihrt import scenes
st = torch.cuda.Stream(); r = bihrt.Renderer(0, stream=st.cuda_stream)
tri = scenes.displaced_sphere(708); cam = scenes.pinhole_camera(aspect=3840 / 2160)
r.load_models(torch.from_numpy(tri).cuda()).build(); r.sync()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def bench(fn):
    ts = []
    for _ in range(4):
        with torch.cuda.stream(st):
            flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); fn(); e1.record(st)
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
W, H = 3840, 2160
for g, q in ((1, -1), (1, 0), (1, 1)):
    r.set_option("trace_lane_groups", g); r.set_option("trace_sm_queues", q); print("queues", q)
    print("groups", g, "full 16spp %.2f ms" % bench(lambda: r.render(cam, W, H, spp=16, jitter=True)),
          "| tiles (0,2) %.2f ms (1,2) %.2f ms" % (bench(lambda: r.render(cam, W, H, spp=16, jitter=True, shard=(0, 2))), bench(lambda: r.render(cam, W, H, spp=16, jitter=True, shard=(1, 2)))),
          "| samples [0,8) %.2f ms" % bench(lambda: r.render_samples(cam, W, H, 16, 0, 8, jitter=True)),
          "| samples [0,2) %.2f ms" % bench(lambda: r.render_samples(cam, W, H, 16, 0, 2, jitter=True)))
