"""Fixed cost of a trace launch: time vs number of rays (image height), with and without the L2 flush."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bih-gpu-raytracer_b200"))
import bihrt
from bihrt import scenes
st = torch.cuda.Stream()
r = bihrt.Renderer(0, stream=st.cuda_stream)
tri = scenes.displaced_sphere(scenes.SPHERE_NSEG["1m"])
r.load_models(torch.from_numpy(tri).cuda()).build(); r.sync()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, do_flush, reps=7):
    ts = []
    for _ in range(reps):
        with torch.cuda.stream(st):
            if do_flush: flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); fn(); e1.record(st)
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
import sys as _s
for oset in (_s.argv[1] if len(_s.argv) > 1 else "").split(";"):
  for kv in filter(None, oset.split(",")):
      r.set_option(kv.split("=")[0], int(kv.split("=")[1]))
  print("options:", oset or "(default)")
  for spp in (1, 16):
    for (w, h) in ( (1920, 1080), (960, 540), (480, 270), (240, 135)):
        cam = scenes.pinhole_camera(aspect=w / h)
        fn = lambda: r.render(cam, w, h, spp=spp, jitter=spp > 1)
        fn(); r.sync()
        a, b = timed(fn, True), timed(fn, False)
        n = w * h * spp
        print("  " + "spp %2d %4dx%4d rays %9d: flushed %.3f ms (%.0f Mr/s)  warm %.3f ms (%.0f Mr/s)" % (spp, w, h, n, a, n / a / 1e3, b, n / b / 1e3), flush=True)
