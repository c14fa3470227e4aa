"""Graph-replay build time (CUDA events on the context's stream, L2 flushed before each build) per scene size.
usage: python tools/build_time.py [70k 260k 1m 10m] [option=value ...]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bih-gpu-raytracer_b200"))
import bihrt
from bihrt import scenes
r = bihrt.Renderer(0)
stream = torch.cuda.Stream()
r.set_stream(stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for kv in [a for a in sys.argv[1:] if "=" in a]:
    r.set_option(kv.split("=")[0], int(kv.split("=")[1]))
for key in ([a for a in sys.argv[1:] if "=" not in a] or ["70k", "260k", "1m", "10m"]):
    tri = scenes.displaced_sphere(scenes.SPHERE_NSEG[key])
    r.load_models(torch.from_numpy(tri).cuda())
    ts = []
    with torch.cuda.stream(stream):
        for it in range(12):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); r.build(); e1.record(stream); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts[3:]))
    print("%s n=%d build %.4f ms = %.4f ms/Mtri (min %.4f)" % (key, len(tri), ms, ms / (len(tri) / 1e6), min(ts[3:]) / (len(tri) / 1e6)), flush=True)
