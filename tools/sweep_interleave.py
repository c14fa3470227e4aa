"""One rank's share of an N-way unit-interleaved frame, timed on ONE GPU for several run lengths
(development aid: emulates what each GPU of an N-GPU job does, at 1/N of the GPU cost)."""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bih-gpu-raytracer_b200"))
import bihrt
from bihrt import scenes
ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="1m"); ap.add_argument("--w", type=int, default=3840); ap.add_argument("--h", type=int, default=2160)
ap.add_argument("--spp", type=int, default=16); ap.add_argument("--count", type=int, default=8); ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--chunks", default="1,2,4,8,16,32,64")
a = ap.parse_args()
st = torch.cuda.Stream()
r = bihrt.Renderer(0, stream=st.cuda_stream)
tri = scenes.displaced_sphere(scenes.SPHERE_NSEG[a.scene]); cam = scenes.pinhole_camera(aspect=a.w / a.h)
r.load_models(torch.from_numpy(tri).cuda()).build(); r.sync()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn):
    ts = []
    for _ in range(a.reps):
        with torch.cuda.stream(st):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); fn(); e1.record(st)
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))
full = timed(lambda: r.render(cam, a.w, a.h, spp=a.spp, jitter=True))
print("full frame: min %.3f ms med %.3f ms -> ideal share %.3f ms" % (full[0], full[1], full[0] / a.count), flush=True)
for c in [int(x) for x in a.chunks.split(",")]:
    r.set_option("interleave_chunk", c)
    res = [timed(lambda k=k: r.render_interleaved_to(cam, a.w, a.h, a.spp, k, a.count, None, jitter=True)) for k in (0, a.count - 1)]
    print("chunk %3d: rank 0 min %.3f med %.3f   rank %d min %.3f med %.3f   (x ideal %.3f)" % (
        c, res[0][0], res[0][1], a.count - 1, res[1][0], res[1][1], max(res[0][0], res[1][0]) / (full[0] / a.count)), flush=True)
