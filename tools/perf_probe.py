"""Quick GPU timing of build + render for the synthetic configs (development aid, not the bench)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "bih-gpu-raytracer_b200"))
import bihrt
from bihrt import scenes


STREAM = None


def timed(fn, reps, flush=None):
    ts = []
    for _ in range(reps):
        with torch.cuda.stream(STREAM):
            if flush is not None:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(STREAM)
            fn()
            e1.record(STREAM)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return np.array(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", default="70k,260k,1m")
    ap.add_argument("--w", type=int, default=1920)
    ap.add_argument("--h", type=int, default=1080)
    ap.add_argument("--spp", default="1,4")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--opts", default="")
    a = ap.parse_args()
    torch.cuda.init()
    r = bihrt.Renderer(0)
    global STREAM
    STREAM = torch.cuda.Stream()
    r.set_stream(STREAM.cuda_stream)
    for kv in filter(None, a.opts.split(",")):
        k, v = kv.split("=")
        r.set_option(k, int(v))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cam = scenes.pinhole_camera(aspect=a.w / a.h)
    for name in a.scenes.split(","):
        if name == "atrium":
            tri, c = scenes.atrium(), scenes.atrium_camera(a.w / a.h)
        elif name.startswith("soup"):
            tri, c = scenes.random_soup(260000), cam
        else:
            tri, c = scenes.displaced_sphere(scenes.SPHERE_NSEG[name]), cam
        d = torch.from_numpy(tri).cuda()
        r.load_models(d)
        r.build(); r.sync()
        tb = timed(lambda: r.build(), a.reps, flush)
        info = r.build_info()
        n = info["n"]
        print("%-7s n=%d nu=%d build: min %.3f ms med %.3f ms -> %.3f ms/Mtri (min)" % (
            name, n, info["nu"], tb.min(), np.median(tb), tb.min() / (n / 1e6)), flush=True)
        for spp in [int(s) for s in a.spp.split(",")]:
            nr = a.w * a.h * spp
            r.render(c, a.w, a.h, spp=spp, jitter=spp > 1); r.sync()
            tt = timed(lambda: r.render(c, a.w, a.h, spp=spp, jitter=spp > 1), a.reps, flush)
            print("        render %dx%d spp=%d: min %.3f ms med %.3f ms -> %.1f Mrays/s (min) %.1f (med)" % (
                a.w, a.h, spp, tt.min(), np.median(tt), nr / tt.min() / 1e3, nr / np.median(tt) / 1e3), flush=True)
        # counters on a 1/16 sample of the primary rays
        rays = scenes.camera_rays(c, a.w // 4, a.h // 4)
        _, s, _, cnt = r.trace(rays, counted=True)
        print("        per ray: nodes %.1f tris %.1f max_stack %d hit %.3f" % (
            cnt["nodes"] / len(rays), cnt["tris"] / len(rays), cnt["max_stack"], (s >= 0).mean()), flush=True)


if __name__ == "__main__":
    main()
