"""Smallest program that launches the trace kernel on a named scene (for ncu captures)."""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bih-gpu-raytracer_b200"))
import bihrt
from bihrt import scenes
ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="1m"); ap.add_argument("--w", type=int, default=1920); ap.add_argument("--h", type=int, default=1080)
ap.add_argument("--spp", type=int, default=4); ap.add_argument("--reps", type=int, default=3); ap.add_argument("--opts", default="")
a = ap.parse_args()
r = bihrt.Renderer(0)
for kv in filter(None, a.opts.split(",")):
    k, v = kv.split("="); r.set_option(k, int(v))
tri = scenes.atrium() if a.scene == "atrium" else scenes.displaced_sphere(scenes.SPHERE_NSEG[a.scene])
cam = scenes.atrium_camera(a.w / a.h) if a.scene == "atrium" else scenes.pinhole_camera(aspect=a.w / a.h)
r.load_models(torch.from_numpy(tri).cuda()).build()
for _ in range(a.reps):
    r.render(cam, a.w, a.h, spp=a.spp, jitter=a.spp > 1)
r.sync()
print("ok", r.build_info())
