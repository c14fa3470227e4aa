"""Per-stage device time of bihrt_build from CUDA events recorded between the kernels (option profile=1)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bih-gpu-raytracer_b200"))
import bihrt
from bihrt import scenes
NAMES = ["init+memsets", "scene_box", "morton", "sort0", "sort1", "sort2", "sort3", "rle", "reorder+boxes", "heap_up", "nodes", "end"]
r = bihrt.Renderer(0)
r.set_option("profile", 1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
OPTS = [a for a in sys.argv[1:] if "=" in a]
for kv in OPTS:
    r.set_option(kv.split("=")[0], int(kv.split("=")[1]))
for key in ([a for a in sys.argv[1:] if "=" not in a] or ["1m"]):
    tri = scenes.displaced_sphere(scenes.SPHERE_NSEG[key])
    r.load_models(torch.from_numpy(tri).cuda())
    acc = []
    for it in range(8):
        flush.zero_(); torch.cuda.synchronize()
        r.build(); r.sync()
        acc.append([r.get_stat("build_stage_ns_%d" % i) / 1e3 for i in range(12)])
    a = np.array(acc[2:])
    med = np.median(a, axis=0)
    print(key, "n=%d" % len(tri), " ".join("%s=%.1f" % (n, v) for n, v in zip(NAMES, med)), "| sum %.1f us" % med.sum(), flush=True)
