"""ctypes binding of the CPU oracle (oracle/bih_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module; the product package (bih-gpu-raytracer_b200/bihrt) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbih_oracle.so")
_lib = None


def build(force=False):
    """Compile oracle/libbih_oracle.so with gcc (a few hundred ms)."""
    src = os.path.join(_HERE, "bih_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libbih_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_build.restype = C.c_int64
        _lib.orc_rle.restype = C.c_int64
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Bih:
    """Result of orc_build: the reference's array model, trimmed to Nu."""

    def __init__(self, tri9):
        tri9 = np.ascontiguousarray(tri9, dtype=np.float32).reshape(-1, 9)
        n = tri9.shape[0]
        self.tri9, self.n = tri9, n
        m = max(n, 1)
        self.lo = np.zeros((m, 3), np.float32)
        self.hi = np.zeros((m, 3), np.float32)
        self.cnorm = np.zeros((m, 3), np.float32)
        self.scene_lo = np.zeros(3, np.float32)
        self.scene_hi = np.zeros(3, np.float32)
        self.codes = np.zeros(m, np.uint32)          # sorted Morton codes
        self.tris_idx = np.zeros(m, np.uint32)       # slot -> prim
        self.umc = np.zeros(m, np.uint32)
        self.cnt = np.zeros(m, np.uint32)
        self.first = np.zeros(m, np.int32)
        self.clip = np.zeros((m, 2), np.float32)
        self.axis = np.zeros(m, np.int32)
        self.is_leaf = np.zeros((m, 2), np.uint8)
        self.children = np.zeros((m, 2), np.int32)
        self.parent = np.zeros(m, np.int32)
        self.leaf_parents = np.zeros(m, np.int32)
        self.nu = int(lib().orc_build(
            _p(tri9), C.c_int64(n), _p(self.lo), _p(self.hi), _p(self.cnorm), _p(self.scene_lo),
            _p(self.scene_hi), _p(self.codes), _p(self.tris_idx), _p(self.umc), _p(self.cnt),
            _p(self.first), _p(self.clip), _p(self.axis), _p(self.is_leaf), _p(self.children),
            _p(self.parent), _p(self.leaf_parents)))
        nu, ni = self.nu, max(self.nu - 1, 0)
        self.codes, self.tris_idx = self.codes[:n], self.tris_idx[:n]
        self.umc, self.cnt, self.first = self.umc[:nu], self.cnt[:nu], self.first[:nu]
        self.leaf_parents = self.leaf_parents[:nu]
        self.clip, self.axis, self.is_leaf = self.clip[:ni], self.axis[:ni], self.is_leaf[:ni]
        self.children, self.parent = self.children[:ni], self.parent[:ni]

    def trace(self, rays6, mode="ref", threads=0, want_counters=False):
        """mode: 'ref' (literal TraverseTree), 'proper' (pruned, same results), 'box' (proper + children boxes: what the
        shipped kernel does, bit for bit), 'brute'.  Returns t, slot, prim[, counters]."""
        rays6 = np.ascontiguousarray(rays6, dtype=np.float32).reshape(-1, 6)
        nr = rays6.shape[0]
        t = np.empty(nr, np.float32)
        slot = np.empty(nr, np.int32)
        prim = np.empty(nr, np.int32)
        counters = np.zeros(3, np.uint64)
        keep = [np.ascontiguousarray(x) for x in (self.tris_idx, self.cnt, self.first, self.clip,
                                                 self.axis, self.is_leaf, self.children)]
        lib().orc_trace(C.c_int({"ref": 0, "proper": 1, "brute": 2, "box": 3}[mode]), _p(self.tri9),
                        C.c_int64(self.n), C.c_int64(self.nu), *[_p(k) for k in keep],
                        _p(self.scene_lo), _p(self.scene_hi), _p(rays6), C.c_int64(nr), _p(t),
                        _p(slot), _p(prim), _p(counters), C.c_int(threads))
        if want_counters:
            return t, slot, prim, {"nodes": int(counters[0]), "tris": int(counters[1]),
                                   "max_stack": int(counters[2])}
        return t, slot, prim


def camera_rays(cam12, w, h, spp=1, jitter=False, seed=1984):
    cam12 = np.ascontiguousarray(cam12, dtype=np.float32).reshape(12)
    rays = np.empty((h * w * spp, 6), np.float32)
    lib().orc_camera_rays(_p(cam12), C.c_int(w), C.c_int(h), C.c_int(spp), C.c_int(int(jitter)),
                          C.c_uint64(seed), _p(rays))
    return rays


def camera_rays_window(cam12, w, h, x0, y0, ww, wh, spp=1, jitter=False, seed=1984):
    """Rays of the pixels [x0, x0+ww) x [y0, y0+wh) of the w x h frame (frame pixel ids / u,v), window-row-major."""
    cam12 = np.ascontiguousarray(cam12, dtype=np.float32).reshape(12)
    rays = np.empty((wh * ww * spp, 6), np.float32)
    lib().orc_camera_rays_window(_p(cam12), C.c_int(w), C.c_int(h), C.c_int(spp), C.c_int(int(jitter)), C.c_uint64(seed),
                                 C.c_int(x0), C.c_int(y0), C.c_int(ww), C.c_int(wh), _p(rays))
    return rays


def pack_framebuffer(hit_slot, w, h, spp):
    hit_slot = np.ascontiguousarray(hit_slot, dtype=np.int32)
    fb = np.empty(w * h, np.uint32)
    lib().orc_pack_framebuffer(_p(hit_slot), C.c_int(w), C.c_int(h), C.c_int(spp), _p(fb))
    return fb


def max_threads():
    return int(lib().orc_max_threads())
