"""ctypes binding of oracle/_ref/libref_harness.so: the reference's own CUDA kernels (BuildTree,
FindClipPlanes, TraverseTree), compiled unmodified for sm_100a.  TEST INFRASTRUCTURE ONLY (GPU box)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libref_harness.so")


def available():
    return os.path.exists(LIB)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefScene:
    def __init__(self, ob):
        """ob: oracle.Bih (its prep arrays are what App::LoadModels uploads, R/src/App.cpp:158-164)."""
        self.lib = C.CDLL(LIB)
        self.lib.refh_create.restype = C.c_void_p
        self.lib.refh_trace.restype = C.c_float
        self.lib.refh_last_build_ms.restype = C.c_float
        self.h = C.c_void_p(self.lib.refh_create())
        self.n = ob.n
        keep = [np.ascontiguousarray(x, np.float32) for x in (ob.tri9, ob.lo, ob.hi, ob.cnorm, ob.scene_lo, ob.scene_hi)]
        rc = self.lib.refh_load(self.h, *[_p(k) for k in keep], C.c_int(ob.n))
        assert rc == 0

    def build(self):
        self.nu = self.lib.refh_build(self.h)
        assert self.nu >= 0, "reference kernels failed"
        return self.nu

    def build_ms(self):
        return float(self.lib.refh_last_build_ms(self.h))

    def export(self):
        n, m = self.n, max(self.n, 1)
        out = {"morton_codes": np.zeros(m, np.uint32), "tris_indexes": np.zeros(m, np.uint32),
               "unique_morton_codes": np.zeros(m, np.uint32), "duplicates_cnts": np.zeros(m, np.uint32),
               "first_idxs": np.zeros(m, np.int32), "clip_planes": np.zeros((m, 2), np.float32),
               "axis": np.zeros(m, np.int32), "is_leaf": np.zeros((m, 2), np.uint8),
               "children": np.zeros((m, 2), np.int32), "parent": np.zeros(m, np.int32),
               "leaf_parents": np.zeros(m, np.int32)}
        o = out
        nu = self.lib.refh_export(self.h, _p(o["morton_codes"]), _p(o["tris_indexes"]), _p(o["unique_morton_codes"]),
                                  _p(o["duplicates_cnts"]), _p(o["first_idxs"]), _p(o["clip_planes"]), _p(o["axis"]),
                                  _p(o["is_leaf"]), _p(o["children"]), _p(o["parent"]), _p(o["leaf_parents"]))
        ni = max(nu - 1, 0)
        trim = {"morton_codes": n, "tris_indexes": n, "unique_morton_codes": nu, "duplicates_cnts": nu, "first_idxs": nu,
                "leaf_parents": nu, "clip_planes": ni, "axis": ni, "is_leaf": ni, "children": ni, "parent": ni}
        res = {k: out[k][:trim[k]] for k in out}
        res.update(n=n, nu=nu)
        return res

    def trace(self, rays6, reps=1):
        rays6 = np.ascontiguousarray(rays6, np.float32).reshape(-1, 6)
        t = np.empty(len(rays6), np.float32)
        slot = np.empty(len(rays6), np.int32)
        ms = self.lib.refh_trace(self.h, _p(rays6), C.c_int(len(rays6)), _p(t), _p(slot), C.c_int(reps))
        assert ms >= 0, "reference trace kernel failed"
        return t, slot, float(ms)

    def close(self):
        if self.h:
            self.lib.refh_destroy(self.h)
            self.h = None
