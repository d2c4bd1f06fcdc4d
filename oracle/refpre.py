"""ctypes loader for oracle/_ref/libdvins_refpre*.so - the REFERENCE's own CUDA pre/post-processing kernels compiled
from /root/reference by oracle/ref_pre/build_ref.py (TEST INFRASTRUCTURE; needs a GPU to run; only tests/ import it).

`RefPre("ref")`  : built with the reference's nvcc flags (--use_fast_math, loop_fusion/CMakeLists.txt:92)
`RefPre("ieee")` : same sources, -fmad=false and no fast-math = the arithmetic the fp32 oracle restates
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def lib_path(kind: str = "ref") -> str:
    return os.path.join(_DIR, "libdvins_refpre.so" if kind == "ref" else "libdvins_refpre_ieee.so")


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class RefPre:
    def __init__(self, kind: str = "ref"):
        self.lib = C.CDLL(lib_path(kind))

    def sp(self, img_u8: np.ndarray, d2i: np.ndarray, h_adj: int, w_adj: int) -> np.ndarray:
        img = np.ascontiguousarray(img_u8, np.uint8)
        rows, cols = img.shape[:2]
        ch = 1 if img.ndim == 2 else img.shape[2]
        out = np.zeros((h_adj, w_adj), np.float32)
        m = np.ascontiguousarray(d2i, np.float32)
        rc = self.lib.refpre_sp(_p(img, C.c_uint8), ch, rows, cols, h_adj, w_adj, _p(m, C.c_float), _p(out, C.c_float))
        assert rc == 0, rc
        return out

    def mix(self, img_bgr_u8: np.ndarray, d2i: np.ndarray) -> np.ndarray:
        img = np.ascontiguousarray(img_bgr_u8, np.uint8)
        assert img.ndim == 3 and img.shape[2] == 3      # the caller cvtColor's gray to BGR first (deep_net.cpp:1259)
        rows, cols = img.shape[:2]
        out = np.zeros((3, 320, 320), np.float32)
        m = np.ascontiguousarray(d2i, np.float32)
        rc = self.lib.refpre_mix(_p(img, C.c_uint8), 3, rows, cols, _p(m, C.c_float), _p(out, C.c_float))
        assert rc == 0, rc
        return out

    def normalize_kpts(self, kpts: np.ndarray, shift_w: float, shift_h: float, scale: float) -> np.ndarray:
        k = np.ascontiguousarray(kpts)
        out = np.zeros(k.shape, np.float32)
        if k.dtype == np.int32:
            rc = self.lib.refpre_normalize_kpts_i32(_p(k, C.c_int32), k.size, C.c_float(shift_w), C.c_float(shift_h),
                                                    C.c_float(scale), _p(out, C.c_float))
        else:
            k = k.astype(np.float32)
            rc = self.lib.refpre_normalize_kpts_f32(_p(k, C.c_float), k.size, C.c_float(shift_w), C.c_float(shift_h),
                                                    C.c_float(scale), _p(out, C.c_float))
        assert rc == 0, rc
        return out

    def matches_post(self, kn0, kn1, matches, shift_w, shift_h, sw0=1.0, sh0=1.0, sw1=1.0, sh1=1.0):
        kn0 = np.ascontiguousarray(kn0, np.float32); kn1 = np.ascontiguousarray(kn1, np.float32)
        ma = np.ascontiguousarray(matches, np.int32)
        k = ma.shape[0]
        mk0 = np.zeros((k, 2), np.float32); mk1 = np.zeros((k, 2), np.float32)
        rc = self.lib.refpre_matches_post(_p(kn0, C.c_float), kn0.shape[0], _p(kn1, C.c_float), kn1.shape[0],
                                          _p(ma, C.c_int32), k, C.c_float(shift_w), C.c_float(shift_h), C.c_float(sw0),
                                          C.c_float(sh0), C.c_float(sw1), C.c_float(sh1), _p(mk0, C.c_float),
                                          _p(mk1, C.c_float))
        assert rc == 0, rc
        return mk0, mk1
