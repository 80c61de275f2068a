"""ctypes wrapper of oracle/kb_oracle.c (TEST INFRASTRUCTURE / CPU BASELINE ONLY): the C port of
the per-iteration hot path used as ``cpu_baseline`` / ``--impl reference`` in bench.py and as a
fast oracle at sizes where oracle/epdd.py (scipy) is too slow."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkb_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "kb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", src, "-o", _SO, "-lm"])
    return _SO


def _L():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.ko_new.restype = C.c_void_p
        lib.ko_new.argtypes = [C.c_int32, C.c_int32, C.c_void_p]
        lib.ko_set_dom.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 12
        lib.ko_finish.argtypes = [C.c_void_p]
        lib.ko_free.argtypes = [C.c_void_p]
        lib.ko_apply_schur.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.ko_apply_nn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.ko_pcg.restype = C.c_int32
        lib.ko_pcg.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_int, C.c_int]
        _lib = lib
    return _lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class CpuModel:
    def __init__(self, sd, blocks, factors, PiSd=None):
        lib = _L()
        cnt = _i32(sd.node_gamma_cnt)
        self.n = sd.n_gamma
        self.m = lib.ko_new(sd.ndom, sd.n_gamma, cnt.ctypes.data)
        for d in range(sd.ndom):
            perm, L = factors[d]
            L = L.tocsc(); G = blocks.A_IG[d].tocsc(); H = blocks.A_GG[d].tocsc()
            arrs = [_i32(perm), _i32(L.indptr), _i32(L.indices), _f64(L.data), _i32(G.indptr), _i32(G.indices),
                    _f64(G.data), _i32(H.indptr), _i32(H.indices), _f64(H.data), _i32(sd.gamma_l2g[d])]
            pis = np.asfortranarray(PiSd[d], dtype=np.float64) if PiSd is not None else None
            lib.ko_set_dom(self.m, d, G.shape[0], G.shape[1], *[a.ctypes.data for a in arrs],
                           pis.ctypes.data if pis is not None else None)
        lib.ko_finish(self.m)

    def apply_schur(self, x, nthreads=1):
        x = _f64(x); y = np.empty(self.n)
        _L().ko_apply_schur(self.m, x.ctypes.data, y.ctypes.data, nthreads)
        return y

    def apply_nn(self, r, nthreads=1):
        r = _f64(r); z = np.empty(self.n)
        _L().ko_apply_nn(self.m, r.ctypes.data, z.ctypes.data, nthreads)
        return z

    def pcg(self, b, x, maxit=0, eps=1e-7, nthreads=1, precond=True):
        b = _f64(b); x = _f64(x).copy()
        cap = (maxit if maxit else self.n) + 1
        res = np.empty(cap)
        it = _L().ko_pcg(self.m, b.ctypes.data, x.ctypes.data, maxit, eps, res.ctypes.data, nthreads, int(precond))
        return x, it, res[:it].copy()

    def __del__(self):
        try:
            _L().ko_free(self.m)
        except Exception:
            pass
