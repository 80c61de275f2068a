"""ctypes wrapper of oracle/kb_oracle.c (TEST INFRASTRUCTURE / CPU BASELINE ONLY): the C port of
the per-iteration hot path used as ``cpu_baseline`` / ``--impl reference`` in bench.py and as a
fast oracle at sizes where oracle/epdd.py (scipy) is too slow."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# Arithmetic variants of the same source (rounding-floor probes of DESIGN.md section 2; the default build has none):
#   fma   fused multiply-adds everywhere (products enter the sums unrounded, as on the GPU)
#   ldot  the three inner products of a PCG iteration accumulated in extended precision
#   tdot  ... in plain doubles, but summed as a tree (the shape of the GPU's block partials)
#   bdot  ... in plain doubles with the 16 running sums of a SIMD BLAS ddot (what Julia's dot() does at least)
# KB_ORACLE_VARIANT selects the variant every call uses; CpuModel.pcg(variant=...) runs one solve with another variant
# on the same model (the variants share every data structure).
_FLAGS = {"": [], "fma": ["-mfma", "-ffp-contract=fast"], "ldot": ["-DKO_LONG_DOT"], "tdot": ["-DKO_TREE_DOT"], "bdot": ["-DKO_BLAS_DOT"]}
_VARIANT = os.environ.get("KB_ORACLE_VARIANT", "")
_libs = {}


def _so(variant):
    return os.path.join(_HERE, "libkb_oracle" + ("_" + variant if variant else "") + ".so")


_SO = _so(_VARIANT)


def build(force=False, variant=None):
    variant = _VARIANT if variant is None else variant
    src, so = os.path.join(_HERE, "kb_oracle.c"), _so(variant)
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-fPIC", "-shared"] + _FLAGS[variant] + [src, "-o", so, "-lm"])
    return so


def _L(variant=None):
    variant = _VARIANT if variant is None else variant
    if variant not in _libs:
        lib = C.CDLL(build(variant=variant))
        lib.ko_new.restype = C.c_void_p
        lib.ko_new.argtypes = [C.c_int32, C.c_int32, C.c_void_p]
        lib.ko_set_dom.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 12
        lib.ko_finish.argtypes = [C.c_void_p]
        lib.ko_free.argtypes = [C.c_void_p]
        lib.ko_apply_schur.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.ko_apply_nn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.ko_set_reverse.argtypes = [C.c_void_p, C.c_int]
        lib.ko_set_pis.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        lib.ko_assemble_local_schurs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.ko_schur_rhs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.ko_pcg.restype = C.c_int32
        lib.ko_pcg.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_int, C.c_int]
        _libs[variant] = lib
    return _libs[variant]


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

        self.n_gd = [int(g.size) for g in sd.gamma_l2g]
        self.n_I = [int(blocks.A_IG[d].shape[0]) for d in range(sd.ndom)]

    def set_reverse_order(self, flag):
        """Rounding-floor probe: interface sums in descending subdomain order (same mathematics)."""
        _L().ko_set_reverse(self.m, int(bool(flag)))

    def set_pinv(self, PiSd):
        """NeumannNeumannSchurPreconditioner blocks from another (e.g. the median) realization."""
        for d, P in enumerate(PiSd):
            P = np.asfortranarray(P, dtype=np.float64)
            _L().ko_set_pis(self.m, d, P.ctypes.data)

    def assemble_local_schurs(self, nthreads=1):
        """assemble_local_schurs (:667-695): list of dense symmetric S_d."""
        off = np.concatenate([[0], np.cumsum([n * n for n in self.n_gd])]).astype(np.int64)
        buf = np.empty(int(off[-1]))
        _L().ko_assemble_local_schurs(self.m, buf.ctypes.data, off.ctypes.data, nthreads)
        return [buf[off[d]:off[d + 1]].reshape(self.n_gd[d], self.n_gd[d], order="F") for d in range(len(self.n_gd))]

    def schur_rhs(self, b_I, b_G, nthreads=1):
        """get_schur_rhs (:835-864)."""
        off = np.concatenate([[0], np.cumsum(self.n_I)]).astype(np.int64)
        cat = _f64(np.concatenate(b_I)) if len(b_I) else np.zeros(0)
        bG = _f64(b_G)
        out = np.empty(self.n)
        _L().ko_schur_rhs(self.m, cat.ctypes.data, off.ctypes.data, bG.ctypes.data, out.ctypes.data, nthreads)
        return out

    def apply_schur(self, x, nthreads=1):
        x = _f64(x); y = np.empty(self.n)
        _L().ko_apply_schur(self.m, x.ctypes.data, y.ctypes.data, nthreads)
        return y

    def apply_nn(self, r, nthreads=1):
        r = _f64(r); z = np.empty(self.n)
        _L().ko_apply_nn(self.m, r.ctypes.data, z.ctypes.data, nthreads)
        return z

    def pcg(self, b, x, maxit=0, eps=1e-7, nthreads=1, precond=True, variant=None):
        b = _f64(b); x = _f64(x).copy()
        cap = (maxit if maxit else self.n) + 1
        res = np.empty(cap)
        it = _L(variant).ko_pcg(self.m, b.ctypes.data, x.ctypes.data, maxit, eps, res.ctypes.data, nthreads, int(precond))
        return x, it, res[:it].copy()

    def __del__(self):
        try:
            _L().ko_free(self.m)
        except Exception:
            pass
