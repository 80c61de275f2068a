"""ctypes wrapper of hostprep/chol.c: nested-dissection ordering + sparse Cholesky of A_II,d.

Stands in for the Julia host side's ``F = cholesky(A_IIdd[d])`` -> (``F.p``, ``sparse(F.L)``)
(SuiteSparse CHOLMOD; EllipticPdeDomainDecomposition.jl:1536), which is absent from this image.
"""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libhostprep.so")
    src = os.path.join(_HERE, "chol.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-fPIC", "-shared", "-fopenmp", src, "-o", so, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = C.CDLL(build())
        i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
        f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
        lib.hp_nd_order.argtypes = [C.c_int32, i32p, i32p, f64p, f64p, C.c_int32, i32p]
        lib.hp_nd_order.restype = C.c_int
        lib.hp_chol_count.argtypes = [C.c_int32, i32p, i32p, i32p, i32p, i32p]
        lib.hp_chol_count.restype = C.c_int64
        lib.hp_chol_factor.argtypes = [C.c_int32, i32p, i32p, f64p, i32p, i32p, i32p, i32p, f64p]
        lib.hp_chol_factor.restype = C.c_int
        _LIB = lib
    return _LIB


def nd_order(A, x, y, leaf=8):
    """Geometric nested-dissection permutation (perm[new] = old) of the graph of A."""
    A = sp.csc_matrix(A)
    n = A.shape[0]
    perm = np.empty(n, dtype=np.int32)
    rc = _lib().hp_nd_order(n, A.indptr.astype(np.int32), A.indices.astype(np.int32),
                            np.ascontiguousarray(x, dtype=np.float64),
                            np.ascontiguousarray(y, dtype=np.float64), leaf, perm)
    if rc:
        raise RuntimeError(f"hp_nd_order failed ({rc})")
    return perm


def cholesky(A, perm):
    """L (CSC, sorted, diagonal first) with L L^T = A[perm, perm]; also the elimination tree."""
    A = sp.csc_matrix(A)
    A.sort_indices()
    n = A.shape[0]
    Ap, Ai = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    perm = np.ascontiguousarray(perm, dtype=np.int32)
    parent = np.empty(n, dtype=np.int32)
    cc = np.empty(n, dtype=np.int32)
    nnz = _lib().hp_chol_count(n, Ap, Ai, perm, parent, cc)
    if nnz < 0:
        raise RuntimeError("hp_chol_count failed")
    Lp = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(cc, out=Lp[1:])
    Li = np.empty(nnz, dtype=np.int32)
    Lx = np.empty(nnz, dtype=np.float64)
    rc = _lib().hp_chol_factor(n, Ap, Ai, np.ascontiguousarray(A.data, dtype=np.float64), perm,
                               parent, Lp, Li, Lx)
    if rc:
        raise np.linalg.LinAlgError(f"matrix not positive definite at pivot {rc - 1}")
    return sp.csc_matrix((Lx, Li, Lp), shape=(n, n)), parent


def mmd_order(A):
    """Minimum-degree permutation (perm[new] = old) of A' + A from SuperLU's ordering step -- the kind of ordering
    CHOLMOD's default AMD gives the reference's `cholesky(A_IIdd[d])` (EllipticPdeDomainDecomposition.jl:1536)."""
    import scipy.sparse.linalg as spla
    lu = spla.splu(sp.csc_matrix(A), permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
    return np.argsort(np.asarray(lu.perm_c)).astype(np.int32)


def factor_subdomains(A_II, points, node_Id, leaf=8, threads=None, ordering=None):
    """[(perm, L)] for every subdomain.  ordering "nd" (default): geometric nested dissection from the interior nodes'
    coordinates; "mmd": minimum degree (KB200_HOST_ORDERING selects it when the argument is not given)."""
    ordering = ordering or os.environ.get("KB200_HOST_ORDERING", "nd")
    if ordering not in ("nd", "mmd"):
        raise ValueError("ordering is 'nd' or 'mmd'")

    def one(d):
        nodes = node_Id[d]
        if A_II[d].shape[0] == 0 or A_II[d].nnz == 0:
            return np.zeros(0, np.int32), sp.csc_matrix((0, 0))
        p = mmd_order(A_II[d]) if ordering == "mmd" else nd_order(A_II[d], points[0, nodes], points[1, nodes], leaf)
        L, _ = cholesky(A_II[d], p)
        return p, L
    threads = threads or min(32, os.cpu_count() or 1)
    with ThreadPoolExecutor(threads) as ex:
        return list(ex.map(one, range(len(A_II))))
