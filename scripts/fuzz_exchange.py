"""Fuzzing of the owner-ordered compact interface exchange (csrc/xchg_tables.h + the emulation of exchange_gather in
tests/planemu) on synthetic topologies: random numbers of subdomains and interface nodes, every node touched by 1..8
randomly chosen subdomains (not necessarily neighbours in the numbering), 1..8 ranks with the library's contiguous
sharding; every rank must assemble exactly the bits of the sequential sum over ascending subdomains
(EllipticPdeDomainDecomposition.jl:740-742).  No GPU needed.   usage: fuzz_exchange.py [seconds=180] [first_seed=1]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from planemu.build import build as build_emu  # noqa: E402

lib = C.CDLL(build_emu())
f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
lib.planemu_exchange.restype = C.c_int64
lib.planemu_exchange.argtypes = [C.c_int, C.c_int64, C.c_int64, np.ctypeslib.ndpointer(np.int64, flags="C"),
                                 np.ctypeslib.ndpointer(np.int32, flags="C"), f64p, C.c_void_p, C.c_int, f64p,
                                 np.ctypeslib.ndpointer(np.int64, flags="C")]
t_end = time.time() + (float(sys.argv[1]) if len(sys.argv) > 1 else 180.0)
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ncase = nrun = nfail = 0
while time.time() < t_end:
    rng = np.random.default_rng(seed)
    ndom = int(rng.integers(1, 41))
    ng = int(rng.integers(1, 3000))
    maxt = int(rng.integers(1, min(8, ndom) + 1))
    touch = [[] for _ in range(ndom)]
    for g in range(ng):
        k = int(rng.integers(1, maxt + 1))
        for d in rng.choice(ndom, size=k, replace=False):
            touch[int(d)].append(g)
    l2g_d = [rng.permutation(np.array(t, dtype=np.int32)) for t in touch]      # local numbering in any order
    n_gd = np.array([t.size for t in l2g_d], dtype=np.int64)
    l2g = np.concatenate(l2g_d).astype(np.int32) if ndom else np.zeros(0, np.int32)
    loc = rng.standard_normal(l2g.size) * 10.0 ** rng.integers(-8, 8, l2g.size)
    base = rng.standard_normal(ng)
    for with_base in (False, True):
        seq = base.copy() if with_base else np.zeros(ng)
        off = 0
        for d in range(ndom):
            seq[l2g_d[d]] += loc[off:off + l2g_d[d].size]
            off += l2g_d[d].size
        for world in range(1, min(8, ndom) + 1):
            y = np.zeros(world * ng)
            sent = np.zeros(world, dtype=np.int64)
            threads = int(rng.choice([32, 128, 256, 1024]))
            m = lib.planemu_exchange(world, ndom, ng, n_gd, l2g, loc, base.ctypes.data if with_base else None, threads, y, sent)
            ok = m >= 0 and sent.sum() == m and m <= max(l2g.size, ng) and all(
                np.array_equal(y[r * ng:(r + 1) * ng], seq) for r in range(world))
            nrun += 1
            if not ok:
                nfail += 1
                print("FAIL", seed, ndom, ng, maxt, with_base, world, threads, m, flush=True)
    ncase += 1
    seed += 1
print("done: topologies", ncase, "exchanges", nrun, "failures", nfail, "last seed", seed, flush=True)
