"""World-size-2 CPU test (gloo) of the multi-rank host logic: contiguous subdomain shards, partial
interface sums combined by all-reduce, rank-0-only addition of b_Γ, identical scalars on every
rank.  The same sharding rule is used by kb200_set_topology (api.cu) and bench.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import krylov_b200  # noqa: F401
    from krylov_b200.hostprep import problem as PB
    import oracle.epdd as E
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pr = PB.build_problem(33, 4)
    sd, lb = pr.sd, pr.blocks
    inner = E.ExactInterior(lb.A_II)
    d0, d1 = (sd.ndom * rank) // world, (sd.ndom * (rank + 1)) // world      # kb200_set_topology's rule
    x = np.random.default_rng(7).standard_normal(sd.n_gamma)
    part = np.zeros(sd.n_gamma)
    for d in range(d0, d1):
        g = sd.gamma_l2g[d]
        part[g] += E.apply_local_schur(lb, inner, d, x[g])
    t = torch.from_numpy(part)
    dist.all_reduce(t)
    full = E.apply_local_schurs(lb, inner, sd, x)
    # Schur rhs: b_Γ enters once (rank 0), subdomain terms are sharded
    rhs = lb.b_G.copy() if rank == 0 else np.zeros(sd.n_gamma)
    for d in range(d0, d1):
        rhs[sd.gamma_l2g[d]] -= lb.A_IG[d].T @ inner.solve(d, lb.b_I[d])
    tr = torch.from_numpy(rhs)
    dist.all_reduce(tr)
    ref_rhs = E.get_schur_rhs(lb, inner, sd)
    # max-over-ranks timing reduction used by bench.py
    tm = torch.tensor([float(rank + 1)])
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    q.put((rank, float(abs(t.numpy() - full).max() / abs(full).max()),
           float(abs(tr.numpy() - ref_rhs).max() / abs(ref_rhs).max()), float(tm.item()), (d0, d1)))
    dist.destroy_process_group()


def test_two_rank_sharded_apply_matches_single_rank():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ranges = sorted(r[4] for r in res)
    assert ranges[0][0] == 0 and ranges[-1][1] == 4 and ranges[0][1] == ranges[1][0]
    for rank, e_apply, e_rhs, tmax, _ in res:
        assert e_apply <= 1e-13 and e_rhs <= 1e-13 and tmax == float(world)


def test_exchange_slot_counter_alternates_over_real_exchanges_only():
    """The two receive slots of the peer-memory exchange must alternate between consecutive REAL exchanges although
    launches skipped on the device after convergence sit between them.  The count of real exchanges lives on the
    device in a two-word ping-pong counter (kernels.cuh: launch L reads cnt[L & 1], writes cnt[(L + 1) & 1]); this
    replays that rule on random launch sequences."""
    import random
    for seed in range(200):
        rng = random.Random(seed)
        cnt, last_slot, reals = [0, 0], None, 0
        for L in range(rng.randint(1, 400)):
            par = L & 1
            if rng.random() < 0.6:                      # a real exchange
                seq = cnt[par] + 1
                cnt[par ^ 1] = seq
                reals += 1
                assert seq == reals                      # monotone sequence number, identical on every rank
                assert last_slot is None or (seq & 1) != last_slot
                last_slot = seq & 1
            else:                                        # skipped on the device: the counter is handed on
                cnt[par ^ 1] = cnt[par]


def test_compact_exchange_reproduces_the_sequential_sum_bit_for_bit():
    """The owner-ordered compact exchange (csrc/xchg_tables.h, exchange_gather in kernels.cuh) emulated on
    the CPU through the product's own table builder: for 1..8 ranks every rank assembles exactly the bits of
    the reference's sequential sum over ascending subdomains (EllipticPdeDomainDecomposition.jl:740-742),
    with and without a base vector, and all ranks together send little more than n_gamma entries."""
    import ctypes as C
    import krylov_b200  # noqa: F401
    from krylov_b200.hostprep import problem as PB
    from planemu.build import build as build_emu
    lib = C.CDLL(build_emu())
    f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
    lib.planemu_exchange.restype = C.c_int64
    lib.planemu_exchange.argtypes = [C.c_int, C.c_int64, C.c_int64, np.ctypeslib.ndpointer(np.int64, flags="C"),
                                     np.ctypeslib.ndpointer(np.int32, flags="C"), f64p, C.c_void_p, C.c_int, f64p,
                                     np.ctypeslib.ndpointer(np.int64, flags="C")]
    for nx, ndom, jitter in [(33, 4, 0.2), (65, 16, 0.2), (49, 9, 0.0), (65, 13, 0.3)]:
        sd = PB.build_mesh_and_partition(nx, ndom, jitter)[-1]
        ng = sd.n_gamma
        n_gd = np.array([g.size for g in sd.gamma_l2g], dtype=np.int64)
        l2g = np.concatenate(sd.gamma_l2g).astype(np.int32)
        rng = np.random.default_rng(nx)
        # values spread over many magnitudes: any change of the association order changes bits
        loc = rng.standard_normal(l2g.size) * 10.0 ** rng.integers(-8, 8, l2g.size)
        base = rng.standard_normal(ng)
        for with_base in (False, True):
            seq = base.copy() if with_base else np.zeros(ng)
            off = 0
            for d in range(ndom):                       # the reference's loop: ascending subdomains
                g = sd.gamma_l2g[d]
                seq[g] += loc[off:off + g.size]
                off += g.size
            for world in (1, 2, 3, 4, 8):
                if world > ndom:
                    continue
                y = np.zeros(world * ng)
                sent = np.zeros(world, dtype=np.int64)
                m = lib.planemu_exchange(world, ndom, ng, n_gd, l2g, loc, base.ctypes.data if with_base else None, 256, y, sent)
                assert m >= ng and sent.sum() == m
                assert m <= l2g.size                     # never more than one entry per contribution
                if (nx, ndom, world) == (65, 16, 2):     # few rank-crossing nodes: little more than n_gamma entries
                    assert m <= 1.25 * ng
                for r in range(world):
                    assert np.array_equal(y[r * ng:(r + 1) * ng], seq), (nx, ndom, world, r)


def test_interface_rhs_of_a_sharded_run_has_the_bits_of_the_full_assembly():
    """bench.py on several ranks: every rank assembles only its own subdomains' elements; b_Γ of the whole mesh comes
    from the interface-touching elements alone and must equal the 1-GPU assembly bit for bit (so that the iteration
    count of the benchmark solve does not depend on the number of GPUs), as must the local b_I."""
    from krylov_b200.hostprep import problem as PB
    base = PB.build_mesh_and_partition(65, 16)
    full = PB.realization(base, ireal=1, factorize=False)
    bg = PB.interface_rhs(base, full.coeff, full.f_vals, full.u_dir)
    assert np.array_equal(bg, full.blocks.b_G)
    part = PB.realization(base, ireal=1, factorize=False, dom_range=(4, 8))
    for d in range(4, 8):
        assert np.array_equal(part.blocks.b_I[d], full.blocks.b_I[d])
        assert (part.blocks.A_IG[d] != full.blocks.A_IG[d]).nnz == 0
