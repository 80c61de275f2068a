"""Development probe: kernel timings at a given size for several cluster sizes."""
import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import krylov_b200
from krylov_b200.hostprep import problem as PB
from krylov_b200 import fem, solvers

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2001
ndom = int(sys.argv[2]) if len(sys.argv) > 2 else 256
csizes = [int(v) for v in sys.argv[3].split(',')] if len(sys.argv) > 3 else [4]
do_pcg = len(sys.argv) > 4 and sys.argv[4] == 'pcg'
stages = [int(v) for v in os.environ.get('STAGES', '0').split(',')]
extra = [dict(kv.split('=') for kv in cfg.split('+')) for cfg in os.environ.get('OPTS', '').split(',') if cfg] or [{}]
t = time.time(); pr = PB.build_problem(nx, ndom, ireal=1); print('host prep %.1fs' % (time.time() - t), flush=True)
B, parts = PB.algorithmic_bytes_per_iteration(pr)
print('B_iter GB', B / 1e9, {k: v / 1e9 for k, v in parts.items()}, flush=True)
for cs, nst, ex in [(a, b, e) for a in csizes for b in stages for e in extra]:
    t = time.time()
    S = fem.B200SchurOperator(pr.sd, pr.blocks, pr.factors, csize=cs, profile=True, options=dict({'stages': nst}, **{k: int(v) for k, v in ex.items()}))
    S.h.finalize(); tf = time.time() - t
    st = S.h.stats()
    ms = S.h.time_kernel(0, 20)
    alg = parts['trisolve'] + parts['spmv']
    print(f"csize={cs} stages={nst} {ex} finalize {tf:.1f}s schur kernel {ms:.3f} ms  stream {st['stream_bytes']/1e9:.3f} GB -> {st['stream_bytes']/ms/1e6:.0f} GB/s actual, algorithmic {alg/ms/1e6:.0f} GB/s | md {st['merge_depth']:.0f} mb {st['min_block']:.0f} levels {st['max_levels']:.0f} top {st['max_top']:.0f} smem {st['smem_bytes']:.0f} ctas {st['n_cta']:.0f}", flush=True)
    for fl in [int(v) for v in os.environ.get('FLAGS', '').split(',') if v]:
        S.h.set_option('dbg_flags', fl); print(f'   dbg_flags={fl}: schur kernel {S.h.time_kernel(0, 10):.3f} ms', flush=True); S.h.set_option('dbg_flags', 0)
    if os.environ.get('DBG'):
        S.h.set_option('debug_timing', 1)
        S.h.set_option('dbg_flags', int(os.environ.get('DBGFLAGS', '0')))
        S.h.time_kernel(0, 1)
        st_ = S.h.debug_stamps(int(st['n_cta'])).astype(float)
        cs_ = cs
        r0 = st_[0::cs_]; r1 = st_[1::cs_] if cs_ > 1 else st_
        def seg(a, i, j): return (a[:, j] - a[:, i]) / 1.9e3   # us at ~1.9 GHz
        names = [('init', 0, 9), ('INIT sec', 9, 9), ('FWD', 9, 10), ('PART', 10, 1), ('wait sync1', 1, 2), ('TOPF(rank0)', 2, 3), ('TOPB', 3, 4), ('wait sync2', 4, 5), ('BWD', 5, 6), ('GG+OUT', 6, 7), ('epilogue', 7, 8), ('total', 0, 8)]
        for nm, i, j in names:
            a = seg(r0, i, j); b = seg(r1, i, j)
            print(f"   {nm:12s} rank0 mean {a.mean():8.1f} max {a.max():8.1f} us | rank1 mean {b.mean():8.1f} max {b.max():8.1f} us")
        pf = st_[:, 16:48].reshape(-1, 8, 4)
        if pf[:, :, 1].sum() > 0:
            for nm, sel in (('rank0', pf[0::cs_]), ('rank1', pf[1::cs_] if cs_ > 1 else pf)):
                tot = (st_[0::cs_][:, 7] - st_[0::cs_][:, 0]).mean() / 1.9e3
                print(f"   prof {nm}: per warp mean: slices {sel[:, :, 1].mean():.0f} (max {sel[:, :, 1].max():.0f}), "
                      f"in slices {sel[:, :, 0].mean() / 1.9e3:.1f} us (max warp {sel[:, :, 0].max(axis=1).mean() / 1.9e3:.1f}), "
                      f"cycles/slice {sel[:, :, 0].sum() / max(1, sel[:, :, 1].sum()):.0f}, "
                      f"in level barriers {sel[:, :, 2].mean() / 1.9e3:.1f} us, waiting for chunks {sel[:, :, 3].mean() / 1.9e3:.1f} us, kernel {tot:.1f} us")
        S.h.set_option('debug_timing', 0)
        S.h.set_option('dbg_flags', 0)
    if do_pcg:
        t = time.time(); Sd = S.assemble_local_schurs(); print('  assemble S_d %.1fs' % (time.time() - t), flush=True)
        t = time.time(); Pi = [PB.pinv_sym(Sd[d]) for d in range(ndom)]; print('  pinv %.1fs' % (time.time() - t), flush=True)
        M = fem.B200NeumannNeumann(S, Pi)
        print('  nn kernel ms', S.h.time_kernel(1, 20), 'dense schur kernel ms', S.h.time_kernel(2, 20), flush=True)
        bs = S.schur_rhs(pr.blocks.b_I, pr.blocks.b_G)
        for rep in range(2):
            t = time.time(); x, it, res, tm = solvers.pcg(S, bs, np.zeros(pr.sd.n_gamma), M, maxit=5000, return_timings=True); tw = time.time() - t
            print(f"  pcg it {it} wall {tw*1e3:.1f} ms solve_ms {tm.solve_ms:.1f} schur {tm.schur_ms:.1f}/{tm.schur_launches} nn {tm.nn_ms:.1f}/{tm.nn_launches} per-it {tm.solve_ms/(it-1):.3f} ms  roofline(alg) {B*(it-1)/tm.solve_ms/1e6:.0f} GB/s  res {res[-1]/res[0]:.2e}", flush=True)
    del S
