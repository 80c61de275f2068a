import numpy as np, time, sys
sys.path.insert(0, '/root/repo')
import krylov_b200
from krylov_b200.hostprep import problem as PB
from krylov_b200 import fem, solvers
import oracle.epdd as E, oracle.solvers as OS

def run(nx, ndom, csize):
    pr = PB.build_problem(nx, ndom, ireal=1)
    sd, lb = pr.sd, pr.blocks
    inner = E.ExactInterior(lb.A_II)
    S = fem.B200SchurOperator(sd, lb, pr.factors, csize=csize, profile=True)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(sd.n_gamma)
    y = S * x
    yr = E.apply_local_schurs(lb, inner, sd, x)
    print(f"nx={nx} ndom={ndom} cs={csize} apply_schur relerr {abs(y-yr).max()/abs(yr).max():.2e}", S.h.stats())
    bs = S.schur_rhs(lb.b_I, lb.b_G); bsr = E.get_schur_rhs(lb, inner, sd)
    print("  schur_rhs relerr", abs(bs-bsr).max()/abs(bsr).max())
    Sd = S.assemble_local_schurs()
    Sdr = E.assemble_local_schurs(lb, inner)
    print("  dense S_d relerr", max(abs(Sd[d]-Sdr[d]).max()/abs(Sdr[d]).max() for d in range(ndom)))
    Pi = E.prepare_neumann_neumann_schur_precond(Sdr)
    M = fem.B200NeumannNeumann(S, Pi)
    z = M.solve(x); zr = E.apply_neumann_neumann_schur(Pi, sd, x)
    print("  apply_nn relerr", abs(z-zr).max()/abs(zr).max())
    S.assembled = True
    ya = S * x
    print("  assembled apply relerr", abs(ya-yr).max()/abs(yr).max())
    S.assembled = False
    xg, it, res, tm = solvers.pcg(S, bsr, np.zeros(sd.n_gamma), M, return_timings=True)
    xo, ito, reso = OS.pcg(lambda v: E.apply_local_schurs(lb, inner, sd, v), bsr, np.zeros(sd.n_gamma), lambda r: E.apply_neumann_neumann_schur(Pi, sd, r))
    print(f"  pcg it gpu {it} oracle {ito} relerr x {np.linalg.norm(xg-xo)/np.linalg.norm(xo):.2e} res hist rel diff {abs(res[:min(it,ito)]-reso[:min(it,ito)]).max()/reso[0]:.2e}")
    print(f"  timings solve_ms {tm.solve_ms:.3f} schur_ms {tm.schur_ms:.3f} ({tm.schur_launches}) nn_ms {tm.nn_ms:.3f} launches {tm.kernel_launches}")
    uI = S.subdomain_solutions(xg, lb.b_I); uIr = E.get_subdomain_solutions(xg, lb, inner, sd)
    print("  backsub relerr", max(abs(a-b).max() for a,b in zip(uI,uIr))/max(abs(b).max() for b in uIr))
    xc, itc, resc = solvers.cg(S, bsr, np.zeros(sd.n_gamma))
    xco, itco, _ = OS.pcg(lambda v: E.apply_local_schurs(lb, inner, sd, v), bsr, np.zeros(sd.n_gamma), lambda r: r.copy())
    print(f"  cg it gpu {itc} oracle {itco}")

for args in [(33,4,1),(33,4,2),(65,16,4),(129,16,4),(129,16,8),(257,16,1)]:
    run(*args)
