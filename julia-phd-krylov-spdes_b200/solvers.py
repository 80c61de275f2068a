"""RecyclingKrylovSolvers signatures dispatching on the B200 operator / preconditioner types:
``pcg(S, b, x, M; maxit)`` (cg.jl:67-72) runs the whole solve on the device and returns the
reference's tuple ``(x, it, res_norm)``."""
from .fem import B200NeumannNeumann, B200SchurOperator

EPS = 1e-7  # RecyclingKrylovSolvers.jl:21


def pcg(A: B200SchurOperator, b, x, M: B200NeumannNeumann = None, maxit=0, return_timings=False):
    if M is not None and M.S is not A:
        raise ValueError("preconditioner and operator live on different handles")
    x, it, res, tm = A.h.pcg(b, x, maxit=maxit, rtol=EPS, assembled=A.assembled, precond=M is not None)
    return (x, it, res, tm) if return_timings else (x, it, res)


def cg(A: B200SchurOperator, b, x, maxit=0):
    """cg.jl:14-50 (no preconditioner)."""
    return pcg(A, b, x, None, maxit)
