"""CPU restatement of RecyclingKrylovSolvers (TEST INFRASTRUCTURE): pcg, defpcg, eigpcg,
eigdefpcg, following cg.jl:67-109, defcg.jl:242-308, eigcg.jl:143-267 and defcg.jl:337-473.

``A`` and ``M`` are callables (x -> A x, r -> M \\ r), standing for the reference's
LinearMaps.FunctionMap and preconditioner objects.  ``EPS`` is the reference's hard-wired
relative residual tolerance (RecyclingKrylovSolvers.jl:21).  As in the reference, ``it`` starts
at 1 (it = iterations + 1), ``res_norm`` has ``it`` entries and ``x`` is updated in place.

PARITY UNPINNED: the reference ships no tests, golden vectors or data for this path and cannot run here (no Julia);
this restatement is pinned by the reference's in-script identities and self-generated goldens only.
"""
import numpy as np

EPS = 1e-7  # RecyclingKrylovSolvers.jl:21


def pcg(A, b, x, M, maxit=0, eps=EPS):
    """cg.jl:67-109."""
    n = x.size
    maxit = n if maxit == 0 else maxit                       # :81
    res_norm = np.empty(max(n, maxit) + 1)
    it = 1
    r = b - A(x)                                             # :83
    rTr = r @ r
    z = M(r)                                                 # :85
    rTz = r @ z
    p = z.copy()
    res_norm[it - 1] = np.sqrt(rTr)
    tol = eps * np.linalg.norm(b)                            # :89-90
    while it < maxit and res_norm[it - 1] > tol:             # :92
        Ap = A(p)
        d = p @ Ap
        alpha = rTz / d
        beta = 1.0 / rTz
        x += alpha * p
        r -= alpha * Ap
        rTr = r @ r
        z = M(r)
        rTz = r @ z
        beta *= rTz
        p = beta * p + z
        it += 1
        res_norm[it - 1] = np.sqrt(rTr)
    return x, it, res_norm[:it].copy()


def _julia_rank(Y):
    """LinearAlgebra.rank default: #singular values > min(size)*eps*sigma_max."""
    s = np.linalg.svd(Y, compute_uv=False)
    if s.size == 0 or s[0] == 0.0:
        return 0
    tol = min(Y.shape) * np.finfo(np.float64).eps * s[0]
    return int((s > tol).sum())


def _rayleigh_ritz(Tm, Tm_minor, nvec):
    """Shared restart algebra of eigcg.jl:220-228 / defcg.jl:425-432: returns (Q Z, vals, nev)."""
    m = Tm.shape[0]
    Y = np.zeros((m, 2 * nvec))
    Y[:, :nvec] = np.linalg.eigh(Tm)[1][:, :nvec]
    Y[:m - 1, nvec:] = np.linalg.eigh(Tm_minor)[1][:, :nvec]
    nev = _julia_rank(Y)
    Q = np.linalg.svd(Y, full_matrices=False)[0][:, :nev]
    H = Q.T @ (Tm @ Q)
    H = np.triu(H) + np.triu(H, 1).T                         # Symmetric(H)
    vals, Z = np.linalg.eigh(H)
    return Q @ Z, vals, nev


def _sym(U):
    """Symmetric(U): mirror the upper triangle."""
    return np.triu(U) + np.triu(U, 1).T


def defpcg(A, b, x, W, M, maxit=0, eps=EPS):
    """defcg.jl:242-308."""
    n, nvec = W.shape
    WtA = np.empty((nvec, n))
    for iv in range(nvec):                                   # :261-264
        WtA[iv, :] = A(W[:, iv])
    WtAW = WtA @ W                                           # :268
    r = b - A(x)
    mu = np.linalg.solve(WtAW, W.T @ r)                      # :272-274
    x += W @ mu
    maxit = n if maxit == 0 else maxit
    res_norm = np.empty(max(n, maxit) + 1)
    it = 1
    r = b - A(x)
    rTr = r @ r
    res_norm[0] = np.sqrt(rTr)
    z = M(r)
    rTz = r @ z
    mu = np.linalg.solve(WtAW, WtA @ z)                      # :284
    p = z - W @ mu
    tol = eps * np.linalg.norm(b)
    while it < maxit and res_norm[it - 1] > tol:             # :290
        Ap = A(p)
        d = p @ Ap
        alpha = rTz / d
        beta = 1.0 / rTz
        x += alpha * p
        r -= alpha * Ap
        rTr = r @ r
        z = M(r)
        rTz = r @ z
        beta *= rTz
        mu = np.linalg.solve(WtAW, WtA @ z)                  # :301
        p = beta * p + z - W @ mu                            # :302
        it += 1
        res_norm[it - 1] = np.sqrt(rTr)
    return x, it, res_norm[:it].copy()


def eigpcg(A, b, x, M, nvec, spdim, maxit=0, eps=EPS):
    """eigcg.jl:143-267.  Returns (x, it, res_norm, V[:, :nvec])."""
    n = x.size
    V = np.zeros((n, spdim))
    AV = np.empty((n, spdim))
    VtAV = np.zeros((spdim, spdim))
    tvec = np.zeros(n)
    maxit = n if maxit == 0 else maxit
    res_norm = np.empty(max(n, maxit) + 1)
    it = 1
    r = b - A(x)
    rTr = r @ r
    z = M(r)
    rTz = r @ z
    p = z.copy()
    res_norm[0] = np.sqrt(rTr)
    ivec = 0                                                 # 0-based column cursor (:178)
    V[:, ivec] = z / np.sqrt(rTz)
    just_restarted = False
    hlpr = 0.0
    nev = 0
    tol = eps * np.linalg.norm(b)
    while it < maxit and res_norm[it - 1] > tol:             # :185
        Ap = A(p)
        d = p @ Ap
        alpha = rTz / d
        beta = 1.0 / rTz
        x += alpha * p
        r -= alpha * Ap
        rTr = r @ r
        z = M(r)
        if just_restarted:
            hlpr = np.sqrt(rTz)                              # :194-196
        rTz = r @ z
        beta *= rTz
        if ivec == spdim - 1:
            tvec -= beta * Ap                                # :199-201
        p = beta * p + z
        it += 1
        res_norm[it - 1] = np.sqrt(rTr)
        VtAV[ivec, ivec] += 1.0 / alpha                      # :206
        if just_restarted:
            tvec += Ap                                       # :208
            nev = ivec
            VtAV[:nev, ivec] = V[:, :nev].T @ (tvec / hlpr)  # :210
            just_restarted = False
        if ivec == spdim - 1:                                # :214
            for jv in range(spdim):
                AV[:, jv] = A(V[:, jv])
            VtAV[:, :] = V.T @ AV                            # :222
            Tm = _sym(VtAV)
            QZ, vals, nev = _rayleigh_ritz(Tm, Tm[:spdim - 1, :spdim - 1], nvec)
            V[:, :nev] = V @ QZ                              # :231
            ivec = nev
            V[:, ivec] = z / np.sqrt(rTz)
            VtAV[:, :] = 0.0
            VtAV[np.arange(nev), np.arange(nev)] = vals
            VtAV[ivec, ivec] = beta / alpha
            tvec = -beta * Ap                                # :240
            just_restarted = True
        else:
            ivec += 1
            V[:, ivec] = z / np.sqrt(rTz)
            VtAV[ivec - 1, ivec] = -np.sqrt(beta) / alpha
            VtAV[ivec, ivec] = beta / alpha
    if not just_restarted:                                   # :251
        if ivec + 1 > nvec:
            m = ivec                                         # reference: ivec -= 1 (1-based) => m columns
            Tm = _sym(VtAV[:m, :m])
            QZ, vals, nev = _rayleigh_ritz(Tm, Tm[:m - 1, :m - 1], nvec)
            V[:, :nev] = V[:, :m] @ QZ
    return x, it, res_norm[:it].copy(), V[:, :nvec].copy()


def eigdefpcg(A, b, x, M, W, spdim, maxit=0, eps=EPS):
    """defcg.jl:337-473.  Returns (x, it, res_norm, V[:, :nvec])."""
    n, nvec = W.shape
    WtA = np.empty((nvec, n))
    for iv in range(nvec):                                   # :357-360
        WtA[iv, :] = A(W[:, iv])
    WtAW = WtA @ W
    WtW = W.T @ W
    r = b - A(x)
    mu = np.linalg.solve(WtAW, W.T @ r)
    x += W @ mu
    V = np.zeros((n, spdim))
    VtAV = np.zeros((spdim, spdim))
    first_restart = True
    maxit = n if maxit == 0 else maxit
    res_norm = np.empty(max(n, maxit) + 1)
    it = 1
    r = b - A(x)
    rTr = r @ r
    z = M(r)
    rTz = r @ z
    mu = np.linalg.solve(WtAW, WtA @ z)
    p = z - W @ mu
    res_norm[0] = np.sqrt(rTr)
    VtAV[:nvec, :nvec] = WtAW                                # :390
    V[:, :nvec] = W
    just_restarted = False
    ivec = nvec                                              # 0-based cursor (:394)
    V[:, ivec] = z / np.sqrt(rTz)
    tol = eps * np.linalg.norm(b)
    while it < maxit and res_norm[it - 1] > tol:             # :400
        Ap = A(p)
        d = p @ Ap
        alpha = rTz / d
        beta = 1.0 / rTz
        x += alpha * p
        r -= alpha * Ap
        r -= W @ np.linalg.solve(WtW, W.T @ r)               # :407
        rTr = r @ r
        z = M(r)
        rTz = r @ z
        beta *= rTz
        mu = np.linalg.solve(WtAW, WtA @ z)
        p = beta * p + z - W @ mu
        it += 1
        res_norm[it - 1] = np.sqrt(rTr)
        VtAV[ivec, ivec] += 1.0 / alpha                      # :417
        if ivec == spdim - 1:                                # :419
            if first_restart:
                VtAV[:nvec, nvec:spdim] = WtA @ V[:, nvec:spdim]
                first_restart = False
            Tm = _sym(VtAV)
            QZ, vals, nev = _rayleigh_ritz(Tm, Tm[:spdim - 1, :spdim - 1], nvec)
            V[:, :nev] = V @ QZ
            ivec = nev
            V[:, ivec] = z / np.sqrt(rTz)
            VtAV[:, :] = 0.0
            VtAV[np.arange(nev), np.arange(nev)] = vals
            VtAV[ivec, ivec] = beta / alpha
            just_restarted = True
        else:
            just_restarted = False
            ivec += 1
            V[:, ivec] = z / np.sqrt(rTz)
            VtAV[ivec - 1, ivec] = -np.sqrt(beta) / alpha
            VtAV[ivec, ivec] = beta / alpha
    if not just_restarted:                                   # :450
        if ivec + 1 > nvec:
            m = ivec
            if first_restart:
                VtAV[:nvec, nvec:m] = WtA @ V[:, nvec:m]
            Tm = _sym(VtAV[:m, :m])
            QZ, vals, nev = _rayleigh_ritz(Tm, Tm[:m - 1, :m - 1], nvec)
            V[:, :nev] = V[:, :m] @ QZ
    return x, it, res_norm[:it].copy(), V[:, :nvec].copy()


# ---------------------------------------------------------------------------------------------
# Post-hoc Rayleigh-Ritz / harmonic-Ritz recycling: rrpcg / rrdefpcg (rrdefpcg.jl:204-275, :48-151)
# and hrpcg / hrdefpcg (hrdefpcg.jl:217-295, :48-164).
#
# The reference fills the search-space columns only while `it <= cap` (cap = spdim, or spdim - nvec with a deflation
# basis) and then projects with ALL cap columns: when the solve stops earlier the remaining columns are uninitialised
# memory (`Array{T,2}(undef, ...)`).  That is undefined behaviour, not an algorithm; this restatement (and the library)
# projects onto the m = min(it, cap) columns that exist.  eigen(A, B) of two Symmetric matrices is LAPACK's sygvd:
# ascending eigenvalues, B-orthonormal eigenvectors (scipy.linalg.eigh(A, B) is the same routine).
# ---------------------------------------------------------------------------------------------
def _posthoc(A, b, x, M, W, nvec, spdim, maxit, eps, harmonic):
    import scipy.linalg as sla
    n = x.size
    defl = W is not None
    if defl:
        nvec = W.shape[1]
        WtA = np.empty((nvec, n))
        for iv in range(nvec):                               # rrdefpcg.jl:73-76
            WtA[iv, :] = A(W[:, iv])
        WtAW = _sym(WtA @ W)                                 # :80-81 Symmetric()
        WtW = _sym(W.T @ W)                                  # :82-83
        r = b - A(x)                                         # :86-89
        x += W @ np.linalg.solve(WtAW, W.T @ r)
    cap = spdim - nvec if defl else spdim
    C = np.zeros((n, cap))                                   # Z (RR) or P (HR)
    maxit = n if maxit == 0 else maxit
    res_norm = np.empty(max(n, maxit) + 1)
    it = 1
    r = b - A(x)
    rTr = r @ r
    res_norm[0] = np.sqrt(rTr)
    z = M(r)
    rTz = r @ z
    p = z - W @ np.linalg.solve(WtAW, WtA @ z) if defl else z.copy()       # :97-98
    C[:, 0] = p if harmonic else z / rTz ** 0.5                              # :99 / hrdefpcg.jl:101
    tol = eps * np.linalg.norm(b)
    while it < maxit and res_norm[it - 1] > tol:             # :104
        Ap = A(p)
        d = p @ Ap
        alpha = rTz / d
        beta = 1.0 / rTz
        x += alpha * p
        r -= alpha * Ap
        if defl:
            r -= W @ np.linalg.solve(WtW, W.T @ r)           # :111
        rTr = r @ r
        z = M(r)
        rTz = r @ z
        beta *= rTz
        p = beta * p + z - W @ np.linalg.solve(WtAW, WtA @ z) if defl else beta * p + z   # :116-117
        it += 1
        res_norm[it - 1] = np.sqrt(rTr)
        if it <= cap:                                        # :120-122
            C[:, it - 1] = p if harmonic else z / rTz ** 0.5
    m = min(it, cap)
    C = C[:, :m]
    AC = np.column_stack([A(C[:, j]) for j in range(m)])     # :127-133
    k0 = nvec if defl else 0
    T = np.zeros((k0 + m, k0 + m))
    V = np.column_stack([W, C]) if defl else C
    if not harmonic:                                         # rrdefpcg.jl:124-146
        if defl:
            T[:k0, :k0] = WtAW
            T[:k0, k0:] = WtA @ C
        T[k0:, k0:] = C.T @ AC
        vals, vecs = np.linalg.eigh(_sym(T))                 # eigen(VtAV, I)
    else:                                                    # hrdefpcg.jl:124-157
        MAC = np.column_stack([M(AC[:, j]) for j in range(m)])
        G = np.zeros_like(T)
        if defl:
            MAW = np.column_stack([M(WtA[iv, :]) for iv in range(nvec)])
            G[:k0, :k0] = WtA @ MAW
            G[:k0, k0:] = WtA @ MAC
            T[:k0, :k0] = WtAW
        G[k0:, k0:] = AC.T @ MAC
        T[k0 + np.arange(m), k0 + np.arange(m)] = np.einsum("ij,ij->j", C, AC)
        vals, vecs = sla.eigh(_sym(G), _sym(T))              # eigen(VtAinvMAV, VtAV)
    nret = nvec
    W2 = V @ vecs[:, :nret]                                  # :148-150
    return x, it, res_norm[:it].copy(), W2


def rrpcg(A, b, x, M, nvec, spdim, maxit=0, eps=EPS):
    """rrdefpcg.jl:204-275."""
    return _posthoc(A, b, x, M, None, nvec, spdim, maxit, eps, False)


def rrdefpcg(A, b, x, W, M, spdim, maxit=0, eps=EPS):
    """rrdefpcg.jl:48-151 (positional order of the reference: A, b, x, W, M, spdim)."""
    return _posthoc(A, b, x, M, W, W.shape[1], spdim, maxit, eps, False)


def hrpcg(A, b, x, M, nvec, spdim, maxit=0, eps=EPS):
    """hrdefpcg.jl:217-295."""
    return _posthoc(A, b, x, M, None, nvec, spdim, maxit, eps, True)


def hrdefpcg(A, b, x, W, M, spdim, maxit=0, eps=EPS):
    """hrdefpcg.jl:48-164."""
    return _posthoc(A, b, x, M, W, W.shape[1], spdim, maxit, eps, True)


# ---------------------------------------------------------------------------------------------
# Thick-restart Rayleigh-Ritz / harmonic-Ritz recycling: trrrpcg / trrrdefpcg (trrrdefpcg.jl:234-355,
# :48-199) and trhrpcg / trhrdefpcg (trhrdefpcg.jl:247-389, :48-208).
#
# The recurrence is that of the post-hoc variants above (deflated by the ORIGINAL basis W throughout); the search-space
# columns are written cyclically and every time the window is full the recycle basis W2 is replaced by the nvec lowest
# (harmonic) Ritz vectors of [W2 | window].  Without a deflation basis the first window has spdim columns and projects
# onto them alone; later windows have spdim - nvec columns.  A solve that stops before its first window is full
# returns W2 = W (deflated variants) -- the reference's trrrpcg / trhrpcg return uninitialised memory in that case,
# this restatement (and the library) return zeros.  The harmonic variants keep the reference's structural zeros: the
# (W2, P) blocks of V'AV are not computed (trhrdefpcg.jl:170-177), although W2 stops being A-conjugate to the window
# after the first restart.
# ---------------------------------------------------------------------------------------------
def _thick_restart(A, b, x, M, W, nvec, spdim, maxit, eps, harmonic):
    import scipy.linalg as sla
    n = x.size
    defl = W is not None
    if defl:
        nvec = W.shape[1]
        WtA = np.empty((nvec, n))
        for iv in range(nvec):
            WtA[iv, :] = A(W[:, iv])
        WtAW = _sym(WtA @ W)
        WtW = _sym(W.T @ W)
        W2, W2tA, W2tAW2 = W.copy(), WtA.copy(), WtAW.copy()
        r = b - A(x)
        x += W @ np.linalg.solve(WtAW, W.T @ r)
    else:
        W2 = np.zeros((n, nvec))
    wcap = spdim - nvec if defl else spdim                  # columns of the current window
    C = np.zeros((n, spdim))
    restarted = False
    maxit = n if maxit == 0 else maxit
    res_norm = np.empty(max(n, maxit) + 1)
    it = 1
    r = b - A(x)
    rTr = r @ r
    res_norm[0] = np.sqrt(rTr)
    z = M(r)
    rTz = r @ z
    p = z - W @ np.linalg.solve(WtAW, WtA @ z) if defl else z.copy()
    C[:, 0] = p if harmonic else z / rTz ** 0.5
    col = 1                                                  # columns of the window filled so far
    tol = eps * np.linalg.norm(b)
    while it < maxit and res_norm[it - 1] > tol:
        Ap = A(p)
        d = p @ Ap
        alpha = rTz / d
        beta = 1.0 / rTz
        x += alpha * p
        r -= alpha * Ap
        if defl:
            r -= W @ np.linalg.solve(WtW, W.T @ r)
        rTr = r @ r
        z = M(r)
        rTz = r @ z
        beta *= rTz
        p = beta * p + z - W @ np.linalg.solve(WtAW, WtA @ z) if defl else beta * p + z
        it += 1
        res_norm[it - 1] = np.sqrt(rTr)
        C[:, col] = p if harmonic else z / rTz ** 0.5        # ivec of the reference = col + 1
        col += 1
        if col < wcap:
            continue
        # ---- the window is full: project ------------------------------------------------------------------------
        Cw = C[:, :wcap]
        ACw = np.column_stack([A(Cw[:, j]) for j in range(wcap)])
        with_w2 = defl or restarted
        if with_w2 and restarted:                            # W2 changed at the last restart: A W2 again
            for iv in range(nvec):
                W2tA[iv, :] = A(W2[:, iv])
            W2tAW2 = _sym(W2tA @ W2)
        k0 = nvec if with_w2 else 0
        mt = k0 + wcap
        V = np.column_stack([W2, Cw]) if with_w2 else Cw
        T = np.zeros((mt, mt))
        if not harmonic:
            if with_w2:
                T[:k0, :k0] = W2tAW2
                T[:k0, k0:] = W2tA @ Cw
            T[k0:, k0:] = Cw.T @ ACw
            vals, vecs = np.linalg.eigh(_sym(T))             # eigen(VtAV, I)
        else:
            MACw = np.column_stack([M(ACw[:, j]) for j in range(wcap)])
            G = np.zeros((mt, mt))
            if with_w2:
                MAW2 = np.column_stack([M(W2tA[iv, :]) for iv in range(nvec)])
                G[:k0, :k0] = W2tA @ MAW2
                G[:k0, k0:] = W2tA @ MACw
                T[:k0, :k0] = W2tAW2
            G[k0:, k0:] = ACw.T @ MACw
            T[k0 + np.arange(wcap), k0 + np.arange(wcap)] = np.einsum("ij,ij->j", Cw, ACw)
            vals, vecs = sla.eigh(_sym(G), _sym(T))          # eigen(VtAinvMAV, VtAV)
        W2 = V @ vecs[:, :nvec]
        if not defl and not restarted:
            W2tA = np.empty((nvec, n))
        restarted = True
        wcap = spdim - nvec
        col = 0
    return x, it, res_norm[:it].copy(), W2


def trrrpcg(A, b, x, M, nvec, spdim, maxit=0, eps=EPS):
    """trrrdefpcg.jl:234-355."""
    return _thick_restart(A, b, x, M, None, nvec, spdim, maxit, eps, False)


def trrrdefpcg(A, b, x, W, M, spdim, maxit=0, eps=EPS):
    """trrrdefpcg.jl:48-199 (positional order of the reference: A, b, x, W, M, spdim)."""
    return _thick_restart(A, b, x, M, W, W.shape[1], spdim, maxit, eps, False)


def trhrpcg(A, b, x, M, nvec, spdim, maxit=0, eps=EPS):
    """trhrdefpcg.jl:247-389."""
    return _thick_restart(A, b, x, M, None, nvec, spdim, maxit, eps, True)


def trhrdefpcg(A, b, x, W, M, spdim, maxit=0, eps=EPS):
    """trhrdefpcg.jl:48-208."""
    return _thick_restart(A, b, x, M, W, W.shape[1], spdim, maxit, eps, True)


# ---------------------------------------------------------------------------------------------
# Locally-optimal thick-restart variants: lotrrrpcg / lotrrrdefpcg (lotrrrdefpcg.jl:252-389, :48-214)
# and lotrhrpcg / lotrhrdefpcg (lotrhrdefpcg.jl:267-428, :48-229).
#
# As the thick-restart variants above, with the restart of eigCG in place of the plain projection: the new basis spans
# the nvec lowest (harmonic) Ritz vectors of the search space [W2 | window] AND those of its leading (spdim - 1) block
# (eigcg.jl:220-228 applied to the projected matrices), so W2 carries nev in [nvec, 2 nvec] columns, the window has
# spdim - nev and the first nvec columns of W2 are returned.  Reading of the reference where it is inconsistent: the
# blocks of the projected matrices follow nev (lotrrrdefpcg.jl:146 mixes nvec and nev in the mirrored block, which
# Symmetric() never reads); H = Q'(T Q) and G = Q'(B Q) are symmetrised (Julia's eigen takes the symmetric path only
# when they are symmetric to the last bit; otherwise it returns the same eigenvectors with another column scaling --
# the span, which is what a deflation basis is, does not change).
# ---------------------------------------------------------------------------------------------
def _rayleigh_ritz_general(G, B, nvec):
    """eigCG-type restart for the pencil (G, B), B = None for the identity: returns (Q Z, nev)."""
    import scipy.linalg as sla
    m = G.shape[0]
    eig = (lambda g, b: np.linalg.eigh(g)[1]) if B is None else (lambda g, b: sla.eigh(g, b)[1])
    Y = np.zeros((m, 2 * nvec))
    Y[:, :nvec] = eig(G, B)[:, :nvec]
    Y[:m - 1, nvec:] = eig(G[:m - 1, :m - 1], None if B is None else B[:m - 1, :m - 1])[:, :nvec]
    nev = _julia_rank(Y)
    Q = np.linalg.svd(Y, full_matrices=False)[0][:, :nev]
    H = _sym(Q.T @ (G @ Q))
    if B is None:
        Z = np.linalg.eigh(H)[1]
    else:
        Z = sla.eigh(H, _sym(Q.T @ (B @ Q)))[1]
    return Q @ Z, nev


def _lo_thick_restart(A, b, x, M, W, nvec, spdim, maxit, eps, harmonic):
    n = x.size
    defl = W is not None
    if defl:
        nvec = W.shape[1]
        WtA = np.empty((nvec, n))
        for iv in range(nvec):
            WtA[iv, :] = A(W[:, iv])
        WtAW = _sym(WtA @ W)
        WtW = _sym(W.T @ W)
        W2 = W.copy()
        r = b - A(x)
        x += W @ np.linalg.solve(WtAW, W.T @ r)
    else:
        W2 = np.zeros((n, nvec))
    nev = nvec if defl else 0                                # columns of W2 inside the search space
    restarted = False
    C = np.zeros((n, spdim))
    maxit = n if maxit == 0 else maxit
    res_norm = np.empty(max(n, maxit) + 1)
    it = 1
    r = b - A(x)
    rTr = r @ r
    res_norm[0] = np.sqrt(rTr)
    z = M(r)
    rTz = r @ z
    p = z - W @ np.linalg.solve(WtAW, WtA @ z) if defl else z.copy()
    C[:, 0] = p if harmonic else z / rTz ** 0.5
    col = 1
    tol = eps * np.linalg.norm(b)
    while it < maxit and res_norm[it - 1] > tol:
        Ap = A(p)
        d = p @ Ap
        alpha = rTz / d
        beta = 1.0 / rTz
        x += alpha * p
        r -= alpha * Ap
        if defl:
            r -= W @ np.linalg.solve(WtW, W.T @ r)
        rTr = r @ r
        z = M(r)
        rTz = r @ z
        beta *= rTz
        p = beta * p + z - W @ np.linalg.solve(WtAW, WtA @ z) if defl else beta * p + z
        it += 1
        res_norm[it - 1] = np.sqrt(rTr)
        C[:, col] = p if harmonic else z / rTz ** 0.5
        col += 1
        wcap = spdim - nev
        if col < wcap:
            continue
        # ---- the window is full: eigCG-type restart on [W2 | window] --------------------------------------------------
        Cw = C[:, :wcap]
        ACw = np.column_stack([A(Cw[:, j]) for j in range(wcap)])
        V = np.column_stack([W2[:, :nev], Cw]) if nev else Cw
        T = np.zeros((spdim, spdim))
        if nev:
            W2tA = np.vstack([A(W2[:, iv]) for iv in range(nev)])
            W2tAW2 = _sym(W2tA @ W2[:, :nev])
        if not harmonic:
            if nev:
                T[:nev, :nev] = W2tAW2
                T[:nev, nev:] = W2tA @ Cw
            T[nev:, nev:] = Cw.T @ ACw
            QZ, nev_new = _rayleigh_ritz_general(_sym(T), None, nvec)
        else:
            MACw = np.column_stack([M(ACw[:, j]) for j in range(wcap)])
            G = np.zeros((spdim, spdim))
            if nev:
                MAW2 = np.column_stack([M(W2tA[iv, :]) for iv in range(nev)])
                G[:nev, :nev] = W2tA @ MAW2
                G[:nev, nev:] = W2tA @ MACw
                T[:nev, :nev] = W2tAW2
            G[nev:, nev:] = ACw.T @ MACw
            T[nev + np.arange(wcap), nev + np.arange(wcap)] = np.einsum("ij,ij->j", Cw, ACw)
            QZ, nev_new = _rayleigh_ritz_general(_sym(G), _sym(T), nvec)
        W2 = V @ QZ
        nev = nev_new
        restarted = True
        col = 0
    if not defl and not restarted:
        return x, it, res_norm[:it].copy(), np.zeros((n, nvec))
    return x, it, res_norm[:it].copy(), W2[:, :nvec].copy()


def lotrrrpcg(A, b, x, M, nvec, spdim, maxit=0, eps=EPS):
    """lotrrrdefpcg.jl:252-389."""
    return _lo_thick_restart(A, b, x, M, None, nvec, spdim, maxit, eps, False)


def lotrrrdefpcg(A, b, x, W, M, spdim, maxit=0, eps=EPS):
    """lotrrrdefpcg.jl:48-214 (positional order of the reference: A, b, x, W, M, spdim)."""
    return _lo_thick_restart(A, b, x, M, W, W.shape[1], spdim, maxit, eps, False)


def lotrhrpcg(A, b, x, M, nvec, spdim, maxit=0, eps=EPS):
    """lotrhrdefpcg.jl:267-428."""
    return _lo_thick_restart(A, b, x, M, None, nvec, spdim, maxit, eps, True)


def lotrhrdefpcg(A, b, x, W, M, spdim, maxit=0, eps=EPS):
    """lotrhrdefpcg.jl:48-229."""
    return _lo_thick_restart(A, b, x, M, W, W.shape[1], spdim, maxit, eps, True)
