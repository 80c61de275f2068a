"""Host-side stand-in for the reference's McmcSampler (Fem/Samplers.jl:93-199): a random-walk
Metropolis chain on the latent coordinates xi of a linear-Gaussian expansion
g = sum_k sqrt(Lambda_k) xi_k Psi[:, k], the coefficient being exp(g) (Example09: a chain of correlated
lognormal realizations solved with eigpcg -> eigdefpcg).  The reference feeds it Karhunen-Loeve pairs
read from data/*.kl-eigvals.npz / *.kl-eigvecs.npz; those files do not exist here, so `rff_modes` builds
an expansion of the same form for the squared-exponential covariance sig2 exp(-|d|^2 / L^2)
(Fem/Covariances.jl:23-27) from random Fourier features.  Random streams differ from Julia's, so chains
are statistically, not bit-wise, the reference's.  Stays on the host."""
import numpy as np


def rff_modes(points, m=128, sig2=1.0, L=0.1, seed=481456):
    """(Lambda, Psi): Psi = [cos(w_k.x) | sin(w_k.x)] (nnode x 2m), Lambda_k = sig2 / m, with
    w_k ~ N(0, (2 / L^2) I): E[g g'] = sig2 exp(-|d|^2 / L^2) for xi ~ N(0, I)."""
    rng = np.random.default_rng(seed)
    w = rng.normal(scale=np.sqrt(2.0) / L, size=(m, 2))
    arg = points.T @ w.T                                     # nnode x m
    Psi = np.concatenate([np.cos(arg), np.sin(arg)], axis=1)
    Lam = np.full(2 * m, sig2 / m)
    return Lam, Psi


class McmcSampler:
    """prepare_mcmc_sampler + draw! (Samplers.jl:130-199)."""

    def __init__(self, Lam, Psi, seed=0):
        self.n, self.m = Psi.shape                           # :132
        self.rng = np.random.default_rng(seed)
        self.sig2 = 1.0                                      # :134
        self.xi = self.rng.normal(scale=np.sqrt(self.sig2), size=self.m)            # :135-136
        self.theta2 = 2.38 ** 2 * self.sig2 / self.m         # :138
        self.chi = self.xi + self.rng.normal(scale=np.sqrt(self.theta2), size=self.m)   # :139-140
        self.Lam, self.Psi = np.asarray(Lam, dtype=np.float64), np.asarray(Psi, dtype=np.float64)
        self.g = self.Psi @ (np.sqrt(self.Lam) * self.xi)    # :142-145

    def draw(self):
        """draw! (:163-199): proposes xi + N(0, theta2 I) until accepted with probability
        min(1, exp((|xi|^2 - |chi|^2) / (2 sig2))); returns the number of proposals."""
        sq_xi = float(self.xi @ self.xi)
        cnt = 0
        while True:
            self.chi = self.xi + self.rng.normal(scale=np.sqrt(self.theta2), size=self.m)
            sq_chi = float(self.chi @ self.chi)
            alpha = np.exp((sq_xi - sq_chi) / 2.0 / self.sig2) if sq_xi < sq_chi else 1.0
            cnt += 1
            if self.rng.random() <= alpha:                    # `while u > alpha` (:178)
                break
        self.xi = self.chi.copy()
        self.g = self.Psi @ (np.sqrt(self.Lam) * self.xi)
        return cnt
