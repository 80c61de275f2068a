"""CPU oracle for the Schur-complement PCG hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, on the CPU, the algorithm of
the reference (venkovic/julia-phd-krylov-spdes) for the hot path named in
BASELINE.json.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
only as the checker or the CPU arm -- never as the product path.

PARITY UNPINNED: the reference ships no tests, golden vectors or data files
 and Julia is not installed in the build container, so this
restatement cannot be pinned against reference outputs.  It is pinned instead
against the algebraic identities the reference's own example scripts check
(Example03_EllipticPdeDomainDecomposition.jl:175 and :204) -- see
``tests/test_oracle_*.py`` -- and against the iteration counts a prototype written
independently from the reference during the survey recorded for five configurations
(reproduced exactly for the constant coefficient, within +-2
for the lognormal fields; tests/test_oracle_identities.py).  Every function cites the
reference file:line it follows.

Index convention: everything here is 0-based (nodes, elements, subdomains,
interface ids) except ``node_owner``, which keeps the reference's values
(-1 interface, 0 Dirichlet, d+1 interior of subdomain d).  ``to_julia`` helpers
in ``oracle.epdd`` convert maps to the reference's 1-based values.
"""
