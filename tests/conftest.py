import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import krylov_b200  # noqa: E402  (import alias of julia-phd-krylov-spdes_b200/)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


_cache = {}


def get_problem(nx, ndom, ireal=1, jitter=0.2, median=False):
    """Small workload (mesh, partition, maps, blocks, factors), cached across tests."""
    from krylov_b200.hostprep import problem as PB
    key = (nx, ndom, ireal, jitter, median)
    if key not in _cache:
        _cache[key] = PB.build_problem(nx, ndom, ireal=ireal, jitter=jitter, median=median)
    return _cache[key]


def get_oracle_ops(pr):
    """Exact-interior oracle operators for a problem: (inner, Sd, PiSd, b_schur)."""
    import oracle.epdd as E
    key = ("ops", id(pr))
    if key not in _cache:
        inner = E.ExactInterior(pr.blocks.A_II)
        Sd = E.assemble_local_schurs(pr.blocks, inner)
        Pi = E.prepare_neumann_neumann_schur_precond(Sd)
        bs = E.get_schur_rhs(pr.blocks, inner, pr.sd)
        _cache[key] = (inner, Sd, Pi, bs)
    return _cache[key]


@pytest.fixture(scope="session")
def small_problem():
    return get_problem(33, 4)


@pytest.fixture(scope="session")
def medium_problem():
    return get_problem(65, 16)
