"""The C-ABI library builds for sm_100a, loads on a CPU-only box, exports every symbol declared in
include/krylov_b200.h, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, has_gpu
from krylov_b200 import capi


@pytest.fixture(scope="module")
def built():
    return capi.build()


def test_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "krylov_b200.h")).read()
    declared = set(re.findall(r"\b(kb200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = C.CDLL(built)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(capi.SYMBOLS)


def test_sass_targets_sm_100a(built):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", built], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu(built):
    with pytest.raises(capi.KB200Error) as e:
        capi.Handle(0)
    assert e.value.code == -2 and "CUDA" in str(e.value)
