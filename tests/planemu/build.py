"""Builds tests/planemu/libplanemu.so (CPU emulator of the device program and of the host-side
multi-GPU bookkeeping; TEST TOOLING) when it is older than any of its sources."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "julia-phd-krylov-spdes_b200", "csrc")
SO = os.path.join(HERE, "libplanemu.so")
SRCS = [os.path.join(HERE, "planemu.cpp"), os.path.join(CSRC, "plan.cpp")]
DEPS = SRCS + [os.path.join(CSRC, f) for f in ("plan.h", "xchg_tables.h")]


def build(force=False):
    if force or not os.path.exists(SO) or any(os.path.getmtime(SO) < os.path.getmtime(d) for d in DEPS):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared"] + SRCS + ["-o", SO])
    return SO
