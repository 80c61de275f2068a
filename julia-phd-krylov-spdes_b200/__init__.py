"""krylov_b200: B200-native Schur-complement PCG hot path behind the reference's API.

Layout: ``csrc/`` CUDA kernels + C-ABI (libkrylov_b200.so), ``capi.py`` ctypes binding of
include/krylov_b200.h, ``fem.py`` / ``solvers.py`` the host-side mirror of the reference's
Fem / RecyclingKrylovSolvers call shapes, ``hostprep/`` the host-side preparation the Julia
side keeps (mesh, partition, index maps, block assembly, sparse Cholesky), ``julia/`` the
ccall shim a maintainer adds to the reference.
"""
