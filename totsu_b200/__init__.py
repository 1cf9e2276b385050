"""totsu_b200 — B200 (sm_100a) linear-algebra / cone backend for the Totsu first-order conic solver.

The product is `libtotsu_b200.so` (hand-written CUDA kernels behind the C ABI of include/totsu_b200.h) and the
C++ host layer in totsu_b200/host (mirror of the reference's LinAlg / SliceLike / Operator / Cone / Solver
surface, exported to Python by `libtotsu_b200_host.so`).  This package only loads them.
"""
from . import capi  # noqa: F401
