"""onephase.jl_b200 -- B200-native (sm_100a) replacement for the per-iteration KKT
solve of ohinder/OnePhase.jl, behind the reference's own plugin API.

    csrc/                 CUDA kernels, symbolic analysis, C ABI (libonephase_b200.so)
    _lib.py               ctypes binding of include/onephase_b200.h
    kkt.py                host-side mirror of linear_solver_* / *_KKT_solver / ipopt_strategy!

The seeded synthetic inputs of the BASELINE.json shapes are test data: tests/problems.py.

The directory name contains a dot, so load it with __graft_entry__.package()
(it registers the package as `onephase_jl_b200`).
"""
from . import _lib, kkt  # noqa: F401
from ._lib import Handle, OPBError, build, cache_clear, launch_count  # noqa: F401
from .kkt import (Class_iterate, Class_parameters, DistShard, Schur_B200_KKT_solver, Symmetric_B200_KKT_solver,  # noqa: F401
                  System_rhs, ThreadShard, compute_schur_diag, estimate_y_tilde, eval_diag_J_T_J, ipopt_strategy,
                  linear_solver_B200, pick_KKT_solver, respond_to_failed_step)
