"""wrf_b200 -- B200-native WRF-ARW dry dynamical-core step (solve_em hot path).

Only what the path needs: csrc/ (CUDA kernels + C-ABI), lib.py (host-side mirror
of the reference interface over the C-ABI), ideal.py (initial states for the
shipped ideal configs), decomp.py (RSL_LITE patch decomposition), fortran/ (the
ISO_C_BINDING shim a WRF build links instead of dyn_em/solve_em.o).
"""
from .ideal import Case, Config, Dims, make_case  # noqa: F401
from .lib import Solver, WrfGpuError, comm_init_local, load, library_path, run_ranks  # noqa: F401
from . import decomp  # noqa: F401
