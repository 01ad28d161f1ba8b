"""
oracle/ -- CPU restatement of stat-fem's data-conditioning hot path.  TEST INFRASTRUCTURE ONLY.

This package is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``stat-fem_b200/`` imports it, and the product raises if its CUDA library is missing.

The reference (``/root/reference/stat_fem``) is pure Python that delegates its arithmetic to NumPy/SciPy
and, through Firedrake, to PETSc (un-vendored, un-pinned: ``setup.py:47``, ``docker/Dockerfile:3``).
Every function here follows the reference statement it cites (``file:line`` relative to
``/root/reference/stat_fem``) using NumPy/SciPy only; the sparse solve is ``scipy.sparse.linalg.splu``
(direct LU -- the numerical class of Firedrake's default ``preonly + lu``).

Pinning (see ``tests/golden/make_golden.py``, ``tests/test_oracle_golden.py``):
  * the reference's own modules are executed *unmodified* in the build container on top of a small
    NumPy/SciPy stand-in for Firedrake/PETSc (``oracle/fake_firedrake``) and their outputs are committed
    as fixtures under ``tests/golden/``; the oracle is checked against every one of them;
  * the dense 1-D goldens that the reference's tests hold (``tests/helper_funcs.py:60-239``) are replayed.
Unpinned by any reference test (only the code defines them): G's pattern for ``cutoff > 0``, MAP values,
every 2-D/3-D result.  For those the pin is the unmodified-reference-on-the-stand-in output.
"""
from .covariance import sqexp, sqexp_deriv, calc_r2
from .forcing import compute_G_rows, compute_G_csr, compute_G_csr_fast
from .interp import interpolate_cell, build_interpolation_matrix
from .solver import (solve_forcing_covariance, interp_covariance_to_data, OracleLinearSolver,
                     estimate_params_MAP)
from .obsdata import OracleObsData
