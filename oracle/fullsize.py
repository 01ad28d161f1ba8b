"""Oracle (test infrastructure only): parity checks at the full BASELINE sizes, where the faithful O(n^2) / 2 n_obs-solve
reference cannot run in full ("Parity checks in the same run").

* G: pattern and values of a row subsample against the reference's O(n) row formula (ForcingCovariance.py:159-167);
* sparse part: a subsample of the rows of Cu and the mean mu, recomputed with the reference's loop -- two direct
  solves and one G product per sensor (solving_utils.py:49-53, 139-142; LinearSolver.py:214-220).  The direct solver
  is SciPy SuperLU in 2-D and the DST solver of ``oracle.fast_poisson`` in 3-D (LU fill-in does not fit), whose
  solutions are verified against the assembled matrix;
* dense part: logposterior, its gradient, posterior mean and covariance at the sensors recomputed with SciPy
  (LinearSolver.py:344-373, 592-691) from the Cu under test.
"""
import time
import numpy as np
import scipy.sparse as sp
from .forcing import compute_G_rows
from .solver import OracleLinearSolver, _DirectSolver, interp_covariance_to_data
from .obsdata import OracleObsData
from .fast_poisson import DSTSolver


def rel(a, c):
    a, c = np.asarray(a, dtype=np.float64), np.asarray(c, dtype=np.float64)
    return float(np.max(np.abs(a - c)) / max(np.max(np.abs(c)), 1e-300))


def check_G_rows(coords, int_basis, sigma, l, cutoff, reg, indptr, indices, vals, rows):
    "-> (pattern identical, max relative value error) over ``rows`` of the CSR matrix under test"
    gd, _ = compute_G_rows(coords, int_basis, sigma, l, cutoff, reg, rows=[int(r) for r in rows])
    ok, err = True, 0.
    for r in rows:
        vals_ref, cols_ref = gd[int(r)]
        cols = indices[indptr[r]:indptr[r + 1]]
        same = bool(np.array_equal(cols, cols_ref))
        ok &= same
        if same and len(cols):
            err = max(err, float(np.max(np.abs(vals[indptr[r]:indptr[r + 1]] - vals_ref) / np.abs(vals_ref))))
    return ok, err


def check_prior(A_scipy, bc_nodes, b, G_scipy, P, sensors, y, unc, mu, Cu, columns, nodes_per_side=None, dim=None,
                direct="lu", theta=None, dense=None):
    """``mu``, ``Cu``: the results under test (NumPy).  ``columns``: sensor indices whose rows of Cu are recomputed.
    ``dense``: dict of results under test computed from (mu, Cu) at ``theta`` -- any of logpost, logpost_deriv, muy,
    Cuy -- compared with SciPy on the same Cu.  Returns a dict of relative errors and timings."""
    out = {}
    t = time.perf_counter()
    if direct == "lu":
        solver = _DirectSolver(A_scipy, bc_nodes)
    else:
        solver = DSTSolver(A_scipy, bc_nodes, nodes_per_side, dim)
    out["cpu_solver"] = "scipy splu" if direct == "lu" else "DST-I fast Poisson solver, residual-verified against the assembled A"
    P = sp.csc_matrix(P)
    err_rows, worst_res = 0., 0.
    for cidx in columns:
        row = interp_covariance_to_data(P, G_scipy, solver, P, col_range=(int(cidx), int(cidx) + 1))
        err_rows = max(err_rows, rel(Cu[int(cidx):int(cidx) + 1], row))
        if direct != "lu":
            p = np.asarray(P[:, int(cidx)].todense()).ravel()
            worst_res = max(worst_res, solver.residual(solver.solve(p, False), p))
    x = solver.solve(b, bcs_on=True)
    omu = P.T @ x
    out.update(columns_checked=[int(c) for c in columns], Cu_rows_max_rel=err_rows, mu_rel=rel(mu, omu),
               cpu_solves_s=time.perf_counter() - t)
    if direct != "lu":
        out["cpu_solver_residual_max"] = max(worst_res, solver.residual(x, b, bcs_on=True))
    if dense:
        t = time.perf_counter()
        ols = OracleLinearSolver.__new__(OracleLinearSolver)
        ols.data = OracleObsData(sensors, y, unc)
        ols.mu, ols.Cu, ols.params = omu, np.asarray(Cu), None
        theta = np.asarray(theta, dtype=np.float64)
        if "logpost" in dense:
            ref = float(ols.logposterior(theta))
            out["logpost_rel"] = abs(dense["logpost"] - ref) / abs(ref)
        if "logpost_deriv" in dense:
            out["logpost_deriv_rel"] = rel(dense["logpost_deriv"], ols.logpost_deriv(theta))
        if "muy" in dense or "Cuy" in dense:
            ols.set_params(theta)
            omuy, oCuy = ols.solve_posterior_covariance()
            if "muy" in dense:
                out["posterior_mean_rel"] = rel(dense["muy"], omuy)
            if "Cuy" in dense:
                out["posterior_cov_rel"] = rel(dense["Cuy"], oCuy)
        out["cpu_dense_s"] = time.perf_counter() - t
    return out
