"""Full-size parity check of BASELINE config 2 against the oracle: the GPU computes mu and the full
Cu (4096 x 4096) at 1 048 576 DOF; the CPU oracle (SciPy sparse LU = the reference's direct-solver class) recomputes
a subsample of the columns with the reference's two-solves-per-sensor loop, and all dense quantities (logposterior,
gradient, posterior mean and covariance at the sensors) with SciPy from the GPU's Cu.  Prints one JSON object.

    python tools/parity_c2.py [nodes] [n_obs] [n_check_columns]
"""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import stat_fem_b200 as stat_fem
from stat_fem_b200 import fem
import oracle

nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nobs = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
ncheck = int(sys.argv[3]) if len(sys.argv) > 3 else 8

V, A, b = fem.poisson_problem(nodes, 2)
h = 1. / (nodes - 1)
sig, lf = np.log(2.e-2), np.log(1.5 * h)
reg = 1.e-8 * h ** 4 * np.exp(2 * sig)
rng = np.random.default_rng(1)
s = rng.uniform(0.02, 0.98, (nobs, 2))
y = 0.7 * np.prod(np.sin(2 * np.pi * s), axis=1) + np.random.default_rng(100).normal(scale=2.e-3, size=nobs)
od = stat_fem.ObsData(s, y, 2.e-3)
G = stat_fem.ForcingCovariance(V, sig, lf, 1.e-3, reg)
ls = stat_fem.LinearSolver(A, b, G, od)
t = time.perf_counter()
mu, Cu = ls.solve_prior()
t_gpu = time.perf_counter() - t
theta = np.array([np.log(0.7), np.log(1.e-2), np.log(0.5)])
lp, gr = ls.logposterior(theta), ls.logpost_deriv(theta)
ls.set_params(theta)
muy, Cuy = ls.solve_posterior_covariance()

rel = lambda a, c: float(np.max(np.abs(np.asarray(a) - np.asarray(c))) / np.max(np.abs(c)))
out = dict(nodes=nodes, n=V.dim(), n_obs=nobs, gpu_solve_prior_s=t_gpu, cg_iterations=list(ls.solver.block_iterations))

# ---- G pattern and values on a row subsample against the reference's O(n) row formula
ip, ix, v = G.G.to_host()
rows = np.unique(np.concatenate([rng.choice(V.dim(), 48, replace=False), [0, 1, nodes - 1, nodes, V.dim() - 1]]))
gd, _ = oracle.compute_G_rows(V.coords, V.integrate_basis_functions(), sig, lf, 1.e-3, reg, rows=[int(r) for r in rows])
pat_ok, val_err = True, 0.
for r in rows:
    vals_ref, cols_ref = gd[int(r)]
    cols = ix[ip[r]:ip[r + 1]]
    pat_ok &= bool(np.array_equal(cols, cols_ref))
    if np.array_equal(cols, cols_ref):
        val_err = max(val_err, float(np.max(np.abs(v[ip[r]:ip[r + 1]] - vals_ref) / np.abs(vals_ref))))
out.update(G_rows_checked=int(len(rows)), G_pattern_identical=pat_ok, G_values_max_rel=val_err)

# ---- sparse part: reference loop on a column subsample with a direct solver
n = V.dim()
Gs = sp.csr_matrix((v, ix, ip), shape=(n, n))
P = ls.im.interp        # host-built interpolation matrix (the oracle's brute-force cell search is O(cells) per sensor)
t = time.perf_counter()
ols = oracle.OracleLinearSolver(A.to_scipy(), b.data, Gs, oracle.OracleObsData(s, y, 2.e-3), P, A.bc_nodes())
cols = sorted(int(c) for c in rng.choice(nobs, ncheck, replace=False))
err_rows = 0.
for cidx in cols:
    row = oracle.interp_covariance_to_data(P, Gs, ols.solver, P, col_range=(cidx, cidx + 1))
    err_rows = max(err_rows, rel(Cu[cidx:cidx + 1], row))
omu = P.T @ ols.solver.solve(b.data, bcs_on=True)
out.update(cpu_factor_and_columns_s=time.perf_counter() - t, columns_checked=cols, Cu_rows_max_rel=err_rows, mu_rel=rel(mu, omu))

# ---- dense part: SciPy on the GPU's Cu (the reference's LinearSolver.py:592-691, 344-373)
ols.mu, ols.Cu = omu, Cu
out.update(logpost_gpu=lp, logpost_ref=float(ols.logposterior(theta)))
out["logpost_rel"] = abs(out["logpost_gpu"] - out["logpost_ref"]) / abs(out["logpost_ref"])
out["logpost_deriv_rel"] = rel(gr, ols.logpost_deriv(theta))
ols.set_params(theta)
omuy, oCuy = ols.solve_posterior_covariance()
out.update(posterior_mean_rel=rel(muy, omuy), posterior_cov_rel=rel(Cuy, oCuy))
print(json.dumps(out))
