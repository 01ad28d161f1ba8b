"""
Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference modules
(/root/reference/stat_fem/*.py) on top of oracle/fake_firedrake.py (a NumPy/SciPy stand-in for the
Firedrake/PETSc/mpi4py layer the reference sits on).  Run in the build container only:

    python tests/golden/make_golden.py

/root/reference does not exist on the GPU box, so tests never call this script; they read the .npz files.
Problem inputs (mesh, stiffness, load) come from the repo's host-side generator ``stat_fem_b200.fem`` and
are stored in the fixture, so the fixtures are self-contained.
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import fake_firedrake as ff          # noqa: E402
import stat_fem_b200                              # noqa: E402,F401
from stat_fem_b200 import fem                     # noqa: E402

stat_fem = ff.install("/root/reference")
FC = stat_fem.ForcingCovariance.ForcingCovariance
IM = stat_fem.InterpolationMatrix.InterpolationMatrix
OD = stat_fem.ObsData.ObsData
LS = stat_fem.LinearSolver.LinearSolver
CF = stat_fem.covariance_functions


def g_to_csr(fc):
    G_dict, max_nnz = fc._compute_G_vals()
    n = fc.nx
    indptr = np.zeros(n + 1, dtype=np.int64)
    for i in range(n):
        indptr[i + 1] = indptr[i] + len(G_dict[i][1])
    indices = np.concatenate([G_dict[i][1] for i in range(n)]).astype(np.int32)
    vals = np.concatenate([G_dict[i][0] for i in range(n)])
    return indptr, indices, vals, max_nnz


def run_case(name, dim, nodes, sigma_f, l_f, cutoff, reg, sensors, data, unc, params, predict_coords=None,
             map_start=None, map_kwargs=None):
    V, A, b = fem.poisson_problem(nodes, dim)
    if name == "ref1d":
        # the reference's own 1-D fixture uses f = 4 pi^2 sin(2 pi x) (tests/helper_funcs.py:40-48)
        f = fem.Function(V)
        f.interpolate(lambda X: 4. * np.pi * np.pi * np.sin(2. * np.pi * X[:, 0]))
        b = fem.assemble_load(V, f)
    coords, cells = V.coords, V.mesh().cells
    int_basis = V.integrate_basis_functions()
    bc_nodes = A.bc_nodes()
    fs = ff.FakeSpace(coords, cells, int_basis)
    fA = ff.FakeMatrix(A.to_scipy(), bc_nodes)
    fb = ff.Function(fs, b.data.copy())
    fc = FC(fs, sigma_f, l_f, cutoff, reg)
    indptr, indices, vals, max_nnz = g_to_csr(fc)
    fc.assemble()
    od = OD(sensors, data, unc)
    ls = LS(fA, fb, fc, od)
    mu, Cu = ls.solve_prior()
    Pdense = ls.im.interp._csr.toarray()
    out = dict(dim=dim, nodes=nodes, coords=coords, cells=cells, int_basis=int_basis, bc_nodes=bc_nodes,
               A_indptr=A.indptr, A_indices=A.indices, A_data=A.data, b=b.data,
               sigma_f=sigma_f, l_f=l_f, cutoff=cutoff, reg=reg,
               G_indptr=indptr, G_indices=indices, G_vals=vals, G_max_nnz=max_nnz,
               sensors=np.asarray(sensors, dtype=np.float64), data=np.asarray(data, dtype=np.float64),
               unc=np.asarray(unc, dtype=np.float64), params=np.asarray(params, dtype=np.float64),
               P=Pdense, mu=mu, Cu=Cu, x_prior=ls.x.vector().gather())
    # A^-1 G A^-1 1  (tests/test_solving_utils.py:16-33)
    rhs = ff.Function(fs, np.ones(fs.dim())).vector()
    out["AGA_ones"] = stat_fem.solving_utils.solve_forcing_covariance(fc, ls.solver, rhs).gather()
    y = ff.Function(fs).vector()
    fc.mult(ff.Function(fs, np.ones(fs.dim())).vector(), y)
    out["G_ones"] = y.gather()
    ls.set_params(params)
    u = ff.Function(fs)
    ls.solve_posterior(u)
    out["post_mean"] = u.vector().gather()
    u2 = ff.Function(fs)
    ls.solve_posterior(u2, scale_mean=True)
    out["post_mean_scaled"] = u2.vector().gather()
    out["muy"], out["Cuy"] = ls.solve_posterior_covariance()
    out["muy_scaled"], _ = ls.solve_posterior_covariance(scale_mean=True)
    out["m_eta"], out["C_eta"] = ls.solve_prior_generating()
    if np.asarray(unc).shape == ():
        out["m_etay"], out["C_etay"] = ls.solve_posterior_generating()
    out["logpost"] = ls.logposterior(params)
    out["logpost_deriv"] = ls.logpost_deriv(params)
    pts = np.array([[0.3, -0.2, 0.1], [-1.2, 0.4, -0.8], [0.05, -2.0, 0.7]])
    out["lp_points"] = pts
    out["lp_values"] = np.array([ls.logposterior(p) for p in pts])
    out["lp_derivs"] = np.array([ls.logpost_deriv(p) for p in pts])
    ls.set_params(params)
    if predict_coords is not None:
        out["predict_coords"] = np.asarray(predict_coords, dtype=np.float64)
        out["predict_mean"] = ls.predict_mean(predict_coords)
        out["predict_mean_unscaled"] = ls.predict_mean(predict_coords, scale_mean=False)
        out["predict_cov"] = ls.predict_covariance(predict_coords, 0.1)
    if map_start is not None:
        ls_map = stat_fem.estimation.estimate_params_MAP(fA, fb, fc, od, start=np.array(map_start),
                                                         **(map_kwargs or {}))
        out["map_start"] = np.asarray(map_start, dtype=np.float64)
        out["map_params"] = ls_map.params
        out["map_logpost"] = ls_map.logposterior(ls_map.params)
        out["map_kwargs_ftol"] = (map_kwargs or {}).get("ftol", np.nan)
        out["map_kwargs_gtol"] = (map_kwargs or {}).get("gtol", np.nan)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "n =", fs.dim(), "nnz(G) =", indptr[-1], "max_row_nnz =", max_nnz,
          "logpost =", out["logpost"])


def covariance_goldens():
    rng = np.random.default_rng(7)
    out = {}
    for d in (1, 2, 3):
        x1 = rng.uniform(0, 1, (9, d))
        x2 = rng.uniform(0, 1, (13, d))
        out[f"x1_{d}"], out[f"x2_{d}"] = x1, x2
        out[f"sqexp_{d}"] = CF.sqexp(x1, x2, np.log(0.7), np.log(0.3))
        out[f"deriv_{d}"] = CF.sqexp_deriv(x1, x2, np.log(0.7), np.log(0.3))
        out[f"r2_{d}"] = CF.calc_r2(x1, x2)
        od = OD(x1, rng.normal(size=9), rng.uniform(0.01, 0.2, 9))
        out[f"K_{d}"] = od.calc_K(np.array([-1.1, -0.4]))
        out[f"Ks_{d}"] = od.calc_K_plus_sigma(np.array([-1.1, -0.4]))
        out[f"Kd_{d}"] = od.calc_K_deriv(np.array([-1.1, -0.4]))
        out[f"unc_{d}"] = od.get_unc()
    # interpolate_cell known answers (tests/test_InterpolationMatrix.py:255-291) + random simplices
    ic = stat_fem.InterpolationMatrix.interpolate_cell
    for d in (1, 2, 3):
        nodes = rng.uniform(0, 1, (d + 1, d))
        w = rng.dirichlet(np.ones(d + 1))
        pt = w @ nodes
        out[f"ic_nodes_{d}"], out[f"ic_pt_{d}"], out[f"ic_w_{d}"] = nodes, pt, ic(pt, nodes)
    np.savez_compressed(os.path.join(HERE, "covariance.npz"), **out)
    print("covariance goldens written")


if __name__ == "__main__":
    covariance_goldens()
    # 1. the reference's own 1-D fixture (tests/helper_funcs.py)
    run_case("ref1d", 1, 11, np.log(1.), np.log(0.1), 0., 1.e-8,
             np.array([[0.75], [0.5], [0.25], [0.125]]), np.array([-1.185, -0.06286, 0.8934, 0.7962]), 0.1,
             [-0.9, 0.1, -0.5], predict_coords=np.array([[0.05], [0.825], [0.45], [0.225], [0.775]]),
             map_start=[0., 0., 0.])
    # 2. 2-D, the example's G parameters (examples/poisson.py:45-46): structurally non-symmetric G
    rng = np.random.default_rng(11)
    s = rng.uniform(0.05, 0.95, (20, 2))
    y = 0.7 * np.sin(2 * np.pi * s[:, 0]) * np.sin(2 * np.pi * s[:, 1]) + rng.normal(scale=0.01, size=20)
    run_case("sq11_example", 2, 11, np.log(2.e-2), np.log(0.354), 1.e-3, 1.e-8, s, y, 2.e-3,
             [np.log(0.7), np.log(1.e-2), np.log(0.5)], predict_coords=rng.uniform(0.1, 0.9, (7, 2)),
             map_start=[0., 0., 0.], map_kwargs=dict(ftol=1.e-15, gtol=1.e-10))
    # 3. 2-D, benchmark-style scaled nugget and short correlation length: ~100 nnz/row
    h = 1. / 32.
    sig = np.log(2.e-2)
    s = rng.uniform(0.02, 0.98, (24, 2))
    y = 0.7 * np.sin(2 * np.pi * s[:, 0]) * np.sin(2 * np.pi * s[:, 1]) + rng.normal(scale=0.01, size=24)
    run_case("sq33_cutoff", 2, 33, sig, np.log(1.5 * h), 1.e-3, 1.e-8 * h ** 4 * np.exp(2 * sig), s, y,
             rng.uniform(1.e-3, 4.e-3, 24), [np.log(0.7), np.log(1.e-2), np.log(0.5)],
             map_start=[0., 0., 0.], map_kwargs=dict(ftol=1.e-15, gtol=1.e-10))
    # 4. 3-D Kuhn mesh
    h = 1. / 8.
    s = rng.uniform(0.05, 0.95, (16, 3))
    y = 0.7 * np.prod(np.sin(2 * np.pi * s), axis=1) + rng.normal(scale=0.01, size=16)
    run_case("cube9_cutoff", 3, 9, sig, np.log(0.8 * h), 1.e-3, 1.e-8 * h ** 6 * np.exp(2 * sig), s, y, 2.e-3,
             [np.log(0.7), np.log(1.e-2), np.log(0.5)], predict_coords=rng.uniform(0.1, 0.9, (5, 3)))
