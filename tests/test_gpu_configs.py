"""GPU parity tests at the BASELINE.json configurations:

* C1 end to end (the reference example, examples/poisson.py:15-124: 101^2 P1 nodes, 10 x 10 sensors) in both forcing
  covariance regimes against the CPU oracle -- G pattern, prior, log-posterior, MAP optimum, posterior mean/covariance;
* C2 at full size (1 048 576 DOF, 4096 sensors): G rows, a column subsample of Cu by the reference's two-solve loop
  with SciPy SuperLU, and every dense quantity recomputed with SciPy from the GPU's Cu;
* the MAP optimum must not depend on how Cu was assembled (symmetric split vs plain GEMM: the one-GPU and multi-GPU
  orderings of the same sum);
* the folded posterior-mean path (one solve through the resident W = G Z) against the reference's four-solve chain.
Tolerances (north_star): 1e-8 relative on means and covariances, 1e-6 on the log-posterior, gradient and MAP parameters."""
import os
import sys
import numpy as np
import scipy.sparse as sp
import pytest
from numpy.testing import assert_allclose

from conftest import relerr

pytestmark = pytest.mark.gpu

TOL, TOL_LP = 1.e-8, 1.e-6
MAP_KW = dict(ftol=1.e-15, gtol=1.e-10)       # SURVEY 8(d); estimation.py:92-93 passes them to L-BFGS-B


def c1_problem(regime):
    """examples/poisson.py:15-72 with datagrid = 10 and np.random.seed(0)"""
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import fem
    import oracle
    nodes = 101
    V, A, b = fem.poisson_problem(nodes, 2)
    h = 1. / (nodes - 1)
    sigma_f = np.log(2.e-2)
    if regime == "reference-defaults":          # SURVEY finding 3: G comes out diagonal
        l_f, cutoff, reg = np.log(0.354), 1.e-3, 1.e-8
    else:                                        # benchmark-style scaled nugget: ~100 entries per row
        l_f, cutoff, reg = np.log(1.5 * h), 1.e-3, 1.e-8 * h ** 4 * np.exp(2 * sigma_f)
    np.random.seed(0)
    g = (np.arange(10) + 1.) / 11.
    X, Y = np.meshgrid(g, g)
    sensors = np.stack([X.ravel(), Y.ravel()], axis=1)
    rho, sigma_eta, l_eta, sigma_y = np.log(0.7), np.log(1.e-2), np.log(0.5), 2.e-3
    Keta = oracle.sqexp(sensors, sensors, sigma_eta, l_eta)
    y = (np.exp(rho) * np.sin(2. * np.pi * sensors[:, 0]) * np.sin(2. * np.pi * sensors[:, 1]) +
         np.random.multivariate_normal(np.zeros(100), Keta) + np.random.normal(scale=sigma_y, size=100))
    G = stat_fem.ForcingCovariance(V, sigma_f, l_f, cutoff, reg)
    od = stat_fem.ObsData(sensors, y, sigma_y)
    return V, A, b, G, od, (sigma_f, l_f, cutoff, reg), sensors, y, sigma_y


@pytest.mark.parametrize("regime", ["reference-defaults", "scaled-nugget"])
def test_c1_end_to_end_against_oracle(regime):
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import fem
    import oracle
    V, A, b, G, od, (sigma_f, l_f, cutoff, reg), sensors, y, unc = c1_problem(regime)
    n = V.dim()
    # --- G: bit-identical pattern
    ip, ix, vals = oracle.compute_G_csr_fast(V.coords, V.integrate_basis_functions(), sigma_f, l_f, cutoff, reg)
    G.assemble()
    gip, gix, gv = G.G.to_host()
    assert np.array_equal(gip, ip) and np.array_equal(gix, ix)
    assert_allclose(gv, vals, rtol=1e-13, atol=0.)
    if regime == "reference-defaults":
        assert len(ix) == n          # diagonal
    else:
        assert 60 * n < len(ix) < 110 * n
    # --- oracle objects
    Gs = sp.csr_matrix((vals, ix, ip), shape=(n, n))
    P = oracle.build_interpolation_matrix(V.coords, V.mesh().cells, sensors)
    ols = oracle.OracleLinearSolver(A.to_scipy(), b.data, Gs, oracle.OracleObsData(sensors, y, unc), P, A.bc_nodes())
    omu, oCu = ols.solve_prior()
    ls = stat_fem.LinearSolver(A, b, G, od, solver_parameters={"ksp_rtol": 1e-13})
    mu, Cu = ls.solve_prior()
    assert relerr(mu, omu) < TOL and relerr(Cu, oCu) < TOL
    # --- log-posterior and gradient at a few points
    for th in ([0., 0., 0.], [np.log(0.7), np.log(1.e-2), np.log(0.5)], [-0.3, -4., -1.]):
        th = np.array(th)
        ref = ols.logposterior(th)
        assert abs(ls.logposterior(th) - ref) <= TOL_LP * abs(ref)
        gref = ols.logpost_deriv(th)
        assert_allclose(ls.logpost_deriv(th), gref, rtol=TOL_LP, atol=TOL_LP * max(1., np.max(np.abs(gref))))
    # --- MAP from zero.  The drop-in call (estimation.py:7-110) at SciPy's default tolerances: same optimum value
    _, fmin0 = oracle.estimate_params_MAP(ols, start=np.zeros(3))
    fit0 = stat_fem.estimate_params_MAP(A, b, G, od, start=np.zeros(3), solver_parameters={"ksp_rtol": 1e-13})
    assert abs(fit0.logposterior(fit0.params) - fmin0["fun"]) <= TOL_LP * max(1., abs(fmin0["fun"]))
    # ... and the optimum itself pinned to 1e-6: L-BFGS-B at ftol=1e-15 / gtol=1e-10 followed by the
    # gradient-only Newton polish on both sides (function values carry more rounding noise than 1e-6 in theta allows)
    from stat_fem_b200.estimation import newton_polish
    _, fmin = oracle.estimate_params_MAP(ols, start=np.zeros(3), **MAP_KW)
    theta_ref, info_ref = newton_polish(ols, fmin["x"])
    assert info_ref["converged"]
    fit = stat_fem.fit_snapshots(stat_fem.LinearSolver(A, b, G, od, solver_parameters={"ksp_rtol": 1e-13}), [od], start=np.zeros(3),
                                 polish=True, accept_precision_floor=True, **MAP_KW)[0]
    print(regime, "theta* gpu", fit.params, "oracle", theta_ref, "(L-BFGS-B only:", fmin["x"], ") evals", fit.n_logpost_evals,
          fmin["nfev"], fit.minimize_result["message"], fit.minimize_result["polish"]["last_update"])
    assert fit.minimize_result["polish"]["converged"]
    assert_allclose(fit.params, theta_ref, rtol=TOL_LP, atol=TOL_LP)
    assert abs(fit.logposterior(fit.params) - ols.logposterior(theta_ref)) <= TOL_LP * max(1., abs(fmin["fun"]))
    # --- posterior at the optimum
    ols.set_params(theta_ref)
    fit.set_params(theta_ref)
    x = fem.Function(V)
    fit.solve_posterior(x, scale_mean=True)
    assert relerr(x.data, ols.solve_posterior(scale_mean=True)) < TOL
    muy, Cuy = fit.solve_posterior_covariance()
    omuy, oCuy = ols.solve_posterior_covariance()
    assert relerr(muy, omuy) < TOL and relerr(Cuy, oCuy) < TOL


def test_c2_full_size_parity():
    """BASELINE config 2 at full size on the device against the CPU oracle on a subsample."""
    import stat_fem_b200 as stat_fem
    from oracle import fullsize
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    prob = bench.make_problem("c2")
    V, A, b = prob["V"], prob["A"], prob["b"]
    n, nobs = V.dim(), prob["nobs"]
    assert (n, nobs) == (1048576, 4096)
    G = stat_fem.ForcingCovariance(V, prob["sigma_f"], prob["l_f"], prob["cutoff"], prob["reg"])
    od = stat_fem.ObsData(prob["sensors"], prob["datas"][0], prob["unc"])
    ls = stat_fem.LinearSolver(A, b, G, od)
    mu, Cu = ls.solve_prior()
    assert max(ls.solver.block_iterations) <= 12
    theta = bench.THETA0
    dense = dict(logpost=ls.logposterior(theta), logpost_deriv=ls.logpost_deriv(theta))
    ls.set_params(theta)
    dense["muy"], dense["Cuy"] = ls.solve_posterior_covariance()
    ip, ix, v = G.G.to_host()
    rng = np.random.default_rng(12345)
    rows = np.unique(np.concatenate([rng.choice(n, 48, replace=False), [0, 1, 1023, 1024, n - 1]]))
    pat, verr = fullsize.check_G_rows(V.coords, V.integrate_basis_functions(), prob["sigma_f"], prob["l_f"], prob["cutoff"],
                                      prob["reg"], ip, ix, v, rows)
    assert pat and verr < 1e-13
    Gs = sp.csr_matrix((v, ix, ip), shape=(n, n))
    cols = sorted(int(c) for c in rng.choice(nobs, 8, replace=False))
    out = fullsize.check_prior(A.to_scipy(), A.bc_nodes(), b.data, Gs, ls.im.interp, prob["sensors"], prob["datas"][0], prob["unc"],
                               mu, Cu, cols, direct="lu", theta=theta, dense=dense)
    print("C2 full-size parity:", out)
    assert out["Cu_rows_max_rel"] < TOL and out["mu_rel"] < TOL
    assert out["posterior_mean_rel"] < TOL and out["posterior_cov_rel"] < TOL
    assert out["logpost_rel"] < TOL_LP and out["logpost_deriv_rel"] < TOL_LP
    # the posterior mean on the mesh at full size: the folded path against the reference's four-solve chain
    from stat_fem_b200 import fem
    x1 = fem.Function(V)
    ls.solve_posterior(x1, scale_mean=True)
    ls._W_dev = None                      # no resident W: two solves per term (solving_utils.py:49-53)
    x2 = fem.Function(V)
    ls.solve_posterior(x2, scale_mean=True)
    assert relerr(x1.data, x2.data) < TOL


def test_map_optimum_independent_of_cu_assembly():
    """theta* must agree to 1e-6 whether Cu came from the symmetric split (lower tiles + mirror, the one-GPU path) or
    the plain GEMM (the ordering of the sharded path): 257^2 nodes, 512 sensors, 2 snapshots."""
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import fem
    nodes = 257
    V, A, b = fem.poisson_problem(nodes, 2)
    h = 1. / (nodes - 1)
    sig, l = np.log(2.e-2), np.log(1.5 * h)
    reg = 1.e-8 * h ** 4 * np.exp(2 * sig)
    rng = np.random.default_rng(21)
    s = rng.uniform(0.02, 0.98, (512, 2))
    truth = 0.7 * np.prod(np.sin(2 * np.pi * s), axis=1)
    datas = []
    for k in range(2):          # bench.py's synthetic snapshots: smooth low-rank discrepancy + white noise
        disc = np.zeros(512)
        for m in range(6):
            kvec = rng.uniform(0.5, 2.5, 2)
            disc += 4.e-3 * rng.normal() * np.sin(np.pi * s @ kvec + rng.uniform(0, 2 * np.pi))
        datas.append(stat_fem.ObsData(s, truth + disc + rng.normal(scale=2.e-3, size=512), 2.e-3))
    G = stat_fem.ForcingCovariance(V, sig, l, 1.e-3, reg)
    thetas = {}
    for split in (True, False):
        ls = stat_fem.LinearSolver(A, b, G, datas[0], solver_parameters={"symmetric_split": split})
        fits = stat_fem.fit_snapshots(ls, datas, start=np.zeros(3), polish=True, accept_precision_floor=True, **MAP_KW)
        thetas[split] = np.array([f.params for f in fits])
        print("split", split, thetas[split], [f.n_logpost_evals for f in fits],
              [(f.minimize_result["message"], f.minimize_result["polish"]["iterations"], f.minimize_result["polish"]["last_update"])
               for f in fits])
        for f in fits:
            assert f.minimize_result["polish"]["converged"]
            # the polished point is a stationary point of the analytic gradient
            assert np.max(np.abs(f.logpost_deriv(f.params))) < 1e-4
    assert_allclose(thetas[True], thetas[False], rtol=TOL_LP, atol=TOL_LP)


def test_posterior_mean_folded_path_matches_four_solve_chain():
    "solve_posterior / solve_posterior_snapshots through the resident W against the chain without it, 2-D and 3-D"
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import fem
    for nodes, dim, nobs, lfac in ((65, 2, 300, 1.5), (17, 3, 60, 0.8)):
        V, A, b = fem.poisson_problem(nodes, dim)
        h = 1. / (nodes - 1)
        sig, l = np.log(2.e-2), np.log(lfac * h)
        reg = 1.e-8 * h ** (2 * dim) * np.exp(2 * sig)
        rng = np.random.default_rng(5)
        s = rng.uniform(0.03, 0.97, (nobs, dim))
        od = stat_fem.ObsData(s, 0.7 * np.prod(np.sin(2 * np.pi * s), axis=1) + rng.normal(scale=5e-3, size=nobs), 2.e-3)
        G = stat_fem.ForcingCovariance(V, sig, l, 1.e-3, reg)
        theta = np.array([np.log(0.7), np.log(1.e-2), np.log(0.5)])
        means = []
        for keep in (True, False):
            ls = stat_fem.LinearSolver(A, b, G, od, solver_parameters={"keep_factors": keep, "ksp_rtol": 1e-13})
            ls.set_params(theta)
            x = fem.Function(V)
            ls.solve_posterior(x, scale_mean=True)
            assert (ls._W_dev is not None) == keep
            means.append(x.data.copy())
        assert relerr(means[0], means[1]) < TOL


def test_block_solve_fails_loudly_on_nonfinite_input():
    "a NaN right-hand side must raise, not return 'converged' (the recurrence norm is NaN from the start)"
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import fem, _lib
    from stat_fem_b200.solving_utils import SparseSolver
    V, A, b = fem.poisson_problem(33, 2)
    for pc in ("mg", "jacobi"):
        sp_ = SparseSolver(A, solver_parameters={"pc_type": pc})
        B = sp_.ctx.zeros((V.dim(), 4))
        B[:, 0] = 1.
        B[5, 2] = float("nan")
        X = sp_.ctx.empty((V.dim(), 4))
        with pytest.raises(_lib.NotConvergedError):
            sp_.solve_block(B, X)
        B[5, 2] = 0.
        B[:, 2] = 2.
        sp_.solve_block(B, X)             # and the same solver still works afterwards
        assert np.all(np.isfinite(X.cpu().numpy()))


def _fake_firedrake_objects(c):
    "the golden case's inputs wrapped in the Firedrake stand-ins of oracle/fake_firedrake.py (duck-typed containers)"
    from oracle import fake_firedrake as ff
    V = ff.FakeSpace(c["coords"], c["cells"], c["int_basis"])
    A = sp.csr_matrix((c["A_data"], c["A_indices"], c["A_indptr"]), shape=(len(c["b"]), len(c["b"])))
    return ff, V, ff.FakeMatrix(A, c["bc_nodes"]), ff.Function(V, np.array(c["b"], dtype=np.float64))


@pytest.mark.parametrize("name", ["sq11_example", "sq33_cutoff", "cube9_cutoff"])
def test_adapter_accepts_firedrake_style_objects(name):
    """a WithGeometry-like function space, a MatrixBase-like assembled matrix and Firedrake-like
    Function / Vector objects go through the same API (ForcingCovariance.py:74-79, LinearSolver.py:118-140) and give
    the reference's golden results"""
    from conftest import load_case
    import stat_fem_b200 as stat_fem
    c = load_case(name)
    ff, V, A, b = _fake_firedrake_objects(c)
    G = stat_fem.ForcingCovariance(V, float(c["sigma_f"]), float(c["l_f"]), float(c["cutoff"]), float(c["reg"]))
    x = ff.Function(V, np.ones(V.dim())).vector()
    y = ff.Function(V).vector()
    G.mult(x, y)
    assert relerr(y.gather(), c["G_ones"]) < 1e-13
    od = stat_fem.ObsData(c["sensors"], c["data"], c["unc"])
    ls = stat_fem.LinearSolver(A, b, G, od)
    mu, Cu = ls.solve_prior()
    assert relerr(mu, c["mu"]) < TOL and relerr(Cu, c["Cu"]) < TOL
    ls.set_params(c["params"])
    u = ff.Function(V)
    ls.solve_posterior(u)
    assert relerr(u.vector().gather(), c["post_mean"]) < TOL
    u2 = ff.Function(V)
    stat_fem.solve_posterior(A, u2, b, G, od, c["params"])
    assert relerr(u2.vector().gather(), c["post_mean"]) < TOL
    im = stat_fem.InterpolationMatrix(V, c["sensors"].reshape(len(c["sensors"]), -1))
    v = ff.Function(V, np.arange(V.dim(), dtype=np.float64)).vector()
    assert_allclose(im.interp_mesh_to_data(v), c["P"].T @ v.gather(), atol=1e-10)
    # a bare SciPy matrix (Dirichlet nodes recognised from its identity rows) works as well
    As = sp.csr_matrix((c["A_data"], c["A_indices"], c["A_indptr"]), shape=(V.dim(), V.dim()))
    mu2, Cu2 = stat_fem.LinearSolver(As, b, G, od).solve_prior()
    assert relerr(mu2, c["mu"]) < TOL and relerr(Cu2, c["Cu"]) < TOL
    with pytest.raises(TypeError):
        stat_fem.LinearSolver(np.eye(3), b, G, od)
    with pytest.raises(TypeError):
        stat_fem.ForcingCovariance(object(), 0., 0.)


def test_user_covariance_callable_and_persistence(tmp_path):
    """``cov=`` callables (ForcingCovariance.py:52-53, 161) and saving / restoring G and the prior"""
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import fem
    from stat_fem_b200.covariance_functions import sqexp
    import oracle
    V, A, b = fem.poisson_problem(33, 2)
    h = 1. / 32
    sig, l, reg = np.log(2.e-2), np.log(1.5 * h), 1.e-8 * h ** 4 * np.exp(2 * np.log(2.e-2))

    def my_sqexp(x1, x2, sigma, corr):                 # a different callable with the built-in's values
        return sqexp(x1, x2, sigma, corr)

    def matern32(x1, x2, sigma, corr):
        from scipy.spatial.distance import cdist
        r = cdist(np.atleast_2d(x1), np.atleast_2d(x2)) * np.exp(-corr) * np.sqrt(3.)
        return np.exp(2. * sigma) * (1. + r) * np.exp(-r)

    Gk = stat_fem.ForcingCovariance(V, sig, l, 1.e-3, reg)
    Gk.assemble()
    Gc = stat_fem.ForcingCovariance(V, sig, l, 1.e-3, reg, cov=my_sqexp)
    Gc.assemble()
    a, c_ = Gk.G.to_host(), Gc.G.to_host()
    assert np.array_equal(a[0], c_[0]) and np.array_equal(a[1], c_[1])
    assert_allclose(a[2], c_[2], rtol=1e-13, atol=0.)
    Gm = stat_fem.ForcingCovariance(V, sig, l, 1.e-3, reg, cov=matern32)
    gd, maxnnz = Gm._compute_G_vals()
    ref, refmax = oracle.compute_G_rows(V.coords, V.integrate_basis_functions(), sig, l, 1.e-3, reg, cov=matern32)
    assert maxnnz == refmax
    for i in (0, 17, 500, V.dim() - 1):
        assert np.array_equal(gd[i][1], ref[i][1]) and gd[i][1].dtype == np.int32
        assert_allclose(gd[i][0], ref[i][0], rtol=1e-14, atol=0.)
    # the callable-assembled G drives the solves like any other
    s = np.random.default_rng(3).uniform(0.05, 0.95, (12, 2))
    od = stat_fem.ObsData(s, np.zeros(12), 2.e-3)
    Cu_k = stat_fem.LinearSolver(A, b, Gk, od).solve_prior()[1]
    Cu_c = stat_fem.LinearSolver(A, b, Gc, od).solve_prior()[1]
    assert relerr(Cu_c, Cu_k) < 1e-12
    # persistence
    path = str(tmp_path / "G.npz")
    Gk.save(path)
    Gl = stat_fem.ForcingCovariance.load(V, path)
    assert Gl.is_assembled and all(np.array_equal(p, q) for p, q in zip(Gl.G.to_host(), a))
    ls = stat_fem.LinearSolver(A, b, Gk, od)
    mu, Cu = ls.solve_prior()
    ls.save_prior(str(tmp_path / "prior.npz"))
    ls2 = stat_fem.LinearSolver(A, b, Gl, od)
    mu2, Cu2 = ls2.load_prior(str(tmp_path / "prior.npz"))
    assert np.array_equal(mu, mu2) and np.array_equal(Cu, Cu2)
    th = np.array([np.log(0.7), np.log(1.e-2), np.log(0.5)])
    assert ls2.logposterior(th) == ls.logposterior(th)
    x1, x2 = fem.Function(V), fem.Function(V)
    ls.set_params(th); ls2.set_params(th)
    ls.solve_posterior(x1); ls2.solve_posterior(x2)           # ls2 has no resident W: the two-solve chain
    assert relerr(x2.data, x1.data) < TOL


def test_nonsymmetric_stiffness_is_rejected():
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import fem
    from stat_fem_b200.solving_utils import SparseSolver
    V, A, b = fem.poisson_problem(17, 2)
    bad = fem.Matrix(V, A.indptr, A.indices, A.data.copy(), A.bcs)
    rows = np.repeat(np.arange(V.dim()), np.diff(A.indptr))
    k = np.nonzero((A.indices != rows) & (A.data != 0.))[0][7]
    bad.data[k] *= 1.5                                     # one off-diagonal entry no longer equal to its transpose
    with pytest.raises(ValueError):
        SparseSolver(bad)
    SparseSolver(A)


def test_forcing_covariance_long_rows_fall_back_to_all_pairs():
    """a correlation length that keeps more than 12288 entries per row (the cap of the cell-list sort, csrc/gcov.cu) and a
    diagonal-only 3-D case whose natural cell grid would exceed 2^28 cells: both assemble, rows match the reference formula"""
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import fem
    import oracle
    V, A, b = fem.poisson_problem(113, 2)                  # 12 769 DOF
    sig, l = np.log(2.e-2), np.log(2.)
    G = stat_fem.ForcingCovariance(V, sig, l, 1.e-3, 1.e-12)
    G.assemble()
    assert G.max_row_nnz > 12288
    ip, ix, v = G.G.to_host()
    rows = [0, 56, 6000, V.dim() - 1]
    ref, _ = oracle.compute_G_rows(V.coords, V.integrate_basis_functions(), sig, l, 1.e-3, 1.e-12, rows=rows)
    for i in rows:
        assert np.array_equal(ix[ip[i]:ip[i + 1]], ref[i][1])
        assert_allclose(v[ip[i]:ip[i + 1]], ref[i][0], rtol=1e-13, atol=0.)
    V3, A3, b3 = fem.poisson_problem(9, 3)
    # nugget so large that nothing but the diagonal passes the ratio test: the search radius is empty and the natural
    # "any small cell" grid (2048 per direction) would have 2^33 cells
    G3 = stat_fem.ForcingCovariance(V3, np.log(2.e-2), np.log(0.354), 1.e-3, 1.e-3)
    G3.assemble()
    assert G3.G.nnz == V3.dim()


@pytest.mark.parametrize("nodes,jitter", [(129, 0.25), (200, 0.0), (97, 0.3)])
def test_block_solve_on_jittered_and_non_aligned_meshes(nodes, jitter):
    """The lattice multigrid only sees DOF coordinates and the CSR matrix: an irregular mesh (interior nodes moved by up to
    `jitter` mesh widths, same connectivity, stiffness re-assembled) and a lattice-unaligned size (200^2 nodes: the
    stiffness operator does not get the tensor-core tile form there and runs the scalar tile kernel -- the documented gate
    of csrc/tileop.cu) must still converge in a handful of iterations to the direct solution."""
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import fem
    from stat_fem_b200.solving_utils import SparseSolver
    from oracle.solver import _DirectSolver
    base = fem.UnitSquareMesh(nodes - 1, nodes - 1)
    coords = base.coords.copy()
    rng = np.random.default_rng(8)
    h = 1. / (nodes - 1)
    interior = np.setdiff1d(np.arange(coords.shape[0]), base.boundary_nodes())
    coords[interior] += jitter * h * rng.uniform(-1., 1., (len(interior), 2))
    mesh = fem.Mesh(coords, base.cells)
    V = fem.FunctionSpace(mesh, "CG", 1)
    bc = fem.DirichletBC(V, 0., base.boundary_nodes())
    A = fem.assemble_stiffness(V, bcs=bc)
    sensors = rng.uniform(0.05, 0.95, (48, 2))
    im = stat_fem.InterpolationMatrix(V, sensors)
    im.assemble()
    sp_ = SparseSolver(A)
    Z = sp_.ctx.empty((V.dim(), 48))
    sp_.solve_columns(im.csc, 0, 48, Z)
    assert max(sp_.block_iterations) <= 16, sp_.block_iterations
    assert sp_.last_relres.max() < 1e-11
    lu = _DirectSolver(A.to_scipy(), A.bc_nodes())
    Zh = Z.cpu().numpy()
    P = im.interp.toarray()
    for j in (0, 17, 47):
        ref = lu.solve(P[:, j], bcs_on=False)
        assert relerr(Zh[:, j], ref) < 1e-9
    print("nodes", nodes, "jitter", jitter, "iterations", sp_.block_iterations, "levels", sp_.mg_levels)
