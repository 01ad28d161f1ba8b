"""CPU tests: the oracle (oracle/) against (i) outputs of the unmodified reference modules run on the
Firedrake stand-in (tests/golden/*.npz, made by tests/golden/make_golden.py) and (ii) the dense 1-D goldens
the reference's own tests hold (stat_fem/tests/helper_funcs.py)."""
import numpy as np
import scipy.sparse as sp
from numpy.testing import assert_allclose, assert_array_equal
import pytest

import oracle
from conftest import load_case, relerr


def make_oracle(c):
    n = c["coords"].shape[0]
    A = sp.csr_matrix((c["A_data"], c["A_indices"], c["A_indptr"]), shape=(n, n))
    G = sp.csr_matrix((c["G_vals"], c["G_indices"], c["G_indptr"]), shape=(n, n))
    od = oracle.OracleObsData(c["sensors"], c["data"], c["unc"])
    P = oracle.build_interpolation_matrix(c["coords"], c["cells"], c["sensors"])
    return oracle.OracleLinearSolver(A, c["b"], G, od, P, c["bc_nodes"]), P


def test_covariance_functions_match_reference():
    z = load_case("covariance")
    for d in (1, 2, 3):
        x1, x2 = z[f"x1_{d}"], z[f"x2_{d}"]
        assert_array_equal(oracle.calc_r2(x1, x2), z[f"r2_{d}"])
        assert_array_equal(oracle.sqexp(x1, x2, np.log(0.7), np.log(0.3)), z[f"sqexp_{d}"])
        assert_array_equal(oracle.sqexp_deriv(x1, x2, np.log(0.7), np.log(0.3)), z[f"deriv_{d}"])
        od = oracle.OracleObsData(x1, np.zeros(9), z[f"unc_{d}"])
        assert_array_equal(od.calc_K(np.array([-1.1, -0.4])), z[f"K_{d}"])
        assert_array_equal(od.calc_K_plus_sigma(np.array([-1.1, -0.4])), z[f"Ks_{d}"])
        assert_array_equal(od.calc_K_deriv(np.array([-1.1, -0.4])), z[f"Kd_{d}"])
        assert_allclose(oracle.interpolate_cell(z[f"ic_pt_{d}"], z[f"ic_nodes_{d}"]), z[f"ic_w_{d}"], atol=1e-14)


def test_covariance_closed_forms():
    "reference tests/test_covariance_functions.py: closed forms at atol 1e-10, FD derivative at rtol 1e-5"
    x = np.array([[1.], [2.], [4.]])
    s, l = np.log(2.), np.log(1.5)
    r = np.abs(x - x.T)
    assert_allclose(oracle.sqexp(x, x, s, l), 4. * np.exp(-0.5 * r ** 2 / 1.5 ** 2), atol=1e-10)
    d = oracle.sqexp_deriv(x, x, s, l)
    dx = 1e-6
    assert_allclose(d[0], (oracle.sqexp(x, x, s + dx, l) - oracle.sqexp(x, x, s - dx, l)) / (2 * dx), rtol=1e-5, atol=1e-8)
    assert_allclose(d[1], (oracle.sqexp(x, x, s, l + dx) - oracle.sqexp(x, x, s, l - dx)) / (2 * dx), rtol=1e-5, atol=1e-8)
    assert_allclose(oracle.calc_r2(np.array([1., 2.]), np.array([[1., 2.], [2., 4.]])), [[0., 5.]], atol=1e-12)


def test_interpolate_cell_known_answers():
    "reference tests/test_InterpolationMatrix.py:255-291"
    assert_allclose(oracle.interpolate_cell(np.array([0.5]), np.array([[0.], [1.]])), [0.5, 0.5])
    assert_allclose(oracle.interpolate_cell(np.array([0.5, 0.5]), np.array([[0., 0.], [1., 0.], [0., 1.]])),
                    [0., 0.5, 0.5], atol=1e-10)
    assert_allclose(oracle.interpolate_cell(np.array([1 / 3.] * 3),
                                            np.array([[0., 0., 0.], [1., 0., 0.], [0., 1., 0.], [0., 0., 1.]])),
                    [0., 1 / 3., 1 / 3., 1 / 3.], atol=1e-9)
    with pytest.raises(AssertionError):
        oracle.interpolate_cell(np.array([0.5, 0.5]), np.eye(4)[:, :3])


def test_G_pattern_and_values_bit_exact(case):
    c = case
    ip, ix, v = oracle.compute_G_csr(c["coords"], c["int_basis"], float(c["sigma_f"]), float(c["l_f"]),
                                     float(c["cutoff"]), float(c["reg"]))
    assert_array_equal(ip, c["G_indptr"])
    assert_array_equal(ix, c["G_indices"])
    assert ix.dtype == np.int32
    assert_array_equal(v, c["G_vals"])
    assert int(np.max(np.diff(ip))) == int(c["G_max_nnz"])
    ip2, ix2, v2 = oracle.compute_G_csr_fast(c["coords"], c["int_basis"], float(c["sigma_f"]), float(c["l_f"]),
                                            float(c["cutoff"]), float(c["reg"]))
    assert_array_equal(ip2, c["G_indptr"])
    assert_array_equal(ix2, c["G_indices"])
    assert_array_equal(v2, c["G_vals"])


def test_interpolation_matrix(case):
    _, P = make_oracle(case)
    assert_allclose(P.toarray(), case["P"], atol=1e-12)


def test_prior_posterior_logpost(case):
    c = case
    ls, P = make_oracle(c)
    mu, Cu = ls.solve_prior()
    assert relerr(mu, c["mu"]) < 1e-11
    assert relerr(Cu, c["Cu"]) < 1e-11
    assert relerr(ls.x, c["x_prior"]) < 1e-11
    n = c["coords"].shape[0]
    assert relerr(oracle.solve_forcing_covariance(ls.G, ls.solver, np.ones(n)), c["AGA_ones"]) < 1e-11
    assert relerr(ls.G @ np.ones(n), c["G_ones"]) < 1e-14
    ls.set_params(c["params"])
    assert relerr(ls.solve_posterior(), c["post_mean"]) < 1e-10
    assert relerr(ls.solve_posterior(True), c["post_mean_scaled"]) < 1e-10
    muy, Cuy = ls.solve_posterior_covariance()
    assert relerr(muy, c["muy"]) < 1e-9
    assert relerr(Cuy, c["Cuy"]) < 1e-9
    assert relerr(ls.solve_posterior_covariance(True)[0], c["muy_scaled"]) < 1e-9
    m_eta, C_eta = ls.solve_prior_generating()
    assert relerr(m_eta, c["m_eta"]) < 1e-11 and relerr(C_eta, c["C_eta"]) < 1e-11
    if "m_etay" in c:
        m_etay, C_etay = ls.solve_posterior_generating()
        assert relerr(m_etay, c["m_etay"]) < 1e-9 and relerr(C_etay, c["C_etay"]) < 1e-9
    assert_allclose(ls.logposterior(c["params"]), c["logpost"], rtol=1e-10)
    assert_allclose(ls.logpost_deriv(c["params"]), c["logpost_deriv"], rtol=1e-8, atol=1e-9)
    for p, v, g in zip(c["lp_points"], c["lp_values"], c["lp_derivs"]):
        assert_allclose(ls.logposterior(p), v, rtol=1e-10)
        assert_allclose(ls.logpost_deriv(p), g, rtol=1e-7, atol=1e-8)
    if "predict_coords" in c:
        ls.set_params(c["params"])
        Pn = oracle.build_interpolation_matrix(c["coords"], c["cells"], c["predict_coords"])
        assert relerr(ls.predict_mean(Pn), c["predict_mean"]) < 1e-10
        assert relerr(ls.predict_mean(Pn, scale_mean=False), c["predict_mean_unscaled"]) < 1e-10
        dn = oracle.OracleObsData(c["predict_coords"], np.zeros(len(c["predict_coords"])), 0.1)
        assert relerr(ls.predict_covariance(Pn, dn), c["predict_cov"]) < 1e-9


def test_map_matches_reference(case):
    c = case
    if "map_params" not in c:
        pytest.skip("no MAP golden for this case")
    ls, _ = make_oracle(c)
    kw = {}
    if np.isfinite(c["map_kwargs_ftol"]):
        kw = dict(ftol=float(c["map_kwargs_ftol"]), gtol=float(c["map_kwargs_gtol"]))
    ls, res = oracle.estimate_params_MAP(ls, start=c["map_start"], **kw)
    assert_allclose(ls.logposterior(ls.params), c["map_logpost"], rtol=1e-8)
    # the optimum is flat in some directions; parameters agree to 1e-6 only when tight
    assert_allclose(ls.params, c["map_params"], rtol=1e-4, atol=1e-4)


def test_reference_1d_dense_goldens():
    """The dense goldens the reference's tests hold for its 1-D fixture (tests/helper_funcs.py:60-239;
    test_solving_utils.py:16-33, 74-99; test_LinearSolver.py:105-136, 166-194, 342-383; atol 1e-10)."""
    c = load_case("ref1d")
    nx = 10
    A_numpy = np.diag(np.full(nx + 1, 20.)) + np.diag(np.full(nx, -10.), 1) + np.diag(np.full(nx, -10.), -1)
    A_numpy[0, :] = 0; A_numpy[-1, :] = 0; A_numpy[:, 0] = 0; A_numpy[:, -1] = 0
    A_numpy[0, 0] = 1; A_numpy[-1, -1] = 1
    x = np.linspace(0., 1., nx + 1)
    basis = np.array([0.05] + [0.1] * 9 + [0.05])
    r = np.abs(x[:, None] - x[None, :])
    cov = np.outer(basis, basis) * np.exp(0.) ** 2 * np.exp(-0.5 * r ** 2 / 0.1 ** 2) + np.eye(nx + 1) * 1e-8
    interp = np.zeros((nx + 1, 4))
    interp[7, 0] = interp[8, 0] = 0.5
    interp[5, 1] = 1.
    interp[2, 2] = interp[3, 2] = 0.5
    interp[1, 3], interp[2, 3] = 0.75, 0.25
    sensors = np.array([[0.75], [0.5], [0.25], [0.125]])
    params = np.array([-0.9, 0.1, -0.5])
    rs = np.abs(sensors - sensors.T)
    Ks = np.exp(params[1]) ** 2 * np.exp(-0.5 * rs ** 2 / np.exp(params[2]) ** 2) + np.eye(4) * 0.1 ** 2

    ls, P = make_oracle(c)
    assert_allclose(ls.solver.A.toarray(), A_numpy, atol=1e-12)
    assert_allclose(c["int_basis"], basis, atol=1e-14)                  # test_ForcingCovariance.py:70-81
    assert_allclose(ls.G.toarray(), cov, atol=1e-8, rtol=1e-6)          # test_ForcingCovariance.py:83-110
    assert_allclose(ls.G @ np.ones(nx + 1), cov @ np.ones(nx + 1), atol=1e-8, rtol=1e-6)
    assert_allclose(P.toarray(), interp, atol=1e-10)                    # test_InterpolationMatrix.py:69-88
    exp = np.linalg.solve(A_numpy, cov @ np.linalg.solve(A_numpy, np.ones(nx + 1)))
    assert_allclose(oracle.solve_forcing_covariance(ls.G, ls.solver, np.ones(nx + 1)), exp, atol=1e-10)
    C_expected = interp.T @ np.linalg.solve(A_numpy, cov @ np.linalg.solve(A_numpy, interp))
    mu, Cu = ls.solve_prior()
    assert_allclose(Cu, C_expected, atol=1e-10)
    b0 = c["b"].copy(); b0[[0, -1]] = 0.
    assert_allclose(mu, interp.T @ np.linalg.solve(A_numpy, b0), atol=1e-10)
    rho = np.exp(params[0])
    ls.set_params(params)
    muy, Cuy = ls.solve_posterior_covariance()
    Kinv = np.linalg.inv(Ks)
    Cuy_expected = np.linalg.inv(np.linalg.inv(Cu) + rho ** 2 * Kinv)
    muy_expected = Cuy_expected @ (rho * Kinv @ c["data"] + np.linalg.solve(Cu, mu))
    assert_allclose(muy, muy_expected, atol=1e-10)
    assert_allclose(Cuy, Cuy_expected, atol=1e-10)
    KCu = rho ** 2 * Cu + Ks
    resid = c["data"] - rho * mu
    ll = 0.5 * (resid @ np.linalg.inv(KCu) @ resid + np.log(np.linalg.det(KCu)) + 4 * np.log(2. * np.pi))
    assert_allclose(ls.logposterior(params), ll)
    fd = np.array([(ls.logposterior(params + 1e-8 * np.eye(3)[k]) - ls.logposterior(params)) / 1e-8
                   for k in range(3)])
    assert_allclose(ls.logpost_deriv(params), fd, atol=1e-6, rtol=1e-6)
    # posterior mean on the mesh (test_LinearSolver.py:24-63) via the full-mesh Cu
    Pfull = sp.identity(nx + 1, format="csc")
    Cu_full = oracle.interp_covariance_to_data(Pfull, ls.G, ls.solver, Pfull)
    mu_full = ls.x
    tmp_1 = rho * Cu_full @ interp @ np.linalg.solve(Ks, c["data"]) + mu_full
    tmp_2 = Cu_full @ interp @ np.linalg.inv(Ks / rho ** 2 + Cu) @ interp.T @ tmp_1
    assert_allclose(ls.solve_posterior(), tmp_1 - tmp_2, atol=1e-10)
