"""Oracle (test infrastructure only): the conditioning solves, restating ``solving_utils.py``,
``LinearSolver.py`` and ``estimation.py`` of the reference with SciPy (splu + cho_factor/cho_solve)."""
import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import splu
from scipy.linalg import cho_factor, cho_solve, LinAlgError
from scipy.optimize import minimize


class _DirectSolver(object):
    """Stand-in for Firedrake's ``LinearSolver(A)`` (LinearSolver.py:136-140): direct LU of the assembled
    matrix whose Dirichlet rows *and columns* are identity (tests/helper_funcs.py:104-128).  ``bcs_on``
    mimics Firedrake's RHS lifting for homogeneous Dirichlet conditions (``b[bc] = 0``); solving_utils.py:46-57
    turns that off for the covariance solves."""

    def __init__(self, A, bc_nodes):
        self.A = sp.csc_matrix(A)
        self.bc_nodes = np.asarray(bc_nodes, dtype=np.int64)
        self.lu = splu(self.A)

    def solve(self, rhs, bcs_on):
        rhs = np.array(rhs, dtype=np.float64, copy=True)
        if bcs_on and self.bc_nodes.size:
            rhs[self.bc_nodes] = 0.
        return self.lu.solve(rhs)


def solve_forcing_covariance(G, ls, rhs):
    """``A^-1 G A^-1 rhs`` with BC application off -- solving_utils.py:10-59 (solves at 51, 53; MatMult at 52)."""
    x = ls.solve(rhs, bcs_on=False)
    rhs_working = G @ x
    x = ls.solve(rhs_working, bcs_on=False)
    return x


def interp_covariance_to_data(P_left, G, ls, P_right, col_range=None):
    """solving_utils.py:61-169.  Row ``i`` of the temporary array is ``P_left^T (A^-1 G A^-1 p_i)``
    (lines 139-142) and the flattened array is *reshaped* to ``(n_left, n_right)`` (164-169), so
    ``result[i, j] = (G z_i) . z_j`` -- i.e. ``Cu = (G Z)^T Z`` (SURVEY finding 1).  ``col_range`` = the
    ensemble-rank slice ``[imin, imax)`` of lines 117-123; ``None`` = all columns (serial)."""
    P_left = sp.csc_matrix(P_left)
    P_right = sp.csc_matrix(P_right)
    n_left, n_right = P_left.shape[1], P_right.shape[1]
    imin, imax = (0, n_right) if col_range is None else col_range
    result_tmparray = np.zeros((imax - imin, n_left))
    PlT = P_left.T.tocsr()
    for i in range(imin, imax):
        rhs = np.asarray(P_right[:, i].todense()).ravel()          # get_meshspace_column_vector (IM.py:261-287)
        tmp = solve_forcing_covariance(G, ls, rhs)
        result_tmparray[i - imin] = PlT @ tmp                      # interp_mesh_to_data (IM.py:223-259)
    if col_range is None:
        return np.reshape(result_tmparray.flatten(), (n_left, n_right))
    return result_tmparray


class OracleLinearSolver(object):
    """LinearSolver.py:14-691 on NumPy/SciPy inputs.

    ``A``: scipy sparse, Dirichlet rows/cols = identity.  ``b``: load vector as assembled (BC entries
    untouched, like Firedrake's ``assemble(L)``).  ``G``: scipy CSR.  ``P``: scipy sparse (n x n_obs).
    ``data``: OracleObsData."""

    def __init__(self, A, b, G, data, P, bc_nodes):
        self.solver = _DirectSolver(A, bc_nodes)
        self.b = np.asarray(b, dtype=np.float64)
        self.G = sp.csr_matrix(G)
        self.data = data
        self.P = sp.csc_matrix(P)
        self.PT = self.P.T.tocsr()
        self.params = None
        self.x = None
        self.mu = None
        self.Cu = None

    def set_params(self, params):
        "LinearSolver.py:165-183"
        params = np.array(params, dtype=np.float64)
        assert params.shape == (3,), "bad shape for model discrepancy parameters"
        self.params = params

    def solve_prior(self):
        "LinearSolver.py:185-222"
        self.Cu = interp_covariance_to_data(self.P, self.G, self.solver, self.P)
        self.x = self.solver.solve(self.b, bcs_on=True)            # line 217 (BCs on)
        self.mu = self.PT @ self.x                                 # line 218
        return self.mu, self.Cu

    def solve_posterior(self, scale_mean=False):
        "LinearSolver.py:224-305; returns the mesh-space vector the reference assigns into ``x``"
        if self.Cu is None or self.x is None:
            self.solve_prior()
        if self.params is None:
            raise ValueError("must set parameter values to solve posterior")
        rho = np.exp(self.params[0])
        scalefact = rho if scale_mean else 1.
        Ks = self.data.calc_K_plus_sigma(self.params[1:])
        LK = cho_factor(Ks)
        tmp_dataspace_1 = cho_solve(LK, self.data.get_data())
        tmp_meshspace_1 = self.P @ tmp_dataspace_1
        tmp_meshspace_2 = solve_forcing_covariance(self.G, self.solver, tmp_meshspace_1) * rho + self.x
        tmp_dataspace_1 = self.PT @ tmp_meshspace_2
        L = cho_factor(Ks + rho ** 2 * self.Cu)
        tmp_dataspace_2 = cho_solve(L, tmp_dataspace_1)
        tmp_meshspace_1 = self.P @ tmp_dataspace_2
        tmp_meshspace_1 = solve_forcing_covariance(self.G, self.solver, tmp_meshspace_1) * rho ** 2
        return (tmp_meshspace_2 - tmp_meshspace_1) * scalefact

    def solve_posterior_covariance(self, scale_mean=False):
        "LinearSolver.py:308-373"
        if self.mu is None or self.Cu is None:
            self.solve_prior()
        if self.params is None:
            raise ValueError("must set parameter values to solve posterior")
        rho = np.exp(self.params[0])
        scalefact = rho if scale_mean else 1.
        Ks = self.data.calc_K_plus_sigma(self.params[1:])
        LK = cho_factor(Ks)
        LC = cho_factor(Ks + rho ** 2 * self.Cu)
        muy = rho * np.dot(self.Cu, cho_solve(LK, self.data.get_data())) + self.mu
        muy_tmp = rho ** 2 * np.dot(self.Cu, cho_solve(LC, muy))
        muy = muy - muy_tmp
        Cuy = self.Cu - rho ** 2 * np.dot(self.Cu, cho_solve(LC, self.Cu))
        return scalefact * muy, Cuy

    def solve_prior_generating(self):
        "LinearSolver.py:375-406"
        if self.mu is None or self.Cu is None:
            self.solve_prior()
        if self.params is None:
            raise ValueError("must set parameter values to solve prior of generating process")
        rho = np.exp(self.params[0])
        return rho * self.mu, rho ** 2 * self.Cu + self.data.calc_K(self.params[1:])

    def solve_posterior_generating(self):
        "LinearSolver.py:408-439"
        m_eta, C_eta = self.solve_prior_generating()
        unc = self.data.get_unc()
        L = cho_factor(unc ** 2 * np.eye(self.data.get_n_obs()) + C_eta)
        C_etay = cho_solve(L, unc ** 2 * C_eta)
        m_etay = cho_solve(L, np.dot(C_eta, self.data.get_data()) + unc ** 2 * m_eta)
        return m_etay, C_etay

    def predict_mean(self, P_new, scale_mean=True):
        "LinearSolver.py:441-498 (P_new = interpolation matrix of the new coordinates)"
        if self.params is None:
            raise ValueError("must set parameter values to make predictions")
        rho = np.exp(self.params[0])
        scalefact = rho if scale_mean else 1.
        x = self.solve_posterior(False)
        return scalefact * (sp.csc_matrix(P_new).T @ x)

    def predict_covariance(self, P_new, data_new):
        """LinearSolver.py:500-567, including the reference's reshape of the rectangular block
        (solving_utils.py:137, 164-169) -- for n_new != n_obs the reference scrambles Cucd; here the
        square-consistent orientation ``Cucd[i, j] = (G z_new_i) . z_obs_j`` is used only when shapes agree,
        otherwise the reference's reshape is reproduced literally."""
        if self.Cu is None:
            self.solve_prior()
        if self.params is None:
            raise ValueError("must set parameter values to make predictions")
        rho = np.exp(self.params[0])
        P_new = sp.csc_matrix(P_new)
        if P_new.shape[1] > self.data.get_n_obs():
            Cucd = interp_covariance_to_data(P_new, self.G, self.solver, self.P)
        else:
            Cucd = interp_covariance_to_data(self.P, self.G, self.solver, P_new).T
        Cucc = interp_covariance_to_data(P_new, self.G, self.solver, P_new)
        Ks = self.data.calc_K_plus_sigma(self.params[1:])
        LC = cho_factor(Ks + rho ** 2 * self.Cu)
        Cuy = Cucc - rho ** 2 * np.dot(Cucd, cho_solve(LC, Cucd.T))
        return data_new.calc_K_plus_sigma(self.params[1:]) + rho ** 2 * Cuy

    def logposterior(self, params):
        "LinearSolver.py:569-623 (negative log marginal likelihood; priors are forced None at 130-132)"
        self.set_params(params)
        rho = np.exp(self.params[0])
        if self.Cu is None or self.mu is None:
            self.solve_prior()
        KCu = rho ** 2 * self.Cu + self.data.calc_K_plus_sigma(self.params[1:])
        try:
            L = cho_factor(KCu)
        except LinAlgError:
            raise LinAlgError("Error attempting to factorize the covariance matrix in model_loglikelihood")
        resid = self.data.get_data() - rho * self.mu
        invKCudata = cho_solve(L, resid)
        return 0.5 * (self.data.get_n_obs() * np.log(2. * np.pi) +
                      2. * np.sum(np.log(np.diag(L[0]))) +
                      np.dot(resid, invKCudata))

    def logpost_deriv(self, params):
        "LinearSolver.py:625-691"
        self.set_params(params)
        rho = np.exp(self.params[0])
        if self.Cu is None or self.mu is None:
            self.solve_prior()
        KCu = rho ** 2 * self.Cu + self.data.calc_K_plus_sigma(params[1:])
        L = cho_factor(KCu)
        invKCudata = cho_solve(L, self.data.get_data() - rho * self.mu)
        K_deriv = self.data.calc_K_deriv(self.params[1:])
        deriv = np.zeros(3)
        deriv[0] = (-rho * np.dot(self.mu, invKCudata) -
                    rho ** 2 * np.linalg.multi_dot([invKCudata, self.Cu, invKCudata]) +
                    rho ** 2 * np.trace(cho_solve(L, self.Cu)))
        for i in range(0, 2):
            deriv[i + 1] = -0.5 * (np.linalg.multi_dot([invKCudata, K_deriv[i], invKCudata]) -
                                   np.trace(cho_solve(L, K_deriv[i])))
        return deriv


def estimate_params_MAP(ls, start=None, **minimize_kwargs):
    """estimation.py:78-110 on an already constructed OracleLinearSolver: solve_prior, random start
    ``5*(U[0,1)^3 - 0.5)`` from NumPy's global state (83-90), L-BFGS-B with the analytic gradient (92-93)."""
    ls.solve_prior()
    if start is None:
        start = 5. * (np.random.random(3) - 0.5)
    else:
        assert np.array(start).shape == (3,), "bad shape for starting point"
    fmin_dict = minimize(ls.logposterior, start, method='L-BFGS-B', jac=ls.logpost_deriv,
                         options=minimize_kwargs)
    assert fmin_dict['success'], "minimization routine failed"
    ls.set_params(fmin_dict['x'])
    return ls, fmin_dict
