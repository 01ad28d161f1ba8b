"""Functional one-shot wrappers around LinearSolver (API of solving.py in the reference)."""
from .parallel import COMM_SELF
from .LinearSolver import LinearSolver


def solve_posterior(A, x, b, G, data, params, ensemble_comm=COMM_SELF, scale_mean=False, **kwargs):
    "posterior mean on the mesh, stored in ``x`` (solving.py:9-56)"
    ls = LinearSolver(A, b, G, data, ensemble_comm=ensemble_comm, **kwargs)
    ls.set_params(params)
    ls.solve_posterior(x, scale_mean)


def solve_posterior_covariance(A, b, G, data, params, ensemble_comm=COMM_SELF, scale_mean=False, **kwargs):
    "posterior mean and covariance at the sensors (solving.py:58-104)"
    ls = LinearSolver(A, b, G, data, ensemble_comm=ensemble_comm, **kwargs)
    ls.set_params(params)
    return ls.solve_posterior_covariance(scale_mean)


def solve_prior(A, b, G, data, ensemble_comm=COMM_SELF, **kwargs):
    "prior mean and covariance at the sensors (solving.py:106-147)"
    return LinearSolver(A, b, G, data, ensemble_comm=ensemble_comm, **kwargs).solve_prior()


def solve_prior_generating(A, b, G, data, params, ensemble_comm=COMM_SELF, **kwargs):
    "prior of the generating process (solving.py:149-197)"
    ls = LinearSolver(A, b, G, data, ensemble_comm=ensemble_comm, **kwargs)
    ls.set_params(params)
    return ls.solve_prior_generating()


def solve_posterior_generating(A, b, G, data, params, ensemble_comm=COMM_SELF, **kwargs):
    "posterior of the generating process (solving.py:199-246)"
    ls = LinearSolver(A, b, G, data, ensemble_comm=ensemble_comm, **kwargs)
    ls.set_params(params)
    return ls.solve_posterior_generating()


def predict_mean(A, b, G, data, params, coords, ensemble_comm=COMM_SELF, scale_mean=True, **kwargs):
    "predictive mean at new locations (solving.py:248-305)"
    ls = LinearSolver(A, b, G, data, ensemble_comm=ensemble_comm, **kwargs)
    ls.set_params(params)
    return ls.predict_mean(coords, scale_mean)


def predict_covariance(A, b, G, data, params, coords, unc, ensemble_comm=COMM_SELF, **kwargs):
    "predictive covariance at new locations (solving.py:307-365)"
    ls = LinearSolver(A, b, G, data, ensemble_comm=ensemble_comm, **kwargs)
    ls.set_params(params)
    return ls.predict_covariance(coords, unc)
