"""Oracle (test infrastructure only): squared-exponential covariance, restating
``covariance_functions.py`` of the reference."""
import numpy as np
from scipy.spatial.distance import cdist


def calc_r2(x1, x2):
    """Squared pairwise distance, ``cdist(...)**2`` -- covariance_functions.py:122-145.

    The square of the *rounded* Euclidean distance (sqrt then square), not the raw sum of squares:
    this matters for bit-exact sparsity decisions downstream."""
    x1 = np.array(x1)
    x2 = np.array(x2)
    return cdist(np.atleast_2d(x1), np.atleast_2d(x2)) ** 2


def sqexp(x1, x2, sigma, l):
    """``exp(2 sigma) * exp(-0.5 * r2 * exp(-2 l))`` -- covariance_functions.py:4-37 (line 37 keeps this
    exact operation order: ``-0.5*r2`` first, then ``*exp(-2l)``)."""
    x1 = np.array(x1)
    x2 = np.array(x2)
    r2 = calc_r2(x1, x2)
    return np.exp(2. * sigma) * np.exp(-0.5 * r2 * np.exp(-2. * l))


def sqexp_deriv(x1, x2, sigma, l):
    """Gradient wrt (sigma, l): ``[2 K, K * r2 * exp(-2 l)]`` -- covariance_functions.py:39-77."""
    x1 = np.array(x1)
    x2 = np.array(x2)
    deriv = np.zeros((2, x1.shape[0], x2.shape[0]))
    deriv[0] = 2. * sqexp(x1, x2, sigma, l)
    deriv[1] = sqexp(x1, x2, sigma, l) * calc_r2(x1, x2) * np.exp(-2. * l)
    return deriv
