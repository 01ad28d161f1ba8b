"""Squared-exponential covariance functions of stat-fem's API, evaluated by the CUDA kernel-matrix kernel
(csrc/dense_ops.cu).  Signatures and conventions follow covariance_functions.py of the reference: parameters are
natural logs, 1-D inputs are promoted to a single point."""
import numpy as np
from . import _lib


def _points(x):
    x = np.array(x, dtype=np.float64)
    return np.atleast_2d(x)


def _cov_block(x1, x2, sigma, l, which):
    a, b = _points(x1), _points(x2)
    assert a.shape[1] == b.shape[1], "x1 and x2 must have the same number of spatial dimensions"
    ctx = _lib.get_context()
    da, db = ctx.to_device(a, np.float64), ctx.to_device(b, np.float64)
    out = ctx.empty((a.shape[0], b.shape[0]))
    ctx.check(_lib.lib().sfem_cov_matrix(ctx.handle, a.shape[0], b.shape[0], a.shape[1], _lib.ptr(da), _lib.ptr(db),
                                         float(sigma), float(l), which, _lib.ptr(out)))
    return out.cpu().numpy()


def calc_r2(x1, x2):
    "squared distance between all pairs of points (covariance_functions.py:122-145)"
    return _cov_block(x1, x2, 0., 0., 3)


def sqexp(x1, x2, sigma, l):
    "exp(2 sigma) exp(-r^2 exp(-2 l) / 2) for all pairs (covariance_functions.py:4-37)"
    return _cov_block(x1, x2, sigma, l, 0)


def sqexp_deriv(x1, x2, sigma, l):
    "gradient wrt (sigma, l), shape (2, m, n) (covariance_functions.py:39-77)"
    return np.stack([_cov_block(x1, x2, sigma, l, 1), _cov_block(x1, x2, sigma, l, 2)], axis=0)


def sqexp_hessian(x1, x2, sigma, l):
    "second derivatives wrt (sigma, l), shape (2, 2, m, n) (covariance_functions.py:79-120; unused by the library)"
    K = _cov_block(x1, x2, sigma, l, 0)
    q = _cov_block(x1, x2, 0., 0., 3) * np.exp(-2. * l)
    hess = np.zeros((2, 2) + K.shape)
    hess[0, 0] = 4. * K
    hess[0, 1] = hess[1, 0] = 2. * K * q
    hess[1, 1] = K * (q * q - 2. * q)
    return hess
