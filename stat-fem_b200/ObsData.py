"""Sensor observations: locations, values and measurement errors (API of ObsData.py in the reference).  The container
lives on the host; the model-discrepancy matrices are produced by the CUDA kernel-matrix kernel."""
import numpy as np
from . import _lib


class ObsData(object):
    def __init__(self, coords, data, unc):
        coords = np.array(coords, dtype=np.float64)
        if coords.ndim == 1:
            coords = coords.reshape(-1, 1)
        assert coords.ndim == 2, "coords must be a 1D or 2D array"
        self.n_obs, self.n_dim = coords.shape
        self.coords = coords.copy()
        data = np.array(data, dtype=np.float64)
        assert data.shape == (self.n_obs,), "data must be a 1D array with the same length as coords"
        self.data = data.copy()
        unc = np.nan_to_num(np.array(unc, dtype=np.float64))
        assert unc.shape == (self.n_obs,) or unc.shape == (), "bad shape for unc, must be an array or float"
        assert np.all(unc >= 0.), "all uncertainties must be non-negative"
        self.unc = unc.copy()
        self._dev = None

    def get_n_dim(self):
        return self.n_dim

    def get_n_obs(self):
        return self.n_obs

    def get_coords(self):
        return self.coords

    def get_data(self):
        return self.data

    def get_unc(self):
        return self.unc

    # ---- device mirrors (coords, data, unc) used by LinearSolver ------------------------------------------
    def device(self):
        if self._dev is None:
            ctx = _lib.get_context()
            self._dev = dict(ctx=ctx, coords=ctx.to_device(self.coords, np.float64),
                             data=ctx.to_device(self.data, np.float64),
                             unc=ctx.to_device(np.atleast_1d(self.unc), np.float64),
                             unc_vec=int(self.unc.shape != ()))
        return self._dev

    def _matrix(self, params, which, with_unc):
        params = np.array(params)
        assert params.shape == (2,), "parameters must have length 2"
        d = self.device()
        ctx = d["ctx"]
        out = ctx.empty((self.n_obs, self.n_obs))
        ctx.check(_lib.lib().sfem_kernel_matrix(ctx.handle, self.n_obs, self.n_dim, _lib.ptr(d["coords"]),
                                                float(params[0]), float(params[1]),
                                                _lib.ptr(d["unc"] if with_unc else None), d["unc_vec"], 0.,
                                                _lib.ptr(None), which, _lib.ptr(out)))
        return out.cpu().numpy()

    def calc_K(self, params):
        "model discrepancy covariance sqexp(coords, coords, params) (ObsData.py:131-157)"
        return self._matrix(params, 0, False)

    def calc_K_plus_sigma(self, params):
        "K + diag(unc^2) (ObsData.py:159-185)"
        return self._matrix(params, 0, True)

    def calc_K_deriv(self, params):
        "derivatives of K wrt (sigma, l), shape (2, n_obs, n_obs) (ObsData.py:187-214)"
        return np.stack([self._matrix(params, 1, False), self._matrix(params, 2, False)], axis=0)

    def __str__(self):
        return ("Observational Data:\nNumber of dimensions:\n{}\nNumber of observations:\n{}\nCoordinates:\n{}\n"
                "Data:\n{}\nUncertainty:\n{}".format(self.n_dim, self.n_obs, self.coords, self.data, self.unc))
