"""Oracle (test infrastructure only): observation container + discrepancy kernel matrices,
restating ``ObsData.py`` of the reference."""
import numpy as np
from .covariance import sqexp, sqexp_deriv


class OracleObsData(object):
    """ObsData.py:26-66 (constructor checks), 131-214 (calc_K, calc_K_plus_sigma, calc_K_deriv)."""

    def __init__(self, coords, data, unc):
        coords = np.array(coords, dtype=np.float64)
        if coords.ndim == 1:
            coords = np.reshape(coords, (-1, 1))
        assert coords.ndim == 2
        self.n_obs, self.n_dim = coords.shape
        self.coords = np.copy(coords)
        data = np.array(data, dtype=np.float64)
        assert data.shape == (self.n_obs,)
        self.data = np.copy(data)
        unc = np.nan_to_num(np.array(unc, dtype=np.float64))
        assert unc.shape == (self.n_obs,) or unc.shape == ()
        assert np.all(unc >= 0.)
        self.unc = np.copy(unc)

    def get_n_dim(self): return self.n_dim
    def get_n_obs(self): return self.n_obs
    def get_coords(self): return self.coords
    def get_data(self): return self.data
    def get_unc(self): return self.unc

    def calc_K(self, params):
        "ObsData.py:131-157"
        params = np.array(params)
        assert params.shape == (2,)
        return sqexp(self.coords, self.coords, params[0], params[1])

    def calc_K_plus_sigma(self, params):
        "ObsData.py:159-185: scalar unc -> eye*unc**2, vector unc -> diag(unc**2)"
        if self.unc.shape == ():
            sigma_dat = np.eye(self.n_obs) * self.unc ** 2
        else:
            sigma_dat = np.diag(self.unc ** 2)
        return self.calc_K(params) + sigma_dat

    def calc_K_deriv(self, params):
        "ObsData.py:187-214"
        params = np.array(params)
        assert params.shape == (2,)
        return sqexp_deriv(self.coords, self.coords, params[0], params[1])
