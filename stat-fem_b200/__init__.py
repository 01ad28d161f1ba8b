"""stat-fem_b200: B200-native implementation of stat-fem's data-conditioning hot path behind stat-fem's own Python API
(see DESIGN.md).  Usage: ``import stat_fem_b200 as stat_fem``.

Importing the package does not touch the GPU; the first call that needs arithmetic loads libstatfem_b200.so and
creates a context on the current CUDA device, and raises if either is missing (there is no CPU fallback)."""
__version__ = "0.1.0"

from . import fem
from .parallel import EnsembleComm, COMM_SELF, comm_world
from .ForcingCovariance import ForcingCovariance
from .InterpolationMatrix import InterpolationMatrix, interpolate_cell
from .ObsData import ObsData
from .LinearSolver import LinearSolver, solve_posterior_snapshots
from .solving import solve_posterior, solve_posterior_covariance, solve_prior
from .solving import solve_prior_generating, solve_posterior_generating
from .solving import predict_mean, predict_covariance
from .estimation import estimate_params_MAP, fit_snapshots
from .assemble import assemble
from . import covariance_functions
