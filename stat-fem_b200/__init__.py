"""stat-fem_b200: B200-native implementation of stat-fem's data-conditioning hot path behind stat-fem's
own Python API.  ``import stat_fem_b200 as stat_fem``."""
from . import fem
