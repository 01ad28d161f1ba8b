import sys, os, time, cProfile, pstats
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import stat_fem_b200 as stat_fem
from stat_fem_b200 import fem, _lib
from stat_fem_b200.solving_utils import SparseSolver
V, A, b = fem.poisson_problem(1024, 2)
_lib.DEVICE_CACHE[0] = True
s = SparseSolver(A); del s
torch.cuda.synchronize()
t = time.perf_counter(); s = SparseSolver(A); torch.cuda.synchronize(); print("SparseSolver init (cached uploads): %.3f s" % (time.perf_counter() - t)); del s
pr = cProfile.Profile(); pr.enable(); s = SparseSolver(A); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
