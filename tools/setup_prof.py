"""cProfile of the host-side setup through the public API with host buffers (the e2e arm's LinearSolver construction)."""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import stat_fem_b200 as stat_fem
from stat_fem_b200 import fem, _lib
V, A, b = fem.poisson_problem(1024, 2)
h = 1. / 1023
sig = np.log(2.e-2); reg = 1.e-8 * h ** 4 * np.exp(2 * sig)
s = np.random.default_rng(1).uniform(0.02, 0.98, (4096, 2))
od = stat_fem.ObsData(s, np.zeros(4096), 2.e-3)
G = stat_fem.ForcingCovariance(V, sig, np.log(1.5 * h), 1.e-3, reg); G.assemble()
for rep in range(2):
    ls = stat_fem.LinearSolver(A, b, G, od); ls.im.assemble(); torch.cuda.synchronize(); del ls
pr = cProfile.Profile(); t = time.perf_counter(); pr.enable()
ls = stat_fem.LinearSolver(A, b, G, od); ls.im.assemble(); torch.cuda.synchronize()
pr.disable(); print("LinearSolver + P assemble, host buffers: %.3f s" % (time.perf_counter() - t))
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
