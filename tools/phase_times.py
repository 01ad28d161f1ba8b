"""Phase timing of the time-to-posterior pipeline through the public API (host buffers), for orientation."""
import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import stat_fem_b200 as stat_fem
from stat_fem_b200 import fem

nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 513
nobs = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 2
t0 = time.time()
V, A, b = fem.poisson_problem(nodes, dim)
h = 1. / (nodes - 1)
sig = np.log(2.e-2); lf = np.log((1.5 if dim == 2 else 0.8) * h)
reg = 1.e-8 * h ** (2 * dim) * np.exp(2 * sig)
rng = np.random.default_rng(1)
s = rng.uniform(0.02, 0.98, (nobs, dim))
y = 0.7 * np.prod(np.sin(2 * np.pi * s), axis=1) + rng.normal(scale=0.01, size=nobs)
od = stat_fem.ObsData(s, y, 2.e-3)
print("host setup %.2fs n=%d" % (time.time() - t0, V.dim()), flush=True)

def timed(name, fn):
    torch.cuda.synchronize(); t = time.time(); r = fn(); torch.cuda.synchronize()
    print("%-28s %8.3f s" % (name, time.time() - t), flush=True); return r

for rep in range(2):
    print("--- rep", rep)
    G = stat_fem.ForcingCovariance(V, sig, lf, 1.e-3, reg)
    timed("G assemble", G.assemble)
    print("   nnz(G) = %d (%.1f/row)" % (G.G.nnz, G.G.nnz / V.dim()))
    ls = timed("LinearSolver init (+MG setup)", lambda: stat_fem.LinearSolver(A, b, G, od))
    print("   levels", ls.solver.mg_levels)
    timed("P assemble", ls.im.assemble)
    ctx = ls.solver.ctx
    n = V.dim()
    Z = ctx.empty((n, nobs))
    timed("Z = A^-1 P", lambda: ls.solver.solve_columns(ls.im.csc, 0, nobs, Z))
    print("   iterations per block", ls.solver.block_iterations, "block", ls.solver.pick_block_size(nobs))
    W = ctx.empty((n, nobs))
    timed("W = G Z", lambda: G.G.matmul(Z, out=W))
    Cu = ctx.empty((nobs, nobs))
    from stat_fem_b200 import _lib
    timed("Cu = W^T Z", lambda: ctx.check(_lib.lib().sfem_dgemm(ctx.handle, 1, 0, nobs, nobs, n, 1.0, _lib.ptr(W), nobs, _lib.ptr(Z), nobs, 0.0, _lib.ptr(Cu), nobs)))
    del Z, W, Cu
    timed("solve_prior (all)", ls.solve_prior)
    p0 = np.array([np.log(0.7), np.log(1e-2), np.log(0.5)])
    timed("logposterior", lambda: ls.logposterior(p0))
    timed("logpost_deriv", lambda: ls.logpost_deriv(p0 + 1e-3))
    timed("solve_posterior_covariance", lambda: (ls.set_params(p0), ls.solve_posterior_covariance()))
    x = fem.Function(V)
    timed("solve_posterior (mesh)", lambda: ls.solve_posterior(x))
    del ls, G
    torch.cuda.empty_cache()
