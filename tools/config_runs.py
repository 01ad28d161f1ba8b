"""Runs of the BASELINE.json stress configurations that are not the bench line: prints one JSON object
per configuration with device timings (CUDA events) and the achieved rates against the algorithmic work figures.

    python tools/config_runs.py c4 [nodes]     forcing-covariance stress: 2048^2 DOF, ~300 nnz/row in G, assembly + W = G Z (512 columns)
    python tools/config_runs.py c5 [n_obs]     dense-end stress: 16384 observations, solve_posterior_covariance + 200 value+gradient evaluations
    python tools/config_runs.py c3lite [nodes] 3-D Poisson (Kuhn tetrahedra), one GPU's share of the columns
"""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import stat_fem_b200 as stat_fem
from stat_fem_b200 import fem, _lib


def dev_ms(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), r


def c4(nodes=2048, nobs=512):
    V, A, b = fem.poisson_problem(nodes, 2)
    h = 1. / (nodes - 1)
    sig, lf = np.log(2.e-2), np.log(2.63 * h)
    reg = 1.e-8 * h ** 4 * np.exp(2 * sig)
    G = stat_fem.ForcingCovariance(V, sig, lf, 1.e-3, reg)
    G.assemble(); G.destroy()
    G = stat_fem.ForcingCovariance(V, sig, lf, 1.e-3, reg)
    ms_g, _ = dev_ms(G.assemble)
    n, nnz = V.dim(), G.G.nnz
    ctx = G.G.ctx
    sensors = np.random.default_rng(3).uniform(0.02, 0.98, (nobs, 2))
    od = stat_fem.ObsData(sensors, np.zeros(nobs), 2.e-3)
    ls = stat_fem.LinearSolver(A, b, G, od)
    ls.im.assemble()
    Z = ls.solver.ctx.empty((n, nobs))
    ms_z, _ = dev_ms(lambda: ls.solver.solve_columns(ls.im.csc, 0, nobs, Z))
    W = ctx.empty((n, nobs))
    _lib.USE_BLOCK_FORM[0] = False                    # scalar CSR kernel (spmm.cu)
    G.G.matmul(Z, out=W)
    ms_w0, _ = dev_ms(lambda: G.G.matmul(Z, out=W))
    _lib.USE_BLOCK_FORM[0] = True                     # 8x4-block form on FP64 tensor cores (bsr.cu)
    ms_build, _ = dev_ms(G.G.block_form)
    G.G.matmul(Z, out=W)
    ms_w, _ = dev_ms(lambda: G.G.matmul(Z, out=W))
    out = dict(config="c4", n=n, nnzG=int(nnz), nnz_per_row=nnz / n, n_obs=nobs,
               GZ_scalar_csr_ms=ms_w0, block_form_build_ms=ms_build, block_fill=nnz / (32. * G.G.block_tiles),
               G_assembly_ms=ms_g, G_assembly_GBs=(12. * nnz + 8. * (n + 1) + 24. * n) / ms_g * 1e-6,
               solve_ms=ms_z, cg_iterations=list(ls.solver.block_iterations), relres_max=float(ls.solver.last_relres.max()),
               GZ_ms=ms_w, GZ_TFLOPs=2. * nnz * nobs / ms_w * 1e-9, GZ_GBs=(16. * n * nobs + 12. * nnz * ((nobs + 63) // 64)) / ms_w * 1e-6,
               n_borderline=int(G.n_borderline))
    print(json.dumps(out))


def c5(nobs=16384, nodes=256, pairs=200):
    V, A, b = fem.poisson_problem(nodes, 2)
    h = 1. / (nodes - 1)
    sig, lf = np.log(2.e-2), np.log(1.5 * h)
    reg = 1.e-8 * h ** 4 * np.exp(2 * sig)
    rng = np.random.default_rng(4)
    sensors = rng.uniform(0.02, 0.98, (nobs, 2))
    y = 0.7 * np.prod(np.sin(2 * np.pi * sensors), axis=1) + rng.normal(scale=2.e-3, size=nobs)
    od = stat_fem.ObsData(sensors, y, 2.e-3)
    G = stat_fem.ForcingCovariance(V, sig, lf, 1.e-3, reg)
    ls = stat_fem.LinearSolver(A, b, G, od)
    ls.host_results = False
    ms_prior, _ = dev_ms(ls.solve_prior)
    th0 = np.array([np.log(0.7), np.log(1.e-2), np.log(0.5)])
    ls.fuse_gradient = True
    ls.logposterior(th0 + 1e-4)          # workspace allocation
    t = time.perf_counter()
    ms_map, vals = dev_ms(lambda: [(ls.logposterior(th0 + 0.01 * k / pairs * np.array([1., -1., 1.])),
                                    ls.logpost_deriv(th0 + 0.01 * k / pairs * np.array([1., -1., 1.]))) for k in range(pairs)])
    wall_map = time.perf_counter() - t
    ls.fuse_gradient = False
    ls.set_params(th0)
    ls.host_results = False               # device-resident result: the 2.1 GB copy to the host is not part of the kernel time
    ms_cov, (muy, Cuy) = dev_ms(ls.solve_posterior_covariance)
    Cuy = Cuy.cpu().numpy()
    out = dict(config="c5", n=V.dim(), n_obs=nobs, solve_prior_ms=ms_prior, pairs=pairs, map_ms=ms_map, map_wall_s=wall_map,
               ms_per_value_gradient=ms_map / pairs, value_gradient_TFLOPs_on_n3=float(nobs) ** 3 / (ms_map / pairs) * 1e-9,
               posterior_cov_ms=ms_cov, posterior_cov_TFLOPs=(2. / 3. + 4.) * float(nobs) ** 3 / ms_cov * 1e-9,
               logpost_first=vals[0][0], logpost_last=vals[-1][0], Cuy_trace=float(np.trace(Cuy)))
    print(json.dumps(out))


def c3lite(nodes=65, nobs=256):
    V, A, b = fem.poisson_problem(nodes, 3)
    h = 1. / (nodes - 1)
    sig, lf = np.log(2.e-2), np.log(0.8 * h)
    reg = 1.e-8 * h ** 6 * np.exp(2 * sig)
    sensors = np.random.default_rng(2).uniform(0.02, 0.98, (nobs, 3))
    od = stat_fem.ObsData(sensors, np.zeros(nobs), 2.e-3)
    for rep in range(2):         # the first pass pays one-off allocations
        G = stat_fem.ForcingCovariance(V, sig, lf, 1.e-3, reg)
        ms_g, _ = dev_ms(G.assemble)
        ls = stat_fem.LinearSolver(A, b, G, od)
        ls.host_results = False
        ms_prior, _ = dev_ms(ls.solve_prior)
    out = dict(config="c3lite", n=V.dim(), n_obs=nobs, nnzG_per_row=G.G.nnz / V.dim(), G_assembly_ms=ms_g, solve_prior_ms=ms_prior,
               cg_iterations=list(ls.solver.block_iterations), relres_max=float(ls.solver.last_relres.max()), levels=ls.solver.mg_levels)
    print(json.dumps(out))


if __name__ == "__main__":
    which = sys.argv[1]
    args = [int(a) for a in sys.argv[2:]]
    {"c4": c4, "c5": c5, "c3lite": c3lite}[which](*args)
