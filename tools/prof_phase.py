"""Short single-phase drivers at BASELINE config-2 shapes, for `ncu` captures and quick A/B timing.

    python tools/prof_phase.py solve [nodes] [ncols] [dim]   one block PCG solve of `ncols` sensor columns
    python tools/prof_phase.py dense [n_obs] [reps]          logposterior + gradient evaluations at n_obs sensors
    python tools/prof_phase.py gcov  [nodes] [dim] [lfac] [ncol]   forcing-covariance assembly + W = G Z, both kernel forms

Prints device time per phase (CUDA events around the call) and the library's per-kernel-family table."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ctypes as C
import stat_fem_b200 as stat_fem
from stat_fem_b200 import fem, _lib
from stat_fem_b200.solving_utils import SparseSolver


def dev_time(fn, reps=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


def show(ctx):
    prof = ctx.profile_read()
    for k, v in prof.items():
        if v["count"]:
            gbs = v["bytes"] / (v["ms"] * 1e6) if v["ms"] > 0 else 0.
            tfs = v["flops"] / (v["ms"] * 1e9) if v["ms"] > 0 else 0.
            print("  %-18s %9.3f ms  %6d launches  %8.1f GB/s  %7.2f TFLOP/s" % (k, v["ms"], v["count"], gbs, tfs))


def phase_solve(argv):
    nodes = int(argv[0]) if len(argv) > 0 else 1024
    ncols = int(argv[1]) if len(argv) > 1 else 384
    dim = int(argv[2]) if len(argv) > 2 else 2
    params = {}
    for kv in argv[3:]:
        k, v = kv.split("=")
        params[k] = float(v) if "." in v or "e" in v else int(v)
    V, A, b = fem.poisson_problem(nodes, dim)
    sensors = np.random.default_rng(1).uniform(0.02, 0.98, (max(ncols, 8), dim))
    im = stat_fem.InterpolationMatrix(V, sensors)
    im.assemble()
    ls = SparseSolver(A, solver_parameters=dict(block_size=ncols, **params))
    print("levels", ls.mg_levels)
    n = V.dim()
    Z = ls.ctx.empty((n, ncols))
    ls.solve_columns(im.csc, 0, ncols, Z)          # warm-up
    ls.ctx.profile_enable(0xffffffff)
    ms, _ = dev_time(lambda: ls.solve_columns(im.csc, 0, ncols, Z))
    print("solve %d columns at n=%d: %.2f ms, iterations %s, max relres %.2e" %
          (ncols, n, ms, ls.block_iterations, ls.last_relres.max()))
    show(ls.ctx)


def phase_dense(argv):
    nobs = int(argv[0]) if len(argv) > 0 else 4096
    reps = int(argv[1]) if len(argv) > 1 else 3
    graph = len(argv) > 2 and argv[2] == "graph"      # private stream, no per-family timers: the CUDA-graph replay path
    ctx = _lib.new_stream_context() if graph else _lib.new_context()
    if graph:
        torch.cuda.set_stream(ctx.tstream)
    lib = _lib.lib()
    rng = np.random.default_rng(0)
    Xs = rng.uniform(0.02, 0.98, (nobs, 2))
    X = rng.normal(size=(nobs, 64))
    Cu = ctx.to_device(1e-6 * (X @ X.T) / 64, np.float64)
    y = ctx.to_device(rng.normal(size=nobs) * 1e-2, np.float64)
    mu = ctx.to_device(rng.normal(size=nobs) * 1e-2, np.float64)
    dXs = ctx.to_device(Xs, np.float64)
    unc = ctx.to_device(np.array([2e-3]), np.float64)
    need = lib.sfem_dense_workspace_bytes(nobs)
    work = ctx.workspace(need)
    base = work.data_ptr()
    off = (-base) % 256
    theta = (C.c_double * 3)(np.log(0.7), np.log(1e-2), np.log(0.5))
    val = C.c_double(0.)
    grad = (C.c_double * 3)()

    def ev(with_grad):
        ctx.check(lib.sfem_logpost(ctx.handle, nobs, 2, _lib.ptr(dXs), _lib.ptr(y), _lib.ptr(unc), 0, _lib.ptr(mu), _lib.ptr(Cu),
                                   theta, C.byref(val), grad if with_grad else None, C.c_void_p(base + off), int(work.numel() - off)))
    ev(True)
    ev(True)
    if not graph:
        ctx.profile_enable(0xffffffff)
    ms_v, _ = dev_time(lambda: ev(False), reps)
    ms_g, _ = dev_time(lambda: ev(True), reps)
    print("n_obs=%d  logposterior %.3f ms (%.2f TFLOP/s on n^3/3)   value+gradient %.3f ms (%.2f TFLOP/s on n^3)  value %.12g" %
          (nobs, ms_v, nobs ** 3 / 3. / ms_v * 1e-9, ms_g, float(nobs) ** 3 / ms_g * 1e-9, val.value))
    show(ctx)


def phase_gcov(argv):
    nodes = int(argv[0]) if len(argv) > 0 else 1024
    dim = int(argv[1]) if len(argv) > 1 else 2
    lfac = float(argv[2]) if len(argv) > 2 else (1.5 if dim == 2 else 0.8)
    V, A, b = fem.poisson_problem(nodes, dim)
    h = 1. / (nodes - 1)
    sig = np.log(2.e-2)
    reg = 1.e-8 * h ** (2 * dim) * np.exp(2 * sig)
    G = stat_fem.ForcingCovariance(V, sig, np.log(lfac * h), 1.e-3, reg)
    G.assemble()
    G.destroy()
    G = stat_fem.ForcingCovariance(V, sig, np.log(lfac * h), 1.e-3, reg)
    t = time.perf_counter()
    ms, _ = dev_time(G.assemble)
    print("G assembly n=%d: %.2f ms device, %.2f ms wall, nnz %d (%.1f/row)" % (V.dim(), ms, 1e3 * (time.perf_counter() - t),
                                                                                  G.G.nnz, G.G.nnz / V.dim()))
    ctx = G.G.ctx
    ncol = int(argv[3]) if len(argv) > 3 else 512
    Z = ctx.empty((V.dim(), ncol)).normal_()
    W = ctx.empty((V.dim(), ncol))
    nnz = G.G.nnz
    for form in (False, True):
        _lib.USE_BLOCK_FORM[0] = form
        if form:
            ms, _ = dev_time(G.G.block_form)
            print("8x4-block form: %.2f ms to build, %d blocks, fill %.2f" % (ms, G.G.block_tiles, nnz / (32. * G.G.block_tiles)))
        G.G.matmul(Z, out=W)
        ms, _ = dev_time(lambda: G.G.matmul(Z, out=W), 3)
        print("W = G Z (%d columns, %s): %.2f ms  %.1f GB/s algorithmic  %.2f TFLOP/s" %
              (ncol, "DMMA block form" if form else "scalar CSR kernel", ms,
               (16. * V.dim() * ncol + 12. * nnz * ((ncol + 63) // 64)) / ms * 1e-6, 2. * nnz * ncol / ms * 1e-9))


if __name__ == "__main__":
    {"solve": phase_solve, "dense": phase_dense, "gcov": phase_gcov}[sys.argv[1]](sys.argv[2:])
