"""The sharded covariance assembly through the C ABI alone (sfem_comm_init + sfem_prior_cov, include/statfem_b200.h):
no torch.distributed anywhere -- the NCCL unique id travels through a file, as an MPI_Bcast would carry it.  Every rank
computes its columns and rank 0 compares the gathered Cu with the single-process result of the Python API.

    python tools/cabi_sharded_check.py NRANKS [nodes] [n_obs]        (spawns NRANKS processes, one GPU each)
"""
import ctypes as C
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def worker(rank, nranks, nodes, nobs, idfile):
    import torch
    torch.cuda.set_device(rank)
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import fem, _lib
    from stat_fem_b200.parallel import ownership_range
    from stat_fem_b200.solving_utils import SparseSolver
    lib = _lib.lib()
    V, A, b = fem.poisson_problem(nodes, 2)
    h = 1. / (nodes - 1)
    sig, l = np.log(2.e-2), np.log(1.5 * h)
    reg = 1.e-8 * h ** 4 * np.exp(2 * sig)
    sensors = np.random.default_rng(1).uniform(0.02, 0.98, (nobs, 2))
    G = stat_fem.ForcingCovariance(V, sig, l, 1.e-3, reg)
    G.assemble()
    sp = SparseSolver(A)                        # stiffness + multigrid on this rank's context
    ctx = sp.ctx
    im = stat_fem.InterpolationMatrix(V, sensors)
    im.assemble()
    # ---- rendezvous: 128 bytes from rank 0 (a file here; MPI_Bcast in an mpi4py program)
    uid = (C.c_char * 128)()
    if rank == 0:
        assert lib.sfem_nccl_unique_id(uid) == 0
        with open(idfile + ".tmp", "wb") as f:
            f.write(bytes(uid))
        os.rename(idfile + ".tmp", idfile)
    else:
        while not os.path.exists(idfile):
            time.sleep(0.05)
        uid = (C.c_char * 128).from_buffer_copy(open(idfile, "rb").read())
    ctx.check(lib.sfem_comm_init(ctx.handle, uid, rank, nranks))
    lo, hi = ownership_range(nobs, rank, nranks)
    Cu = ctx.empty((nobs, nobs)) if rank == 0 else None
    iters = C.c_int(0)
    torch.cuda.synchronize()
    t = time.perf_counter()
    ctx.check(lib.sfem_prior_cov(ctx.handle, nobs, _lib.ptr(im.csc.indptr), _lib.ptr(im.csc.indices), _lib.ptr(im.csc.data),
                                 _lib.ptr(G.G.indptr), _lib.ptr(G.G.indices), _lib.ptr(G.G.data), lo, hi, 0, 1, 1.e-12, 0., 500,
                                 None, None, _lib.ptr(Cu), C.byref(iters)))
    dt = time.perf_counter() - t
    if rank == 0:
        od = stat_fem.ObsData(sensors, np.zeros(nobs), 2.e-3)
        ref = stat_fem.LinearSolver(A, b, G, od).solve_prior()[1]
        got = Cu.cpu().numpy()
        err = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
        print("C-ABI sharded prior covariance: %d ranks, %d DOF, %d sensors, %.3f s, CG iterations %d, max rel diff vs the "
              "single-process API %.2e  %s" % (nranks, V.dim(), nobs, dt, iters.value, err, "OK" if err < 1e-10 else "FAIL"))
        assert err < 1e-10
    ctx.check(lib.sfem_comm_destroy(ctx.handle))


if __name__ == "__main__":
    nranks = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    nodes = int(sys.argv[2]) if len(sys.argv) > 2 else 513
    nobs = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    import torch.multiprocessing as mp
    idfile = "/tmp/sfem_nccl_id_%d" % os.getpid()
    procs = [mp.get_context("spawn").Process(target=worker, args=(r, nranks, nodes, nobs, idfile)) for r in range(nranks)]
    for p in procs:
        p.start()
    code = 0
    for p in procs:
        p.join(600)
        code |= (p.exitcode or 0)
    if os.path.exists(idfile):
        os.remove(idfile)
    sys.exit(code)
