"""World-size 2-4 CPU tests (gloo) of the host-side sharding logic: PETSc-style column ownership, the all-gather /
accumulate / gather-to-root exchange of the sharded covariance assembly (solving_utils.sharded_wtz, the step that
replaces solving_utils.py:117-159 of the reference), object broadcast and allreduce.  The GEMM is a torch.mm stand-in
here; on the GPU the same function is driven with the DMMA GEMM of the library."""
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, n_left, n_right, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import stat_fem_b200  # noqa: F401  (import shim; no GPU is touched by the functions used here)
        from stat_fem_b200.parallel import EnsembleComm, ownership_range
        from stat_fem_b200.solving_utils import sharded_wtz
        comm = EnsembleComm()
        assert (comm.rank, comm.size) == (rank, world)
        rng = np.random.default_rng(7)
        Z_left = torch.from_numpy(rng.normal(size=(n, n_left)))
        W_right = torch.from_numpy(rng.normal(size=(n, n_right)))
        imin, imax = comm.column_range(n_right)
        lmin, lmax = comm.column_range(n_left)
        assert (imin, imax) == ownership_range(n_right, rank, world)

        def gemm_tn(Wb, Zb, out, beta):
            out.copy_(Wb.t() @ Zb + beta * out)

        full = sharded_wtz(comm, W_right[:, imin:imax].contiguous(), Z_left[:, lmin:lmax].contiguous(), n_left, n_right,
                           gemm_tn, lambda shape: torch.zeros(shape, dtype=torch.float64), chunk_bytes=8. * n_left * 1500)
        got = comm.bcast({"k": rank} if rank == 0 else None, root=0)
        assert got == {"k": 0}
        tot = comm.allreduce(np.array([float(rank + 1)]))
        assert tot[0] == sum(range(1, world + 1))
        comm.barrier()
        if rank == 0:
            ref = (W_right.t() @ Z_left).numpy()
            err = float(np.max(np.abs(full.numpy() - ref)) / np.max(np.abs(ref)))
            out_q.put(("ok", err, tuple(full.shape)))
        else:
            assert full is None
            out_q.put(("ok", 0.0, None))
    except Exception as exc:     # pragma: no cover - reported to the parent
        out_q.put(("fail", repr(exc), None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,n_left,n_right", [(2, 4000, 7, 7), (2, 3001, 5, 9), (2, 2500, 1, 3), (3, 2200, 8, 8), (4, 2100, 6, 10),
                                                    (4, 1500, 3, 3)])
def test_sharded_wtz(world, n, n_left, n_right):
    "world sizes 2-4, uneven column splits, ranks that own no column at all (3 columns over 4 ranks)"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, n_left, n_right, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for status, val, shape in results:
        assert status == "ok", val
        assert val < 1e-13
    shapes = [s for _, _, s in results if s is not None]
    assert shapes == [(n_right, n_left)]


class _TorchOps(object):
    "CPU stand-in for the DMMA kernels driven by sharded_symmetric_wtz"
    @staticmethod
    def zeros(shape):
        return torch.zeros(shape, dtype=torch.float64)

    @staticmethod
    def empty(shape):
        return torch.full(shape, float("nan"), dtype=torch.float64)      # stale data must never be read

    @staticmethod
    def gemm_tn(Wb, Zb, out, beta):
        out.copy_(Wb.t() @ Zb + (beta * out if beta != 0.0 else 0.0))

    @staticmethod
    def gemm_tn_lower(Wb, Zb, out):
        full = torch.tril(Wb.t() @ Zb)                                   # lower triangle only, then mirrored
        out.copy_(full + torch.tril(full, -1).t())

    @staticmethod
    def sym_fold(T, nblocks):
        from stat_fem_b200.parallel import ownership_range
        n = T.shape[0]
        blk = np.zeros(n, dtype=np.int64)
        for s in range(nblocks):
            a, b = ownership_range(n, s, nblocks)
            blk[a:b] = s
        other = torch.from_numpy(blk[:, None] != blk[None, :])
        T.copy_(torch.where(other, T + T.t(), T))

    @staticmethod
    def add(T, E):
        return T + E


def _worker_sym(rank, world, port, n, n_obs, chunk_rows, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import stat_fem_b200  # noqa: F401
        from stat_fem_b200.parallel import EnsembleComm
        from stat_fem_b200.solving_utils import sharded_symmetric_wtz
        comm = EnsembleComm()
        rng = np.random.default_rng(11)
        S = rng.normal(size=(n, n))
        S = torch.from_numpy(S + S.T)
        rows = np.sort(rng.choice(n, 37, replace=False))
        Efull = np.zeros((n, n))
        Efull[rows] = rng.normal(size=(37, n)) * (rng.uniform(size=(37, n)) < 0.02)
        Efull = torch.from_numpy(Efull)
        Z = torch.from_numpy(rng.normal(size=(n, n_obs)))
        lo, hi = comm.column_range(n_obs)
        Zr = Z[:, lo:hi].contiguous()
        Ws = S @ Zr
        We = (Efull @ Zr)[rows].contiguous()
        Ze = Zr[rows].contiguous()
        full = sharded_symmetric_wtz(comm, Ws, Zr, We, Ze, n_obs, _TorchOps, chunk_bytes=8. * chunk_rows * max(hi - lo, 1))
        if rank == 0:
            ref = (((S + Efull) @ Z).t() @ Z).numpy()
            err = float(np.max(np.abs(full.numpy() - ref)) / np.max(np.abs(ref)))
            sym = (S @ Z).t() @ Z
            out_q.put(("ok", err, tuple(full.shape)))
        else:
            assert full is None
            out_q.put(("ok", 0.0, None))
    except Exception as exc:     # pragma: no cover - reported to the parent
        import traceback
        out_q.put(("fail", traceback.format_exc(), None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,n_obs,chunk_rows", [(2, 1300, 9, 1024), (2, 1100, 8, 4000), (3, 1200, 10, 1024), (4, 1250, 13, 1024),
                                                      (4, 1100, 3, 4000), (5, 1100, 11, 1024)])
def test_sharded_symmetric_half_ring(world, n, n_obs, chunk_rows):
    """the half-ring schedule of the symmetric covariance assembly: even and odd world sizes (the split pair at distance
    size/2), uneven ownership, several row chunks per ring step, ranks without columns"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sym, args=(r, world, port, n, n_obs, chunk_rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for status, val, shape in results:
        assert status == "ok", val
        assert val < 1e-12
    assert [s for _, _, s in results if s is not None] == [(n_obs, n_obs)]


def test_ownership_ranges_match_petsc_rule():
    from stat_fem_b200.parallel import ownership_range
    for n in (0, 1, 7, 100, 4096):
        for size in (1, 2, 3, 8):
            ranges = [ownership_range(n, r, size) for r in range(size)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(size - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
