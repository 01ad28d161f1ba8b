"""Ensemble (observation-column) parallelism: one process per GPU, ``torch.distributed`` for the plumbing.

The reference splits the sensor columns over a Firedrake ``Ensemble.ensemble_comm`` with PETSc's contiguous ownership
rule and gathers the rows of the result on ensemble rank 0 (solving_utils.py:117-123, 146-159).  ``EnsembleComm`` is
the equivalent object here; with no process group it degenerates to a single rank."""
import numpy as np


def ownership_range(n, rank, size):
    "PETSc's default contiguous split of range(n): the first n % size ranks own one extra entry"
    base, extra = divmod(int(n), int(size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class EnsembleComm(object):
    """Communicator over which the independent FEM solves are divided (mpi4py-like surface used by the reference:
    rank, size, bcast, barrier, allreduce)."""

    def __init__(self, group=None, use_dist=None):
        self._dist = None
        self.group = group
        if use_dist is None:
            try:
                import torch.distributed as dist
                use_dist = dist.is_available() and dist.is_initialized()
            except Exception:
                use_dist = False
        if use_dist:
            import torch.distributed as dist
            self._dist = dist
            self.rank = dist.get_rank(group)
            self.size = dist.get_world_size(group)
        else:
            self.rank, self.size = 0, 1

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def column_range(self, n):
        return ownership_range(n, self.rank, self.size)

    def bcast(self, obj, root=0):
        if self._dist is None or self.size == 1:
            return obj
        box = [obj]
        self._dist.broadcast_object_list(box, src=root, group=self.group)
        return box[0]

    def barrier(self):
        if self._dist is not None and self.size > 1:
            self._dist.barrier(group=self.group)

    def allreduce(self, x, op=None):
        if self._dist is None or self.size == 1:
            return x
        import torch
        t = torch.as_tensor(np.asarray(x, dtype=np.float64)).clone()
        backend = self._dist.get_backend(self.group)
        if backend == "nccl":
            t = t.cuda()
        self._dist.all_reduce(t, group=self.group)
        out = t.cpu().numpy()
        return out.item() if out.ndim == 0 else out

    def broadcast_tensor(self, t, root=0):
        "in-place broadcast of a (device) tensor"
        if self._dist is not None and self.size > 1:
            self._dist.broadcast(t, src=root, group=self.group)
        return t

    def send_tensor(self, t, dst):
        self._dist.send(t.contiguous(), dst=dst, group=self.group)

    def recv_tensor(self, t, src):
        self._dist.recv(t, src=src, group=self.group)
        return t

    def allreduce_tensor(self, t):
        "in-place sum of a (device) tensor over the ensemble"
        if self._dist is not None and self.size > 1:
            self._dist.all_reduce(t, group=self.group)
        return t

    def reduce_tensor(self, t, root=0):
        "in-place sum onto ``root`` (the other ranks' buffers are left unspecified)"
        if self._dist is not None and self.size > 1:
            self._dist.reduce(t, dst=root, group=self.group)
        return t

    def all_gather_into(self, out, local):
        "out[s] = rank s's ``local`` (equal shapes; ``out`` is a persistent (size, ...) buffer owned by the caller)"
        if self._dist is None or self.size == 1:
            out[0].copy_(local)
            return None
        flat = out.view((out.shape[0] * out.shape[1],) + tuple(out.shape[2:]))       # concatenation along dim 0
        return self._dist.all_gather_into_tensor(flat, local, group=self.group, async_op=True)

    def exchange_start(self, send, dst, recv, src):
        """post one grouped send (to ``dst``) + receive (from ``src``); either tensor may be None.  Returns the list of
        request handles for ``exchange_finish``.  The two directions are grouped (ncclGroupStart/End under NCCL) so
        that a ring of such calls cannot deadlock."""
        ops = []
        if send is not None:
            ops.append(self._dist.P2POp(self._dist.isend, send, dst, group=self.group))
        if recv is not None:
            ops.append(self._dist.P2POp(self._dist.irecv, recv, src, group=self.group))
        return self._dist.batch_isend_irecv(ops) if ops else []

    @staticmethod
    def exchange_finish(reqs):
        for q in reqs:
            q.wait()

    # ---- tensor collectives used by the sharded Cu assembly -------------------------------------------------
    def all_gather_rows(self, local, counts):
        """local: tensor (m x n_loc).  Returns a list of tensors (m x counts[s]) -- every rank's column block of the
        same row range.  One collective per call; callers chunk over rows so the receive buffers stay small."""
        if self._dist is None or self.size == 1:
            return [local]
        import torch
        cmax = int(max(counts))
        m = int(local.shape[0])
        if int(local.shape[1]) == cmax:
            send = local.contiguous()
        else:       # collectives need equal shapes: pad the narrower blocks (PETSc ownership differs by at most one)
            send = torch.zeros((m, cmax), dtype=local.dtype, device=local.device)
            send[:, :local.shape[1]] = local
        outs = [torch.empty((m, cmax), dtype=local.dtype, device=local.device) for _ in counts]
        self._dist.all_gather(outs, send, group=self.group)
        return [o[:, :int(cnt)] for o, cnt in zip(outs, counts)]

    def all_gather_rows_start(self, local, counts, buffers=None):
        """asynchronous all_gather_rows: returns a handle for all_gather_rows_finish (the exchange overlaps the caller's
        GEMM).  ``buffers``: a dict the caller keeps between calls; it holds the persistent receive buffer (size, m,
        cmax) and the padded send buffer of each parity, so that the steady state allocates nothing and one
        all_gather_into_tensor lands every rank's block in place (no per-piece copies)."""
        if self._dist is None or self.size == 1:
            return ([local], None, counts)
        import torch
        cmax = int(max(counts))
        m = int(local.shape[0])
        store = buffers if buffers is not None else {}
        slot = store.get("next", 0)
        store["next"] = 1 - slot
        key = ("recv", slot)
        if key not in store or store[key].shape[1] < m or store[key].shape[2] != cmax:
            store[key] = torch.empty((self.size, m, cmax), dtype=local.dtype, device=local.device)
            store[("send", slot)] = torch.zeros((m, cmax), dtype=local.dtype, device=local.device)
        recv = store[key][:, :m, :]
        if not recv.is_contiguous():       # a shorter last chunk: use the leading part of the flat buffer instead
            recv = store[key].view(-1)[:self.size * m * cmax].view(self.size, m, cmax)
        if int(local.shape[1]) == cmax and local.is_contiguous():
            send = local
        else:       # collectives need equal shapes: pad the narrower blocks (PETSc ownership differs by at most one)
            send = store[("send", slot)].view(-1)[:m * cmax].view(m, cmax)
            send[:, :local.shape[1]] = local
        # (size * m, cmax): the concatenation along dim 0, the one output layout every backend accepts
        work = self._dist.all_gather_into_tensor(recv.view(self.size * m, cmax), send, group=self.group, async_op=True)
        return ([recv[s] for s in range(self.size)], work, counts, send)

    def all_gather_rows_finish(self, handle):
        outs, work, counts = handle[0], handle[1], handle[2]
        if work is None:
            return outs
        work.wait()
        return [o[:, :int(cnt)] for o, cnt in zip(outs, counts)]

    def gather_rows_to_root(self, local, counts, root=0):
        """local: tensor (counts[rank] x m) of result rows; returns the (sum(counts) x m) concatenation on root, None
        elsewhere (the Scatter.toZero of solving_utils.py:152-155)."""
        if self._dist is None or self.size == 1:
            return local
        import torch
        if self.rank == root:
            outs = [torch.empty((int(cnt), local.shape[1]), dtype=local.dtype, device=local.device) for cnt in counts]
        else:
            outs = None
        backend = self._dist.get_backend(self.group)
        if backend == "nccl":
            # NCCL gather: point-to-point sends to the root
            if self.rank == root:
                outs[root].copy_(local)
                reqs = [self._dist.irecv(outs[s], src=s, group=self.group) for s in range(self.size) if s != root]
                for r in reqs:
                    r.wait()
            else:
                self._dist.send(local.contiguous(), dst=root, group=self.group)
        else:
            rmax = int(max(counts))
            send = torch.zeros((rmax, local.shape[1]), dtype=local.dtype, device=local.device)
            send[:local.shape[0]] = local
            bufs = [torch.empty_like(send) for _ in counts] if self.rank == root else None
            self._dist.gather(send, bufs, dst=root, group=self.group)
            if self.rank == root:
                outs = [b[:int(cnt)] for b, cnt in zip(bufs, counts)]
        return torch.cat(outs, dim=0) if self.rank == root else None


COMM_SELF = EnsembleComm(use_dist=False)


def comm_world():
    "communicator over every launched process (the role of COMM_WORLD in LinearSolver.py:617, estimation.py:88-106)"
    return EnsembleComm()
