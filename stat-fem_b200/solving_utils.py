"""The conditioning solves of stat-fem (API of solving_utils.py in the reference) on the GPU.

``interp_covariance_to_data`` is the hot loop of the reference: for every sensor it performs two PETSc KSP solves and
one sparse mat-vec, one sensor at a time (solving_utils.py:139-142).  Here all sensor columns owned by this rank are
advanced together:  Z = A^-1 P  (block PCG, csrc/cg.cu + mg.cu, using A = A^T to halve the solves),  W = G Z
(csrc/spmm.cu),  result = W^T Z_left  (FP64 tensor-core GEMM, csrc/dgemm.cu)."""
import ctypes as C
import warnings
import numpy as np
from . import _lib
from . import fem
from . import adapter
from .parallel import EnsembleComm, COMM_SELF, ownership_range
from .ForcingCovariance import ForcingCovariance
from .InterpolationMatrix import InterpolationMatrix

_KNOWN_SOLVER_KEYS = {"ksp_rtol", "ksp_atol", "ksp_max_it", "pc_type", "block_size", "mg_nu", "mg_omega", "mg_omega2",
                      "mg_max_coarse", "workspace_gb", "symmetric_split", "keep_factors"}


def _lattice_nodes(coords, bbox=None):
    """level-1 lattice: 2^p cells per direction.  Preferred: a spacing of exactly TWICE the estimated mesh width,
    anchored at the lower corner of the bounding box and overhanging the upper corner (by at most 25 %): on lattice-like
    meshes every DOF then sits on a lattice node, edge midpoint or cell centre, the transfer weights are exactly 0, 1/2,
    1/4, 1, and all Galerkin coarse operators stay compact (3^d points) instead of filling their 5^d stencils.
    Otherwise the lattice is stretched over the box with a spacing of 1.5 to 3 mesh widths."""
    n, d = coords.shape
    lo, hi = bbox if bbox is not None else (coords.min(axis=0), coords.max(axis=0))
    ext = np.where(hi > lo, hi - lo, 1.)
    geo = float(np.prod(ext) ** (1. / d))
    per_side = float(n) ** (1. / d)
    if abs(per_side - round(per_side)) < 1.e-6:
        per_side = float(round(per_side))
    m1, hi_lat = [], []
    for k in range(d):
        cells = max(per_side * ext[k] / geo - 1., 2.)          # estimated number of mesh cells along direction k
        if abs(cells - round(cells)) < 1.e-6:
            cells = float(round(cells))
        r = cells / 2.                                         # lattice cells of width two mesh cells
        p_up = max(int(np.ceil(np.log2(r) - 1.e-9)), 1)
        # aligned lattice if it overhangs the box by at most 25 % (60 % in 3-D, where a stretched lattice costs 5^3-point
        # coarse stencils instead of 3^3)
        if 2. ** p_up / r - 1. <= (0.25 if d <= 2 else 0.6):
            H = 2. * ext[k] / cells
            m1.append(2 ** p_up + 1)
            hi_lat.append(lo[k] + H * 2 ** p_up)
        else:                                                  # stretched to the box: spacing in [1.5, 3) mesh widths
            p = max(int(np.floor(np.log2(cells / 1.5) + 1.e-9)), 1)
            m1.append(2 ** p + 1)
            hi_lat.append(hi[k])
    return lo, np.array(hi_lat), m1


class SparseSolver(object):
    """Stands where the reference has ``firedrake.LinearSolver(A, ...)`` (LinearSolver.py:136-140): owns the device
    copy of the stiffness matrix and the multigrid hierarchy, and solves blocks of right-hand sides.

    ``solver_parameters`` keys understood: ksp_rtol (default 1e-12), ksp_atol (0), ksp_max_it (500), pc_type ('mg' or
    'jacobi'), block_size (columns advanced together), mg_nu, mg_omega.  Other PETSc keys (e.g. ``pc_type: lu``) are
    ignored with a warning -- there is no direct-solver or CPU path."""

    def __init__(self, A, P=None, solver_parameters=None, nullspace=None, transpose_nullspace=None,
                 near_nullspace=None, options_prefix=None, function_space=None):
        A = adapter.as_matrix(A, function_space)      # fem.Matrix, SciPy sparse, or a duck-typed assembled matrix
        self.A = A
        params = dict(solver_parameters or {})
        for key in params:
            if key not in _KNOWN_SOLVER_KEYS:
                warnings.warn("solver parameter %r is not used by the GPU solver" % key)
        pc = params.get("pc_type", "mg")
        if pc not in ("mg", "jacobi"):
            warnings.warn("pc_type %r is not available on the GPU; using the multigrid V-cycle" % pc)
            pc = "mg"
        self.precond = 1 if pc == "mg" else 0
        self.rtol = float(params.get("ksp_rtol", 1.e-12))
        self.atol = float(params.get("ksp_atol", 0.))
        self.max_it = int(params.get("ksp_max_it", 500 if self.precond else 20000))
        self.block_size = params.get("block_size", None)
        self.workspace_gb = float(params.get("workspace_gb", 24.))
        self._workspace_gb_given = "workspace_gb" in params
        self.mg_nu = int(params.get("mg_nu", 2))
        # damped Jacobi by default (measured on the Poisson meshes of BASELINE.json: as few CG iterations as degree-2
        # Chebyshev weights); mg_omega2 sets a different weight for every second sweep (polynomial smoothing)
        self.mg_omega = float(params.get("mg_omega", 0.8))
        self.mg_omega2 = float(params.get("mg_omega2", self.mg_omega))
        self.mg_max_coarse = int(params.get("mg_max_coarse", 600))
        self.symmetric_split = bool(params.get("symmetric_split", True))
        # keep W = G A^-1 P resident after solve_prior (8 n n_obs / ranks bytes): posterior means then cost one
        # mesh-space solve instead of four.  None: decide from the free device memory at solve_prior time
        self.keep_factors = params.get("keep_factors", None)
        self.n = A.shape[0]
        self.ctx = _lib.new_context()
        self._work = None
        self.last_iterations = 0
        self.last_relres = np.zeros(0)
        self.total_iterations = 0
        ctx = self.ctx
        self.dA = _lib.cached_upload(A, "csr", ctx, lambda: _lib.DeviceCSR(ctx, A.shape, A.indptr, A.indices, A.data))
        ctx.check(_lib.lib().sfem_set_stiffness(ctx.handle, self.n, _lib.ptr(self.dA.indptr), _lib.ptr(self.dA.indices),
                                                _lib.ptr(self.dA.data)))
        self.bc_nodes = A.bc_nodes()
        self._check_symmetric()
        if self.precond == 1:
            self._setup_multigrid()

    def _check_symmetric(self):
        """The covariance path uses A = A^T twice: conjugate gradients, and (G Z)^T Z standing for the reference's
        P^T A^-1 G A^-1 P (which a direct LU would also give for a non-symmetric A).  One randomised bilinear test on the
        device, v^T (A w) = w^T (A v), rejects a non-symmetric operator (or non-symmetric Dirichlet elimination) up
        front instead of returning a silently different covariance."""
        import torch
        gen = torch.Generator(device=self.ctx.tdev)
        gen.manual_seed(12345)
        vw = torch.rand((self.n, 2), dtype=torch.float64, device=self.ctx.tdev, generator=gen) + 0.5
        Avw = self.ctx.empty((self.n, 2))        # through THIS solver's context (the uploaded matrix may be a cached one)
        self.ctx.check(_lib.lib().sfem_spmm_csr(self.ctx.handle, self.n, self.dA.nnz, _lib.ptr(self.dA.indptr),
                                                _lib.ptr(self.dA.indices), _lib.ptr(self.dA.data), 2, _lib.ptr(vw), 2,
                                                _lib.ptr(Avw), 2))
        a = float((vw[:, 0] * Avw[:, 1]).sum())
        b = float((vw[:, 1] * Avw[:, 0]).sum())
        scale = float((vw[:, 0] * Avw[:, 0]).sum().abs() + (vw[:, 1] * Avw[:, 1]).sum().abs()) + abs(a) + abs(b)
        if not abs(a - b) <= 1.e-9 * scale:
            raise ValueError("A must be symmetric (assemble it with the Dirichlet conditions eliminated symmetrically): "
                             "v^T A w = %.12g but w^T A v = %.12g" % (a, b))

    def _setup_multigrid(self):
        ctx, lib = self.ctx, _lib.lib()
        coords = np.ascontiguousarray(self.A.function_space.coords, dtype=np.float64)
        n, d = coords.shape
        lo, hi, m1 = _lattice_nodes(coords, self.A.function_space.bounding_box())
        mask = np.zeros(n, dtype=np.uint8)
        mask[self.bc_nodes] = 1
        self._d_coords = _lib.cached_upload(self.A.function_space, "coords", ctx, lambda: ctx.to_device(coords, np.float64))
        self._d_mask = _lib.cached_upload(self.A, "bcmask", ctx, lambda: ctx.to_device(mask, np.uint8))
        lo3 = (C.c_double * 3)(*(list(lo) + [0.] * (3 - d)))
        hi3 = (C.c_double * 3)(*(list(hi) + [0.] * (3 - d)))
        while True:
            m3 = (C.c_int * 3)(*(m1 + [1] * (3 - d)))
            rc = lib.sfem_mg_setup(ctx.handle, d, _lib.ptr(self._d_coords), _lib.ptr(self._d_mask), lo3, hi3, m3,
                                   self.mg_nu, self.mg_omega, self.mg_omega2, self.mg_max_coarse)
            if rc == _lib.SFEM_ERR_LIMIT and max(m1) > 3:
                m1 = [max((m - 1) // 2 + 1, 3) for m in m1]     # mesh edges longer than a lattice cell: coarser lattice
                continue
            ctx.check(rc)
            break
        self.lattice = m1
        self.mg_levels = [int(lib.sfem_mg_level_size(ctx.handle, l)) for l in range(lib.sfem_mg_num_levels(ctx.handle))]

    def __del__(self):
        try:
            _lib.destroy_context(self.ctx)
        except Exception:
            pass

    # ---- block solves -----------------------------------------------------------------------------------------
    def pick_block_size(self, ncols):
        if self.block_size is not None:
            return max(1, min(int(self.block_size), ncols))
        per_col = self.n * 8 * 7.
        kb = int(self.workspace_gb * 1.e9 / per_col)
        if kb < 192 and not self._workspace_gb_given:
            # large meshes: blocks narrower than three 64-column slabs leave the per-chunk costs of the tile kernels
            # unamortised (4M DOF: 64-column blocks ran at 0.59 of the HBM roofline); take up to 40 % of the memory that
            # is free right now to reach 192 columns
            import torch
            free = torch.cuda.mem_get_info(self.ctx.device)[0]       # plus what torch's allocator holds but does not use
            free += torch.cuda.memory_reserved(self.ctx.device) - torch.cuda.memory_allocated(self.ctx.device)
            kb = max(kb, min(int(0.4 * free / per_col), 192))
        kb = max(8, min(kb, 512))
        if kb >= 64:
            kb -= kb % 64
        return max(1, min(kb, ncols))

    @staticmethod
    def block_ranges(ncols, kb):
        """(first column, width) of the blocks ``ncols`` columns are solved in, at most ``kb`` wide.  The tile kernels
        work through a block in slabs of 64 columns and their per-chunk costs are shared by the slabs of a block, so the
        slabs are dealt out evenly over the fewest blocks (512 columns: 256 + 256 rather than 384 + 128) without
        adding a slab."""
        if kb < 64 or ncols <= kb:
            return [(j0, min(kb, ncols - j0)) for j0 in range(0, ncols, kb)]
        per = kb // 64
        slabs = -(-ncols // 64)
        nb = -(-slabs // per)
        base, extra = divmod(slabs, nb)
        out, j0 = [], 0
        for i in range(nb):
            k = min(64 * (base + (1 if i < extra else 0)), ncols - j0)
            out.append((j0, k))
            j0 += k
        return out

    def _workspace(self, k):
        need = _lib.lib().sfem_solve_workspace_bytes(self.ctx.handle, k, self.precond)
        if self._work is None or self._work.numel() < need + 256:
            self._work = None
            self._work = self.ctx.workspace(need)
        return self._work, need

    def solve_block(self, B, X):
        """X (n x k device block, any row stride) = A^-1 B.  Raises NotConvergedError on stagnation."""
        k = int(B.shape[1])
        work, need = self._workspace(k)
        iters = C.c_int(0)
        relres = (C.c_double * k)()
        base = work.data_ptr()
        off = (-base) % 256
        ctx = self.ctx
        ctx.check(_lib.lib().sfem_solve_block(ctx.handle, k, _lib.ptr(B), int(B.stride(0)), _lib.ptr(X), int(X.stride(0)),
                                              self.precond, self.rtol, self.atol, self.max_it, C.byref(iters), relres,
                                              C.c_void_p(base + off), int(work.numel() - off)))
        self.last_iterations = iters.value
        self.total_iterations += iters.value
        self.last_relres = np.frombuffer(relres, dtype=np.float64).copy()
        return X

    def solve_columns(self, P_csc, col_begin, ncols, out):
        "out (n x ncols device) = A^-1 P[:, col_begin:col_begin+ncols], advancing block_size columns at a time"
        ctx = self.ctx
        kb = self.pick_block_size(ncols)
        rhs = ctx.empty((self.n, kb + (kb & 1)))
        self.block_iterations = []
        for j0, k in self.block_ranges(ncols, kb):
            B = rhs[:, :k]
            ctx.check(_lib.lib().sfem_scatter_columns(ctx.handle, self.n, _lib.ptr(P_csc.indptr), _lib.ptr(P_csc.indices),
                                                      _lib.ptr(P_csc.data), col_begin + j0, k, _lib.ptr(B), int(B.stride(0))))
            self.solve_block(B, out[:, j0:j0 + k])
            self.block_iterations.append(self.last_iterations)
        return out

    def solve_vector(self, rhs_dev, apply_bcs):
        "single right-hand side on the device (length n) -> new device vector"
        ctx = self.ctx
        B = rhs_dev.clone().reshape(-1, 1)
        if apply_bcs and self.bc_nodes.size:
            B[ctx.to_device(self.bc_nodes.astype(np.int64), np.int64), 0] = 0.
        X = ctx.empty((self.n, 1))
        self.solve_block(B, X)
        return X.reshape(-1)

    def solve(self, x, b):
        """firedrake-style ``solver.solve(x, b)``: homogeneous Dirichlet lifting of the right-hand side iff ``A.bcs`` is
        set -- the switch solving_utils.py:46-57 flips around the covariance solves."""
        sol = self.solve_vector(self.ctx.to_device(adapter.function_array(b), np.float64), apply_bcs=self.A.bcs is not None)
        adapter.assign_function(x, sol.cpu().numpy())


def _aga_device(G, ls, rhs_dev):
    "A^-1 G A^-1 rhs on device vectors, Dirichlet lifting off (solving_utils.py:46-57)"
    x = ls.solve_vector(rhs_dev, apply_bcs=False)
    w = G.G.matmul(x)
    return ls.solve_vector(w, apply_bcs=False)


def solve_forcing_covariance(G, ls, rhs):
    "A^-1 G A^-1 rhs for one mesh-space vector (solving_utils.py:10-59)"
    if not isinstance(G, ForcingCovariance):
        raise TypeError("G must be a ForcingCovariance object")
    if not isinstance(ls, SparseSolver):
        raise TypeError("ls must be a firedrake LinearSolver")
    if not adapter.is_vector_like(rhs):
        raise TypeError("rhs must be a firedrake vector")
    if not G.is_assembled:
        G.assemble()
    out = _aga_device(G, ls, ls.ctx.to_device(adapter.function_array(rhs), np.float64))
    f = fem.Function(G.function_space, out.cpu().numpy())
    return f.vector()


def sharded_wtz(ensemble_comm, Wr, Zl, n_left, n_right, gemm_tn, zeros, chunk_bytes=1.e9):
    """The exchange step of the sharded covariance assembly.  Rank r holds ``Wr`` (n x its columns of the right
    matrix) and ``Zl`` (n x its columns of the left matrix).  Row chunks of ``Zl`` are all-gathered and every rank
    accumulates its rows ``Wr^T Z_left`` of the (n_right x n_left) result with ``gemm_tn(W_chunk, Z_chunk, out, beta)``
    (out = W_chunk^T Z_chunk + beta out); the row blocks are then gathered on rank 0 in ownership order -- the
    ``Scatter.toZero`` of solving_utils.py:146-159.  Returns the full array on rank 0, None elsewhere."""
    size = ensemble_comm.size
    n = int(Wr.shape[0])
    nloc = int(Wr.shape[1])
    res = zeros((max(nloc, 1), n_left))[:nloc]
    if size == 1:
        if nloc and n_left:
            gemm_tn(Wr, Zl, res, 0.0)
        return res
    lcounts = [ownership_range(n_left, s, size) for s in range(size)]
    chunk = max(1024, min(n, int(chunk_bytes / (8. * max(n_left, 1)))))
    widths = [b - a for a, b in lcounts]
    starts = list(range(0, n, chunk))
    # double-buffered: the all-gather of row chunk i+1 (NCCL, its own stream) is in flight while chunk i is multiplied;
    # both receive buffers are allocated once and every rank's block lands in place (all_gather_into_tensor)
    bufs = {}
    handle = ensemble_comm.all_gather_rows_start(Zl[starts[0]:min(n, starts[0] + chunk), :], widths, bufs)
    for i, k0 in enumerate(starts):
        k1 = min(n, k0 + chunk)
        pieces = ensemble_comm.all_gather_rows_finish(handle)
        if i + 1 < len(starts):
            handle = ensemble_comm.all_gather_rows_start(Zl[starts[i + 1]:min(n, starts[i + 1] + chunk), :], widths, bufs)
        for (a, b), piece in zip(lcounts, pieces):
            if nloc and b > a:
                gemm_tn(Wr[k0:k1], piece, res[:, a:b], 1.0)
    rcounts = [ownership_range(n_right, s, size) for s in range(size)]
    return ensemble_comm.gather_rows_to_root(res.contiguous(), [b - a for a, b in rcounts])


def sharded_symmetric_wtz(ensemble_comm, Ws, Z, We, Ze, n_obs, ops, chunk_bytes=1.e9):
    """Sharded covariance assembly for the symmetric case (left and right sensors identical): Cu = (S Z)^T Z + (E Z)^T Z
    with G = S + E as in ``_symmetric_split_wtz``.  Rank r holds ``Ws`` = S Z_r and ``Z`` = Z_r (n x its columns) and
    the rows ``We`` = (E Z_r), ``Ze`` = Z_r at the non-empty rows of E (ne x its columns; None if E is empty).

    Symmetric term, half-ring schedule: block (r, s) = Ws_r^T Z_s equals block (s, r)^T, so every unordered pair of
    ranks is computed once -- rank r multiplies its diagonal block (tiles on and below the diagonal) and the blocks
    (r, r+d) for ring distances d = 1 .. size/2, receiving Z_{r+d} row chunk by row chunk from rank r+d while it sends
    its own chunks to rank r-d (double-buffered, the transfer of chunk i+1 overlaps the GEMM of chunk i).  For an even
    size the pairs at distance size/2 are split between their two ranks (the lower rank takes the first half of its
    own columns, the upper rank the rest), so every rank does size/2 block-GEMMs instead of the size of the plain
    all-gather schedule.  The row blocks are gathered on rank 0 (the ``Scatter.toZero`` of solving_utils.py:146-159),
    folded into the full symmetric matrix there (``ops.sym_fold``) and the small unsymmetric term is added.

    ``ops``: the arithmetic -- gemm_tn(W, Z, out, beta), gemm_tn_lower(W, Z, out) (out = lower tiles of W^T Z,
    mirrored), zeros(shape), empty(shape), sym_fold(T, nblocks), add(T, E) -- DMMA kernels on the GPU, torch on CPU in
    the gloo tests.  Returns the (n_obs x n_obs) array on rank 0, None elsewhere."""
    R, r = ensemble_comm.size, ensemble_comm.rank
    ranges = [ownership_range(n_obs, s, R) for s in range(R)]
    widths = [b - a for a, b in ranges]
    lo, hi = ranges[r]
    nloc = hi - lo
    n = int(Ws.shape[0])
    resS = ops.zeros((max(nloc, 1), n_obs))[:nloc]
    if nloc:
        ops.gemm_tn_lower(Ws, Z, resS[:, lo:hi])
    # ---- half ring ------------------------------------------------------------------------------------------
    cmax = max(widths)
    chunk = max(1024, min(n, int(chunk_bytes / (8. * max(cmax, 1)))))
    plan = [(d, k0, min(n, k0 + chunk)) for d in range(1, R // 2 + 1) for k0 in range(0, n, chunk)]
    bufs = [ops.empty((chunk * max(cmax, 1),)) for _ in range(2)] if plan else []

    def post(i):
        d, k0, k1 = plan[i]
        dst, src = (r - d) % R, (r + d) % R
        send = Z[k0:k1] if (nloc and widths[dst]) else None
        recv = bufs[i % 2][:(k1 - k0) * widths[src]].view(k1 - k0, widths[src]) if (widths[src] and nloc) else None
        if send is not None and not send.is_contiguous():
            send = send.contiguous()
        return ensemble_comm.exchange_start(send, dst, recv, src), recv

    pending = post(0) if plan else None
    for i, (d, k0, k1) in enumerate(plan):
        reqs, piece = pending
        ensemble_comm.exchange_finish(reqs)
        if i + 1 < len(plan):
            pending = post(i + 1)
        s = (r + d) % R
        if piece is None:
            continue
        a, b = ranges[s]
        if R % 2 == 0 and d == R // 2:         # the pair (r, s) meets twice at this distance: split it
            if r < s:
                h = nloc // 2
                if h:
                    ops.gemm_tn(Ws[k0:k1, :h], piece, resS[:h, a:b], 1.0)
            else:
                h = widths[s] // 2
                if widths[s] - h:
                    ops.gemm_tn(Ws[k0:k1], piece[:, h:], resS[:, a + h:b], 1.0)
        else:
            ops.gemm_tn(Ws[k0:k1], piece, resS[:, a:b], 1.0)
    # ---- unsymmetric remainder: rows [lo, hi) of (E Z)^T Z -----------------------------------------------------
    ne = 0 if We is None else int(We.shape[0])
    resE = None
    if ne:
        resE = ops.zeros((max(nloc, 1), n_obs))[:nloc]
        local = ops.zeros((ne, cmax))
        if nloc:
            local[:, :nloc] = Ze
        gathered = ops.empty((R, ne, cmax))
        work = ensemble_comm.all_gather_into(gathered, local)
        if work is not None:
            work.wait()
        for s, (a, b) in enumerate(ranges):
            if nloc and b > a:
                ops.gemm_tn(We, gathered[s][:, :b - a], resE[:, a:b], 0.0)
    TS = ensemble_comm.gather_rows_to_root(resS.contiguous(), widths)
    TE = ensemble_comm.gather_rows_to_root(resE.contiguous(), widths) if ne else None
    if r != 0:
        return None
    ops.sym_fold(TS, R)
    return ops.add(TS, TE) if ne else TS


def _symmetric_split_wtz(ctx, lib, G, W, Z, n, k, restore_W=False):
    """W^T Z for W = G Z on one process, using G = S + E (E: entries without a stored transpose, csrc/symsplit.cu):
    (S Z)^T Z is symmetric, so only the tiles on and below the diagonal are computed and mirrored; (E Z)^T Z is a GEMM
    over the few non-empty rows of E.  W is overwritten with S Z (and put back to G Z at the end if ``restore_W``)."""
    rows, E = G.unsymmetric_part()
    ne = int(rows.shape[0])
    res = ctx.empty((k, k))
    if ne:
        We = E.matmul(Z)                                             # (ne x k): rows of E Z
        ctx.check(lib.sfem_rows_axpy(ctx.handle, ne, _lib.ptr(rows), k, -1.0, _lib.ptr(We), int(We.stride(0)),
                                     _lib.ptr(W), int(W.stride(0))))          # W <- S Z
    ctx.check(lib.sfem_dgemm_lower(ctx.handle, _lib.OP_T, _lib.OP_N, k, n, 1.0, _lib.ptr(W), int(W.stride(0)),
                                   _lib.ptr(Z), int(Z.stride(0)), 0.0, _lib.ptr(res), k))
    ctx.check(lib.sfem_mirror_lower(ctx.handle, k, _lib.ptr(res), k))
    if ne:
        Ze = ctx.empty((ne, k))
        ctx.check(lib.sfem_rows_gather(ctx.handle, ne, _lib.ptr(rows), k, _lib.ptr(Z), int(Z.stride(0)), _lib.ptr(Ze), k))
        ctx.check(lib.sfem_dgemm(ctx.handle, _lib.OP_T, _lib.OP_N, k, k, ne, 1.0, _lib.ptr(We), int(We.stride(0)),
                                 _lib.ptr(Ze), k, 1.0, _lib.ptr(res), k))
        if restore_W:
            ctx.check(lib.sfem_rows_axpy(ctx.handle, ne, _lib.ptr(rows), k, 1.0, _lib.ptr(We), int(We.stride(0)),
                                         _lib.ptr(W), int(W.stride(0))))      # W <- S Z + E Z = G Z
    return res


class _DeviceOps(object):
    "the arithmetic of sharded_symmetric_wtz on the GPU (DMMA GEMMs of csrc/dgemm.cu, csrc/symsplit.cu)"

    def __init__(self, ctx, lib):
        self.ctx, self.lib = ctx, lib
        self.zeros, self.empty = ctx.zeros, ctx.empty

    def gemm_tn(self, Wb, Zb, out, beta):
        ctx = self.ctx
        ctx.check(self.lib.sfem_dgemm(ctx.handle, _lib.OP_T, _lib.OP_N, int(Wb.shape[1]), int(Zb.shape[1]), int(Wb.shape[0]), 1.0,
                                      _lib.ptr(Wb), int(Wb.stride(0)), _lib.ptr(Zb), int(Zb.stride(0)), float(beta),
                                      _lib.ptr(out), int(out.stride(0))))

    def gemm_tn_lower(self, Wb, Zb, out):
        ctx, k = self.ctx, int(Wb.shape[1])
        ctx.check(self.lib.sfem_dgemm_lower(ctx.handle, _lib.OP_T, _lib.OP_N, k, int(Wb.shape[0]), 1.0, _lib.ptr(Wb),
                                            int(Wb.stride(0)), _lib.ptr(Zb), int(Zb.stride(0)), 0.0, _lib.ptr(out),
                                            int(out.stride(0))))
        ctx.check(self.lib.sfem_mirror_lower(ctx.handle, k, _lib.ptr(out), int(out.stride(0))))

    def sym_fold(self, T, nblocks):
        ctx = self.ctx
        ctx.check(self.lib.sfem_sym_fold(ctx.handle, int(T.shape[0]), int(nblocks), _lib.ptr(T), int(T.stride(0))))

    def add(self, T, E):
        ctx = self.ctx
        ctx.check(self.lib.sfem_axpby(ctx.handle, T.numel(), 1.0, _lib.ptr(T), 1.0, _lib.ptr(E), _lib.ptr(T)))
        return T


def _sharded_symmetric_device(ensemble_comm, ctx, lib, G, W, Z, n, nloc, n_obs, restore_W=False):
    "split W = G Z_r into S Z_r and the rows of E Z_r on this rank, then run the half-ring assembly"
    rows, E = G.unsymmetric_part()
    ne = int(rows.shape[0])
    We = Ze = None
    if ne:
        We = ctx.zeros((ne, max(nloc, 1)))[:, :nloc]
        Ze = ctx.zeros((ne, max(nloc, 1)))[:, :nloc]
        if nloc:
            E.matmul(Z, out=We)
            ctx.check(lib.sfem_rows_axpy(ctx.handle, ne, _lib.ptr(rows), nloc, -1.0, _lib.ptr(We), int(We.stride(0)),
                                         _lib.ptr(W), int(W.stride(0))))          # W <- S Z
            ctx.check(lib.sfem_rows_gather(ctx.handle, ne, _lib.ptr(rows), nloc, _lib.ptr(Z), int(Z.stride(0)), _lib.ptr(Ze),
                                           int(Ze.stride(0))))
    full = sharded_symmetric_wtz(ensemble_comm, W, Z, We, Ze, n_obs, _DeviceOps(ctx, lib))
    if ne and nloc and restore_W:
        ctx.check(lib.sfem_rows_axpy(ctx.handle, ne, _lib.ptr(rows), nloc, 1.0, _lib.ptr(We), int(We.stride(0)),
                                     _lib.ptr(W), int(W.stride(0))))              # W <- G Z again
    return full


def interp_covariance_to_data(im_left, G, ls, im_right, ensemble_comm=COMM_SELF, return_device=False, keep=None):
    """Phi_l^T A^-1 G A^-1 Phi_r on the root process, ``(0, 0)`` elsewhere (solving_utils.py:61-169).

    Row i of the reference's temporary array is ``P_left^T (A^-1 G A^-1 p_i)`` and the flattened array is *reshaped*
    to ``(n_left, n_right)``; with Z = A^-1 P and W = G Z_right that array is W^T Z_left.  The sensor columns of the
    right matrix are split over ``ensemble_comm`` exactly as PETSc would (lines 117-123); each rank solves its columns,
    all-gathers row-chunks of Z_left and accumulates its rows of W^T Z_left, which are then gathered on rank 0.

    ``keep``: a dict that receives this rank's block ``W = G Z_right`` (n x its columns) and its column range
    ``cols`` -- LinearSolver keeps it resident so that later products A^-1 G A^-1 P c are one solve of W c."""
    if not isinstance(im_left, InterpolationMatrix):
        raise TypeError("first argument to interp_covariance_to_data must be an InterpolationMatrix")
    if not isinstance(ls, SparseSolver):
        raise TypeError("ls must be a firedrake LinearSolver")
    if not isinstance(G, ForcingCovariance):
        raise TypeError("G must be a ForcingCovariance class")
    if not isinstance(im_right, InterpolationMatrix):
        raise TypeError("fourth argument to interp_covariance_to_data must be an InterpolationMatrix")
    if not isinstance(ensemble_comm, EnsembleComm):
        raise TypeError("ensemble_comm must be an MPI communicator created from a firedrake Ensemble")
    for obj in (im_left, im_right, G):
        if not obj.is_assembled:
            obj.assemble()
    ctx = ls.ctx
    lib = _lib.lib()
    n = ls.n
    n_left, n_right = im_left.n_data, im_right.n_data
    rank, size = ensemble_comm.rank, ensemble_comm.size
    imin, imax = ensemble_comm.column_range(n_right)
    nloc = imax - imin
    same = (im_left is im_right) or (n_left == n_right and np.array_equal(im_left.coords, im_right.coords))

    # Z_right (this rank's columns) and W = G Z_right
    Zr = ctx.empty((n, max(nloc, 1)))[:, :nloc]
    if nloc:
        ls.solve_columns(im_right.csc, imin, nloc, Zr)
    Wr = ctx.empty((n, max(nloc, 1)))[:, :nloc]
    if nloc:
        G.G.matmul(Zr, out=Wr)
    # Z_left: this rank's share of the left columns
    if same:
        lmin, lmax, Zl = imin, imax, Zr
    else:
        lmin, lmax = ensemble_comm.column_range(n_left)
        Zl = ctx.empty((n, max(lmax - lmin, 1)))[:, :lmax - lmin]
        if lmax > lmin:
            ls.solve_columns(im_left.csc, lmin, lmax - lmin, Zl)
    # rows [imin, imax) of the (n_right x n_left) array  W^T Z_left, gathered on the ensemble root
    def gemm_tn(Wb, Zb, out, beta):
        ctx.check(lib.sfem_dgemm(ctx.handle, _lib.OP_T, _lib.OP_N, int(Wb.shape[1]), int(Zb.shape[1]), int(Wb.shape[0]), 1.0,
                                 _lib.ptr(Wb), int(Wb.stride(0)), _lib.ptr(Zb), int(Zb.stride(0)), beta,
                                 _lib.ptr(out), int(out.stride(0))))
    if size == 1 and same and nloc >= 256 and ls.symmetric_split:
        full = _symmetric_split_wtz(ctx, lib, G, Wr, Zr, n, nloc, restore_W=keep is not None)
    elif size > 1 and same and n_right >= 256 * size and ls.symmetric_split:
        full = _sharded_symmetric_device(ensemble_comm, ctx, lib, G, Wr, Zr, n, nloc, n_right, restore_W=keep is not None)
    else:
        full = sharded_wtz(ensemble_comm, Wr, Zl, n_left, n_right, gemm_tn, ctx.zeros)
    if keep is not None:
        keep["W"], keep["cols"] = Wr, (imin, imax)
    if im_left.comm.rank == 0 and rank == 0:
        out = full.reshape(-1).reshape(n_left, n_right)      # the reference's reshape (lines 164-169)
        return out if return_device else out.cpu().numpy()
    if return_device:
        return None
    return np.zeros((0, 0))
