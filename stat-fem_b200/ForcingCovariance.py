"""Sparse forcing covariance G (API of ForcingCovariance.py in the reference), assembled on the GPU.

The reference computes every one of the n^2 candidate entries in a Python loop and keeps those whose value relative
to the (nugget-regularised) diagonal exceeds ``cutoff`` (ForcingCovariance.py:159-167).  Here the same rule is applied
by the cell-list kernels of csrc/gcov.cu; the resulting CSR pattern (int32 ascending columns) is bit-identical and
the values agree to the last ulp or two of ``exp``."""
import ctypes as C
import numpy as np
from . import _lib
from . import fem
from . import adapter
from .covariance_functions import sqexp

_BORDER_CAP = 4096


class ForcingCovariance(object):
    def __init__(self, function_space, sigma, l, cutoff=1.e-3, regularization=1.e-8, cov=sqexp):
        # native P1Space, or any duck-typed P1 function space (Firedrake WithGeometry): see adapter.py
        function_space = adapter.as_function_space(function_space)
        self.function_space = function_space
        self.comm = function_space.comm
        self.nx = function_space.dim()
        self.nx_local = function_space.dim()
        assert regularization >= 0., "regularization parameter must be non-negative"
        self.sigma = sigma
        self.l = l
        self.cutoff = cutoff
        self.regularization = regularization
        self.cov = cov
        self.local_startind, self.local_endind = 0, self.nx
        self.is_assembled = False
        self.G = None
        self.max_row_nnz = 0
        self.n_borderline = 0

    def _integrate_basis_functions(self):
        "int phi_i dx for every DOF (ForcingCovariance.py:111-128)"
        return np.array(self.function_space.integrate_basis_functions())

    # ---- host-side scalars for the kernel -----------------------------------------------------------------
    def _search_radius(self, int_basis):
        """largest distance at which row[j]/row[i] > cutoff can hold for any row (0 -> all pairs)"""
        if self.cutoff <= 0.:
            return 0.
        b = int_basis
        e2s = np.exp(2. * self.sigma)
        arg = b * b.max() * e2s / (self.cutoff * (b * b * e2s + self.regularization))
        amax = float(arg.max())
        if amax <= 1.:
            return -1.      # nothing but the diagonal can pass
        return float(np.sqrt(2. * np.exp(2. * self.l) * np.log(amax)) * (1. + 1.e-5))

    def _host_rows(self, rows, coords, int_basis):
        """Re-evaluate complete rows with the reference's NumPy/SciPy expression (ForcingCovariance.py:159-164).  Only
        used for the handful of rows the kernel flags as containing a ratio within 1e-14 of the cutoff."""
        from scipy.spatial.distance import cdist
        out = {}
        e2s, em2l = np.exp(2. * self.sigma), np.exp(-2. * self.l)
        for i in rows:
            r2 = cdist(np.atleast_2d(coords[i]), coords) ** 2
            row = (int_basis[i] * int_basis * (e2s * np.exp(-0.5 * r2 * em2l)))[0]
            row[i] += self.regularization
            keep = (row / row[i] > self.cutoff)
            out[int(i)] = (row[keep], np.nonzero(keep)[0].astype(np.int32))
        return out

    def assemble(self):
        "compute G on the GPU (idempotent, ForcingCovariance.py:171-197)"
        if self.is_assembled:
            return
        if self.cov is not sqexp:
            return self._assemble_callable()
        ctx = _lib.get_context()
        lib = _lib.lib()
        coords = np.ascontiguousarray(self.function_space.coords, dtype=np.float64)
        n, d = coords.shape
        int_basis = self._integrate_basis_functions()
        R = self._search_radius(int_basis)
        lo, hi = self.function_space.bounding_box()
        ext = float((hi - lo).max())
        # the cell list holds at most 2^28 cells and 4096 per direction (csrc/gcov.cu); a cell edge LARGER than the search
        # radius is always valid (the 3^d neighbourhood still covers it), so small radii just get coarser cells
        min_edge = ext / float(min(4000, int(2. ** (28. / d)) - 2)) if ext > 0. else 1.
        if R < 0.:
            cell_edge = max(min_edge, ext / 2048., 1.e-300)            # diagonal only: any small cell will do
        elif R > 0.:
            cell_edge = max(R, min_edge)
        else:
            cell_edge = 0.                                             # cutoff <= 0: all pairs
        d_coords = _lib.cached_upload(self.function_space, "coords", ctx, lambda: ctx.to_device(coords, np.float64))
        d_b = _lib.cached_upload(self.function_space, "int_basis", ctx, lambda: ctx.to_device(int_basis, np.float64))
        import torch
        d_indptr = ctx.empty((n + 1,), torch.int64)
        d_border = ctx.empty((_BORDER_CAP,), torch.int32)
        nnz, maxnnz, nborder = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        lo3 = (C.c_double * 3)(*(list(lo) + [0.] * (3 - d)))
        hi3 = (C.c_double * 3)(*(list(hi) + [0.] * (3 - d)))
        def count(edge):
            return lib.sfem_gcov_count(ctx.handle, n, d, _lib.ptr(d_coords), _lib.ptr(d_b),
                                       float(np.exp(2. * self.sigma)), float(np.exp(-2. * self.l)), float(self.cutoff),
                                       float(self.regularization), float(int_basis.max()), float(edge), lo3, hi3,
                                       _lib.ptr(d_indptr), C.byref(nnz), C.byref(maxnnz), _lib.ptr(d_border),
                                       _BORDER_CAP, C.byref(nborder))
        rc = count(cell_edge)
        if rc == _lib.SFEM_ERR_LIMIT and cell_edge > 0.:
            # a row keeps more entries than the cell-list kernels sort in shared memory (long correlation length on a
            # fine mesh): the all-pairs mode has no such cap (rows are visited in ascending column order)
            rc = count(0.)
        ctx.check(rc)
        d_indices = ctx.empty((max(nnz.value, 1),), torch.int32)
        d_vals = ctx.empty((max(nnz.value, 1),), torch.float64)
        ctx.check(lib.sfem_gcov_fill(ctx.handle, _lib.ptr(d_indptr), _lib.ptr(d_indices), _lib.ptr(d_vals)))
        d_indices, d_vals = d_indices[:nnz.value], d_vals[:nnz.value]
        self.max_row_nnz = int(maxnnz.value)
        self.n_borderline = int(nborder.value)
        if self.n_borderline > 0:
            d_indptr, d_indices, d_vals = self._redecide(ctx, d_indptr, d_indices, d_vals, d_border, coords, int_basis)
        self.G = _lib.DeviceCSR(ctx, (n, n), d_indptr, d_indices, d_vals, on_device=True)
        self.is_assembled = True

    def _assemble_callable(self, chunk_entries=2.e7):
        """G for a user-supplied covariance callable ``cov(x1, x2, sigma, l) -> (len(x1), len(x2))`` (the ``cov=``
        argument of ForcingCovariance.py:52-53, used at line 161).  A Python callable can only be evaluated on the host:
        the reference's row rule (ForcingCovariance.py:159-167) is applied here to blocks of rows at a time with NumPy,
        in the reference's operation order, and the resulting CSR (int32 ascending columns) is uploaded.  Everything
        downstream of G (G x, W = G Z, Cu) runs on the GPU as for the built-in kernel; only ``sqexp`` has a device
        assembly kernel (csrc/gcov.cu)."""
        ctx = _lib.get_context()
        coords = np.ascontiguousarray(self.function_space.coords, dtype=np.float64)
        n = coords.shape[0]
        b = self._integrate_basis_functions()
        rows_per = int(max(1, min(n, chunk_entries // max(n, 1))))
        ptr = np.zeros(n + 1, dtype=np.int64)
        idx_parts, val_parts = [], []
        maxnnz = 0
        for r0 in range(0, n, rows_per):
            r1 = min(n, r0 + rows_per)
            K = np.asarray(self.cov(coords[r0:r1], coords, self.sigma, self.l), dtype=np.float64)
            assert K.shape == (r1 - r0, n), "cov must return an array of shape (len(x1), len(x2))"
            row = (b[r0:r1, None] * b[None, :]) * K
            ar = np.arange(r1 - r0)
            row[ar, r0 + ar] += self.regularization
            keep = (row / row[ar, r0 + ar][:, None]) > self.cutoff
            cnt = keep.sum(axis=1)
            ptr[r0 + 1:r1 + 1] = cnt
            maxnnz = max(maxnnz, int(cnt.max()))
            rr, cc = np.nonzero(keep)            # row-major: ascending columns inside every row
            idx_parts.append(cc.astype(np.int32))
            val_parts.append(row[rr, cc])
        np.cumsum(ptr, out=ptr)
        self.max_row_nnz = maxnnz
        self.n_borderline = 0
        self.G = _lib.DeviceCSR(ctx, (n, n), ptr, np.concatenate(idx_parts), np.concatenate(val_parts))
        self.is_assembled = True

    # ---- persistence: the assembled matrix with the parameters that produced it -------------
    def save(self, path):
        "write the assembled G (CSR) and its parameters to ``path`` (.npz)"
        if not self.is_assembled:
            self.assemble()
        ip, ix, v = self.G.to_host()
        np.savez(path, indptr=ip, indices=ix, data=v, sigma=self.sigma, l=self.l, cutoff=self.cutoff,
                 regularization=self.regularization, nx=self.nx, max_row_nnz=self.max_row_nnz,
                 cov=getattr(self.cov, "__name__", "callable"))

    @classmethod
    def load(cls, function_space, path, cov=sqexp):
        "a ForcingCovariance over ``function_space`` with G read back from ``save`` (no re-assembly)"
        z = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
        G = cls(function_space, float(z["sigma"]), float(z["l"]), float(z["cutoff"]), float(z["regularization"]), cov)
        assert int(z["nx"]) == G.nx, "the saved matrix belongs to a function space of another size"
        ctx = _lib.get_context()
        G.G = _lib.DeviceCSR(ctx, (G.nx, G.nx), z["indptr"], z["indices"], z["data"])
        G.max_row_nnz = int(z["max_row_nnz"])
        G.is_assembled = True
        return G

    def unsymmetric_part(self):
        """(rows, E) with E the compact-row device CSR of the entries of G whose transpose is not stored, and ``rows``
        the (device, int32) indices of its non-empty rows; cached.  G - E is exactly symmetric (see csrc/symsplit.cu)."""
        if not self.is_assembled:
            self.assemble()
        if self.__dict__.get("_unsym") is None:
            import torch
            ctx, lib, G = self.G.ctx, _lib.lib(), self.G
            n = self.nx
            cnt, flag = ctx.empty((n,), torch.int64), ctx.empty((n,), torch.int64)
            efull, rowpos = ctx.empty((n + 1,), torch.int64), ctx.empty((n + 1,), torch.int64)
            nrows, nnz = C.c_int64(0), C.c_int64(0)
            ctx.check(lib.sfem_csr_unsym_count(ctx.handle, n, _lib.ptr(G.indptr), _lib.ptr(G.indices), _lib.ptr(cnt), _lib.ptr(flag),
                                               _lib.ptr(efull), _lib.ptr(rowpos), C.byref(nrows), C.byref(nnz)))
            rows = ctx.empty((max(nrows.value, 1),), torch.int32)
            eptr = ctx.zeros((nrows.value + 1,), torch.int64)
            eidx = ctx.empty((max(nnz.value, 1),), torch.int32)
            evals = ctx.empty((max(nnz.value, 1),), torch.float64)
            if nrows.value:
                ctx.check(lib.sfem_csr_unsym_fill(ctx.handle, n, _lib.ptr(G.indptr), _lib.ptr(G.indices), _lib.ptr(G.data),
                                                  _lib.ptr(efull), _lib.ptr(rowpos), _lib.ptr(rows), _lib.ptr(eptr), _lib.ptr(eidx),
                                                  _lib.ptr(evals)))
            E = _lib.DeviceCSR(ctx, (nrows.value, n), eptr, eidx[:nnz.value], evals[:nnz.value], on_device=True)
            self._unsym = (rows[:nrows.value], E)
        return self._unsym

    def _redecide(self, ctx, d_indptr, d_indices, d_vals, d_border, coords, int_basis):
        """Splice host-decided rows into the CSR arrays (rare: a ratio landed within 1e-14 of the cutoff)."""
        n = self.nx
        if self.n_borderline > _BORDER_CAP:
            rows = np.arange(n)
        else:
            rows = np.unique(d_border[:self.n_borderline].cpu().numpy())
        fixed = self._host_rows(rows, coords, int_basis)
        indptr, indices, vals = d_indptr.cpu().numpy(), d_indices.cpu().numpy(), d_vals.cpu().numpy()
        counts = np.diff(indptr)
        for i, (v, cidx) in fixed.items():
            counts[i] = len(cidx)
        new_ptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(counts, out=new_ptr[1:])
        new_idx = np.empty(new_ptr[-1], dtype=np.int32)
        new_val = np.empty(new_ptr[-1], dtype=np.float64)
        for i in range(n):
            if i in fixed:
                new_val[new_ptr[i]:new_ptr[i + 1]], new_idx[new_ptr[i]:new_ptr[i + 1]] = fixed[i]
            else:
                new_idx[new_ptr[i]:new_ptr[i + 1]] = indices[indptr[i]:indptr[i + 1]]
                new_val[new_ptr[i]:new_ptr[i + 1]] = vals[indptr[i]:indptr[i + 1]]
        self.max_row_nnz = int(counts.max())
        return ctx.to_device(new_ptr, np.int64), ctx.to_device(new_idx, np.int32), ctx.to_device(new_val, np.float64)

    def _compute_G_vals(self):
        """(dict row -> (values, int32 columns), max_row_nnz) like ForcingCovariance.py:130-169, read back from the
        GPU-assembled matrix."""
        self.assemble()
        indptr, indices, vals = self.G.to_host()
        G_dict = {i: (vals[indptr[i]:indptr[i + 1]], indices[indptr[i]:indptr[i + 1]]) for i in range(self.nx)}
        return G_dict, self.max_row_nnz

    def destroy(self):
        if not self.is_assembled:
            return
        self.G = None
        self._unsym = None
        self.is_assembled = False

    def mult(self, x, y):
        "y <- G x for mesh-space vectors (ForcingCovariance.py:212-242)"
        assert adapter.is_vector_like(x), "x must be a firedrake vector"
        assert adapter.is_vector_like(y), "y must be a firedrake vector"
        if not self.is_assembled:
            self.assemble()
        ctx = self.G.ctx
        dx = ctx.to_device(adapter.function_array(x), np.float64)
        y.set_local(self.G.matmul(dx).cpu().numpy())

    def get_nx(self):
        return self.nx

    def get_nx_local(self):
        return self.nx_local

    def __str__(self):
        return "Forcing Covariance with {} mesh points".format(self.get_nx())
