"""
Host-side stand-ins for the few Firedrake objects stat-fem's API touches.

The reference receives a Firedrake ``FunctionSpace``, an assembled ``Matrix`` ``A`` (with Dirichlet
conditions), a load ``Function``/``Vector`` ``b`` and fills a ``Function`` ``x`` (LinearSolver.py:63-66,
224; ForcingCovariance.py:52).  Firedrake/PETSc do not exist in this environment and are outside the hot
path ("lower boundary"), so this module provides light NumPy containers with the same
attribute names the reference's call sites use, plus a P1 mesh/assembly generator for the synthetic
benchmark problems (UnitInterval/UnitSquare/UnitCube, nodes-per-side = cells + 1 as in
examples/poisson.py:15-18).  Nothing here is on the GPU hot path: it produces the *inputs* (A as CSR, DOF
coordinates, lumped basis integrals, Dirichlet node list, load vector) that the CUDA library consumes.
"""
import itertools
import math
import numpy as np
import scipy.sparse as sp


class SelfComm(object):
    """Single-process communicator stand-in (rank 0 of 1): what ``COMM_SELF``/``COMM_WORLD`` look like
    when the script runs on one process.  The multi-GPU ensemble communicator lives in ``parallel.py``."""
    rank = 0
    size = 1

    def bcast(self, obj, root=0):
        return obj

    def barrier(self):
        pass

    def allreduce(self, x, op=None):
        return x

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1


COMM_SELF = SelfComm()
COMM_WORLD = COMM_SELF


class Mesh(object):
    """Simplicial mesh: ``coords`` (n_nodes, d) float64, ``cells`` (n_cells, d+1) int32."""

    def __init__(self, coords, cells, comm=COMM_SELF):
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        if coords.ndim == 1:
            coords = coords.reshape(-1, 1)
        self.coords = coords
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        assert self.cells.shape[1] == self.coords.shape[1] + 1, "cells must be simplices"
        self.comm = comm
        self._locator = None

    def cell_dimension(self):
        return self.coords.shape[1]

    def boundary_nodes(self):
        "nodes on the boundary of the axis-aligned bounding box (exact for the Unit* meshes)"
        lo = self.coords.min(axis=0)
        hi = self.coords.max(axis=0)
        on = np.zeros(self.coords.shape[0], dtype=bool)
        for k in range(self.coords.shape[1]):
            on |= (self.coords[:, k] == lo[k]) | (self.coords[:, k] == hi[k])
        return np.nonzero(on)[0].astype(np.int32)

    # -- point location (the role of Firedrake's mesh.locate_cell, InterpolationMatrix.py:123) --------
    def _build_locator(self):
        d = self.cell_dimension()
        V = self.coords[self.cells]                       # (nc, d+1, d)
        lo = V.min(axis=1)
        hi = V.max(axis=1)
        glo = self.coords.min(axis=0)
        ghi = self.coords.max(axis=0)
        nc = self.cells.shape[0]
        nb = max(1, int(round((nc / 2.0) ** (1.0 / d))))
        nb = min(nb, 2048 if d == 1 else (1024 if d == 2 else 128))
        width = np.where(ghi > glo, (ghi - glo) / nb, 1.0)
        ilo = np.clip(np.floor((lo - glo) / width).astype(np.int64), 0, nb - 1)
        ihi = np.clip(np.floor((hi - glo) / width).astype(np.int64), 0, nb - 1)
        span = ihi - ilo + 1
        # expand every cell over the bins its bounding box touches
        counts = np.prod(span, axis=1)
        total = int(counts.sum())
        cell_rep = np.repeat(np.arange(nc, dtype=np.int64), counts)
        offs = np.arange(total, dtype=np.int64) - np.repeat(np.cumsum(counts) - counts, counts)
        binid = np.zeros(total, dtype=np.int64)
        stride = 1
        rem = offs
        for k in range(d - 1, -1, -1):
            sk = np.repeat(span[:, k], counts)
            ik = rem % sk
            rem = rem // sk
            binid += (np.repeat(ilo[:, k], counts) + ik) * stride
            stride *= nb
        order = np.argsort(binid, kind="stable")
        self._locator = dict(glo=glo, width=width, nb=nb, d=d,
                             bin_cells=cell_rep[order],
                             bin_ptr=np.searchsorted(binid[order], np.arange(nb ** d + 1)))

    def locate_cells(self, pts, tol=1.e-12):
        """For each point: (cell index or -1, barycentric weights (d+1)).  Picks the candidate cell with
        the largest minimum barycentric coordinate, requiring it >= -tol."""
        pts = np.asarray(pts, dtype=np.float64)
        if pts.ndim == 1:
            pts = pts.reshape(-1, self.cell_dimension())
        if self._locator is None:
            self._build_locator()
        L = self._locator
        d, nb = L["d"], L["nb"]
        ib = np.clip(np.floor((pts - L["glo"]) / L["width"]).astype(np.int64), 0, nb - 1)
        binid = np.zeros(pts.shape[0], dtype=np.int64)
        for k in range(d):
            binid = binid * nb + ib[:, k]
        npts = pts.shape[0]
        cells_out = np.full(npts, -1, dtype=np.int64)
        w_out = np.zeros((npts, d + 1))
        best_min = np.full(npts, -np.inf)
        start = L["bin_ptr"][binid]
        count = L["bin_ptr"][binid + 1] - start
        # k-th candidate cell of every point at once (vectorised over the points; bins hold a handful of cells)
        for k in range(int(count.max()) if npts else 0):
            sel = np.nonzero(count > k)[0]
            cand = L["bin_cells"][start[sel] + k]
            V = self.coords[self.cells[cand]]                                # (m, d+1, d)
            T = np.transpose(V[:, 1:, :] - V[:, :1, :], (0, 2, 1))
            lam = np.linalg.solve(T, (pts[sel] - V[:, 0, :])[:, :, None])[:, :, 0]
            lam = np.concatenate([(1. - lam.sum(axis=1))[:, None], lam], axis=1)
            lmin = lam.min(axis=1)
            better = lmin > best_min[sel]                                    # first best candidate wins, as argmax did
            idx = sel[better]
            best_min[idx] = lmin[better]
            cells_out[idx] = cand[better]
            w_out[idx] = lam[better]
        outside = best_min < -tol
        cells_out[outside] = -1
        w_out[outside] = 0.
        return cells_out, w_out

    def locate_cell(self, x, tol=1.e-12):
        c, _ = self.locate_cells(np.asarray(x, dtype=np.float64).reshape(1, -1), tol)
        return None if c[0] < 0 else int(c[0])


def UnitIntervalMesh(ncells, comm=COMM_SELF):
    x = np.linspace(0., 1., ncells + 1)
    cells = np.stack([np.arange(ncells), np.arange(1, ncells + 1)], axis=1)
    return Mesh(x.reshape(-1, 1), cells, comm)


def UnitSquareMesh(nx, ny, comm=COMM_SELF):
    """(nx+1)*(ny+1) nodes, lexicographic (x fastest); each square split along its (0,0)-(1,1) diagonal."""
    xs = np.linspace(0., 1., nx + 1)
    ys = np.linspace(0., 1., ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v00 = (j * (nx + 1) + i).ravel()
    v10 = v00 + 1
    v01 = v00 + (nx + 1)
    v11 = v01 + 1
    cells = np.concatenate([np.stack([v00, v10, v11], axis=1), np.stack([v00, v11, v01], axis=1)], axis=0)
    return Mesh(coords, cells, comm)


def UnitCubeMesh(nx, ny, nz, comm=COMM_SELF):
    """(nx+1)(ny+1)(nz+1) nodes, lexicographic (x fastest); each cube split into 6 Kuhn tetrahedra."""
    xs = np.linspace(0., 1., nx + 1)
    ys = np.linspace(0., 1., ny + 1)
    zs = np.linspace(0., 1., nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = ((k * (ny + 1) + j) * (nx + 1) + i).ravel()
    step = np.array([1, nx + 1, (nx + 1) * (ny + 1)])
    tets = []
    for perm in itertools.permutations(range(3)):
        v0 = base
        v1 = v0 + step[perm[0]]
        v2 = v1 + step[perm[1]]
        v3 = v2 + step[perm[2]]
        tets.append(np.stack([v0, v1, v2, v3], axis=1))
    return Mesh(coords, np.concatenate(tets, axis=0), comm)


class P1Space(object):
    """Continuous piecewise-linear space on a simplicial mesh: one DOF per node, DOF i sits at
    ``mesh.coords[i]``.  Stands in for the Firedrake ``FunctionSpace`` argument of
    ForcingCovariance/InterpolationMatrix (ForcingCovariance.py:74-79, InterpolationMatrix.py:60-65)."""

    def __init__(self, mesh):
        self._mesh = mesh
        self.comm = mesh.comm
        self.cell_node_list = mesh.cells
        self._int_basis = None

    def mesh(self):
        return self._mesh

    def ufl_domain(self):
        return self._mesh

    def dim(self):
        return self._mesh.coords.shape[0]

    node_count = property(lambda self: self._mesh.coords.shape[0])

    @property
    def coords(self):
        return self._mesh.coords

    def bounding_box(self):
        "(lo, hi) of the DOF coordinates, computed once (two passes over n x d doubles)"
        if self.__dict__.get("_bbox") is None:
            c = self._mesh.coords
            self._bbox = (np.array([c[:, k].min() for k in range(c.shape[1])]), np.array([c[:, k].max() for k in range(c.shape[1])]))
        return self._bbox[0].copy(), self._bbox[1].copy()

    def cell_volumes(self):
        V = self._mesh.coords[self._mesh.cells]
        T = V[:, 1:, :] - V[:, :1, :]
        d = T.shape[1]
        return np.abs(np.linalg.det(T)) / math.factorial(d)

    def integrate_basis_functions(self):
        "b_i = int phi_i dx = sum over cells containing i of |cell|/(d+1)  (ForcingCovariance.py:111-128)"
        if self._int_basis is None:
            vol = self.cell_volumes()
            d1 = self._mesh.cells.shape[1]
            out = np.zeros(self.dim())
            for a in range(d1):
                out += np.bincount(self._mesh.cells[:, a], weights=vol / d1, minlength=self.dim())
            self._int_basis = out
        return self._int_basis


def FunctionSpace(mesh, family="CG", degree=1):
    if family not in ("CG", "Lagrange", "P") or degree != 1:
        raise NotImplementedError("only continuous P1 spaces are provided by the host stand-in")
    return P1Space(mesh)


class DirichletBC(object):
    "homogeneous Dirichlet condition on a node set ('on_boundary' or an explicit index array)"

    def __init__(self, V, value, where="on_boundary"):
        if value != 0.:
            raise NotImplementedError("only homogeneous Dirichlet conditions (as in the reference's "
                                      "example and tests)")
        self.function_space = V
        if isinstance(where, str):
            assert where == "on_boundary"
            self.nodes = V.mesh().boundary_nodes()
        else:
            self.nodes = np.asarray(where, dtype=np.int32)


class Vector(object):
    """View of a Function's coefficient array with the handful of Firedrake ``Vector`` methods the
    reference and its tests call (get_local/set_local/gather/copy/size/local_size)."""

    def __init__(self, function):
        self.function = function

    @property
    def array(self):
        return self.function.data

    def get_local(self):
        return self.function.data.copy()

    def set_local(self, values):
        self.function.data[:] = np.asarray(values, dtype=np.float64)

    def gather(self):
        return self.function.data.copy()

    def size(self):
        return self.function.data.shape[0]

    def local_size(self):
        return self.function.data.shape[0]

    def copy(self):
        f = Function(self.function.function_space())
        f.data[:] = self.function.data
        return f.vector()

    def _scale(self, a):
        self.function.data *= a
        return self

    def __add__(self, other):
        out = self.copy()
        out.function.data += other.function.data
        return out

    def __sub__(self, other):
        out = self.copy()
        out.function.data -= other.function.data
        return out


class Function(object):
    "coefficient vector on a P1Space (stands in for firedrake.Function)"

    def __init__(self, V, data=None):
        self._V = V
        self.data = np.zeros(V.dim()) if data is None else np.array(data, dtype=np.float64)
        assert self.data.shape == (V.dim(),)

    def function_space(self):
        return self._V

    def vector(self):
        return Vector(self)

    def assign(self, other):
        if isinstance(other, Vector):
            other = other.function
        self.data[:] = other.data if isinstance(other, Function) else np.asarray(other, dtype=np.float64)
        return self

    def interpolate(self, fn):
        "fn(coords) -> nodal values (P1 interpolation is nodal evaluation)"
        self.data[:] = fn(self._V.coords)
        return self


class Matrix(object):
    """Assembled sparse operator in CSR with its Dirichlet conditions: rows *and* columns of constrained
    DOFs are identity (as Firedrake assembles with ``bcs=``; tests/helper_funcs.py:104-128).  ``bcs`` is the
    attribute the reference toggles around the covariance solves (solving_utils.py:46-57)."""

    def __init__(self, V, indptr, indices, data, bcs=None):
        self.function_space = V
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.data = np.ascontiguousarray(data, dtype=np.float64)
        self.bcs = bcs

    @property
    def shape(self):
        n = self.indptr.shape[0] - 1
        return (n, n)

    def bc_nodes(self):
        if self.bcs is None:
            return np.zeros(0, dtype=np.int32)
        bcs = self.bcs if isinstance(self.bcs, (list, tuple)) else [self.bcs]
        if len(bcs) == 0:
            return np.zeros(0, dtype=np.int32)
        return np.unique(np.concatenate([bc.nodes for bc in bcs])).astype(np.int32)

    def to_scipy(self):
        return sp.csr_matrix((self.data, self.indices, self.indptr), shape=self.shape)


def _p1_gradients(V):
    mesh = V.mesh()
    X = mesh.coords[mesh.cells]                              # (nc, d+1, d)
    T = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))  # columns = edge vectors
    d = T.shape[1]
    Tinv = np.linalg.inv(T)                                  # rows = grad lambda_1..d
    grads = np.concatenate([-Tinv.sum(axis=1, keepdims=True), Tinv], axis=1)  # (nc, d+1, d)
    vol = np.abs(np.linalg.det(T)) / math.factorial(d)
    return grads, vol


def assemble_stiffness(V, bcs=None, chunk=2_000_000):
    """``int grad(v).grad(u) dx`` on P1 (examples/poisson.py:27, tests/helper_funcs.py:33) as CSR.
    Structural zeros of the element pattern are kept (as PETSc does).  With ``bcs`` the constrained rows
    and columns are replaced by identity."""
    mesh = V.mesh()
    n = V.dim()
    d1 = mesh.cells.shape[1]
    A = sp.csr_matrix((n, n))
    S = sp.csr_matrix((n, n))
    for s in range(0, mesh.cells.shape[0], chunk):
        sub = Mesh(mesh.coords, mesh.cells[s:s + chunk])
        grads, vol = _p1_gradients(P1Space(sub))
        K = np.einsum("cad,cbd->cab", grads, grads) * vol[:, None, None]
        rows = np.repeat(sub.cells, d1, axis=1).ravel()
        cols = np.tile(sub.cells, (1, d1)).ravel()
        A = A + sp.coo_matrix((K.ravel(), (rows, cols)), shape=(n, n)).tocsr()
        S = S + sp.coo_matrix((np.ones(K.size), (rows, cols)), shape=(n, n)).tocsr()
    # scipy's sparse addition drops entries that sum to exactly zero; put them back on the element pattern
    A.sort_indices()
    S.sort_indices()
    keyS = np.repeat(np.arange(n, dtype=np.int64), np.diff(S.indptr)) * n + S.indices
    keyA = np.repeat(np.arange(n, dtype=np.int64), np.diff(A.indptr)) * n + A.indices
    data = np.zeros(S.nnz)
    data[np.searchsorted(keyS, keyA)] = A.data
    A = sp.csr_matrix((data, S.indices, S.indptr), shape=(n, n))
    # exact symmetry (element sums arrive in different orders for (i,j) and (j,i)); the pattern is symmetric
    At = sp.csr_matrix((np.arange(1, S.nnz + 1, dtype=np.float64), S.indices, S.indptr), shape=(n, n)).T.tocsr()
    At.sort_indices()
    perm = At.data.astype(np.int64) - 1
    A = sp.csr_matrix((0.5 * (data + data[perm]), S.indices, S.indptr), shape=(n, n))
    A.sort_indices()
    if bcs is not None:
        bc_list = bcs if isinstance(bcs, (list, tuple)) else [bcs]
        nodes = np.unique(np.concatenate([bc.nodes for bc in bc_list]))
        isbc = np.zeros(n, dtype=bool)
        isbc[nodes] = True
        rows = np.repeat(np.arange(n), np.diff(A.indptr))
        kill = isbc[rows] | isbc[A.indices]
        data = A.data.copy()
        data[kill] = 0.
        data[kill & (rows == A.indices)] = 1.
        A = sp.csr_matrix((data, A.indices, A.indptr), shape=(n, n))
        bcs = list(bc_list)
    return Matrix(V, A.indptr, A.indices, A.data, bcs)


def assemble_load(V, f):
    """``int f v dx`` with ``f`` a P1 Function (examples/poisson.py:22-29): consistent mass matrix times the
    nodal values.  Dirichlet entries are left untouched, as Firedrake's ``assemble(L)`` does."""
    mesh = V.mesh()
    n = V.dim()
    d1 = mesh.cells.shape[1]
    vol = V.cell_volumes()
    fv = (f.data if isinstance(f, Function) else np.asarray(f, dtype=np.float64))[mesh.cells]   # (nc, d+1)
    # M_e f = vol/((d+1)(d+2)) * (f_a + sum_b f_b)
    loc = (fv + fv.sum(axis=1, keepdims=True)) * (vol / (d1 * (d1 + 1)))[:, None]
    out = np.zeros(n)
    for a in range(d1):
        out += np.bincount(mesh.cells[:, a], weights=loc[:, a], minlength=n)
    return Function(V, out)


def poisson_problem(nodes_per_side, dim=2):
    """The reference example's PDE (examples/poisson.py:15-34) on a unit square/cube with
    ``nodes_per_side`` nodes per direction: returns (V, A, b)."""
    nc = nodes_per_side - 1
    if dim == 1:
        mesh = UnitIntervalMesh(nc)
    elif dim == 2:
        mesh = UnitSquareMesh(nc, nc)
    else:
        mesh = UnitCubeMesh(nc, nc, nc)
    V = FunctionSpace(mesh, "CG", 1)
    f = Function(V)
    scale = 4. * dim * np.pi * np.pi
    f.interpolate(lambda X: scale * np.prod(np.sin(2. * np.pi * X), axis=1))
    bc = DirichletBC(V, 0., "on_boundary")
    A = assemble_stiffness(V, bcs=bc)
    b = assemble_load(V, f)
    return V, A, b
