"""
Oracle support (test infrastructure only, used ONLY by tests/golden/make_golden.py in the build container):
a single-process NumPy/SciPy stand-in for the slice of Firedrake / petsc4py / mpi4py that stat-fem imports,
so that the reference's modules under /root/reference/stat_fem can be executed UNMODIFIED to produce golden
vectors.  The stand-in supplies only what lies *below* the reference's own code: mesh data, the lumped basis
integrals, a direct sparse solve with Firedrake's homogeneous-Dirichlet RHS lifting, and PETSc Vec/Mat/Scatter
containers.  Everything the reference itself computes (G row rule, interpolation weights, the Cu loop and its
reshape, all dense algebra, the MAP driver) runs from the reference's files.

``install(reference_root)`` registers fake modules in ``sys.modules`` and returns the imported ``stat_fem``.
"""
import contextlib
import sys
import types
import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import splu


# ------------------------------------------------------------------ communicators / mpi4py
class FakeComm(object):
    rank = 0
    size = 1

    def bcast(self, obj, root=0):
        return obj

    def barrier(self):
        pass

    def allreduce(self, x, op=None):
        return x


COMM_WORLD = FakeComm()
COMM_SELF = COMM_WORLD


# ------------------------------------------------------------------ PETSc containers
class FakeVec(object):
    def __init__(self, arr=None):
        self._arr = arr
        self._n = None if arr is None else arr.shape[0]

    def create(self, comm=None):
        return self

    def setSizes(self, sizes):
        local, glob = sizes
        self._n = glob if (local is None or local < 0) else local
        self._arr = np.zeros(self._n)

    def setFromOptions(self):
        if self._arr is None and self._n is not None:
            self._arr = np.zeros(self._n)

    def setUp(self):
        pass

    def getOwnershipRange(self):
        return 0, self._n

    def getSizes(self):
        return self._n, self._n

    def destroy(self):
        pass

    @property
    def array(self):
        return self._arr

    @array.setter
    def array(self, v):
        self._arr[:] = np.asarray(v, dtype=np.float64).ravel()


class FakeScatter(object):
    @staticmethod
    def toZero(vec):
        return FakeScatter(), FakeVec(np.zeros(vec._n))

    def scatter(self, src, dst, mode=None):
        dst._arr[:] = src._arr


class FakeMat(object):
    def create(self, comm=None):
        self._entries = {}
        self._csr = None
        return self

    def setType(self, t):
        pass

    def setSizes(self, sizes):
        (nl, _), (ml, _) = sizes
        self._shape = (nl, ml)

    def setPreallocationNNZ(self, nnz):
        pass

    def setFromOptions(self):
        pass

    def setUp(self):
        pass

    def getOwnershipRange(self):
        return 0, self._shape[0]

    def setValue(self, i, j, v):
        self._entries[(int(i), int(j))] = float(v)

    def setValues(self, rows, cols, vals):
        rows = np.atleast_1d(rows)
        vals = np.asarray(vals, dtype=np.float64).reshape(len(rows), -1)
        for a, r in enumerate(rows):
            for c, v in zip(cols, vals[a]):
                self._entries[(int(r), int(c))] = float(v)

    def assemble(self):
        if self._entries:
            keys = np.array(list(self._entries.keys()), dtype=np.int64)
            vals = np.array(list(self._entries.values()), dtype=np.float64)
            self._csr = sp.csr_matrix((vals, (keys[:, 0], keys[:, 1])), shape=self._shape)
        else:
            self._csr = sp.csr_matrix(self._shape)
        self._csr.sort_indices()

    def mult(self, x, y):
        y._arr[:] = self._csr @ x._arr

    def multTranspose(self, x, y):
        y._arr[:] = self._csr.T @ x._arr

    def getColumnVector(self, idx):
        return FakeVec(np.asarray(self._csr[:, idx].todense()).ravel().copy())

    def getValue(self, i, j):
        return self._csr[i, j]

    def destroy(self):
        pass


class _PETSc(object):
    IntType = np.int32
    Vec = FakeVec
    Mat = FakeMat
    Scatter = FakeScatter

    class ScatterMode(object):
        SCATTER_FORWARD = 0
        SCATTER_REVERSE = 1


# ------------------------------------------------------------------ Firedrake objects
class WithGeometry(object):
    pass


class FakeMesh(object):
    def __init__(self, coords, cells):
        self.coords = np.asarray(coords, dtype=np.float64).reshape(len(coords), -1)
        self.cells = np.asarray(cells)
        self.coordinates = ("coordinates-of", self)

    def cell_dimension(self):
        return self.coords.shape[1]

    def locate_cell(self, x):
        from .interp import locate_cell
        return locate_cell(self.coords, self.cells, np.asarray(x, dtype=np.float64).ravel())


class FakeSpace(WithGeometry):
    """function space built from plain arrays: coords, cells, int_basis (lumped basis integrals)"""

    def __init__(self, coords, cells, int_basis, vector=False):
        self._mesh = coords if isinstance(coords, FakeMesh) else FakeMesh(coords, cells)
        self.comm = COMM_WORLD
        self.cell_node_list = self._mesh.cells
        self.int_basis = int_basis
        self.vector = vector

    def mesh(self):
        return self._mesh

    def ufl_domain(self):
        return self._mesh

    def ufl_element(self):
        return "P1"

    def dim(self):
        return self._mesh.coords.shape[0]


class _Dat(object):
    def __init__(self, owner):
        self._o = owner

    @property
    def data(self):
        return self._o._data

    @property
    def data_with_halos(self):
        return self._o._data

    @property
    @contextlib.contextmanager
    def vec_ro(self):
        yield FakeVec(self._o._data.reshape(-1))

    @property
    @contextlib.contextmanager
    def vec(self):
        yield FakeVec(self._o._data.reshape(-1))


class Function(object):
    def __init__(self, fs, data=None):
        self._fs = fs
        self._data = np.zeros(fs.dim()) if data is None else data
        self.dat = _Dat(self)

    def function_space(self):
        return self._fs

    def vector(self):
        return Vector(self)

    def assign(self, other):
        self._data[:] = other._data
        return self


class Vector(object):
    def __init__(self, function):
        self.function = function
        self.dat = function.dat

    def size(self):
        return self.function._data.size

    def local_size(self):
        return self.function._data.size

    def get_local(self):
        return self.function._data.reshape(-1).copy()

    def set_local(self, v):
        self.function._data.reshape(-1)[:] = v

    def gather(self):
        return self.function._data.reshape(-1).copy()

    def copy(self):
        return Function(self.function._fs, self.function._data.copy()).vector()

    def _scale(self, a):
        self.function._data *= a
        return self

    def __add__(self, o):
        return Function(self.function._fs, self.function._data + o.function._data).vector()

    def __sub__(self, o):
        return Function(self.function._fs, self.function._data - o.function._data).vector()


class Constant(object):
    def __init__(self, v):
        self.v = v

    def __mul__(self, test):
        return _OneForm(test.fs, self.v, False)


class TestFunction(object):
    __test__ = False

    def __init__(self, fs):
        self.fs = fs


class _OneForm(object):
    def __init__(self, fs, scale, measured):
        self.fs, self.scale, self.measured = fs, scale, measured

    def __mul__(self, measure):
        return _OneForm(self.fs, self.scale, True)


dx = "dx"


def assemble(form):
    assert isinstance(form, _OneForm) and form.measured
    return Function(form.fs, form.scale * np.array(form.fs.int_basis, dtype=np.float64))


def VectorFunctionSpace(mesh, element):
    return FakeSpace(mesh, None, None, vector=True)


def interpolate(expr, W):
    assert isinstance(expr, tuple) and expr[0] == "coordinates-of"
    c = expr[1].coords
    data = c.copy() if c.shape[1] > 1 else c[:, 0].copy()   # Firedrake: 1-D coordinate data is a flat array
    f = Function.__new__(Function)
    f._fs, f._data, f.dat = W, data, None
    f.dat = _Dat(f)
    return f


class MatrixBase(object):
    pass


class FakeMatrix(MatrixBase):
    "assembled operator with identity Dirichlet rows/cols + the bcs attribute the reference toggles"

    def __init__(self, A, bc_nodes):
        self.M = sp.csc_matrix(A)
        self.bcs = (np.asarray(bc_nodes, dtype=np.int64),)
        self.bc_nodes = np.asarray(bc_nodes, dtype=np.int64)


class FDLinearSolver(object):
    """firedrake.linear_solver.LinearSolver: direct LU (Firedrake's default preonly+lu); ``solve`` lifts the
    RHS for homogeneous Dirichlet conditions iff ``A.bcs`` is set (solving_utils.py:46-57 relies on that)."""

    def __init__(self, A, P=None, solver_parameters=None, nullspace=None, transpose_nullspace=None,
                 near_nullspace=None, options_prefix=None):
        self.A = A
        self._lu = splu(A.M)

    def solve(self, x, b):
        xv = x.function if isinstance(x, Vector) else x
        bv = b.function if isinstance(b, Vector) else b
        rhs = bv._data.copy()
        if self.A.bcs:
            rhs[self.A.bc_nodes] = 0.
        xv._data[:] = self._lu.solve(rhs)


def install(reference_root="/root/reference"):
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    fd = mod("firedrake", COMM_WORLD=COMM_WORLD, COMM_SELF=COMM_SELF, Function=Function)
    fd.__path__ = []
    mod("firedrake.assemble", assemble=assemble)
    mod("firedrake.constant", Constant=Constant)
    mod("firedrake.function", Function=Function)
    mod("firedrake.functionspace", VectorFunctionSpace=VectorFunctionSpace)
    mod("firedrake.functionspaceimpl", WithGeometry=WithGeometry)
    mod("firedrake.interpolation", interpolate=interpolate)
    mod("firedrake.ufl_expr", TestFunction=TestFunction)
    mod("firedrake.vector", Vector=Vector)
    mod("firedrake.petsc", PETSc=_PETSc)
    mod("firedrake.linear_solver", LinearSolver=FDLinearSolver)
    mod("firedrake.matrix", MatrixBase=MatrixBase)
    mod("ufl", dx=dx)
    mpi = mod("mpi4py")
    mpi.__path__ = []
    mpi.MPI = mod("mpi4py.MPI", SUM="sum")
    pkg = types.ModuleType("stat_fem")
    pkg.__path__ = [reference_root + "/stat_fem"]
    sys.modules["stat_fem"] = pkg
    mod("stat_fem.version", version="0.3.0-reference")
    import importlib
    for name in ("covariance_functions", "ObsData", "ForcingCovariance", "InterpolationMatrix",
                 "solving_utils", "LinearSolver", "solving", "estimation", "assemble"):
        setattr(pkg, name, importlib.import_module("stat_fem." + name))
    return pkg
