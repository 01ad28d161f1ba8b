"""Duck-typed adapter for the lower boundary of the API: what the reference receives from Firedrake --
a ``WithGeometry`` function space (ForcingCovariance.py:74-79, InterpolationMatrix.py:60-65), an assembled ``MatrixBase``
with its Dirichlet conditions (LinearSolver.py:118-119, 136-140), ``Function`` / ``Vector`` objects -- reduced to the
arrays the CUDA library consumes: DOF coordinates, cell -> node lists, lumped basis integrals, A as CSR, the Dirichlet
node set, the load vector.

Accepted without Firedrake being importable (everything is attribute-based):

* function space: ``fem.P1Space``; or any object with ``mesh()`` whose mesh exposes the vertex coordinates as
  ``coords`` (array) or ``coordinates.dat.data_ro`` (Firedrake), cells through ``cell_node_list`` (space) or ``cells``
  (mesh), and optionally the lumped integrals as ``int_basis`` / ``integrate_basis_functions()`` (else they are
  computed from the P1 cells, which is what ``assemble(Constant(1) * TestFunction(V) * dx)`` returns for P1,
  ForcingCovariance.py:126-128);
* matrix: ``fem.Matrix``; a SciPy sparse matrix; or an object carrying ``petscmat.getValuesCSR()`` (Firedrake) or a
  SciPy matrix in ``M``.  Dirichlet nodes come from ``bcs`` (objects with ``nodes``, or index arrays), ``bc_nodes``, or,
  failing both, from the identity rows of the matrix (how Firedrake leaves constrained rows, helper_funcs.py:104-128);
* function / vector: ``fem.Function`` / ``fem.Vector``; a NumPy array; or an object with ``dat.data_ro`` / ``dat.data``
  (Firedrake), ``get_local()`` / ``set_local()``, or ``vector()`` returning one of those.

Anything else raises ``TypeError`` with the reference's message, so the error conventions its tests assert are kept.
Only P1 spaces are supported: the DOFs must be the mesh vertices (one coordinate row per DOF)."""
import numpy as np
import scipy.sparse as sp
from . import fem


def _mesh_coords(mesh):
    if hasattr(mesh, "coords"):
        c = mesh.coords
    elif hasattr(mesh, "coordinates") and hasattr(mesh.coordinates, "dat"):
        dat = mesh.coordinates.dat
        c = dat.data_ro if hasattr(dat, "data_ro") else dat.data
    else:
        raise TypeError("mesh exposes neither coords nor coordinates.dat")
    c = np.asarray(c, dtype=np.float64)
    return c.reshape(-1, 1) if c.ndim == 1 else c


def as_function_space(V, message="bad input type for function_space: must be a FunctionSpace"):
    """-> fem.P1Space (the native container); the original object is kept as ``.origin``"""
    if isinstance(V, fem.P1Space):
        return V
    cached = getattr(V, "_sfem_space", None)
    if isinstance(cached, fem.P1Space):
        return cached
    try:
        mesh = V.mesh()
        coords = _mesh_coords(mesh)
        cells = getattr(V, "cell_node_list", None)
        if cells is None:
            cells = mesh.cells
        cells = np.asarray(cells)
        if cells.ndim != 2 or cells.shape[1] != coords.shape[1] + 1:
            raise TypeError("only P1 simplices are supported")
        ndof = V.dim() if callable(getattr(V, "dim", None)) else coords.shape[0]
        if int(ndof) != coords.shape[0]:
            raise TypeError("the DOFs of the space are not the mesh vertices (P1 only)")
    except TypeError:
        raise TypeError(message)
    except Exception:
        raise TypeError(message)
    space = fem.P1Space(fem.Mesh(coords, cells, getattr(V, "comm", fem.COMM_SELF) if _is_comm(getattr(V, "comm", None)) else fem.COMM_SELF))
    ib = getattr(V, "int_basis", None)
    if ib is None and callable(getattr(V, "integrate_basis_functions", None)):
        ib = V.integrate_basis_functions()
    if ib is not None:
        ib = np.asarray(ib, dtype=np.float64).reshape(-1)
        assert ib.shape[0] == coords.shape[0], "int_basis must have one entry per DOF"
        space._int_basis = ib.copy()
    space.origin = V
    try:
        V._sfem_space = space          # one native space per foreign space: G, P and A then share device uploads
    except Exception:
        pass
    return space


def _is_comm(c):
    return c is not None and hasattr(c, "rank") and hasattr(c, "size")


def _bc_nodes_from(A, n):
    bcs = getattr(A, "bcs", None)
    nodes = []
    if bcs is not None:
        for bc in (bcs if isinstance(bcs, (list, tuple)) else [bcs]):
            if hasattr(bc, "nodes"):
                nodes.append(np.asarray(bc.nodes, dtype=np.int64).reshape(-1))
            else:
                nodes.append(np.asarray(bc, dtype=np.int64).reshape(-1))
    elif getattr(A, "bc_nodes", None) is not None and not callable(A.bc_nodes):
        nodes.append(np.asarray(A.bc_nodes, dtype=np.int64).reshape(-1))
    if nodes:
        return np.unique(np.concatenate(nodes)) if len(nodes) else np.zeros(0, dtype=np.int64)
    return None


def identity_rows(csr):
    "rows (and columns) that are exactly e_i: how an assembled operator carries its Dirichlet conditions"
    n = csr.shape[0]
    counts = np.diff(csr.indptr)
    nz = csr.copy()
    nz.eliminate_zeros()
    cnt = np.diff(nz.indptr)
    diag = nz.diagonal()
    return np.nonzero((cnt == 1) & (diag == 1.))[0].astype(np.int64) if n else np.zeros(0, dtype=np.int64)


def looks_like_matrix(A):
    "is ``A`` one of the assembled-operator containers as_matrix understands?"
    return (isinstance(A, fem.Matrix) or sp.issparse(A) or hasattr(getattr(A, "petscmat", None), "getValuesCSR") or
            sp.issparse(getattr(A, "M", None)))


def as_matrix(A, V=None, message="A must be a firedrake matrix"):
    """-> fem.Matrix over the native space of ``V`` (or of ``A.function_space`` / ``A.a`` when omitted)"""
    if isinstance(A, fem.Matrix):
        return A
    csr = None
    try:
        if sp.issparse(A):
            csr = sp.csr_matrix(A)
        elif hasattr(A, "petscmat") and hasattr(A.petscmat, "getValuesCSR"):
            ip, ix, v = A.petscmat.getValuesCSR()
            n = len(ip) - 1
            csr = sp.csr_matrix((np.asarray(v, dtype=np.float64), np.asarray(ix), np.asarray(ip)), shape=(n, n))
        elif hasattr(A, "M") and sp.issparse(A.M):
            csr = sp.csr_matrix(A.M)
    except Exception:
        csr = None
    if csr is None or csr.shape[0] != csr.shape[1]:
        raise TypeError(message)
    csr.sort_indices()
    n = csr.shape[0]
    if V is None:
        V = getattr(A, "function_space", None)
        V = V() if callable(V) else V
    if V is None:
        raise TypeError(message)
    space = as_function_space(V, message)
    if space.dim() != n:
        raise TypeError(message)
    nodes = _bc_nodes_from(A, n)
    if nodes is None:
        nodes = identity_rows(csr)
    bcs = [fem.DirichletBC(space, 0., nodes.astype(np.int32))] if len(nodes) else []
    M = fem.Matrix(space, csr.indptr, csr.indices, csr.data, bcs)
    M.origin = A
    return M


def function_array(f):
    "read access to the coefficient array of a function / vector / array (a view where possible)"
    if isinstance(f, fem.Vector):
        return f.function.data
    if isinstance(f, fem.Function):
        return f.data
    if isinstance(f, np.ndarray):
        return f.reshape(-1)
    dat = getattr(f, "dat", None)
    if dat is not None and (hasattr(dat, "data_ro") or hasattr(dat, "data")):
        return np.asarray(dat.data_ro if hasattr(dat, "data_ro") else dat.data, dtype=np.float64).reshape(-1)
    if callable(getattr(f, "get_local", None)):
        return np.asarray(f.get_local(), dtype=np.float64).reshape(-1)
    if callable(getattr(f, "vector", None)):
        return function_array(f.vector())
    raise TypeError("not a function, vector or array")


def is_function_like(f):
    try:
        function_array(f)
        return not isinstance(f, (list, tuple, str, bytes, float, int))
    except TypeError:
        return False


def is_vector_like(f):
    "a Firedrake-style *Vector*: what solve_forcing_covariance / ForcingCovariance.mult / interp_mesh_to_data insist on"
    if isinstance(f, fem.Vector):
        return True
    if isinstance(f, (np.ndarray, fem.Function, list, tuple, str, bytes, float, int)):
        return False
    return callable(getattr(f, "get_local", None)) and callable(getattr(f, "set_local", None))


def assign_function(f, values):
    "write ``values`` into a function / vector in place (LinearSolver.py:305 assigns into the caller's Function)"
    values = np.asarray(values, dtype=np.float64).reshape(-1)
    if isinstance(f, fem.Vector):
        f.function.data[:] = values
    elif isinstance(f, fem.Function):
        f.data[:] = values
    elif isinstance(f, np.ndarray):
        f.reshape(-1)[:] = values
    elif callable(getattr(f, "set_local", None)):
        f.set_local(values)
    elif getattr(f, "dat", None) is not None and hasattr(f.dat, "data"):
        np.asarray(f.dat.data).reshape(-1)[:] = values
    elif callable(getattr(f, "vector", None)):
        assign_function(f.vector(), values)
    else:
        raise TypeError("cannot assign into this object")
