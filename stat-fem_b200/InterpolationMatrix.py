"""Interpolation matrix P between the FEM mesh and the sensor locations (API of InterpolationMatrix.py).

P is tiny (d+1 non-zeros per sensor) and is built on the host from the mesh cells; it is then kept in HBM both as
CSC (column k = sensor k, used to densify right-hand-side blocks and as the CSR of P^T) and as CSR (for P d)."""
import numpy as np
import scipy.sparse as sp
from . import _lib
from . import fem
from . import adapter


def interpolate_cell(data_coord, nodal_points):
    """Weights of the cell's nodal basis functions at ``data_coord`` (InterpolationMatrix.py:299-349): the
    least-squares solution of  [X^T; 1] w = [x; 1],  which for P1 simplices is the barycentric coordinate vector."""
    nodal_points = np.asarray(nodal_points, dtype=np.float64)
    if nodal_points.ndim == 1:
        nodal_points = nodal_points.reshape(-1, 1)
    npts, ndim = nodal_points.shape
    assert len(data_coord) == ndim, "data provided for interpolation has the wrong number of spatial dimensions"
    lhs = np.vstack([nodal_points.T, np.ones((1, npts))])
    rhs = np.append(np.asarray(data_coord, dtype=np.float64), 1.)
    val = np.linalg.lstsq(lhs, rhs, rcond=None)[0]
    assert abs(val.sum() - 1.) <= 1.e-10, "interpolation between data and coordinates failed"
    return val


class InterpolationMatrix(object):
    def __init__(self, function_space, coords):
        function_space = adapter.as_function_space(function_space)
        self.coords = np.copy(coords)
        self.function_space = function_space
        self.comm = function_space.comm
        self.n_data = coords.shape[0]
        assert coords.shape[1] == function_space.mesh().cell_dimension(), \
            "shape of coordinates does not match mesh dimension"
        self.n_data_local = self.n_data
        self.n_mesh = function_space.dim()
        self.n_mesh_local = self.n_mesh
        self.meshspace_vector = fem.Function(function_space).vector()
        self.interp = None          # scipy CSC on the host (n_mesh x n_data)
        self.csc = None             # device CSC  (== CSR of P^T)
        self.csr = None             # device CSR
        self.is_assembled = False

    def assemble(self):
        "locate every sensor and store the barycentric weights of its cell (InterpolationMatrix.py:101-134)"
        if self.is_assembled:
            return
        mesh = self.function_space.mesh()
        cells, weights = mesh.locate_cells(self.coords)
        nodes = self.function_space.cell_node_list
        inside = np.nonzero(cells >= 0)[0]
        rows = nodes[cells[inside]].ravel()
        cols = np.repeat(inside, nodes.shape[1])
        vals = weights[inside].ravel()
        P = sp.csc_matrix((vals, (rows, cols)), shape=(self.n_mesh, self.n_data))
        P.sort_indices()
        self.interp = P
        ctx = _lib.get_context()
        self.csc = _lib.DeviceCSR(ctx, (self.n_data, self.n_mesh), P.indptr, P.indices, P.data)
        Pr = P.tocsr()
        Pr.sort_indices()
        self.csr = _lib.DeviceCSR(ctx, (self.n_mesh, self.n_data), Pr.indptr, Pr.indices, Pr.data)
        self.is_assembled = True

    def _gather(self):
        """InterpolationMatrix.py:136-148 gathers the distributed data-space PETSc Vec on rank 0.  Here every process
        holds the whole mesh (the parallel axis is the ensemble of sensor columns, not the mesh), so the data-space
        vector already lives where it is needed: nothing to move."""

    def _scatter(self):
        "InterpolationMatrix.py:150-162 (root -> distributed data-space Vec): nothing to move, see _gather"

    def destroy(self):
        self.interp = self.csc = self.csr = None
        self.is_assembled = False

    def interp_data_to_mesh(self, data_array):
        "P d as a mesh-space Vector (InterpolationMatrix.py:177-221)"
        if not self.is_assembled:
            self.assemble()
        data_array = np.array(data_array)
        assert data_array.ndim == 1, "input data must be a 1D array of sensor values"
        assert data_array.shape[0] == self.n_data, "data_array has an incorrect number of sensor measurements"
        ctx = self.csr.ctx
        out = self.csr.matmul(ctx.to_device(data_array, np.float64))
        f = fem.Function(self.function_space, out.cpu().numpy())
        return f.vector()

    def interp_mesh_to_data(self, input_mesh_vector):
        "P^T v as a NumPy array (InterpolationMatrix.py:223-259)"
        if not self.is_assembled:
            self.assemble()
        if not adapter.is_vector_like(input_mesh_vector):
            raise TypeError("input_mesh_vector must be a firedrake vector")
        assert adapter.function_array(input_mesh_vector).shape[0] == self.n_mesh, \
            "input vector must be the same length as the FEM DOFs and be distributed across processes in the same way"
        ctx = self.csc.ctx
        return self.csc.matmul(ctx.to_device(adapter.function_array(input_mesh_vector), np.float64)).cpu().numpy()

    def get_meshspace_column_vector(self, idx):
        "column idx of P as a mesh-space Vector (InterpolationMatrix.py:261-287)"
        assert idx >= 0 and idx < self.n_data, "idx out of range"
        if not self.is_assembled:
            self.assemble()
        f = fem.Function(self.function_space)
        f.data[:] = np.asarray(self.interp[:, idx].todense()).ravel()
        return f.vector()

    def __str__(self):
        return "Interpolation matrix from {} mesh points to {} data points".format(self.n_mesh, self.n_data)
