"""Oracle (test infrastructure only): interpolation matrix P, restating ``InterpolationMatrix.py``."""
import numpy as np
import scipy.sparse as sp


def interpolate_cell(data_coord, nodal_points):
    """Least-squares barycentric weights -- InterpolationMatrix.py:299-349 (lstsq at 339-343, the
    sum-to-one check with eps=1e-10 at 345-347)."""
    if nodal_points.ndim == 1:
        nodal_points = np.reshape(nodal_points, (-1, 1))
    npts, ndim = nodal_points.shape
    assert len(data_coord) == ndim, "data provided for interpolation has the wrong number of spatial dimensions"
    A = np.ones((ndim + 1, npts))
    A[0:-1, :] = np.transpose(nodal_points)
    b = np.array(list(data_coord) + [1.])
    val = np.linalg.lstsq(A, b, rcond=None)[0]
    assert np.abs(np.sum(val) - 1.) <= 1.e-10, "interpolation between data and coordinates failed"
    return val


def locate_cell(mesh_coords, cells, x, tol=1.e-12):
    """Stand-in for Firedrake's ``mesh.locate_cell`` (InterpolationMatrix.py:123; not in the reference
    tree): first cell whose barycentric coordinates of ``x`` are all >= -tol.  Brute force."""
    d = mesh_coords.shape[1]
    V = mesh_coords[cells]                       # (nc, d+1, d)
    T = np.transpose(V[:, 1:, :] - V[:, :1, :], (0, 2, 1))  # (nc, d, d)
    rhs = (x[None, :] - V[:, 0, :])
    lam = np.linalg.solve(T, rhs[:, :, None])[:, :, 0]
    lam0 = 1. - lam.sum(axis=1)
    ok = (lam.min(axis=1) >= -tol) & (lam0 >= -tol)
    idx = np.nonzero(ok)[0]
    return None if idx.size == 0 else int(idx[0])


def build_interpolation_matrix(mesh_coords, cells, sensor_coords):
    """P (n_mesh x n_data), column i = weights of the cell containing sensor i --
    InterpolationMatrix.py:101-134 (loop 122-130)."""
    mesh_coords = np.asarray(mesh_coords, dtype=np.float64)
    if mesh_coords.ndim == 1:
        mesh_coords = mesh_coords.reshape(-1, 1)
    sensor_coords = np.asarray(sensor_coords, dtype=np.float64)
    if sensor_coords.ndim == 1:
        sensor_coords = sensor_coords.reshape(-1, 1)
    n = mesh_coords.shape[0]
    rows, cols, vals = [], [], []
    for i in range(sensor_coords.shape[0]):
        cell = locate_cell(mesh_coords, cells, sensor_coords[i])
        if cell is not None:
            nodes = cells[cell]
            w = interpolate_cell(sensor_coords[i], mesh_coords[nodes])
            for node, val in zip(nodes, w):
                rows.append(node); cols.append(i); vals.append(val)
    return sp.csc_matrix((vals, (rows, cols)), shape=(n, sensor_coords.shape[0]))
