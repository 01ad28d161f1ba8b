"""Oracle (test infrastructure only): forcing covariance G, restating
``ForcingCovariance._compute_G_vals`` (ForcingCovariance.py:130-169) of the reference."""
import numpy as np
from .covariance import sqexp


def compute_G_rows(coords, int_basis, sigma, l, cutoff, regularization, rows=None, cov=sqexp):
    """The reference's row loop, statement for statement (ForcingCovariance.py:159-167):

        row = (int_basis[i]*int_basis*cov(meshvals[i], meshvals, sigma, l))[0]
        row[i] += regularization
        above_cutoff = (row/row[i] > cutoff)
        G_dict[i] = (row[above_cutoff], arange(nx, dtype=int32)[above_cutoff])

    O(n) work per row (O(n^2) in total, exactly as the reference).  Returns ``(G_dict, max_row_nnz)``
    like the reference's method (ForcingCovariance.py:169)."""
    coords = np.asarray(coords, dtype=np.float64)
    if coords.ndim == 1:
        coords = coords.reshape(-1, 1)
    int_basis = np.asarray(int_basis, dtype=np.float64)
    nx = coords.shape[0]
    G_dict = {}
    nnz = 0
    if rows is None:
        rows = range(nx)
    all_cols = np.arange(0, nx, dtype=np.int32)
    for i in rows:
        row = (int_basis[i] * int_basis * cov(coords[i], coords, sigma, l))[0]
        row[i] += regularization
        above_cutoff = (row / row[i] > cutoff)
        G_dict[i] = (row[above_cutoff], all_cols[above_cutoff])
        new_nnz = int(np.sum(above_cutoff))
        if new_nnz > nnz:
            nnz = new_nnz
    return G_dict, nnz


def _dict_to_csr(G_dict, nx):
    indptr = np.zeros(nx + 1, dtype=np.int64)
    for i in range(nx):
        indptr[i + 1] = indptr[i] + len(G_dict[i][1])
    indices = np.empty(indptr[-1], dtype=np.int32)
    vals = np.empty(indptr[-1], dtype=np.float64)
    for i in range(nx):
        indices[indptr[i]:indptr[i + 1]] = G_dict[i][1]
        vals[indptr[i]:indptr[i + 1]] = G_dict[i][0]
    return indptr, indices, vals


def compute_G_csr(coords, int_basis, sigma, l, cutoff, regularization):
    """All rows through :func:`compute_G_rows`, packed the way ``ForcingCovariance.assemble`` fills the
    PETSc AIJ matrix (ForcingCovariance.py:185-195): CSR, ascending int32 columns."""
    coords = np.asarray(coords, dtype=np.float64)
    if coords.ndim == 1:
        coords = coords.reshape(-1, 1)
    G_dict, _ = compute_G_rows(coords, int_basis, sigma, l, cutoff, regularization)
    return _dict_to_csr(G_dict, coords.shape[0])


def search_radius2(int_basis, sigma, l, cutoff, regularization):
    """Per-row squared radius outside which ``row[j]/row[i] > cutoff`` cannot hold:
    ``b_i b_j e^{2s} e^{-r2 e^{-2l}/2} / (b_i^2 e^{2s} + reg) > cutoff`` with ``b_j <= b_max``.
    Returns +inf where the test can never exclude by distance (cutoff <= 0)."""
    b = np.asarray(int_basis, dtype=np.float64)
    if cutoff <= 0.:
        return np.full(b.shape, np.inf)
    e2s = np.exp(2. * sigma)
    arg = b * b.max() * e2s / (cutoff * (b * b * e2s + regularization))
    with np.errstate(divide="ignore", invalid="ignore"):
        r2 = 2. * np.exp(2. * l) * np.log(arg)
    r2 = np.where(arg > 1., r2, 0.)
    return r2 * (1. + 1.e-9) + 1.e-300


def compute_G_csr_fast(coords, int_basis, sigma, l, cutoff, regularization, rows=None):
    """Same decisions and values as :func:`compute_G_rows` but only evaluating candidates inside the
    conservative radius of :func:`search_radius2` (cKDTree).  The arithmetic on each candidate is the
    reference's expression, evaluated elementwise by the same NumPy/SciPy calls, so pattern and values
    are identical (checked against the O(n^2) restatement in tests/test_oracle_golden.py).  Used as the
    checker at sizes where O(n^2) does not finish in seconds, and as the 'optimised CPU' baseline."""
    from scipy.spatial import cKDTree
    coords = np.asarray(coords, dtype=np.float64)
    if coords.ndim == 1:
        coords = coords.reshape(-1, 1)
    b = np.asarray(int_basis, dtype=np.float64)
    nx = coords.shape[0]
    R2 = search_radius2(b, sigma, l, cutoff, regularization)
    if rows is None:
        rows = np.arange(nx)
    rows = np.asarray(rows)
    if not np.all(np.isfinite(R2)):
        G_dict, _ = compute_G_rows(coords, b, sigma, l, cutoff, regularization, rows=rows)
        cand_lists = None
    else:
        tree = cKDTree(coords)
        cand_lists = tree.query_ball_point(coords[rows], np.sqrt(R2[rows]) * (1. + 1.e-9) + 1.e-12,
                                           return_sorted=True)
        G_dict = {}
        for i, cand in zip(rows, cand_lists):
            cand = np.asarray(cand, dtype=np.int64)
            if cand.size == 0 or i not in cand:
                cand = np.union1d(cand, [i])
            row = (b[i] * b[cand] * sqexp(coords[i], coords[cand], sigma, l))[0]
            ii = int(np.searchsorted(cand, i))
            row[ii] += regularization
            above = (row / row[ii] > cutoff)
            G_dict[int(i)] = (row[above], cand[above].astype(np.int32))
    indptr = np.zeros(len(rows) + 1, dtype=np.int64)
    for k, i in enumerate(rows):
        indptr[k + 1] = indptr[k] + len(G_dict[int(i)][1])
    indices = np.empty(indptr[-1], dtype=np.int32)
    vals = np.empty(indptr[-1], dtype=np.float64)
    for k, i in enumerate(rows):
        indices[indptr[k]:indptr[k + 1]] = G_dict[int(i)][1]
        vals[indptr[k]:indptr[k + 1]] = G_dict[int(i)][0]
    return indptr, indices, vals
