"""Oracle (test infrastructure only): exact CPU solver for the benchmark stiffness matrices at sizes where a sparse
direct LU no longer fits (3-D, 2 097 152 DOF: the fill-in of ``splu`` is tens of GB).

The benchmark problems (``fem.poisson_problem``, following examples/poisson.py:15-34 of the reference) are P1 on the
uniform unit square / cube with the Kuhn split; their stiffness matrix is the (2d+1)-point Laplacian scaled by
``h^(d-2)`` with homogeneous Dirichlet rows *and columns* replaced by identity.  That operator is diagonalised by the
type-I discrete sine transform, so ``A^-1`` is two DSTs and a division -- the numerical class of a direct solve
(error ~ cond(A) * eps).  It stands where ``oracle.solver._DirectSolver`` (SciPy SuperLU, the class of the
reference's Firedrake default) stands at smaller sizes, is checked against it in tests/test_oracle_golden.py, and every
solve can be verified against the *assembled* matrix with :meth:`residual` so that no property of the stencil is taken
on trust."""
import numpy as np
from scipy.fft import dstn


class DSTSolver(object):
    """Drop-in for ``_DirectSolver`` on ``fem.poisson_problem(nodes_per_side, dim)``: ``solve(rhs, bcs_on)``."""

    def __init__(self, A, bc_nodes, nodes_per_side, dim):
        self.A = A.tocsr()
        self.bc_nodes = np.asarray(bc_nodes, dtype=np.int64)
        self.N, self.d = int(nodes_per_side), int(dim)
        assert self.A.shape[0] == self.N ** self.d
        m = self.N - 2
        h = 1. / (self.N - 1)
        lam1 = 2. - 2. * np.cos(np.arange(1, m + 1) * np.pi / (m + 1))
        lam = np.zeros((m,) * self.d)
        for k in range(self.d):
            shape = [1] * self.d
            shape[k] = m
            lam = lam + lam1.reshape(shape)
        self.lam = lam * h ** (self.d - 2)
        self._interior = (slice(1, -1),) * self.d

    def solve(self, rhs, bcs_on):
        rhs = np.array(rhs, dtype=np.float64, copy=True)
        if bcs_on and self.bc_nodes.size:
            rhs[self.bc_nodes] = 0.
        full = rhs.reshape((self.N,) * self.d)
        out = full.copy()                                   # Dirichlet rows are identity: x_bc = rhs_bc
        inner = dstn(full[self._interior], type=1, norm="ortho")
        inner /= self.lam
        out[self._interior] = dstn(inner, type=1, norm="ortho")
        return out.reshape(-1)

    def residual(self, x, rhs, bcs_on=False):
        "|| rhs - A x || / || rhs || against the assembled matrix"
        rhs = np.array(rhs, dtype=np.float64, copy=True)
        if bcs_on and self.bc_nodes.size:
            rhs[self.bc_nodes] = 0.
        return float(np.linalg.norm(rhs - self.A @ x) / np.linalg.norm(rhs))
