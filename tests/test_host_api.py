"""CPU tests: the C-ABI library loads and exports every symbol the header declares; host-side logic (mesh generator,
argument checking, error classes, column ownership) behaves like the reference.  No compute calls."""
import os
import re
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    import stat_fem_b200  # noqa: F401
    from stat_fem_b200 import _lib
    header = open(os.path.join(ROOT, "include", "statfem_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(sfem_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 25
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), "libstatfem_b200.so does not export " + name
    assert sorted(_lib.EXPORTS) == declared
    assert lib.sfem_version() == 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import fem
    V, A, b = fem.poisson_problem(5, 2)
    G = stat_fem.ForcingCovariance(V, 0., 0.)
    with pytest.raises(RuntimeError):
        G.assemble()
    with pytest.raises(RuntimeError):
        stat_fem.covariance_functions.sqexp(np.zeros((2, 1)), np.zeros((2, 1)), 0., 0.)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "stat-fem_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_mesh_generator_matches_reference_fixture():
    "tests/helper_funcs.py: A_numpy (104-128), int basis [0.05, 0.1 x 9, 0.05] (test_ForcingCovariance.py:70-81), b_numpy (130-146)"
    from stat_fem_b200 import fem
    mesh = fem.UnitIntervalMesh(10)
    V = fem.FunctionSpace(mesh, "CG", 1)
    bc = fem.DirichletBC(V, 0., "on_boundary")
    A = fem.assemble_stiffness(V, bcs=bc).to_scipy().toarray()
    ref = np.diag(np.full(11, 20.)) + np.diag(np.full(10, -10.), 1) + np.diag(np.full(10, -10.), -1)
    ref[0, :] = ref[-1, :] = 0; ref[:, 0] = ref[:, -1] = 0; ref[0, 0] = ref[-1, -1] = 1
    np.testing.assert_allclose(A, ref, atol=1e-12)
    np.testing.assert_allclose(V.integrate_basis_functions(), [0.05] + [0.1] * 9 + [0.05], atol=1e-15)
    f = fem.Function(V).interpolate(lambda X: 4. * np.pi ** 2 * np.sin(2. * np.pi * X[:, 0]))
    b = fem.assemble_load(V, f).data
    b_ref = np.array([3.8674719419474740e-01, 2.1727588820397470e+00, 3.5155977204985351e+00, 3.5155977204985343e+00,
                      2.1727588820397470e+00, -2.7755575615628914e-16, -2.1727588820397479e+00, -3.5155977204985338e+00,
                      -3.5155977204985334e+00, -2.1727588820397470e+00, -3.8674719419474768e-01])
    np.testing.assert_allclose(b, b_ref, atol=1e-14)


@pytest.mark.parametrize("dim,nodes,nnz_row", [(2, 9, 7), (3, 6, 15)])
def test_stiffness_is_symmetric_with_full_pattern(dim, nodes, nnz_row):
    from stat_fem_b200 import fem
    V, A, b = fem.poisson_problem(nodes, dim)
    M = A.to_scipy()
    assert abs(M - M.T).max() == 0.
    assert np.diff(A.indptr).max() == nnz_row
    bcn = A.bc_nodes()
    D = M.toarray()
    assert np.all(D[bcn][:, bcn] == np.eye(len(bcn)))
    free = np.setdiff1d(np.arange(V.dim()), bcn)
    assert np.all(np.linalg.eigvalsh(D[np.ix_(free, free)]) > 0)
    # constant functions are in the kernel of the un-constrained operator
    A0 = fem.assemble_stiffness(V).to_scipy()
    assert abs(A0 @ np.ones(V.dim())).max() < 1e-12
    assert abs(V.integrate_basis_functions().sum() - 1.) < 1e-12


def test_locate_cells_partition_of_unity():
    from stat_fem_b200 import fem
    for dim, nodes in ((1, 11), (2, 7), (3, 5)):
        V, A, b = fem.poisson_problem(nodes, dim)
        rng = np.random.default_rng(dim)
        pts = rng.uniform(0, 1, (50, dim))
        cells, w = V.mesh().locate_cells(pts)
        assert np.all(cells >= 0)
        np.testing.assert_allclose(w.sum(axis=1), 1., atol=1e-12)
        assert w.min() >= -1e-12
        rec = np.einsum("pa,pad->pd", w, V.coords[V.mesh().cells[cells]])
        np.testing.assert_allclose(rec, pts, atol=1e-12)
    c, _ = V.mesh().locate_cells(np.array([[2., 2., 2.]]))
    assert c[0] == -1


def test_interpolate_cell_known_answers():
    "tests/test_InterpolationMatrix.py:255-291"
    from stat_fem_b200 import interpolate_cell
    np.testing.assert_allclose(interpolate_cell(np.array([0.5]), np.array([[0.], [1.]])), [0.5, 0.5])
    np.testing.assert_allclose(interpolate_cell(np.array([0.5, 0.5]), np.array([[0., 0.], [1., 0.], [0., 1.]])), [0., 0.5, 0.5], atol=1e-10)
    np.testing.assert_allclose(interpolate_cell(np.array([1 / 3.] * 3), np.array([[0., 0., 0.], [1., 0., 0.], [0., 1., 0.], [0., 0., 1.]])),
                               [0., 1 / 3., 1 / 3., 1 / 3.], atol=1e-9)
    with pytest.raises(AssertionError):
        interpolate_cell(np.array([0.5, 0.5]), np.eye(4)[:, :3])


def test_obsdata_container_checks():
    "tests/test_ObsData.py:7-86: shapes, getters and failure asserts"
    from stat_fem_b200 import ObsData
    od = ObsData(np.array([[1., 2.], [3., 4.], [5., 6.]]), np.array([1., 2., 3.]), 0.1)
    assert od.get_n_dim() == 2 and od.get_n_obs() == 3 and od.get_unc().shape == ()
    od = ObsData(np.array([1., 2., 3.]), np.array([1., 2., 3.]), np.array([0.1, 0.2, 0.3]))
    assert od.get_coords().shape == (3, 1) and od.get_unc().shape == (3,)
    np.testing.assert_array_equal(od.get_data(), [1., 2., 3.])
    with pytest.raises(AssertionError):
        ObsData(np.zeros((2, 2, 2)), np.zeros(2), 0.1)
    with pytest.raises(AssertionError):
        ObsData(np.zeros((3, 1)), np.zeros(2), 0.1)
    with pytest.raises(AssertionError):
        ObsData(np.zeros((3, 1)), np.zeros(3), np.array([0.1, 0.2]))
    with pytest.raises(AssertionError):
        ObsData(np.zeros((3, 1)), np.zeros(3), -0.1)


def test_constructor_type_checks_match_reference():
    "ForcingCovariance.py:74-75, 88; InterpolationMatrix.py:60-61, 68-69; LinearSolver.py:118-134; assemble.py:21-23"
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import fem
    V, A, b = fem.poisson_problem(4, 2)
    with pytest.raises(TypeError):
        stat_fem.ForcingCovariance(1., 0., 0.)
    with pytest.raises(AssertionError):
        stat_fem.ForcingCovariance(V, 0., 0., regularization=-1.)
    G = stat_fem.ForcingCovariance(V, 0., 0.)
    assert G.cutoff == 1.e-3 and G.regularization == 1.e-8 and not G.is_assembled and G.G is None
    with pytest.raises(TypeError):
        stat_fem.InterpolationMatrix(1., np.zeros((2, 2)))
    with pytest.raises(AssertionError):
        stat_fem.InterpolationMatrix(V, np.zeros((2, 3)))
    od = stat_fem.ObsData(np.array([[0.5, 0.5]]), np.array([1.]), 0.1)
    with pytest.raises(TypeError):
        stat_fem.LinearSolver(1., b, G, od)
    with pytest.raises(TypeError):
        stat_fem.LinearSolver(A, np.zeros(16), G, od)
    with pytest.raises(TypeError):
        stat_fem.LinearSolver(A, b, 1., od)
    with pytest.raises(TypeError):
        stat_fem.LinearSolver(A, b, G, 1.)
    with pytest.raises(TypeError):
        stat_fem.LinearSolver(A, b, G, od, priors=1.)
    with pytest.raises(ValueError):
        stat_fem.LinearSolver(A, b, G, od, priors=[None, None])
    with pytest.raises(TypeError):
        stat_fem.LinearSolver(A, b, G, od, priors=[1., None, None])
    with pytest.raises(TypeError):
        stat_fem.LinearSolver(A, b, G, od, ensemble_comm=1.)
    with pytest.raises(NotImplementedError):
        stat_fem.assemble(A)


def test_ownership_range_is_petsc_split():
    "solving_utils.py:117-123: contiguous split, first n % size ranks get one extra column"
    from stat_fem_b200.parallel import ownership_range
    for n in (0, 1, 4, 5, 4096, 8191):
        for size in (1, 2, 3, 4, 8):
            rs = [ownership_range(n, r, size) for r in range(size)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(size - 1))
            lens = [b - a for a, b in rs]
            assert max(lens) - min(lens) <= 1 and lens == sorted(lens, reverse=True)
    assert ownership_range(5, 0, 2) == (0, 3) and ownership_range(5, 1, 2) == (3, 5)


def test_lattice_choice():
    from stat_fem_b200.solving_utils import _lattice_nodes
    from stat_fem_b200 import fem
    for nodes, expect in ((1024, 513), (1025, 513), (130, 65), (101, 65), (33, 17), (11, 5)):
        xs = np.linspace(0, 1, nodes)
        lo, hi, m1 = _lattice_nodes(np.stack(np.meshgrid(xs, xs), -1).reshape(-1, 2)) if nodes < 200 else \
            _lattice_nodes(np.stack([np.tile(xs, nodes), np.repeat(xs, nodes)], 1))
        assert m1 == [expect, expect], (nodes, m1)


def test_adapter_reduces_firedrake_style_objects_to_arrays():
    """adapter.py on the Firedrake stand-ins of oracle/fake_firedrake.py (no GPU involved): coordinates, cells, lumped
    integrals, CSR, Dirichlet nodes, coefficient arrays; anything else is a TypeError with the reference's message"""
    import scipy.sparse as sp
    from conftest import load_case
    from stat_fem_b200 import adapter, fem
    from oracle import fake_firedrake as ff
    c = load_case("sq11_example")
    V = ff.FakeSpace(c["coords"], c["cells"], c["int_basis"])
    S = adapter.as_function_space(V)
    assert isinstance(S, fem.P1Space) and S.origin is V and adapter.as_function_space(V) is S      # cached on the object
    assert np.array_equal(S.coords, c["coords"]) and np.array_equal(S.cell_node_list, c["cells"])
    assert np.array_equal(S.integrate_basis_functions(), c["int_basis"])
    V2 = ff.FakeSpace(c["coords"], c["cells"], None)          # no lumped integrals supplied: computed from the P1 cells
    np.testing.assert_allclose(adapter.as_function_space(V2).integrate_basis_functions(), c["int_basis"], rtol=1e-12)
    n = len(c["b"])
    A = sp.csr_matrix((c["A_data"], c["A_indices"], c["A_indptr"]), shape=(n, n))
    M = adapter.as_matrix(ff.FakeMatrix(A, c["bc_nodes"]), V)
    assert isinstance(M, fem.Matrix) and np.array_equal(M.bc_nodes(), np.unique(c["bc_nodes"]))
    assert np.array_equal(M.indptr, c["A_indptr"]) and np.array_equal(M.data, c["A_data"])
    M2 = adapter.as_matrix(A, V)                                # bare SciPy matrix: Dirichlet nodes = its identity rows
    assert np.array_equal(M2.bc_nodes(), np.unique(c["bc_nodes"]))

    class PetscLike(object):                                    # the Firedrake way: A.petscmat.getValuesCSR(), A.bcs[i].nodes
        class _Mat(object):
            def getValuesCSR(self_inner):
                return A.indptr, A.indices, A.data
        petscmat = _Mat()
        bcs = [type("BC", (), {"nodes": np.asarray(c["bc_nodes"])})()]
    M3 = adapter.as_matrix(PetscLike(), V)
    assert np.array_equal(M3.bc_nodes(), np.unique(c["bc_nodes"])) and np.array_equal(M3.indices, c["A_indices"])
    f = ff.Function(V, np.arange(n, dtype=np.float64))
    assert np.array_equal(adapter.function_array(f), np.arange(n)) and np.array_equal(adapter.function_array(f.vector()), np.arange(n))
    adapter.assign_function(f, np.ones(n))
    assert np.all(f._data == 1.)
    assert adapter.is_vector_like(f.vector()) and not adapter.is_vector_like(np.ones(n)) and not adapter.is_vector_like(f)
    for bad in (object(), 3, "space", np.ones(4)):
        with pytest.raises(TypeError):
            adapter.as_function_space(bad)
    for bad in (object(), np.eye(3), "A"):
        assert not adapter.looks_like_matrix(bad)
        with pytest.raises(TypeError):
            adapter.as_matrix(bad, V)
    with pytest.raises(TypeError):
        adapter.as_matrix(sp.eye(n + 1).tocsr(), V)             # wrong size for the space


def test_fast_poisson_oracle_solver_matches_direct_lu():
    "oracle/fast_poisson.py (the CPU checker at sizes where a sparse LU does not fit) against SciPy SuperLU"
    from stat_fem_b200 import fem
    from oracle.solver import _DirectSolver
    from oracle.fast_poisson import DSTSolver
    rng = np.random.default_rng(0)
    for nodes, dim in ((33, 2), (13, 3)):
        V, A, b = fem.poisson_problem(nodes, dim)
        As = A.to_scipy()
        lu, ds = _DirectSolver(As, A.bc_nodes()), DSTSolver(As, A.bc_nodes(), nodes, dim)
        for bcs_on in (False, True):
            r = rng.normal(size=V.dim())
            x0, x1 = lu.solve(r, bcs_on), ds.solve(r, bcs_on)
            assert np.max(np.abs(x0 - x1)) <= 1e-12 * np.max(np.abs(x0))
            assert ds.residual(x1, r, bcs_on) < 1e-13


def test_newton_polish_finds_the_stationary_point_of_a_known_function():
    "estimation.newton_polish only needs logpost_deriv: a quadratic-plus-quartic bowl with a known minimiser"
    from stat_fem_b200.estimation import newton_polish

    class Bowl(object):
        x0 = np.array([0.3, -4.1, -1.2])
        H = np.array([[40., 3., 1.], [3., 900., 20.], [1., 20., 15.]])

        def logpost_deriv(self, th):
            d = np.asarray(th) - self.x0
            return self.H @ d + 50. * d ** 3
    x, info = newton_polish(Bowl(), Bowl.x0 + np.array([1e-4, -2e-5, 3e-4]))
    assert info["converged"] and np.max(np.abs(x - Bowl.x0)) < 1e-9
    far = Bowl.x0 + 5.                                            # not near a minimum of a convex bowl? still contracts or is refused
    x2, info2 = newton_polish(Bowl(), far)
    assert info2["converged"] or np.array_equal(x2, far)


def test_block_ranges_deal_slabs_evenly():
    "SparseSolver.block_ranges: contiguous cover, no block wider than the limit, never more 64-column slabs than the plain split"
    from stat_fem_b200.solving_utils import SparseSolver
    for ncols, kb in [(512, 384), (4096, 384), (2048, 384), (1024, 192), (500, 384), (385, 384), (100, 384), (70, 32), (8, 8), (449, 384)]:
        r = SparseSolver.block_ranges(ncols, kb)
        assert r[0][0] == 0 and sum(k for _, k in r) == ncols
        assert all(0 < k <= kb for _, k in r)
        assert all(r[i][0] + r[i][1] == r[i + 1][0] for i in range(len(r) - 1))
        plain = [(j0, min(kb, ncols - j0)) for j0 in range(0, ncols, kb)]
        assert len(r) == len(plain)
        assert sum(-(-k // 64) for _, k in r) <= sum(-(-k // 64) for _, k in plain)
    assert [k for _, k in SparseSolver.block_ranges(512, 384)] == [256, 256]
