"""GPU parity tests of the individual kernels, called through the C ABI (ctypes) and checked against NumPy/SciPy and
the oracle.  Sizes are chosen so the checkers finish in seconds."""
import ctypes as C
import numpy as np
import scipy.sparse as sp
import scipy.linalg as sla
import pytest

from conftest import load_case, relerr, CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available()
    import stat_fem_b200  # noqa: F401
    from stat_fem_b200 import _lib
    ctx = _lib.get_context()
    return _lib, ctx, _lib.lib()


def dev(ctx, a, dt=np.float64):
    return ctx.to_device(a, dt)


@pytest.mark.parametrize("opA,opB", [(1, 0), (0, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (100, 36, 1000), (257, 130, 77), (5, 3, 4099), (64, 200, 20011), (300, 300, 300),
                                   (2500, 2300, 130), (1111, 97, 128)])
def test_dgemm(env, opA, opB, M, N, K):
    _lib, ctx, lib = env
    rng = np.random.default_rng(M * 7 + N * 3 + K + opA * 2 + opB)
    A = rng.normal(size=(K, M) if opA else (M, K))
    B = rng.normal(size=(N, K) if opB else (K, N))
    Cin = rng.normal(size=(M, N))
    ref = 1.5 * (A.T if opA else A) @ (B.T if opB else B) - 0.5 * Cin
    dA, dB, dC = dev(ctx, A), dev(ctx, B), dev(ctx, Cin)
    ctx.check(lib.sfem_dgemm(ctx.handle, opA, opB, M, N, K, 1.5, _lib.ptr(dA), A.shape[1], _lib.ptr(dB), B.shape[1], -0.5,
                             _lib.ptr(dC), N))
    got = dC.cpu().numpy()
    assert relerr(got, ref) < 1e-13 * max(1, np.sqrt(K))


def test_dgemm_strided_and_beta0(env):
    "leading dimensions larger than the logical width, misaligned (odd) ld, beta = 0 with NaN-filled C"
    _lib, ctx, lib = env
    rng = np.random.default_rng(3)
    M, N, K = 70, 45, 513
    for lda, ldb, ldc in [(M + 6, N + 2, N + 4), (M + 3, N + 1, N + 5)]:
        A = rng.normal(size=(K, lda)); B = rng.normal(size=(K, ldb))
        Cb = np.full((M, ldc), np.nan)
        dA, dB, dC = dev(ctx, A), dev(ctx, B), dev(ctx, Cb)
        ctx.check(lib.sfem_dgemm(ctx.handle, 1, 0, M, N, K, 1.0, _lib.ptr(dA), lda, _lib.ptr(dB), ldb, 0.0, _lib.ptr(dC), ldc))
        got = dC.cpu().numpy()
        assert relerr(got[:, :N], A[:, :M].T @ B[:, :N]) < 1e-13
        assert np.all(np.isnan(got[:, N:]))


@pytest.mark.parametrize("n", [1, 7, 100, 128, 129, 257, 300, 384, 513, 640, 1000, 1537])
def test_cholesky_solve_inverse_logdet(env, n):
    _lib, ctx, lib = env
    rng = np.random.default_rng(n)
    X = rng.normal(size=(n, n + 5))
    S = X @ X.T + n * 0.1 * np.eye(n)
    dS = dev(ctx, S)
    dinv = ctx.empty((lib.sfem_potrf_dinv_doubles(n),))
    info = C.c_int(-1)
    ctx.check(lib.sfem_potrf(ctx.handle, n, _lib.ptr(dS), n, _lib.ptr(dinv), C.byref(info)))
    assert info.value == 0
    L = np.tril(dS.cpu().numpy())
    Lref = np.linalg.cholesky(S)
    assert relerr(L, Lref) < 1e-12
    ld = C.c_double(0.)
    ctx.check(lib.sfem_logdet(ctx.handle, n, _lib.ptr(dS), n, C.byref(ld)))
    assert abs(ld.value - np.linalg.slogdet(S)[1]) < 1e-10 * max(1., abs(ld.value))
    for k in (1, 3, n):
        B = rng.normal(size=(n, k))
        dB = dev(ctx, B)
        ctx.check(lib.sfem_potrs(ctx.handle, n, _lib.ptr(dS), n, _lib.ptr(dinv), k, _lib.ptr(dB), k))
        assert relerr(dB.cpu().numpy(), sla.cho_solve((Lref, True), B)) < 1e-10
    W, Kinv = ctx.empty((n, n)), ctx.empty((n, n))
    ctx.check(lib.sfem_potri(ctx.handle, n, _lib.ptr(dS), n, _lib.ptr(dinv), _lib.ptr(W), _lib.ptr(Kinv)))
    assert relerr(Kinv.cpu().numpy(), np.linalg.inv(S)) < 1e-10


def test_cholesky_reports_non_spd(env):
    from scipy.linalg import LinAlgError
    _lib, ctx, lib = env
    n = 200
    S = np.eye(n); S[150, 150] = -1.
    dS = dev(ctx, S)
    dinv = ctx.empty((lib.sfem_potrf_dinv_doubles(n),))
    info = C.c_int(0)
    with pytest.raises(LinAlgError):
        ctx.check(lib.sfem_potrf(ctx.handle, n, _lib.ptr(dS), n, _lib.ptr(dinv), C.byref(info)))
    assert info.value == 151


@pytest.mark.parametrize("k", [1, 2, 7, 64, 100])
def test_spmm_csr(env, k):
    _lib, ctx, lib = env
    rng = np.random.default_rng(k)
    S = sp.random(500, 300, density=0.03, random_state=1, format="csr") + sp.eye(500, 300, format="csr")
    S = sp.csr_matrix(S); S.sort_indices()
    X = rng.normal(size=(300, k))
    dS = _lib.DeviceCSR(ctx, S.shape, S.indptr, S.indices, S.data)
    Y = dS.matmul(dev(ctx, X))
    assert relerr(Y.cpu().numpy(), S @ X) < 1e-14


def test_scatter_columns(env):
    _lib, ctx, lib = env
    P = sp.random(400, 50, density=0.02, random_state=2, format="csc"); P.sort_indices()
    dP = _lib.DeviceCSR(ctx, (50, 400), P.indptr, P.indices, P.data)
    B = ctx.empty((400, 20))
    B.fill_(7.)
    ctx.check(lib.sfem_scatter_columns(ctx.handle, 400, _lib.ptr(dP.indptr), _lib.ptr(dP.indices), _lib.ptr(dP.data), 13, 17,
                                       _lib.ptr(B), 20))
    got = B.cpu().numpy()
    assert np.array_equal(got[:, :17], P[:, 13:30].toarray())
    assert np.all(got[:, 17:] == 7.)


def _assemble_G(stat_fem, fem, coords, cells, sigma, l, cutoff, reg):
    V = fem.P1Space(fem.Mesh(coords, cells))
    G = stat_fem.ForcingCovariance(V, sigma, l, cutoff, reg)
    G.assemble()
    return G, V


@pytest.mark.parametrize("name", CASES)
def test_gcov_matches_reference_golden(env, name):
    "pattern (indptr, int32 ascending indices) array-equal to the unmodified reference's, values to 1e-13"
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import fem
    c = load_case(name)
    G, V = _assemble_G(stat_fem, fem, c["coords"], c["cells"], float(c["sigma_f"]), float(c["l_f"]), float(c["cutoff"]), float(c["reg"]))
    ip, ix, v = G.G.to_host()
    assert np.array_equal(ip, c["G_indptr"])
    assert ix.dtype == np.int32 and np.array_equal(ix, c["G_indices"])
    np.testing.assert_allclose(v, c["G_vals"], rtol=1e-13, atol=0.)
    assert G.max_row_nnz == int(c["G_max_nnz"])
    Gd, mx = G._compute_G_vals()
    assert mx == int(c["G_max_nnz"]) and len(Gd) == V.dim()
    assert np.allclose(V.integrate_basis_functions(), c["int_basis"])


@pytest.mark.parametrize("dim,nodes,lfac", [(2, 129, 1.5), (2, 65, 2.63), (3, 21, 0.8), (1, 500, 3.0)])
def test_gcov_matches_oracle_larger(env, dim, nodes, lfac):
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import fem
    import oracle
    V, A, b = fem.poisson_problem(nodes, dim)
    h = 1. / (nodes - 1)
    sig = np.log(2.e-2)
    reg = 1.e-8 * h ** (2 * dim) * np.exp(2 * sig)
    G = stat_fem.ForcingCovariance(V, sig, np.log(lfac * h), 1.e-3, reg)
    G.assemble()
    ip, ix, v = G.G.to_host()
    oip, oix, ov = oracle.compute_G_csr_fast(V.coords, V.integrate_basis_functions(), sig, np.log(lfac * h), 1.e-3, reg)
    assert np.array_equal(ip, oip) and np.array_equal(ix, oix)
    np.testing.assert_allclose(v, ov, rtol=1e-13, atol=0.)


def test_gcov_example_defaults_are_diagonal(env):
    "with the shipped example's parameters every row keeps only its diagonal (SURVEY finding 3)"
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import fem
    import oracle
    V, A, b = fem.poisson_problem(101, 2)
    G = stat_fem.ForcingCovariance(V, np.log(2.e-2), np.log(0.354))
    G.assemble()
    ip, ix, v = G.G.to_host()
    assert np.array_equal(ip, np.arange(V.dim() + 1)) and np.array_equal(ix, np.arange(V.dim()))
    rows = [0, 50, 5050, 5100, 10200]
    Gd, _ = oracle.compute_G_rows(V.coords, V.integrate_basis_functions(), np.log(2.e-2), np.log(0.354), 1.e-3, 1.e-8, rows=rows)
    for r in rows:
        assert np.array_equal(Gd[r][1], [r]) and np.isclose(v[r], Gd[r][0][0], rtol=1e-15)


def test_gcov_long_rows_sorted(env):
    "rows longer than the in-shared-memory capacity (512) go through the CTA sort"
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import fem
    import oracle
    V, A, b = fem.poisson_problem(65, 2)
    h = 1. / 64
    sig = np.log(2.e-2)
    reg = 1.e-8 * h ** 4 * np.exp(2 * sig)
    G = stat_fem.ForcingCovariance(V, sig, np.log(4.5 * h), 1.e-3, reg)
    G.assemble()
    assert G.max_row_nnz > 512
    ip, ix, v = G.G.to_host()
    oip, oix, ov = oracle.compute_G_csr_fast(V.coords, V.integrate_basis_functions(), sig, np.log(4.5 * h), 1.e-3, reg)
    assert np.array_equal(ip, oip) and np.array_equal(ix, oix)
    np.testing.assert_allclose(v, ov, rtol=1e-13, atol=0.)


@pytest.mark.parametrize("shuffle", [False, True])
def test_gcov_bitmap_rank_and_bitonic_sort_agree(env, shuffle, monkeypatch):
    """The fill kernel orders a row's entries by bitmap rank when the widest column span fits its per-warp bitmap
    (structured numbering) and by the bitonic sort otherwise (here: randomly renumbered nodes, span ~ n); both orders,
    and the forced bitonic path, give the oracle's CSR."""
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import fem
    import oracle
    V0, A, b = fem.poisson_problem(300 if shuffle else 129, 2)
    coords, cells = V0.mesh().coords, V0.mesh().cells
    if shuffle:
        perm = np.random.default_rng(5).permutation(coords.shape[0])
        inv = np.empty_like(perm)
        inv[perm] = np.arange(perm.size)
        coords, cells = coords[perm], inv[cells]
    h = 1. / (np.sqrt(coords.shape[0]) - 1)
    sig, l = np.log(2.e-2), np.log((1.5 if shuffle else 2.2) * h)
    reg = 1.e-8 * h ** 4 * np.exp(2 * sig)
    out = []
    for force_bitonic in (False, True):
        if force_bitonic:
            monkeypatch.setenv("SFEM_GCOV_BITONIC", "1")
        else:
            monkeypatch.delenv("SFEM_GCOV_BITONIC", raising=False)
        G, V = _assemble_G(stat_fem, fem, coords, cells, sig, l, 1.e-3, reg)
        out.append(G.G.to_host())
    oip, oix, ov = oracle.compute_G_csr_fast(V.coords, V.integrate_basis_functions(), sig, l, 1.e-3, reg)
    for ip, ix, v in out:
        assert np.array_equal(ip, oip) and np.array_equal(ix, oix)
        np.testing.assert_allclose(v, ov, rtol=1e-13, atol=0.)
    assert np.array_equal(out[0][2], out[1][2])


@pytest.mark.parametrize("dim,nodes,pc", [(1, 11, "mg"), (2, 33, "jacobi"), (2, 33, "mg"), (2, 130, "mg"), (2, 257, "mg"), (3, 18, "mg"), (3, 33, "mg")])
def test_block_solve_matches_direct(env, dim, nodes, pc):
    "Z = A^-1 P for random sensors against SuperLU, every column to 1e-9 relative"
    from stat_fem_b200 import fem
    from stat_fem_b200.solving_utils import SparseSolver
    from stat_fem_b200.InterpolationMatrix import InterpolationMatrix
    from scipy.sparse.linalg import splu
    V, A, b = fem.poisson_problem(nodes, dim)
    rng = np.random.default_rng(nodes)
    ncol = 37
    s = rng.uniform(0.0, 1.0, (ncol, dim))
    im = InterpolationMatrix(V, s)
    im.assemble()
    ls = SparseSolver(A, solver_parameters={"pc_type": pc, "block_size": 16})
    Z = ls.ctx.empty((V.dim(), ncol))
    ls.solve_columns(im.csc, 0, ncol, Z)
    ref = splu(sp.csc_matrix(A.to_scipy())).solve(im.interp.toarray())
    got = Z.cpu().numpy()
    err = np.max(np.abs(got - ref), axis=0) / np.max(np.abs(ref), axis=0)
    print("dim=%d nodes=%d pc=%s levels=%s iterations=%s max relres=%.2e max err=%.2e" %
          (dim, nodes, pc, getattr(ls, "mg_levels", None), ls.block_iterations, ls.last_relres.max(), err.max()))
    assert err.max() < 1e-9
    assert ls.last_relres.max() < 1e-10
    if pc == "mg":
        assert max(ls.block_iterations) <= 60


@pytest.mark.parametrize("dim,nodes,lfac,k", [(2, 65, 1.5, 256), (2, 50, 2.2, 72), (3, 14, 0.8, 130), (1, 203, 3.0, 64)])
def test_block_form_spmm(env, dim, nodes, lfac, k):
    """W = G Z through the 8x4-block (DMMA) form of csrc/bsr.cu against SciPy's CSR product and against the scalar CSR
    kernel: ragged row groups (n not a multiple of 8), column counts that are not multiples of 64, 1-D / 2-D / 3-D G."""
    import torch
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import fem
    _lib, ctx, lib = env
    V, A, b = fem.poisson_problem(nodes, dim)
    h = 1. / (nodes - 1)
    sig = np.log(2.e-2)
    G = stat_fem.ForcingCovariance(V, sig, np.log(lfac * h), 1.e-3, 1.e-8 * h ** (2 * dim) * np.exp(2 * sig))
    G.assemble()
    ip, ix, v = G.G.to_host()
    n = V.dim()
    S = sp.csr_matrix((v, ix, ip), shape=(n, n))
    rng = np.random.default_rng(k)
    X = rng.normal(size=(n, k))
    Xd = G.G.ctx.to_device(X, np.float64)
    Y = torch.full((n, k), float("nan"), dtype=torch.float64, device="cuda")
    G.G.ctx.check(lib.sfem_b84_spmm(G.G.ctx.handle, G.G.block_form(), k, _lib.ptr(Xd), k, _lib.ptr(Y), k))
    ref = S @ X
    assert G.G.block_tiles * 32 >= S.nnz and G.G.block_tiles > 0
    assert relerr(Y.cpu().numpy(), ref) < 1e-13
    was = _lib.USE_BLOCK_FORM[0]
    try:
        _lib.USE_BLOCK_FORM[0] = False
        Y0 = G.G.matmul(Xd).cpu().numpy()          # scalar CSR kernel
    finally:
        _lib.USE_BLOCK_FORM[0] = was
    assert relerr(Y.cpu().numpy(), Y0) < 1e-13
    print("n=%d nnz/row=%.1f block fill %.2f" % (n, S.nnz / n, S.nnz / (32. * G.G.block_tiles)))


# modes of the staged-tile kernels (csrc/tileop.cuh): name -> (code, needs right-hand side, returns dot partials)
TILE_MODES = {"mult": (0, False, False), "mult_add": (1, True, False), "mult_dot": (2, False, True), "jacobi": (3, True, False),
              "jacobi_dot": (4, True, True), "resid": (5, True, False), "presmooth2": (6, False, False)}


@pytest.mark.parametrize("nodes,dim,k", [(129, 2, 96), (200, 2, 70), (65, 2, 384), (65, 2, 448), (33, 3, 192), (33, 3, 70)])
def test_tile_tensor_form_matches_scalar_form(env, nodes, dim, k):
    """Every operator of the multigrid hierarchy that has an FP64 tensor-core (DMMA) form gives the same block result
    and the same dot products through tileop_dmma_kernel -- in both of its shapes: 64-column slabs (2-D meshes aligned
    with the lattice) and 32-column slabs with up to 128 staged rows per chunk (3-D, non-aligned 2-D; forced on every
    operator here) -- as through the scalar tileop_kernel: ragged column slabs, chunks of fewer than 32 rows, several
    chunks per CTA, dot products accumulated in registers (at most 6 slabs) and through shared memory (more)."""
    import torch
    from stat_fem_b200 import fem
    from stat_fem_b200.solving_utils import SparseSolver
    _lib, _, lib = env
    V, A, b = fem.poisson_problem(nodes, dim)
    ls = SparseSolver(A)
    ctx = ls.ctx
    gen = torch.Generator(device="cuda").manual_seed(nodes + k)
    checked = 0
    fine_has_form = False
    restriction_has_form = False
    for level in range(len(ls.mg_levels)):
        for which, modes in ((0, ("mult_dot", "jacobi", "jacobi_dot", "resid", "presmooth2", "mult")), (1, ("mult_add",)), (2, ("mult",))):
            rows, cols, form = C.c_int64(0), C.c_int64(0), C.c_int(0)
            ctx.check(lib.sfem_tile_apply_one(ctx.handle, level, which, 0, 0, None, None, None, None, None, 0, C.byref(rows),
                                              C.byref(cols), C.byref(form)))
            if form.value == 0 or rows.value == 0:
                continue
            fine_has_form |= (level == 0 and which == 0)
            restriction_has_form |= (level == 0 and which == 2)
            X = torch.rand((cols.value, k), generator=gen, dtype=torch.float64, device="cuda") - 0.5
            B = torch.rand((rows.value, k), generator=gen, dtype=torch.float64, device="cuda") - 0.5
            for name in modes:
                code, has_b, has_dot = TILE_MODES[name]
                out = []
                for tensor_form in (0, 1, 2):       # scalar kernel, tensor-core kernel (automatic shape), narrow shape forced
                    Y = B.clone() if name == "mult_add" else torch.full((rows.value, k), float("nan"), dtype=torch.float64, device="cuda")
                    parts = torch.zeros((1184, k), dtype=torch.float64, device="cuda")
                    npart = C.c_int(0)
                    ctx.check(lib.sfem_tile_apply_one(ctx.handle, level, which, code, k, _lib.ptr(X), _lib.ptr(B), _lib.ptr(Y),
                                                      _lib.ptr(parts), C.byref(npart), tensor_form, None, None, None))
                    out.append((Y.cpu().numpy(), parts[:npart.value].sum(dim=0).cpu().numpy()))
                y0, d0 = out[0]
                for y1, d1 in out[1:]:
                    assert np.all(np.isfinite(y0)) and np.all(np.isfinite(y1)), (level, which, name)
                    assert relerr(y1, y0) < 1e-13, (level, which, name, relerr(y1, y0))
                    if has_dot:
                        assert relerr(d1, d0) < 1e-12, (level, which, name, relerr(d1, d0))
                checked += 1
    assert fine_has_form, "the stiffness operator has no tensor-core form on this mesh"
    if dim == 2 and nodes in (65, 129):
        assert restriction_has_form, "the 2-D restriction (153 staged rows per 32-row chunk) has no tensor-core form"
    assert checked >= 6


def test_covariance_functions_and_obsdata(env):
    import stat_fem_b200 as stat_fem
    cf = stat_fem.covariance_functions
    z = load_case("covariance")
    for d in (1, 2, 3):
        x1, x2 = z[f"x1_{d}"], z[f"x2_{d}"]
        assert np.array_equal(cf.calc_r2(x1, x2), z[f"r2_{d}"])
        np.testing.assert_allclose(cf.sqexp(x1, x2, np.log(0.7), np.log(0.3)), z[f"sqexp_{d}"], rtol=1e-14)
        np.testing.assert_allclose(cf.sqexp_deriv(x1, x2, np.log(0.7), np.log(0.3)), z[f"deriv_{d}"], rtol=1e-14, atol=1e-300)
        od = stat_fem.ObsData(x1, np.zeros(9), z[f"unc_{d}"])
        np.testing.assert_allclose(od.calc_K(np.array([-1.1, -0.4])), z[f"K_{d}"], rtol=1e-14)
        np.testing.assert_allclose(od.calc_K_plus_sigma(np.array([-1.1, -0.4])), z[f"Ks_{d}"], rtol=1e-14)
        np.testing.assert_allclose(od.calc_K_deriv(np.array([-1.1, -0.4])), z[f"Kd_{d}"], rtol=1e-14, atol=1e-300)
    # 1-D point promotion (covariance_functions.py: atleast_2d)
    assert cf.sqexp(np.array([1., 2.]), np.array([[1., 2.], [2., 4.]]), 0., 0.).shape == (1, 2)
    h = cf.sqexp_hessian(z["x1_2"], z["x2_2"], np.log(0.7), np.log(0.3))
    assert h.shape == (2, 2, 9, 13) and np.allclose(h[0, 0], 4. * z["sqexp_2"])


def test_prior_cov_entry_point_single_rank():
    """sfem_prior_cov (the sharded covariance assembly of the C ABI, include/statfem_b200.h) with one rank -- no
    communicator needed -- against LinearSolver.solve_prior; the multi-rank exchange is exercised by
    tools/cabi_sharded_check.py on a multi-GPU box (profiles/r02_cabi_sharded.log)."""
    import ctypes as C
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import fem, _lib
    from stat_fem_b200.solving_utils import SparseSolver
    lib = _lib.lib()
    V, A, b = fem.poisson_problem(65, 2)
    h = 1. / 64
    sig, l = np.log(2.e-2), np.log(1.5 * h)
    reg = 1.e-8 * h ** 4 * np.exp(2 * sig)
    for nobs in (40, 300):           # CSR kernel / 8x4-block form of W = G Z
        sensors = np.random.default_rng(1).uniform(0.02, 0.98, (nobs, 2))
        G = stat_fem.ForcingCovariance(V, sig, l, 1.e-3, reg)
        G.assemble()
        sp_ = SparseSolver(A)
        ctx = sp_.ctx
        im = stat_fem.InterpolationMatrix(V, sensors)
        im.assemble()
        ctx.check(lib.sfem_comm_init(ctx.handle, None, 0, 1))
        Cu = ctx.empty((nobs, nobs))
        Z = ctx.empty((V.dim(), nobs))
        iters = C.c_int(0)
        ctx.check(lib.sfem_prior_cov(ctx.handle, nobs, _lib.ptr(im.csc.indptr), _lib.ptr(im.csc.indices), _lib.ptr(im.csc.data),
                                     _lib.ptr(G.G.indptr), _lib.ptr(G.G.indices), _lib.ptr(G.G.data), 0, nobs, 128, 1, 1.e-12, 0.,
                                     500, _lib.ptr(Z), None, _lib.ptr(Cu), C.byref(iters)))
        od = stat_fem.ObsData(sensors, np.zeros(nobs), 2.e-3)
        ls = stat_fem.LinearSolver(A, b, G, od)
        ref = ls.solve_prior()[1]
        assert relerr(Cu.cpu().numpy(), ref) < 1e-10 and 0 < iters.value < 30
        # bad column range: reported, not silently accepted
        rc = lib.sfem_prior_cov(ctx.handle, nobs, _lib.ptr(im.csc.indptr), _lib.ptr(im.csc.indices), _lib.ptr(im.csc.data),
                                _lib.ptr(G.G.indptr), _lib.ptr(G.G.indices), _lib.ptr(G.G.data), 1, nobs, 128, 1, 1.e-12, 0., 500,
                                None, None, _lib.ptr(Cu), None)
        assert rc == _lib.SFEM_ERR_ARG
