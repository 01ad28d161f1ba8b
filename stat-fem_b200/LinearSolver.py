"""LinearSolver: all stat-fem solves for one FEM model + one data set (API of LinearSolver.py in the reference).

Caches the prior solution x, its interpolation mu and the interpolated prior covariance Cu (LinearSolver.py:150-153,
210-222) -- here on the device as well as on the host -- and evaluates every dense n_obs x n_obs quantity with the
FP64 tensor-core kernels (csrc/dgemm.cu, chol.cu, dense_ops.cu)."""
import ctypes as C
from collections import OrderedDict
import numpy as np
from scipy.linalg import LinAlgError
from . import _lib
from . import fem
from . import adapter
from .parallel import EnsembleComm, COMM_SELF, comm_world
from .ForcingCovariance import ForcingCovariance
from .InterpolationMatrix import InterpolationMatrix
from .ObsData import ObsData
from .solving_utils import SparseSolver, interp_covariance_to_data, _aga_device

_EVAL_CACHE_POINTS = 8       # recent (theta -> value, gradient) pairs kept by LinearSolver._evaluate


class LinearSolver(object):
    def __init__(self, A, b, G, data, *, priors=[None, None, None], ensemble_comm=COMM_SELF, P=None,
                 solver_parameters=None, nullspace=None, transpose_nullspace=None, near_nullspace=None,
                 options_prefix=None):
        if not adapter.looks_like_matrix(A):
            raise TypeError("A must be a firedrake matrix")
        if not adapter.is_function_like(b) or isinstance(b, np.ndarray):
            raise TypeError("b must be a firedrake function or vector")
        if not isinstance(G, ForcingCovariance):
            raise TypeError("G must be a forcing covariance")
        if not isinstance(data, ObsData):
            raise TypeError("data must be an ObsData type")
        if not isinstance(priors, list):
            raise TypeError("priors must be a list of prior objects or None")
        if not len(priors) == 3:
            raise ValueError("priors must be a list of prior objects or None of length 3")
        for p in priors:
            if p is not None:
                raise TypeError("priors must be a list of prior objects or None")
        if not isinstance(ensemble_comm, EnsembleComm):
            raise TypeError("ensemble_comm must be an MPI communicator created with a firedrake Ensemble")
        # native containers, or duck-typed Firedrake-style ones (adapter.py): A over G's function space, b as a Function
        A = adapter.as_matrix(A, G.function_space)
        if not isinstance(b, (fem.Function, fem.Vector)):
            b = fem.Function(G.function_space, adapter.function_array(b))
        self.solver = SparseSolver(A, P=P, solver_parameters=solver_parameters, nullspace=nullspace,
                                   transpose_nullspace=transpose_nullspace, near_nullspace=near_nullspace,
                                   options_prefix=options_prefix)
        self.b = b
        self.G = G
        self.data = data
        self.ensemble_comm = ensemble_comm
        self.priors = list(priors)
        self.params = None
        self.im = InterpolationMatrix(G.function_space, data.get_coords())
        self.x = None
        self.mu = None
        self.Cu = None
        self.current_logpost = None
        # device mirrors + evaluation cache
        self._x_dev = self._mu_dev = self._Cu_dev = None
        self._W_dev = self._W_cols = None      # this rank's block of W = G A^-1 P, kept by solve_prior (keep_factors)
        self._dense_work = None
        self._cache = None
        self.fuse_gradient = False      # set by estimate_params_MAP: evaluate value and gradient in one pass
        self.n_logpost_evals = 0
        self.n_logpost_cache_hits = 0
        self.host_results = True        # False: leave Cu on the device in solve_prior (bench.py's device-resident arm)

    def clone_for_data(self, data, private_stream=False):
        """A LinearSolver for another data set at the SAME sensor locations that shares this one's stiffness solver,
        interpolation matrix and cached prior (x, mu, Cu).  Extension over the reference, whose caches are per object
        (LinearSolver.py:150-153): several data snapshots on one sensor array need the n_obs solves only once.
        ``private_stream``: give the clone its own CUDA stream + context for the dense n_obs x n_obs work, so that
        clones driven from different host threads run concurrently on the GPU (estimation.fit_snapshots)."""
        if not isinstance(data, ObsData):
            raise TypeError("data must be an ObsData type")
        assert np.array_equal(data.get_coords(), self.data.get_coords()), "snapshots must share the sensor locations"
        other = object.__new__(LinearSolver)
        other.__dict__.update(self.__dict__)
        other.data = data
        other._owns_im = False
        other.params = None
        other._cache = None
        other._dense_work = None if private_stream else self._dense_work
        other._dense_ctx = _lib.acquire_stream_context() if private_stream else self.__dict__.get("_dense_ctx")
        other._owns_dense_ctx = bool(private_stream)
        other._data_dev = None
        other.n_logpost_evals = 0
        other.n_logpost_cache_hits = 0
        other._parent = self          # the clone borrows im / solver / prior: keep their owner (and its __del__) alive
        return other

    @property
    def dctx(self):
        "context (device + stream) of the dense data-space work"
        return self.__dict__.get("_dense_ctx") or self.solver.ctx

    def _data_device(self):
        "device copies of the sensor coordinates / data / uncertainties, ordered on this solver's dense stream"
        if self.__dict__.get("_dense_ctx") is None:
            return self.data.device()
        if self.__dict__.get("_data_dev") is None:
            ctx, od = self.dctx, self.data
            self._data_dev = dict(ctx=ctx, coords=ctx.to_device(od.coords, np.float64), data=ctx.to_device(od.data, np.float64),
                                  unc=ctx.to_device(np.atleast_1d(od.unc), np.float64), unc_vec=int(od.unc.shape != ()))
        return self._data_dev

    def __del__(self):
        try:
            if self.__dict__.get("_owns_im", True):
                self.im.destroy()
            if self.__dict__.get("_owns_dense_ctx", False):
                self._dense_work = None
                _lib.release_stream_context(self._dense_ctx)      # back to the pool, workspace attached
        except Exception:
            pass

    def set_params(self, params):
        "hyperparameters (log rho, log sigma_eta, log l_eta) (LinearSolver.py:165-183)"
        params = np.array(params, dtype=np.float64)
        assert params.shape == (3,), "bad shape for model discrepancy parameters"
        self.params = params

    # ---------------------------------------------------------------------------------------------------------
    @property
    def _root(self):
        return self.ensemble_comm.rank == 0 and self.G.comm.rank == 0

    def _work(self):
        n = self.data.get_n_obs()
        need = _lib.lib().sfem_dense_workspace_bytes(n)
        if self._dense_work is None or self._dense_work.numel() < need + 256:
            pooled = self.dctx.__dict__.get("_work_cache")         # a pooled stream context keeps its workspace
            if pooled is None or pooled.numel() < need + 256:
                pooled = self.dctx.workspace(need)
                self.dctx._work_cache = pooled
            self._dense_work = pooled
        w = self._dense_work
        off = (-w.data_ptr()) % 256
        return C.c_void_p(w.data_ptr() + off), int(w.numel() - off)

    def solve_prior(self):
        "prior mean and covariance at the sensors (LinearSolver.py:185-222)"
        ctx = self.solver.ctx
        keep = {} if self._keep_factors() else None
        Cu_dev = interp_covariance_to_data(self.im, self.G, self.solver, self.im, self.ensemble_comm, return_device=True,
                                           keep=keep)
        self._W_dev, self._W_cols = (keep["W"], keep["cols"]) if keep is not None else (None, None)
        self.x = fem.Function(self.G.function_space)
        if self.ensemble_comm.rank == 0:
            bf = self.b.function if isinstance(self.b, fem.Vector) else self.b
            self._x_dev = self.solver.solve_vector(_lib.cached_upload(bf, "data", ctx, lambda: ctx.to_device(bf.data, np.float64)),
                                                   apply_bcs=self.solver.A.bcs is not None)
            if self.host_results:
                self.x.data[:] = self._x_dev.cpu().numpy()
            self._mu_dev = self.im.csc.matmul(self._x_dev)
            self.mu = self._mu_dev.cpu().numpy()
        else:
            self.mu = np.zeros(0)
        if Cu_dev is not None:
            self._Cu_dev = Cu_dev.contiguous()
            self.Cu = self._Cu_dev.cpu().numpy() if self.host_results else self._Cu_dev
        else:
            self.Cu = np.zeros((0, 0))
        self._cache = None
        return self.mu, self.Cu

    # ---- persistence of the cached prior -----------------------------------------------------
    def save_prior(self, path):
        "write the cached prior (x on the mesh, mu and Cu at the sensors) to ``path`` (.npz); ensemble root only"
        if self.Cu is None or self.mu is None:
            self.solve_prior()
        if self.ensemble_comm.rank == 0:
            np.savez(path, x=self._x_dev.cpu().numpy(), mu=self._mu_dev.cpu().numpy(), Cu=self._Cu_dev.cpu().numpy(),
                     sensors=self.data.get_coords())

    def load_prior(self, path):
        "restore what ``save_prior`` wrote instead of running the n_obs solves (same mesh, same sensors)"
        z = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
        assert np.array_equal(z["sensors"], self.data.get_coords()), "the saved prior belongs to other sensor locations"
        assert z["x"].shape[0] == self.solver.n, "the saved prior belongs to another mesh"
        ctx = self.solver.ctx
        self.x = fem.Function(self.G.function_space, z["x"])
        if self.ensemble_comm.rank == 0:
            self._x_dev, self._mu_dev = ctx.to_device(z["x"], np.float64), ctx.to_device(z["mu"], np.float64)
            self._Cu_dev = ctx.to_device(z["Cu"], np.float64)
            self.mu, self.Cu = z["mu"].copy(), z["Cu"].copy()
        else:
            self.mu, self.Cu = np.zeros(0), np.zeros((0, 0))
        self._W_dev = self._W_cols = None
        self._cache = None
        return self.mu, self.Cu

    def _keep_factors(self):
        "keep this rank's block of W = G A^-1 P after solve_prior?  (solver parameter ``keep_factors``; default: if it fits)"
        flag = self.solver.keep_factors
        if flag is not None:
            return bool(flag)
        import torch
        # the same answer on every rank (the largest share of the columns decides): the posterior-mean path is collective
        widest = -(-self.data.get_n_obs() // max(self.ensemble_comm.size, 1))
        total = torch.cuda.get_device_properties(self.solver.ctx.tdev).total_memory
        return 8. * self.solver.n * widest <= 0.25 * total

    def _aga_p(self, Cmat, collective):
        """A^-1 G A^-1 P C for a block C (n_obs x ns, device) of data-space vectors -> (n x ns) on the ensemble root.
        With the resident W = G A^-1 P (solve_prior) this is W C followed by ONE block solve; the columns of W are
        spread over the ensemble like the sensor columns, so with ``collective`` every rank multiplies its block and the
        partial sums are reduced on the root (all ranks must call).  Without W: the two solves of solving_utils.py:49-53
        around G, on the root alone."""
        sp, ctx, lib = self.solver, self.solver.ctx, _lib.lib()
        comm = self.ensemble_comm
        ns = int(Cmat.shape[1])
        W = self.__dict__.get("_W_dev")
        have_all = W is not None and (comm.size == 1 or collective)
        if have_all:
            lo, hi = self._W_cols
            V = ctx.zeros((sp.n, ns))
            if hi > lo:
                Cl = Cmat[lo:hi]
                ctx.check(lib.sfem_skinny_gemm(ctx.handle, sp.n, hi - lo, ns, _lib.ptr(W), int(W.stride(0)), _lib.ptr(Cl),
                                               int(Cl.stride(0)), _lib.ptr(V), ns))
            if comm.size > 1:
                comm.reduce_tensor(V, root=0)
            if comm.rank != 0:
                return None
            X = ctx.empty((sp.n, ns))
            sp.solve_block(V, X)
            return X
        if comm.rank != 0:
            return None
        if not self.im.is_assembled:
            self.im.assemble()
        X = ctx.empty((sp.n, ns))
        sp.solve_block(self.im.csr.matmul(Cmat.contiguous()), X)
        Wb = self.G.G.matmul(X)
        sp.solve_block(Wb, X)
        return X

    def _cu_t_times(self, T, rhos):
        "rho_s Cu^T T[:, s] + mu for every column s (n_obs x ns): P^T (rho A^-1 G A^-1 P t + x) without touching the mesh"
        import torch
        ctx, lib = self.solver.ctx, _lib.lib()
        n, ns = int(T.shape[0]), int(T.shape[1])
        out = self._mu_dev[:, None].repeat(1, ns).contiguous()
        Ts = (T * torch.as_tensor(rhos, dtype=torch.float64, device=T.device)[None, :]).contiguous()
        ctx.check(lib.sfem_dgemm(ctx.handle, _lib.OP_T, _lib.OP_N, n, ns, n, 1.0, _lib.ptr(self._Cu_dev), n, _lib.ptr(Ts), ns,
                                 1.0, _lib.ptr(out), ns))
        return out

    def _dense_solve(self, rho2, with_Cu, b_dev):
        """(K_eta + unc^2 I [+ rho2 Cu])^-1 b on this solver's dense context.  A clone with a private stream
        (``clone_for_data(private_stream=True)``) runs it there: unless the caller already made that stream current
        (fit_snapshots / solve_posterior_snapshots do), the right-hand side -- produced on the main stream -- is
        awaited first and the result is complete before the main stream goes on."""
        import torch
        ctx = self.dctx
        side = ctx.tstream is not None and torch.cuda.current_stream(ctx.tdev) != ctx.tstream
        if side:
            self.solver.ctx.sync()
            with torch.cuda.stream(ctx.tstream):
                out = self._dense_solve_on(ctx, rho2, with_Cu, b_dev)
            ctx.tstream.synchronize()
            return out
        return self._dense_solve_on(ctx, rho2, with_Cu, b_dev)

    def _dense_solve_on(self, ctx, rho2, with_Cu, b_dev):
        d = self._data_device()
        out = ctx.empty((self.data.get_n_obs(),))
        work, nbytes = self._work()
        ctx.check(_lib.lib().sfem_dense_solve(ctx.handle, self.data.get_n_obs(), self.data.get_n_dim(), _lib.ptr(d["coords"]),
                                              _lib.ptr(d["unc"]), d["unc_vec"], float(self.params[1]), float(self.params[2]),
                                              float(rho2), _lib.ptr(self._Cu_dev if with_Cu else None), _lib.ptr(b_dev),
                                              _lib.ptr(out), work, nbytes))
        return out

    def _axpby(self, a, x, b, y):
        ctx = self.solver.ctx
        z = ctx.empty((x.shape[0],))
        ctx.check(_lib.lib().sfem_axpby(ctx.handle, x.shape[0], float(a), _lib.ptr(x), float(b), _lib.ptr(y), _lib.ptr(z)))
        return z

    def solve_posterior(self, x, scale_mean=False):
        "posterior mean on the mesh, written into ``x`` on the ensemble root (LinearSolver.py:224-305)"
        if not isinstance(bool(scale_mean), bool):
            raise TypeError("scale_mean argument must be boolean-like")
        if self.Cu is None or self.x is None:
            self.solve_prior()
        if self.params is None:
            raise ValueError("must set parameter values to solve posterior")
        rho = np.exp(self.params[0])
        scalefact = rho if scale_mean else 1.
        if self.ensemble_comm.rank == 0:
            # The reference's chain (LinearSolver.py:259-305)
            #   t1 = Kd^-1 y,  m2 = rho A^-1 G A^-1 P t1 + x,  d2 = (Kd + rho^2 Cu)^-1 P^T m2,  m3 = A^-1 G A^-1 P d2,
            #   mean = sf (m2 - rho^2 m3)
            # with P^T A^-1 G A^-1 P = Cu^T (the cached prior) and both mesh-space terms folded into one:
            #   mean = sf (x + A^-1 G A^-1 P (rho t1 - rho^2 d2)),   P^T m2 = rho Cu^T t1 + mu.
            d = self.data.device()
            try:
                t1 = self._dense_solve(0., False, d["data"])
            except LinAlgError:
                raise LinAlgError("Error attempting to compute the Cholesky factorization of the model discrepancy")
            d1 = self._cu_t_times(t1.reshape(-1, 1), [rho]).reshape(-1)
            try:
                d2 = self._dense_solve(rho ** 2, True, d1)
            except LinAlgError:
                raise LinAlgError("Error attempting to compute the Cholesky factorization of the model discrepancy "
                                  "plus forcing covariance")
            coef = self._axpby(rho, t1, -rho ** 2, d2)
            m = self._aga_p(coef.reshape(-1, 1), collective=False).reshape(-1)
            out = self._axpby(scalefact, m, scalefact, self._x_dev)
            adapter.assign_function(x, out.cpu().numpy())

    def solve_posterior_covariance(self, scale_mean=False):
        "posterior mean and covariance at the sensors (LinearSolver.py:308-373)"
        if not isinstance(bool(scale_mean), bool):
            raise TypeError("scale_mean argument must be boolean-like")
        if self.mu is None or self.Cu is None:
            self.solve_prior()
        if self.params is None:
            raise ValueError("must set parameter values to solve posterior")
        rho = np.exp(self.params[0])
        scalefact = rho if scale_mean else 1.
        if self._root:
            n = self.data.get_n_obs()
            d = self.data.device()
            ctx = self.solver.ctx
            muy, Cuy = ctx.empty((n,)), ctx.empty((n, n))
            theta = (C.c_double * 3)(*self.params)
            work, nbytes = self._work()
            try:
                ctx.check(_lib.lib().sfem_posterior_cov(ctx.handle, n, self.data.get_n_dim(), _lib.ptr(d["coords"]),
                                                        _lib.ptr(d["data"]), _lib.ptr(d["unc"]), d["unc_vec"],
                                                        _lib.ptr(self._mu_dev), _lib.ptr(self._Cu_dev), theta,
                                                        _lib.ptr(muy), _lib.ptr(Cuy), work, nbytes))
            except LinAlgError:
                raise LinAlgError("Cholesky factorization of one of the covariance matrices failed")
            if not self.host_results:          # bench.py's device-resident arm: results stay in HBM
                return scalefact * muy, Cuy
            return scalefact * muy.cpu().numpy(), Cuy.cpu().numpy()
        return scalefact * np.zeros(0), np.zeros((0, 0))

    def solve_prior_generating(self):
        "prior of the generating process rho u + eta at the sensors (LinearSolver.py:375-406)"
        if self.mu is None or self.Cu is None:
            self.solve_prior()
        if self.params is None:
            raise ValueError("must set parameter values to solve prior of generating process")
        rho = np.exp(self.params[0])
        if self._root:
            return rho * self.mu, self._kcu_matrix(rho ** 2, with_unc=False)
        return np.zeros(0), np.zeros((0, 0))

    def _kcu_matrix(self, rho2, with_unc, symmetric_upper=False):
        """K_eta [+ diag(unc^2)] + rho2 * Cu on the device -> host.  The generating-process prior adds the full
        (slightly non-symmetric) Cu, as NumPy does at LinearSolver.py:401."""
        d = self.data.device()
        ctx = self.solver.ctx
        n = self.data.get_n_obs()
        K = ctx.empty((n, n))
        ctx.check(_lib.lib().sfem_kernel_matrix(ctx.handle, n, self.data.get_n_dim(), _lib.ptr(d["coords"]),
                                                float(self.params[1]), float(self.params[2]),
                                                _lib.ptr(d["unc"] if with_unc else None), d["unc_vec"], 0., _lib.ptr(None), 0,
                                                _lib.ptr(K)))
        return rho2 * self.Cu + K.cpu().numpy()

    def solve_posterior_generating(self):
        "posterior of the generating process at the sensors (LinearSolver.py:408-439)"
        m_eta, C_eta = self.solve_prior_generating()
        if self._root:
            n = self.data.get_n_obs()
            ctx = self.solver.ctx
            lib = _lib.lib()
            unc2 = self.data.get_unc() ** 2
            # SciPy's default cho_factor reads the upper triangle: mirror it so the lower factorisation sees the same matrix
            Mh = unc2 * np.eye(n) + C_eta
            M = ctx.to_device(np.triu(Mh) + np.triu(Mh, 1).T, np.float64)
            dinv = ctx.empty((lib.sfem_potrf_dinv_doubles(n),))
            info = C.c_int(0)
            try:
                ctx.check(lib.sfem_potrf(ctx.handle, n, _lib.ptr(M), n, _lib.ptr(dinv), C.byref(info)))
            except LinAlgError:
                raise LinAlgError("Cholesky factorization of the covariance matrix failed")
            rhs = np.concatenate([unc2 * C_eta, (np.dot(C_eta, self.data.get_data()) + unc2 * m_eta)[:, None]], axis=1)
            B = ctx.to_device(rhs, np.float64)
            ctx.check(lib.sfem_potrs(ctx.handle, n, _lib.ptr(M), n, _lib.ptr(dinv), n + 1, _lib.ptr(B), n + 1))
            sol = B.cpu().numpy()
            return sol[:, n].copy(), sol[:, :n].copy()
        return np.zeros(0), np.zeros((0, 0))

    def predict_mean(self, coords, scale_mean=True):
        "posterior mean at new locations (LinearSolver.py:441-498)"
        if not isinstance(bool(scale_mean), bool):
            raise TypeError("scale_mean argument must be boolean-like")
        coords = np.array(coords, dtype=np.float64)
        if coords.ndim == 1:
            coords = np.reshape(coords, (-1, 1))
        assert coords.ndim == 2, "coords must be a 1d or 2d array"
        assert coords.shape[1] == self.data.get_n_dim(), "axis 1 of coords must be the same length as the FEM dimension"
        if self.Cu is None:
            self.solve_prior()
        if self.params is None:
            raise ValueError("must set parameter values to make predictions")
        scalefact = np.exp(self.params[0]) if scale_mean else 1.
        x = fem.Function(self.G.function_space)
        self.solve_posterior(x)
        im = InterpolationMatrix(self.G.function_space, coords)
        mu = scalefact * im.interp_mesh_to_data(x.vector())
        im.destroy()
        return mu

    def predict_covariance(self, coords, unc):
        "predictive covariance at new locations (LinearSolver.py:500-567)"
        coords = np.array(coords, dtype=np.float64)
        if coords.ndim == 1:
            coords = np.reshape(coords, (-1, 1))
        assert coords.ndim == 2, "coords must be a 1d or 2d array"
        assert coords.shape[1] == self.data.get_n_dim(), "axis 1 of coords must be the same length as the FEM dimension"
        if self.Cu is None:
            self.solve_prior()
        if self.params is None:
            raise ValueError("must set parameter values to make predictions")
        rho = np.exp(self.params[0])
        im_coords = InterpolationMatrix(self.G.function_space, coords)
        if coords.shape[0] > self.data.get_n_obs():
            Cucd = interp_covariance_to_data(im_coords, self.G, self.solver, self.im, self.ensemble_comm)
        else:
            Cucd = interp_covariance_to_data(self.im, self.G, self.solver, im_coords, self.ensemble_comm).T
        Cucc = interp_covariance_to_data(im_coords, self.G, self.solver, im_coords, self.ensemble_comm)
        if self._root:
            n = self.data.get_n_obs()
            d = self.data.device()
            ctx = self.solver.ctx
            lib = _lib.lib()
            M = ctx.empty((n, n))
            ctx.check(lib.sfem_kernel_matrix(ctx.handle, n, self.data.get_n_dim(), _lib.ptr(d["coords"]),
                                             float(self.params[1]), float(self.params[2]), _lib.ptr(d["unc"]), d["unc_vec"],
                                             float(rho ** 2), _lib.ptr(self._Cu_dev), 0, _lib.ptr(M)))
            dinv = ctx.empty((lib.sfem_potrf_dinv_doubles(n),))
            info = C.c_int(0)
            try:
                ctx.check(lib.sfem_potrf(ctx.handle, n, _lib.ptr(M), n, _lib.ptr(dinv), C.byref(info)))
            except LinAlgError:
                raise LinAlgError("Cholesky factorization of one of the covariance matrices failed")
            nc = coords.shape[0]
            B = ctx.to_device(np.ascontiguousarray(Cucd.T), np.float64)          # n x nc
            ctx.check(lib.sfem_potrs(ctx.handle, n, _lib.ptr(M), n, _lib.ptr(dinv), nc, _lib.ptr(B), nc))
            Cd = ctx.to_device(Cucd, np.float64)                                   # nc x n
            Cuy = ctx.to_device(Cucc, np.float64)
            ctx.check(lib.sfem_dgemm(ctx.handle, _lib.OP_N, _lib.OP_N, nc, nc, n, -rho ** 2, _lib.ptr(Cd), n, _lib.ptr(B), nc,
                                     1.0, _lib.ptr(Cuy), nc))
            Knew = ObsData(coords, np.zeros(coords.shape[0]), unc).calc_K_plus_sigma(self.params[1:])
            out = Knew + rho ** 2 * Cuy.cpu().numpy()
        else:
            out = np.zeros((0, 0))
        im_coords.destroy()
        return out

    # ---------------------------------------------------------------------------------------------------------
    def _evaluate(self, params, need_grad):
        """negative log marginal likelihood (and gradient) on the device; identical result on every process.  In the
        reference every process runs the optimiser in lock-step and the value is broadcast from rank 0 at each
        evaluation (LinearSolver.py:615-621).  After ``share_prior`` every process holds mu and Cu, evaluates locally
        and no per-evaluation collective is needed (``self._local_eval``, used by estimation.fit_snapshots)."""
        self.set_params(params)
        if self.Cu is None or self.mu is None:
            self.solve_prior()
        # A few recent points are remembered, not just the last one: when L-BFGS-B can no longer resolve a decrease
        # (tolerances below the rounding noise of the objective) its line searches and restarts return to the same
        # theta again and again; the value is a pure function of (theta, prior, data), so those cost nothing.
        key = self.params.tobytes()
        if self._cache is not None:
            hit = self._cache.get(key)
            if hit is not None and (hit[1] is not None or not need_grad):
                self._cache.move_to_end(key)
                self.n_logpost_cache_hits += 1
                return hit
        world = comm_world()
        local = bool(self.__dict__.get("_local_eval", False)) and self._Cu_dev is not None
        if world.rank == 0 or local:
            n = self.data.get_n_obs()
            d = self._data_device()
            ctx = self.dctx
            theta = (C.c_double * 3)(*self.params)
            value = C.c_double(0.)
            grad = (C.c_double * 3)()
            work, nbytes = self._work()
            try:
                ctx.check(_lib.lib().sfem_logpost(ctx.handle, n, self.data.get_n_dim(), _lib.ptr(d["coords"]),
                                                  _lib.ptr(d["data"]), _lib.ptr(d["unc"]), d["unc_vec"], _lib.ptr(self._mu_dev),
                                                  _lib.ptr(self._Cu_dev), theta, C.byref(value),
                                                  grad if need_grad else None, work, nbytes))
            except LinAlgError:
                raise LinAlgError("Error attempting to factorize the covariance matrix in model_loglikelihood")
            self.n_logpost_evals += 1
            result = (float(value.value), np.array(list(grad), dtype=np.float64) if need_grad else None)
        else:
            result = None
        if not local:
            result = world.bcast(result, root=0)
            assert result is not None, "error in broadcasting the log likelihood"
            world.barrier()
        if self._cache is None:
            self._cache = OrderedDict()
        self._cache[key] = (result[0], result[1])
        self._cache.move_to_end(key)
        while len(self._cache) > _EVAL_CACHE_POINTS:
            self._cache.popitem(last=False)
        return result

    def share_prior(self):
        """Broadcast the cached prior (x, mu, Cu) from ensemble rank 0 to every rank's GPU (NCCL over NVLink; 134 MB
        at 4096 sensors), so that every rank can run the dense data-space work -- e.g. the MAP fits of different data
        snapshots -- on its own.  No counterpart in the reference, whose dense end runs on rank 0 only."""
        comm = self.ensemble_comm
        if comm.size == 1 or self.__dict__.get("_prior_shared", False):
            return
        if self.Cu is None or self.mu is None:
            self.solve_prior()
        ctx = self.solver.ctx
        n_obs, n = self.data.get_n_obs(), self.solver.n
        if comm.rank != 0:
            self._Cu_dev = ctx.empty((n_obs, n_obs))
            self._mu_dev = ctx.empty((n_obs,))
            self._x_dev = ctx.empty((n,))
        for t in (self._Cu_dev, self._mu_dev, self._x_dev):
            comm.broadcast_tensor(t, root=0)
        if comm.rank != 0:
            self.mu = self._mu_dev.cpu().numpy()
            self.Cu = self._Cu_dev if not self.host_results else self._Cu_dev.cpu().numpy()
        self._prior_shared = True

    def logposterior(self, params):
        "negative log posterior (LinearSolver.py:569-623); priors are always None (LinearSolver.py:130-132)"
        value, _ = self._evaluate(params, need_grad=self.fuse_gradient)
        return value

    def logpost_deriv(self, params):
        "gradient of the negative log posterior wrt the three log-parameters (LinearSolver.py:625-691)"
        _, grad = self._evaluate(params, need_grad=True)
        return grad.copy()


def solve_posterior_snapshots(solvers, xs, scale_mean=False):
    """Posterior means on the mesh for several fitted LinearSolvers that share one prior (``clone_for_data``): the
    arithmetic of ``solve_posterior`` (LinearSolver.py:259-305) for every snapshot.  With W = G A^-1 P resident the
    four mesh-space solves of a snapshot collapse into one (see ``solve_posterior``), and the solves of all snapshots
    advance together as ONE block solve on the root.  On several processes whose prior has been shared
    (``share_prior``, done by ``fit_snapshots``) the dense data-space solves of the snapshots are dealt out round-robin;
    every rank multiplies its column block of W and the partial sums are reduced on rank 0 (collective: every rank of
    the ensemble must call).  ``xs[s]`` receives the mean of snapshot s on the ensemble root.  Extension over the reference API (which
    handles one data set per LinearSolver)."""
    import torch
    if not isinstance(bool(scale_mean), bool):
        raise TypeError("scale_mean argument must be boolean-like")
    assert len(solvers) == len(xs) and len(solvers) > 0
    base = solvers[0]
    for ls in solvers:
        if ls.Cu is None or ls.x is None:
            ls.solve_prior()
        if ls.params is None:
            raise ValueError("must set parameter values to solve posterior")
        assert ls.solver is base.solver and ls.im is base.im, "snapshots must share the prior (clone_for_data)"
    comm = base.ensemble_comm
    shared = comm.size > 1 and bool(base.__dict__.get("_prior_shared", False))
    if shared:
        mine = [i for i in range(len(solvers)) if i % comm.size == comm.rank]
    else:
        mine = list(range(len(solvers))) if comm.rank == 0 else []
    sp, G, im, ctx = base.solver, base.G, base.im, base.solver.ctx
    n, nobs = sp.n, base.data.get_n_obs()
    outs = {}
    if mine:
        sub = [solvers[i] for i in mine]
        ns = len(sub)
        rhos = [float(np.exp(ls.params[0])) for ls in sub]

        def dense_cols(rho2s, with_Cu, rhs_cols):
            """one dense data-space solve per snapshot; clones with a private stream run theirs concurrently (one host
            thread each), the others one after the other on the main stream"""
            import threading
            out = ctx.empty((nobs, ns))
            ctx.sync()                   # right-hand sides were produced on the main stream; clones may use their own
            cols, errs = [None] * ns, [None] * ns

            def one(s_, use_stream):
                try:
                    if use_stream:
                        with torch.cuda.stream(sub[s_].dctx.tstream):
                            cols[s_] = sub[s_]._dense_solve(rho2s[s_], with_Cu, rhs_cols[s_])
                            sub[s_].dctx.sync()
                    else:
                        cols[s_] = sub[s_]._dense_solve(rho2s[s_], with_Cu, rhs_cols[s_])
                except BaseException as exc:
                    errs[s_] = exc

            threads = [threading.Thread(target=one, args=(s_, True)) for s_ in range(ns) if sub[s_].dctx.tstream is not None]
            for t in threads:
                t.start()
            for s_ in range(ns):
                if sub[s_].dctx.tstream is None:
                    one(s_, False)
            for t in threads:
                t.join()
            for exc in errs:
                if isinstance(exc, LinAlgError):
                    raise LinAlgError("Error attempting to compute the Cholesky factorization of the model discrepancy" +
                                      (" plus forcing covariance" if with_Cu else ""))
                if exc is not None:
                    raise exc
            ctx.sync()
            for s_ in range(ns):
                out[:, s_] = cols[s_]
            return out

        T1 = dense_cols([0.] * ns, False, [ls._data_device()["data"] for ls in sub])
        D1 = base._cu_t_times(T1, rhos)                                 # P^T (rho A^-1 G A^-1 P t1 + x), (n_obs x ns)
        D2 = dense_cols([r * r for r in rhos], True, [D1[:, s].contiguous() for s in range(ns)])
        rt = torch.tensor(rhos, dtype=torch.float64, device=T1.device)[None, :]
        coef_mine = T1 * rt - D2 * (rt * rt)                            # rho t1 - rho^2 d2 per snapshot
    # every snapshot's coefficient vector on every rank (the blocks of W are spread over the ranks), then ONE block
    # solve for all snapshots on the root:  mean_s = sf_s (x + A^-1 W coef_s)
    nsnap = len(solvers)
    coef = ctx.zeros((nobs, nsnap))
    if mine:
        coef[:, mine] = coef_mine
    if comm.size > 1:
        comm.allreduce_tensor(coef)
    X = base._aga_p(coef, collective=True)
    if comm.rank == 0:
        for i, (x, ls) in enumerate(zip(xs, solvers)):
            sf = float(np.exp(ls.params[0])) if scale_mean else 1.
            out = sf * (X[:, i] + base._x_dev)
            adapter.assign_function(x, out.cpu().numpy())
