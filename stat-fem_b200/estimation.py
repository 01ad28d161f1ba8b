"""MAP / maximum-likelihood estimation of the model-discrepancy hyperparameters (API of estimation.py)."""
import numpy as np
from scipy.optimize import minimize
from .parallel import COMM_SELF, comm_world
from .LinearSolver import LinearSolver

_SOLVER_KWARGS = ("P", "solver_parameters", "nullspace", "transpose_nullspace", "near_nullspace", "options_prefix")


def estimate_params_MAP(A, b, G, data, priors=[None, None, None], start=None, ensemble_comm=COMM_SELF, **kwargs):
    """Minimise the negative log posterior with L-BFGS-B from ``start`` (random in [-2.5, 2.5)^3 if omitted) and return
    the LinearSolver with ``params`` set to the optimum (estimation.py:7-110).  Solver keywords go to LinearSolver,
    everything else becomes SciPy ``options`` (estimation.py:65-76).  The three-parameter quasi-Newton driver stays on
    the host; each evaluation is one fused GPU pass that returns value and gradient together."""
    ls_kwargs = {k: v for k, v in kwargs.items() if k in _SOLVER_KWARGS}
    minimize_kwargs = {k: v for k, v in kwargs.items() if k not in _SOLVER_KWARGS}
    ls = LinearSolver(A, b, G, data, priors=priors, ensemble_comm=ensemble_comm, **ls_kwargs)
    ls.solve_prior()
    return fit_linear_solver(ls, start=start, **minimize_kwargs)


def fit_linear_solver(ls, start=None, **minimize_kwargs):
    """The optimisation half of estimate_params_MAP (estimation.py:83-110) on an existing LinearSolver whose prior is
    already cached -- e.g. one obtained with ``LinearSolver.clone_for_data`` for a further data snapshot."""
    world = comm_world()
    if start is None:
        start = 5. * (np.random.random(3) - 0.5) if world.rank == 0 else None
        start = world.bcast(start, root=0)
    else:
        assert np.array(start).shape == (3,), "bad shape for starting point"
    ls.fuse_gradient = True
    try:
        fmin_dict = minimize(ls.logposterior, start, method='L-BFGS-B', jac=ls.logpost_deriv, options=minimize_kwargs)
    finally:
        ls.fuse_gradient = False
    assert fmin_dict['success'], "minimization routine failed"
    result = fmin_dict['x']
    root_result = world.bcast(result, root=0)
    same_result = np.allclose(root_result, result)
    diff_arg = world.allreduce(int(not same_result))
    assert not diff_arg, "minimization did not produce identical results across all processes"
    world.barrier()
    ls.set_params(root_result)
    ls.minimize_result = fmin_dict
    return ls


def fit_snapshots(ls, datas, start=None, concurrent=True, polish=False, accept_precision_floor=False, **minimize_kwargs):
    """MAP fits for several data snapshots taken at the SAME sensors: the prior (x, mu, Cu) of ``ls`` is shared
    (``LinearSolver.clone_for_data``).  On one process every snapshot's L-BFGS-B loop runs in its own host thread on its
    own CUDA stream, so that the latency-bound parts of one fit (Cholesky panels, small GEMMs) overlap with the
    throughput-bound parts of the others.  On several processes (one per GPU) the prior is first broadcast to every
    GPU (``share_prior``) and the snapshots are dealt out round-robin; the optima are exchanged at the end, so every
    process returns the same list of fitted LinearSolvers.  Extension over the reference API.

    ``accept_precision_floor``: with tolerances below what FP64 can resolve at a few thousand sensors (the objective
    carries a relative rounding noise of ~1e-13; estimation.py:92-93 forwards e.g. ftol=1e-15) L-BFGS-B ends in
    "ABNORMAL_TERMINATION_IN_LNSRCH" -- no lower function value can be *measured* any more -- and the reference's
    ``assert success`` (estimation.py:95) would fire.  With this flag that termination is accepted as converged.
    ``polish``: afterwards refine the optimum with a few chord-Newton steps on the ANALYTIC gradient (Hessian by
    central differences of ``logpost_deriv``): the gradient is resolved far below the noise of the function values, so
    the optimum becomes reproducible to ~1e-8 -- independent of the summation order in which Cu was assembled (one
    GPU or several).  ``minimize_result`` of every solver records what happened."""
    import threading
    import torch
    if ls.Cu is None or ls.mu is None:
        ls.solve_prior()
    world = comm_world()
    ndata = len(datas)
    if start is None:       # one common random start, as estimate_params_MAP draws it (estimation.py:83-88)
        start = 5. * (np.random.random(3) - 0.5) if world.rank == 0 else None
        start = world.bcast(start, root=0)
    if world.size > 1:
        ls.share_prior()
    mine = [i for i in range(ndata) if i % world.size == world.rank]
    # every fit gets its own stream context, even a single one: a private (non-default) stream is what lets the library
    # replay the evaluation as a CUDA graph (csrc/dense_ops.cu)
    conc = bool(concurrent) and len(mine) > 0
    solvers = []
    for i, od in enumerate(datas):
        own_stream = conc and i in mine
        solvers.append(ls if (od is ls.data and not own_stream) else ls.clone_for_data(od, private_stream=own_stream))
    for s in solvers:
        s._local_eval = True          # every rank holds the prior: no per-evaluation broadcast
    errors = [None] * ndata
    results = np.zeros((ndata, 5))

    def work(i, use_stream):
        try:
            if use_stream:
                with torch.cuda.stream(solvers[i].dctx.tstream):
                    _minimize_local(solvers[i], start, minimize_kwargs, polish, accept_precision_floor)
                    solvers[i].dctx.sync()
            else:
                _minimize_local(solvers[i], start, minimize_kwargs, polish, accept_precision_floor)
            results[i, :3] = solvers[i].params
            results[i, 3] = solvers[i].n_logpost_evals
            results[i, 4] = solvers[i].n_logpost_cache_hits
        except BaseException as exc:      # re-raised in the caller
            errors[i] = exc

    if conc:
        torch.cuda.synchronize()              # the shared prior is complete before the side streams read it
        threads = [threading.Thread(target=work, args=(i, True)) for i in mine]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    else:
        for i in mine:
            work(i, False)
    failed = world.allreduce(float(any(e is not None for e in errors)))
    for exc in errors:
        if exc is not None:
            raise exc
    assert not failed, "a MAP fit failed on another process"
    if world.size > 1:
        results = np.asarray(world.allreduce(results))        # rows owned by other ranks are zero here
    for i, s in enumerate(solvers):
        s.set_params(results[i, :3])
        s.n_logpost_evals = int(results[i, 3])
        s.n_logpost_cache_hits = int(results[i, 4])
        s._local_eval = False
    return solvers


def newton_polish(ls, theta, step=1.e-3, max_iter=6, tol=1.e-9):
    """Chord-Newton refinement of a stationary point of the negative log posterior using only its analytic gradient:
    H = central differences of ``logpost_deriv`` (6 evaluations), then theta <- theta - H^-1 g until the update is
    below ``tol``.  Returns (theta, info); theta is returned unchanged if H is not positive definite or the iteration
    does not contract (then the L-BFGS-B point was not near a minimum)."""
    theta = np.array(theta, dtype=np.float64)
    H = np.zeros((3, 3))
    for k in range(3):
        e = np.zeros(3)
        e[k] = step
        H[:, k] = (ls.logpost_deriv(theta + e) - ls.logpost_deriv(theta - e)) / (2. * step)
    H = 0.5 * (H + H.T)
    info = dict(hessian=H, iterations=0, converged=False, last_update=None)
    if not np.all(np.linalg.eigvalsh(H) > 0.):
        return theta, info
    x = theta.copy()
    prev = np.inf
    for it in range(max_iter):
        dx = np.linalg.solve(H, ls.logpost_deriv(x))
        size = float(np.max(np.abs(dx)))
        if not np.isfinite(size) or size > 0.1 or size > 2. * prev:
            return theta, info                    # not contracting: keep the optimiser's point
        x = x - dx
        prev = size
        info.update(iterations=it + 1, last_update=size)
        if size < tol:
            info["converged"] = True
            break
    return x, info


def _minimize_local(ls, start, minimize_kwargs, polish=False, accept_precision_floor=False):
    "the L-BFGS-B loop of fit_linear_solver without its cross-process checks (used per snapshot / per thread)"
    ls.fuse_gradient = True
    try:
        fmin_dict = minimize(ls.logposterior, start, method='L-BFGS-B', jac=ls.logpost_deriv, options=minimize_kwargs)
        at_floor = accept_precision_floor and "ABNORMAL" in str(fmin_dict['message']).upper()
        assert fmin_dict['success'] or at_floor, "minimization routine failed: %s" % (fmin_dict['message'],)
        x = fmin_dict['x']
        if polish:
            x, info = newton_polish(ls, x)
            fmin_dict['polish'] = info
    finally:
        ls.fuse_gradient = False
    ls.set_params(x)
    ls.minimize_result = fmin_dict
