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


def fit_snapshots(ls, datas, start=None, concurrent=True, **minimize_kwargs):
    """MAP fits for several data snapshots taken at the SAME sensors: the prior (x, mu, Cu) of ``ls`` is shared
    (``LinearSolver.clone_for_data``).  On one process every snapshot's L-BFGS-B loop runs in its own host thread on its
    own CUDA stream, so that the latency-bound parts of one fit (Cholesky panels, small GEMMs) overlap with the
    throughput-bound parts of the others.  On several processes (one per GPU) the prior is first broadcast to every
    GPU (``share_prior``) and the snapshots are dealt out round-robin; the optima are exchanged at the end, so every
    process returns the same list of fitted LinearSolvers.  Extension over the reference API."""
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
    conc = bool(concurrent) and len(mine) > 1
    solvers = []
    for i, od in enumerate(datas):
        own_stream = conc and i in mine
        solvers.append(ls if (od is ls.data and not own_stream) else ls.clone_for_data(od, private_stream=own_stream))
    for s in solvers:
        s._local_eval = True          # every rank holds the prior: no per-evaluation broadcast
    errors = [None] * ndata
    results = np.zeros((ndata, 4))

    def work(i, use_stream):
        try:
            if use_stream:
                with torch.cuda.stream(solvers[i].dctx.tstream):
                    _minimize_local(solvers[i], start, minimize_kwargs)
                    solvers[i].dctx.sync()
            else:
                _minimize_local(solvers[i], start, minimize_kwargs)
            results[i, :3] = solvers[i].params
            results[i, 3] = solvers[i].n_logpost_evals
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
        s._local_eval = False
    return solvers


def _minimize_local(ls, start, minimize_kwargs):
    "the L-BFGS-B loop of fit_linear_solver without its cross-process checks (used per snapshot / per thread)"
    ls.fuse_gradient = True
    try:
        fmin_dict = minimize(ls.logposterior, start, method='L-BFGS-B', jac=ls.logpost_deriv, options=minimize_kwargs)
    finally:
        ls.fuse_gradient = False
    assert fmin_dict['success'], "minimization routine failed"
    ls.set_params(fmin_dict['x'])
    ls.minimize_result = fmin_dict
