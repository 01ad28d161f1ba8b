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
    (``LinearSolver.clone_for_data``) and, on a single process, every snapshot's L-BFGS-B loop runs in its own host
    thread on its own CUDA stream, so that the latency-bound parts of one fit (Cholesky panels, small GEMMs) overlap
    with the throughput-bound parts of the others.  Returns the list of fitted LinearSolvers (``ls`` itself is used
    for ``datas[0]`` when that is its own data set).  Extension over the reference API."""
    import threading
    import torch
    if ls.Cu is None or ls.mu is None:
        ls.solve_prior()
    world = comm_world()
    conc = bool(concurrent) and world.size == 1 and len(datas) > 1
    solvers = []
    for od in datas:
        solvers.append(ls if (od is ls.data and not conc) else ls.clone_for_data(od, private_stream=conc))
    if not conc:
        for s in solvers:
            fit_linear_solver(s, start=start, **minimize_kwargs)
        return solvers
    torch.cuda.synchronize()                  # the shared prior is complete before the side streams read it
    errors = [None] * len(solvers)

    def work(i):
        try:
            with torch.cuda.stream(solvers[i].dctx.tstream):
                fit_linear_solver(solvers[i], start=start, **minimize_kwargs)
                solvers[i].dctx.sync()
        except BaseException as exc:      # re-raised in the caller
            errors[i] = exc

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(solvers))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for exc in errors:
        if exc is not None:
            raise exc
    return solvers
