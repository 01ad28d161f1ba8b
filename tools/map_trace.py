import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import stat_fem_b200 as stat_fem
from scipy.optimize import minimize
prob = bench.make_problem("c2")
G = stat_fem.ForcingCovariance(prob["V"], prob["sigma_f"], prob["l_f"], prob["cutoff"], prob["reg"])
ods = [stat_fem.ObsData(prob["sensors"], y, prob["unc"]) for y in prob["datas"]]
ls = stat_fem.LinearSolver(prob["A"], prob["b"], G, ods[0])
ls.host_results = False
ls.solve_prior()
for i in range(4):
    c = ls.clone_for_data(ods[i], private_stream=True)
    c._local_eval = True
    c.fuse_gradient = True
    trace = []
    def f(th, c=c):
        v = c.logposterior(th); trace.append((v, th.copy())); return v
    with torch.cuda.stream(c.dctx.tstream):
        t = time.perf_counter()
        r = minimize(f, np.zeros(3), method='L-BFGS-B', jac=c.logpost_deriv, options=dict(ftol=1e-15, gtol=1e-10))
        dt = time.perf_counter() - t
    vals = np.array([v for v, _ in trace])
    best = vals.min()
    first_best = int(np.argmax(vals <= best + 1e-12 * abs(best)))
    first_close = int(np.argmax(vals <= best + 1e-9 * abs(best)))
    print("fit", i, r.message, "nfev", r.nfev, "nit", r.nit, "evals", len(trace), "first within 1e-9 rel of best at", first_close,
          "within 1e-12 at", first_best, "time %.3f" % dt, "ms/eval %.2f" % (1e3 * dt / len(trace)),
          "device evaluations", c.n_logpost_evals, "cache hits", c.n_logpost_cache_hits)
    tail = vals[-25:] - best
    print("   last 25 f - best:", " ".join("%.1e" % x for x in tail))
