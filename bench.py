#!/usr/bin/env python
"""bench.py -- time-to-posterior of stat-fem's data-conditioning hot path on B200.

One *step* = one full pass of the hot path over one synthetic problem (BASELINE.json config 2 by default):
    G assembly -> stiffness upload + multigrid setup -> Z = A^-1 P (n_obs columns) -> W = G Z -> Cu = W^T Z, mu
    -> for each of the data snapshots: MAP estimate (L-BFGS-B from 0, ftol=1e-15, gtol=1e-10: SURVEY 8d) + posterior
       mean on the mesh (the snapshots' fits run concurrently on separate streams; their means are one block solve).
`value`  : seconds per step with every input already resident in HBM and no result copied back (device-timed).
`e2e`    : seconds per step through the public drop-in API with HOST buffers (NumPy in, NumPy out).
`roofline`: the stiffness-operator kernel family (tileop_dmma_kernel on the fine level: q = A p, Jacobi sweeps, residual),
            timed live with CUDA events inside the timed region; algorithmic bytes per launch are documented in DESIGN.md.
`--impl reference`: the NumPy/SciPy restatement of the reference (oracle/) on the box's host cores; every step is a
            bounded sample of the workload, extrapolated linearly (the faithful O(n^2) reference cannot run config 2 in
            full); the sparse LU is factored once per process and its time is part of every step's value.
`--workload`: c2 (default, BASELINE configs[1]), c1 (configs[0] geometry), c3 (configs[2]: 3-D, 8192 observations, needs
            >= 4 GPUs), c4 (configs[3]: G stress, no MAP), c5 (configs[4]: dense-end stress), mid / tiny (smoke sizes).
`--parity`: after the timed steps rank 0 recomputes a subsample with the CPU oracle (oracle/fullsize.py) and adds a
            `parity` object to the line (G rows, rows of Cu by the two-solve loop, dense quantities from the GPU's Cu).

Launch: python bench.py --gpus N --steps K --warmup W          (N > 1: under torch.distributed.run, one rank/GPU)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # kind "posterior": the full path incl. MAP + posterior means; "gstress": prior only (G assembly, solve, W = G Z, Cu);
    # "dense": prior + 200 paired logposterior/logpost_deriv evaluations on a fixed path + solve_posterior_covariance
    "c2": dict(dim=2, nodes=1024, nobs=4096, nsnap=8, lfac=1.5, kind="posterior", seed=1),     # BASELINE.json configs[1]
    "c1": dict(dim=2, nodes=101, nobs=100, nsnap=1, lfac=1.5, kind="posterior", seed=1),       # configs[0] geometry, scaled nugget
    "c3": dict(dim=3, nodes=128, nobs=8192, nsnap=8, lfac=0.8, kind="posterior", seed=2, min_gpus=4),   # configs[2]
    "c4": dict(dim=2, nodes=2048, nobs=512, nsnap=0, lfac=2.63, kind="gstress", seed=3),       # configs[3]
    "c5": dict(dim=2, nodes=256, nobs=16384, nsnap=1, lfac=1.5, kind="dense", seed=4, pairs=200),       # configs[4]
    "mid3": dict(dim=3, nodes=65, nobs=2048, nsnap=4, lfac=0.8, kind="posterior", seed=2),
    "mid": dict(dim=2, nodes=513, nobs=1024, nsnap=4, lfac=1.5, kind="posterior", seed=1),
    "tiny": dict(dim=2, nodes=65, nobs=64, nsnap=2, lfac=1.5, kind="posterior", seed=1),
    "tiny3": dict(dim=3, nodes=17, nobs=64, nsnap=2, lfac=0.8, kind="posterior", seed=2),
    "tiny4": dict(dim=2, nodes=129, nobs=64, nsnap=0, lfac=2.63, kind="gstress", seed=3),
    "tiny5": dict(dim=2, nodes=65, nobs=512, nsnap=1, lfac=1.5, kind="dense", seed=4, pairs=10),
}
MAP_OPTIONS = dict(ftol=1.e-15, gtol=1.e-10)
THETA0 = np.array([np.log(0.7), np.log(1.e-2), np.log(0.5)])


def make_problem(name):
    from stat_fem_b200 import fem
    w = WORKLOADS[name]
    dim, nodes, nobs, nsnap, lfac = w["dim"], w["nodes"], w["nobs"], w["nsnap"], w["lfac"]
    V, A, b = fem.poisson_problem(nodes, dim)
    h = 1. / (nodes - 1)
    sigma_f = np.log(2.e-2)
    l_f = np.log(lfac * h)
    reg = 1.e-8 * h ** (2 * dim) * np.exp(2 * sigma_f)        # nugget scaled with (h^d)^2 e^{2 sigma}
    sensors = np.random.default_rng(w["seed"]).uniform(0.02, 0.98, (nobs, dim))
    truth = 0.7 * np.prod(np.sin(2. * np.pi * sensors), axis=1)
    datas = []
    for s in range(max(nsnap, 1)):
        rng = np.random.default_rng(100 + s)
        disc = np.zeros(nobs)
        for m in range(6):      # smooth low-rank discrepancy ~ O(1e-2)
            kvec = rng.uniform(0.5, 2.5, dim)
            disc += 4.e-3 * rng.normal() * np.sin(np.pi * sensors @ kvec + rng.uniform(0, 2 * np.pi))
        datas.append(truth + disc + rng.normal(scale=2.e-3, size=nobs))
    return dict(name=name, dim=dim, nodes=nodes, nobs=nobs, nsnap=nsnap, kind=w["kind"], pairs=w.get("pairs", 0), V=V, A=A,
                b=b, sigma_f=sigma_f, l_f=l_f, reg=reg, cutoff=1.e-3, sensors=sensors, datas=datas, unc=2.e-3, h=h)


def theta_path(k, pairs):
    "SURVEY 8(d) C5: theta_k = theta_0 + 0.01 k (1, -1, 1) / pairs"
    return THETA0 + 0.01 * k / max(pairs, 1) * np.array([1., -1., 1.])


# ------------------------------------------------------------------------------------------------ GPU arm
def gpu_step(prob, comm, resident, stats=None):
    """one pass of the hot path through the drop-in API; returns a small dict of results"""
    import torch
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import fem
    from stat_fem_b200.estimation import fit_snapshots
    V, A, b = prob["V"], prob["A"], prob["b"]
    _t = [time.perf_counter()]
    mem = prob.setdefault("_min_free", [float("inf")])

    def lap(name):       # SFEM_BENCH_TIMING=1: wall-clock phase report (adds device synchronisations)
        mem[0] = min(mem[0], torch.cuda.mem_get_info()[0])
        if os.environ.get("SFEM_BENCH_TIMING") and comm.rank == 0:
            torch.cuda.synchronize()
            now = time.perf_counter()
            print("    [%s] %-26s %.3f s" % ("resident" if resident else "e2e", name, now - _t[0]), file=sys.stderr, flush=True)
            _t[0] = now
    G = stat_fem.ForcingCovariance(V, prob["sigma_f"], prob["l_f"], prob["cutoff"], prob["reg"])
    G.assemble()
    lap("G assemble")
    ods = [stat_fem.ObsData(prob["sensors"], y, prob["unc"]) for y in prob["datas"]]
    ls = stat_fem.LinearSolver(A, b, G, ods[0], ensemble_comm=comm)
    if resident and "im" in prob:
        ls.im = prob["im"]              # interpolation matrix already assembled and resident in HBM
        ls._owns_im = False
    ls.host_results = not resident
    lap("LinearSolver + MG setup")
    mu, Cu = ls.solve_prior()
    lap("solve_prior")
    out = dict(nnzG=G.G.nnz, iters=list(ls.solver.block_iterations), thetas=[], evals=0)
    solvers = [ls]
    if prob["kind"] == "posterior":
        # the snapshots share the sensors, hence the prior: their MAP fits run concurrently (one stream each) and their
        # posterior means are advanced together as one block solve
        solvers = fit_snapshots(ls, ods, start=np.zeros(3), polish=True, accept_precision_floor=True, **MAP_OPTIONS)
        out["map_messages"] = sorted({str(getattr(s_, "minimize_result", {}).get("message", "(fitted on another rank)")) for s_ in solvers})
        out["evals_per_fit"] = [int(s_.n_logpost_evals) for s_ in solvers]
        out["eval_cache_hits_per_fit"] = [int(s_.n_logpost_cache_hits) for s_ in solvers]
        for lss in solvers:
            out["thetas"].append(lss.params.copy())
            out["evals"] += lss.n_logpost_evals
        lap("MAP fits")
        xs = [fem.Function(V) for _ in ods]
        stat_fem.solve_posterior_snapshots(solvers, xs, scale_mean=True)
        lap("posterior means")
        out["x_norm"] = float(np.linalg.norm(xs[-1].data))
    elif prob["kind"] == "dense":
        if comm.rank == 0:
            ls.fuse_gradient = True
            vals = []
            for k in range(prob["pairs"]):
                th = theta_path(k, prob["pairs"])
                vals.append((ls.logposterior(th), ls.logpost_deriv(th)))
            ls.fuse_gradient = False
            out["evals"] = ls.n_logpost_evals
            out["logpost_first"], out["logpost_last"] = vals[0][0], vals[-1][0]
            lap("value+gradient path")
            ls.set_params(THETA0)
            keep_host = ls.host_results
            muy, Cuy = ls.solve_posterior_covariance()
            out["cuy_trace"] = float(Cuy.diagonal().sum())
            ls.host_results = keep_host
            lap("solve_posterior_covariance")
    out["mu0"] = float(mu[0]) if len(mu) else 0.
    out["cu_checksum"] = float(Cu.sum()) if comm.rank == 0 else 0.       # same value at every N: the sharded path is exact
    if stats is not None:
        # per-family totals over the sparse-solver context and the snapshots' dense contexts
        prof = ls.solver.ctx.profile_read()
        for ctx in {id(s.dctx): s.dctx for s in solvers if s.dctx is not ls.solver.ctx}.values():
            for tag, v in ctx.profile_read().items():
                for key in ("ms", "bytes", "flops", "count"):
                    prof[tag][key] += v[key]
        stats["profile"] = prof
    out["ls"], out["G"], out["solvers"] = ls, G, solvers
    return out


def _drop(res):
    for key in ("ls", "G", "solvers"):
        res.pop(key, None)
    return res


def workload_text(prob):
    "the same description of the workload on both arms"
    base = ("%s: %dD Poisson, %d^%d P1 nodes (%d DOF), %d observations, G cutoff 1e-3; step = G assembly + MG setup + "
            "%d-column PCG solve + W=GZ + Cu=W^T Z + mean" %
            (prob["name"], prob["dim"], prob["nodes"], prob["dim"], prob["V"].dim(), prob["nobs"], prob["nobs"]))
    if prob["kind"] == "posterior":
        base += " + %d snapshots x (MAP L-BFGS-B ftol=1e-15 gtol=1e-10 + posterior mean)" % prob["nsnap"]
    elif prob["kind"] == "dense":
        base += " + %d logposterior+logpost_deriv pairs on a fixed path + solve_posterior_covariance" % prob["pairs"]
    return base


def bytes_h2d(prob):
    V, A = prob["V"], prob["A"]
    n, d = V.coords.shape
    by = A.indptr.nbytes + A.indices.nbytes + A.data.nbytes + V.coords.nbytes + 8 * n + 8 * n + n       # A, coords, int_basis, b, mask
    by += prob["nobs"] * (d + 1) * 12 * 2 + prob["sensors"].nbytes + max(prob["nsnap"], 1) * 8 * prob["nobs"]   # P (CSC + CSR), sensors, data
    return int(by)


def bytes_d2h(prob):
    n = prob["V"].dim()
    by = 8 * prob["nobs"] ** 2 + 8 * prob["nobs"] + 8 * n          # Cu, mu, prior x
    if prob["kind"] == "posterior":
        by += prob["nsnap"] * 8 * n                                 # posterior means
    if prob["kind"] == "dense":
        by += 8 * prob["nobs"] ** 2 + 8 * prob["nobs"]              # C_y, mu_y
    return int(by)


class ClockSampler(threading.Thread):
    "nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        busy = [x for x in sm if x > 0.5 * max(sm)] if sm else []
        return dict(sm_mhz=float(np.median(busy)) if busy else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    need = WORKLOADS[args.workload].get("min_gpus", 1)
    if world < need:
        if rank == 0:
            print(json.dumps({"metric": "time-to-posterior (s)", "unavailable": "workload %s needs >= %d GPUs (Z alone is %.0f GB)" % (
                args.workload, need, 8e-9 * WORKLOADS[args.workload]["nodes"] ** WORKLOADS[args.workload]["dim"] * WORKLOADS[args.workload]["nobs"])}))
        return
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep NCCL's banner off stdout: one JSON line only
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=300))
    import stat_fem_b200 as stat_fem
    from stat_fem_b200 import _lib
    comm = stat_fem.EnsembleComm() if world > 1 else stat_fem.COMM_SELF
    prob = make_problem(args.workload)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    CSR_TAGS = ("csr_mult_A", "csr_jacobi_A", "csr_resid_A")
    csr_mask = 0b111

    def timed_steps(nsteps, resident):
        """K steps bracketed by barrier + synchronize; device time by CUDA events, max over ranks"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = _lib.total_launches()
        prof = {t: dict(ms=0., bytes=0., count=0) for t in CSR_TAGS}
        res = None
        e0.record()
        t0 = time.perf_counter()
        for _ in range(nsteps):
            stats = {}
            hook = _ProfileHook(csr_mask)
            with hook:
                res = gpu_step(prob, comm, resident, stats)
            for t in CSR_TAGS:
                for key in ("ms", "bytes", "count"):
                    prof[t][key] += stats["profile"][t][key]
            _drop(res)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        dev_s = e0.elapsed_time(e1) * 1e-3
        tmax = torch.tensor([dev_s, wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        launches = _lib.total_launches() - launches0
        return float(tmax[0]), float(tmax[1]), launches, prof, res

    # device-resident arm: inputs cached in HBM (uploaded during warm-up), outputs stay on the device
    _lib.DEVICE_CACHE[0] = True
    im = stat_fem.InterpolationMatrix(prob["V"], prob["sensors"])
    im.assemble()
    prob["im"] = im
    for _ in range(args.warmup):
        _drop(gpu_step(prob, comm, True, {}))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_s, wall_s, launches, prof, res = timed_steps(args.steps, True)
    clocks = sampler.stop() if rank == 0 else {}
    # end-to-end arm: host buffers in, host arrays out, nothing cached
    _lib.DEVICE_CACHE[0] = False
    for obj in (prob["V"], prob["A"], prob["b"]):
        obj.__dict__.pop("_dev_cache", None)
    prob.pop("im")
    e2e_steps = max(1, min(args.steps, 2))
    _drop(gpu_step(prob, comm, False, {}))
    _, e2e_wall, _, _, _ = timed_steps(e2e_steps, False)

    # per-family breakdown of one more (untimed) resident step, for profiles/ and the log; --parity checks its results
    _lib.DEVICE_CACHE[0] = True
    hook = _ProfileHook(0xffffffff)
    stats = {}
    with hook:
        r = gpu_step(prob, comm, True, stats)
    full_prof = stats["profile"]
    gprof = r["G"].G.ctx.profile_read()
    _lib.DEVICE_CACHE[0] = False
    total_mem = torch.cuda.get_device_properties(local_rank).total_memory
    peak_gb = [(total_mem - prob["_min_free"][0]) / 1e9, torch.cuda.max_memory_allocated() / 1e9]
    if world > 1:
        allpk = [None] * world
        dist.all_gather_object(allpk, peak_gb)
    else:
        allpk = [peak_gb]
    parity_inputs = _parity_capture(prob, r) if (args.parity and rank == 0) else None
    _drop(r)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    csr_ms = sum(prof[t]["ms"] for t in CSR_TAGS)
    csr_bytes = sum(prof[t]["bytes"] for t in CSR_TAGS)
    csr_count = sum(prof[t]["count"] for t in CSR_TAGS)
    achieved = csr_bytes / (csr_ms * 1e-3) / 1e9 if csr_ms > 0 else 0.
    line = {
        "metric": "time-to-posterior (s)", "value": dev_s / args.steps, "unit": "s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps,
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(prob),
                   "parallelism": "observation columns sharded over %d GPU(s)" % world,
                   "l2": "working set (>= %.1f GB per vector block) exceeds the 126 MB L2; no flush needed" %
                         (8e-9 * prob["V"].dim() * min(prob["nobs"], 384)),
                   "G_nnz_per_row": res["nnzG"] / prob["V"].dim(),
                   "cg_iterations_per_block": res["iters"], "map_evaluations": res["evals"],
                   "cu_checksum": res["cu_checksum"],
                   "theta_snapshot0": [float(t) for t in res["thetas"][0]] if res["thetas"] else None,
                   "peak_device_memory_gb_per_rank": [round(p[0], 2) for p in allpk],
                   "peak_torch_allocated_gb_per_rank": [round(p[1], 2) for p in allpk]},
        "e2e": {"value": e2e_wall / e2e_steps, "unit": "s", "h2d_bytes_per_step": bytes_h2d(prob),
                "d2h_bytes_per_step": bytes_d2h(prob), "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "tileop_dmma_kernel on the stiffness matrix (staged-tile SpMM on FP64 tensor cores, TMA-staged: q=Ap+dot, fused double Jacobi sweep, Jacobi sweep, residual)",
                     "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": _read_traffic(), "launches": int(csr_count), "avg_launch_ms": csr_ms / max(csr_count, 1),
                     "peak_source": peak_src},
        "phases_ms": {k: round(v["ms"], 3) for k, v in full_prof.items() if v["count"]},
        "phases_gcov_ms": {k: round(v["ms"], 3) for k, v in gprof.items() if v["count"]},
        "phases_tflops": {k: round(v["flops"] / v["ms"] * 1e-9, 2) for k, v in full_prof.items() if v["count"] and v["flops"] > 0 and v["ms"] > 0},
    }
    for k, v in res.items():
        if k in ("logpost_first", "logpost_last", "cuy_trace", "x_norm", "mu0", "map_messages", "evals_per_fit", "eval_cache_hits_per_fit"):
            line["config"][k] = v
    if parity_inputs is not None:
        line["parity"] = _parity_check(prob, parity_inputs)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference(prob)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


class _ProfileHook(object):
    """turn the library's per-family event timers on for every context created inside the block"""

    def __init__(self, mask):
        self.mask = mask

    def __enter__(self):
        from stat_fem_b200 import _lib
        self._orig = _lib.Context.__init__
        mask = self.mask
        orig = self._orig

        def init(ctx_self, *a, **kw):
            orig(ctx_self, *a, **kw)
            ctx_self.profile_enable(mask)
        _lib.Context.__init__ = init
        for c in _lib._live:
            if c.handle:
                c.profile_enable(mask)
        return self

    def __exit__(self, *a):
        from stat_fem_b200 import _lib
        _lib.Context.__init__ = self._orig
        return False


def _read_traffic():
    "DRAM bytes per launch of the dominant kernel (tileop_dmma_kernel<TM_JACOBI>, 384 columns) from the committed ncu capture, if any"
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))["tileop_kernel_bytes_per_launch"]
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ parity (checker leg)
def _parity_capture(prob, r):
    "host copies of what the checker compares (taken from the last resident step on rank 0)"
    ls, G = r["ls"], r["G"]
    ls.host_results = True
    ls._local_eval = True           # rank 0 alone evaluates here: no per-evaluation broadcast (it would wait for the other ranks)
    cap = dict(mu=np.asarray(ls.mu), Cu=ls._Cu_dev.cpu().numpy(), G=G.G.to_host(), P=ls.im.interp)
    dense = {}
    if prob["nobs"] <= 16384:
        ls.fuse_gradient = True
        dense["logpost"] = float(ls.logposterior(THETA0))
        dense["logpost_deriv"] = ls.logpost_deriv(THETA0)
        ls.fuse_gradient = False
        ls.set_params(THETA0)
        dense["muy"], dense["Cuy"] = ls.solve_posterior_covariance()
    cap["dense"] = dense
    if prob["kind"] == "posterior":
        cap["thetas"] = [s.params.copy() for s in r["solvers"]]
    return cap


def _parity_check(prob, cap):
    """CPU oracle on a subsample: tolerances 1e-8 on mu / Cu / posterior mean / covariance, 1e-6 on the
    log-posterior and its gradient, G pattern identical"""
    from oracle import fullsize
    import scipy.sparse as sp
    V, A = prob["V"], prob["A"]
    n = V.dim()
    rng = np.random.default_rng(12345)
    ip, ix, v = cap["G"]
    rows = np.unique(np.concatenate([rng.choice(n, 48, replace=False), [0, 1, prob["nodes"] - 1, prob["nodes"], n - 1]]))
    pat, verr = fullsize.check_G_rows(V.coords, V.integrate_basis_functions(), prob["sigma_f"], prob["l_f"], prob["cutoff"],
                                      prob["reg"], ip, ix, v, rows)
    out = dict(G_rows_checked=int(len(rows)), G_pattern_identical=bool(pat), G_values_max_rel=verr)
    Gs = sp.csr_matrix((v, ix, ip), shape=(n, n))
    cols = sorted(int(c) for c in rng.choice(prob["nobs"], 8, replace=False))
    direct = "lu" if (prob["dim"] <= 2 and n <= 1_200_000) else "dst"
    dense = cap["dense"] if prob["nobs"] <= (4096 if prob["kind"] != "dense" else 16384) or os.environ.get("SFEM_PARITY_DENSE") else \
        {k: cap["dense"][k] for k in ("logpost",) if k in cap["dense"]}
    out.update(fullsize.check_prior(A.to_scipy(), A.bc_nodes(), prob["b"].data, Gs, cap["P"], prob["sensors"], prob["datas"][0],
                                    prob["unc"], cap["mu"], cap["Cu"], cols, prob["nodes"], prob["dim"], direct, THETA0, dense))
    tol = {"Cu_rows_max_rel": 1e-8, "mu_rel": 1e-8, "posterior_mean_rel": 1e-8, "posterior_cov_rel": 1e-8,
           "logpost_rel": 1e-6, "logpost_deriv_rel": 1e-6, "G_values_max_rel": 1e-13}
    out["pass"] = bool(pat and all(out[k] <= t for k, t in tol.items() if k in out))
    return out


# ------------------------------------------------------------------------------------------------ CPU reference arm
_G = {}
_REF = {}       # per-process cache of the one-off parts of the reference arm (LU factor, kd-tree variant, MAP evaluation count)


def _g_row_worker(i):
    import oracle
    p = _G["p"]
    d, _ = oracle.compute_G_rows(p["coords"], p["b"], p["sigma_f"], p["l_f"], p["cutoff"], p["reg"], rows=[i])
    return len(d[i][1])


def _solve_worker(j):
    import oracle
    col = np.asarray(_G["P"][:, j].todense()).ravel()
    x = oracle.solve_forcing_covariance(_G["Gs"], _G["solver"], col)     # 2 solves + 1 SpMV, solving_utils.py:49-53
    return _G["PT"] @ x


def _reference_setup(prob):
    """One-off parts (once per process, their time is added to every step's value): the sparse factorisation the
    reference's Firedrake solver performs once per LinearSolver (LinearSolver.py:136-140), the kd-tree G variant, and
    the number of L-BFGS-B evaluations of one MAP fit, measured by running the oracle's estimate_params_MAP at the
    bench tolerances on a reduced sensor set (256 sensors; the count depends on the 3-parameter landscape, not on n_obs)."""
    if _REF.get("name") == prob["name"]:
        return _REF
    import scipy.sparse as sp
    import oracle
    from oracle.solver import _DirectSolver
    from oracle.fast_poisson import DSTSolver
    V, A = prob["V"], prob["A"]
    n, nobs = V.dim(), prob["nobs"]
    cores = len(os.sched_getaffinity(0))
    rng = np.random.default_rng(0)
    _REF.clear()
    _REF.update(name=prob["name"], cores=cores)
    b = V.integrate_basis_functions()
    _G["p"] = dict(coords=V.coords, b=b, sigma_f=prob["sigma_f"], l_f=prob["l_f"], cutoff=prob["cutoff"], reg=prob["reg"])
    rows = rng.choice(n, min(n, 32 + 1024), replace=False)
    _REF["rows"] = rows
    # optimised CPU variant of the G assembly (cKDTree candidates, same arithmetic, same matrix): one tree build + a
    # per-row cost; two calls with different row counts separate the two (reported for information)
    r_small, r_large = 32, min(n, 32 + 1024)
    t = time.perf_counter()
    oracle.compute_G_csr_fast(V.coords, b, prob["sigma_f"], prob["l_f"], prob["cutoff"], prob["reg"], rows=rows[:r_small])
    t_small = time.perf_counter() - t
    t = time.perf_counter()
    ip, ix, v = oracle.compute_G_csr_fast(V.coords, b, prob["sigma_f"], prob["l_f"], prob["cutoff"], prob["reg"], rows=rows[:r_large])
    t_large = time.perf_counter() - t
    per_row = max(t_large - t_small, 0.) / max(r_large - r_small, 1)
    _REF["t_g_fast"] = max(t_small - r_small * per_row, 0.) + per_row * n / cores       # rows are independent: spread over the cores
    # banded stand-in with the same nnz/row is enough for timing the SpMV inside solve_forcing_covariance
    nnz_row = max(1, int(round(len(ix) / r_large)))
    offs = np.arange(-(nnz_row // 2), nnz_row - nnz_row // 2)
    _G["Gs"] = sp.diags([np.full(n - abs(o), 1e-12) for o in offs], offs, shape=(n, n), format="csr")
    t = time.perf_counter()
    if prob["dim"] <= 2 and n <= 1_200_000:
        _G["solver"] = _DirectSolver(A.to_scipy(), A.bc_nodes())
        _REF["solver_kind"] = "SciPy SuperLU (direct LU, the class of the reference's Firedrake default)"
    else:       # LU fill-in does not fit (3-D) or takes minutes (4M DOF): the DST solver understates the reference's cost
        _G["solver"] = DSTSolver(A.to_scipy(), A.bc_nodes(), prob["nodes"], prob["dim"])
        _REF["solver_kind"] = "DST fast Poisson solver (a sparse LU of this size does not fit: this understates the reference's cost)"
    _REF["t_factor"] = time.perf_counter() - t
    ncols = min(max(2 * cores, 256 if prob["kind"] == "posterior" else 0), nobs)
    cells, w = V.mesh().locate_cells(prob["sensors"][:ncols])
    nodes = V.mesh().cells[cells]
    _G["P"] = sp.csc_matrix((w.ravel(), (nodes.ravel(), np.repeat(np.arange(ncols), nodes.shape[1]))), shape=(n, ncols))
    _G["PT"] = _G["P"].T.tocsr()
    _REF["map_evals"] = None
    if prob["kind"] == "posterior":
        # MAP evaluation count: Cu of the first `ncols` sensors by the reference's loop, then the oracle's L-BFGS-B
        import multiprocessing as mp
        t = time.perf_counter()
        with mp.get_context("fork").Pool(min(cores, ncols)) as pool:
            rows_cu = pool.map(_solve_worker, list(range(ncols)))
        Cu = np.reshape(np.array(rows_cu).flatten(), (ncols, ncols))
        x = _G["solver"].solve(prob["b"].data, bcs_on=True)
        ols = oracle.OracleLinearSolver.__new__(oracle.OracleLinearSolver)
        ols.data = oracle.OracleObsData(prob["sensors"][:ncols], prob["datas"][0][:ncols], prob["unc"])
        ols.Cu, ols.mu, ols.params = Cu, _G["PT"] @ x, None
        ols.solve_prior = lambda: (ols.mu, ols.Cu)
        _, fmin = oracle.estimate_params_MAP(ols, start=np.zeros(3), **MAP_OPTIONS)
        _REF["map_evals"] = int(fmin["nfev"])
        _REF["map_sample"] = "oracle estimate_params_MAP on %d sensors: %d evaluations (%.1f s, once per process)" % (
            ncols, _REF["map_evals"], time.perf_counter() - t)
    return _REF


def cpu_reference(prob):
    """The reference algorithm (oracle/: NumPy/SciPy restatement pinned against the unmodified reference) on the host
    cores, bounded sample, extrapolated linearly to the full workload.  Returns the cpu_baseline dict."""
    import multiprocessing as mp
    import oracle
    ref = _reference_setup(prob)
    cores = ref["cores"]
    V = prob["V"]
    n, nobs, nsnap = V.dim(), prob["nobs"], prob["nsnap"]
    g_rows, cols = min(4 * cores, n), min(2 * cores, nobs)
    rows = ref["rows"]
    ctxmp = mp.get_context("fork")
    with ctxmp.Pool(cores) as pool:
        pool.map(_g_row_worker, [int(i) for i in rows[:cores]])       # start the workers before the clock
        t = time.perf_counter()
        pool.map(_g_row_worker, [int(i) for i in rows[:g_rows]])
        t_g = (time.perf_counter() - t) / g_rows * n                  # O(n^2) row loop, ForcingCovariance.py:159-167
    with ctxmp.Pool(min(cores, cols)) as pool:
        pool.map(_solve_worker, list(range(min(cores, cols))))
        t = time.perf_counter()
        pool.map(_solve_worker, list(range(cols)))
        t_cols = (time.perf_counter() - t) / cols * nobs              # 2 n_obs solves, solving_utils.py:139-142
    parts = {"G_assembly_reference_On2": t_g, "G_assembly_kdtree_variant": ref["t_g_fast"], "solver_setup": ref["t_factor"],
             "cu_solves": t_cols}
    sample = "%d of %d G rows (O(n^2) reference loop), solver set-up in full once per process [%s], %d of %d sensor columns (2 solves + SpMV each)" % (
        g_rows, n, ref["solver_kind"], cols, nobs)
    total = t_g + ref["t_factor"] + t_cols
    if prob["kind"] in ("posterior", "dense"):
        # dense end: one logposterior + one logpost_deriv in the reference's formulation (LinearSolver.py:592-691),
        # measured at min(n_obs, 4096) sensors and scaled cubically above that
        ns = min(nobs, 4096)
        scale3 = (float(nobs) / ns) ** 3
        rng = np.random.default_rng(0)
        od = oracle.OracleObsData(prob["sensors"][:ns], prob["datas"][0][:ns], prob["unc"])
        X = rng.normal(size=(ns, 64))
        ols = oracle.OracleLinearSolver.__new__(oracle.OracleLinearSolver)
        ols.data, ols.Cu, ols.mu, ols.params = od, 1e-6 * (X @ X.T) / 64, np.zeros(ns), None
        t = time.perf_counter()
        ols.logposterior(THETA0)
        ols.logpost_deriv(THETA0)
        t_pair = (time.perf_counter() - t) * scale3
        if prob["kind"] == "posterior":
            t_mean = t_cols / nobs * (0.5 + 2. * nsnap)               # mean solve + 2 x (2 solves) per posterior mean
            t_map = t_pair * ref["map_evals"] * nsnap
            parts.update(mean_and_posterior_solves=t_mean, map=t_map)
            total += t_mean + t_map
            sample += ", 1 logposterior+logpost_deriv pair at n_obs=%d%s x %d evaluations per MAP fit [%s] x %d snapshots" % (
                ns, " (scaled cubically to %d)" % nobs if scale3 > 1 else "", ref["map_evals"], ref["map_sample"], nsnap)
        else:
            t_path = t_pair * prob["pairs"]
            t_cov = t_pair * (2. / 3. + 4.) / (1. / 3. + 6. + 1. / 3.)    # 4.67 n^3 against the pair's 6.67 n^3 (LinearSolver.py:351-367)
            parts.update(mean_solve=t_cols / nobs * 0.5, value_gradient_path=t_path, posterior_covariance=t_cov)
            total += t_cols / nobs * 0.5 + t_path + t_cov
            sample += ", 1 logposterior+logpost_deriv pair at n_obs=%d (scaled cubically to %d) x %d pairs + solve_posterior_covariance by flop ratio" % (
                ns, nobs, prob["pairs"])
    return {"value": total, "unit": "s", "cores": cores, "kind": "port",
            "sample": sample + "; extrapolated linearly", "parts_s": parts}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    prob = make_problem(args.workload)
    t0 = time.perf_counter()
    _reference_setup(prob)
    for _ in range(min(args.warmup, 1)):
        cpu_reference(prob)
    vals = []
    last = None
    for _ in range(args.steps):
        last = cpu_reference(prob)
        vals.append(last["value"])
    val = float(np.mean(vals))
    last["value"] = val
    line = {"impl": "reference", "metric": "time-to-posterior (s)", "value": val, "unit": "s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * val, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(prob), "parallelism": "reference algorithm (oracle port) on %d host cores; each step = a bounded sample, extrapolated linearly to the full workload" % last["cores"]},
            "cpu_baseline": last, "gpu_launches": 0,
            "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parity", action="store_true", help="check the last step against the CPU oracle on a subsample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
