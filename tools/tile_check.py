"""Compare the two forms (scalar / FP64 tensor-core) of every staged-tile operator of a multigrid hierarchy, mode by
mode, with a device synchronisation after each launch (debugging aid; the pytest version is
tests/test_gpu_kernels.py::test_tile_tensor_form_matches_scalar_form).   python tools/tile_check.py [nodes] [k]"""
import os
import sys
import ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from stat_fem_b200 import fem, _lib
from stat_fem_b200.solving_utils import SparseSolver

MODES = {"mult": 0, "mult_add": 1, "mult_dot": 2, "jacobi": 3, "jacobi_dot": 4, "resid": 5, "presmooth2": 6}
nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 200
k = int(sys.argv[2]) if len(sys.argv) > 2 else 70
V, A, b = fem.poisson_problem(nodes, 2)
ls = SparseSolver(A)
ctx, lib = ls.ctx, _lib.lib()
gen = torch.Generator(device="cuda").manual_seed(1)
for level in range(len(ls.mg_levels)):
    for which, modes in ((0, ("mult", "resid", "jacobi", "presmooth2", "mult_dot", "jacobi_dot")), (1, ("mult_add",))):
        rows, cols, form = C.c_int64(0), C.c_int64(0), C.c_int(0)
        ctx.check(lib.sfem_tile_apply_one(ctx.handle, level, which, 0, 0, None, None, None, None, None, 0, C.byref(rows), C.byref(cols), C.byref(form)))
        print("level %d op %d: %d x %d, k-tiles %d" % (level, which, rows.value, cols.value, form.value), flush=True)
        if form.value == 0 or rows.value == 0:
            continue
        X = torch.rand((cols.value, k), generator=gen, dtype=torch.float64, device="cuda") - 0.5
        B = torch.rand((rows.value, k), generator=gen, dtype=torch.float64, device="cuda") - 0.5
        for name in modes:
            out = []
            for tf in (0, 1):
                Y = B.clone() if name == "mult_add" else torch.full((rows.value, k), float("nan"), dtype=torch.float64, device="cuda")
                parts = torch.zeros((1184, k), dtype=torch.float64, device="cuda")
                npart = C.c_int(0)
                ctx.check(lib.sfem_tile_apply_one(ctx.handle, level, which, MODES[name], k, _lib.ptr(X), _lib.ptr(B), _lib.ptr(Y),
                                                  _lib.ptr(parts), C.byref(npart), tf, None, None, None))
                torch.cuda.synchronize()
                out.append((Y.cpu().numpy(), parts[:npart.value].sum(dim=0).cpu().numpy()))
            e = np.max(np.abs(out[0][0] - out[1][0])) / np.max(np.abs(out[0][0]))
            d = np.max(np.abs(out[0][1] - out[1][1])) / max(np.max(np.abs(out[0][1])), 1e-300)
            print("   %-11s rel diff %.2e   dots %.2e" % (name, e, d), flush=True)
