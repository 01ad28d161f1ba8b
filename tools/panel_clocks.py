"""Phase timing of the Cholesky panel kernel (clock64 stamps).  Needs a library built with the stamps enabled:
    make -C stat-fem_b200/csrc clean && make -C stat-fem_b200/csrc NVCC_EXTRA=-DSFEM_PANEL_CLOCKS"""
import sys, ctypes as C
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import stat_fem_b200
from stat_fem_b200 import _lib
ctx = _lib.get_context(); lib = _lib.lib()
n = 512
rng = np.random.default_rng(0); X = rng.normal(size=(n, n + 5)); S = X @ X.T + n * 0.1 * np.eye(n)
for rep in range(3):
    dS = ctx.to_device(S, np.float64); dinv = ctx.empty((lib.sfem_potrf_dinv_doubles(n),)); info = C.c_int(0)
    ctx.check(lib.sfem_potrf(ctx.handle, n, _lib.ptr(dS), n, _lib.ptr(dinv), C.byref(info)))
torch.cuda.synchronize()
buf = (C.c_longlong * 16)()
lib.sfem_debug_panel_clocks(buf)
c = list(buf)
names = ["load", "-", "(i) diag factor jb=0", "(ii)+trsm jb=0", "(iii) syrk jb=0", "rest of phase 1 (jb=1..3)", "store L (last block column)", "diag blocks <- inverse", "phase 2B", "store Dinv"]
for i in range(10):
    print("%-28s %7d cycles  %6.2f us" % (names[i], c[i + 1] - c[i], (c[i + 1] - c[i]) / 1965.))
print("total %.2f us" % ((c[10] - c[0]) / 1965.))
