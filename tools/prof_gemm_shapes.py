"""The three GEMM shapes of the dense end, one launch each, for `ncu` (profiles/r02_dgemm_ncu.csv):
  cu    : Cu = W^T Z, symmetric part -- TN, M = N = 4096 (tiles on and below the diagonal), K = 1 048 576 (split-free)
  syrk  : trailing update of the blocked Cholesky -- NT, lower, M = N = 8192, K = 512, beta = 1
  k128  : in-block update -- NT, lower, M = N = 8192, K = 128, beta = 1
python tools/prof_gemm_shapes.py [cu|syrk|k128|all]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stat_fem_b200
from stat_fem_b200 import _lib
ctx = _lib.get_context(); lib = _lib.lib()
which = sys.argv[1] if len(sys.argv) > 1 else "all"
N_, T_ = _lib.OP_N, _lib.OP_T


def timed(fn, label, flops):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%-5s %9.3f ms  %6.2f TFLOP/s" % (label, ms, flops / ms * 1e-9))


if which in ("cu", "all"):
    n, k = 1048576, 4096
    W = ctx.empty((n, k)).normal_(); Z = ctx.empty((n, k)).normal_(); C = ctx.empty((k, k))
    timed(lambda: ctx.check(lib.sfem_dgemm_lower(ctx.handle, T_, N_, k, n, 1.0, _lib.ptr(W), k, _lib.ptr(Z), k, 0.0, _lib.ptr(C), k)),
          "cu", 2. * n * k * k * (k / 128 + 1) / (2. * k / 128))
    del W, Z
for name, K in (("syrk", 512), ("k128", 128)):
    if which in (name, "all"):
        M = 8192
        A = ctx.empty((M, K)).normal_(); C = ctx.zeros((M, M))
        timed(lambda: ctx.check(lib.sfem_dgemm_lower(ctx.handle, N_, T_, M, K, -1.0, _lib.ptr(A), K, _lib.ptr(A), K, 1.0, _lib.ptr(C), M)),
              name, 2. * M * M * K * (M / 128 + 1) / (2. * M / 128))
