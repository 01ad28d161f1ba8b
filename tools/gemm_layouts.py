"""FP64 DMMA GEMM rate by operand layout (sfem_dgemm): NN, NT, TN at the shapes of the blocked Cholesky's trailing update
(M = N, K = 512 / 128) and of Cu = W^T Z.   python tools/gemm_layouts.py [M] """
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import stat_fem_b200
from stat_fem_b200 import _lib
ctx = _lib.get_context(); lib = _lib.lib()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192


def rate(opA, opB, M, N, K, lower=False, beta=1.0, reps=5):
    A = ctx.empty((M, K) if opA == _lib.OP_N else (K, M)).normal_()
    B = ctx.empty((K, N) if opB == _lib.OP_N else (N, K)).normal_()
    C = ctx.zeros((M, N))
    fn = lib.sfem_dgemm_lower if lower else lib.sfem_dgemm

    def go():
        if lower:
            ctx.check(fn(ctx.handle, opA, opB, M, K, 1.0, _lib.ptr(A), int(A.stride(0)), _lib.ptr(B), int(B.stride(0)), beta, _lib.ptr(C), N))
        else:
            ctx.check(fn(ctx.handle, opA, opB, M, N, K, 1.0, _lib.ptr(A), int(A.stride(0)), _lib.ptr(B), int(B.stride(0)), beta, _lib.ptr(C), N))
    go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        go()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2. * M * N * K * (0.5 if lower else 1.)
    return ms, fl / ms * 1e-9


N_, T_ = _lib.OP_N, _lib.OP_T
for K in (128, 512, 4096):
    for name, oa, ob in (("NN", N_, N_), ("NT", N_, T_), ("TN", T_, N_)):
        ms, tf = rate(oa, ob, M, M, K)
        print("M=N=%d K=%d %s full : %8.3f ms %6.2f TFLOP/s" % (M, K, name, ms, tf))
    ms, tf = rate(N_, T_, M, M, K, lower=True)
    print("M=N=%d K=%d NT lower: %8.3f ms %6.2f TFLOP/s (on the lower half)" % (M, K, ms, tf))
