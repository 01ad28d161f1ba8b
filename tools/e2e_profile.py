"""cProfile of one end-to-end step of bench.py (host buffers in, host arrays out): where the host-side time goes.
    python tools/e2e_profile.py [workload]"""
import cProfile
import os
import pstats
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import stat_fem_b200 as stat_fem
from stat_fem_b200 import _lib

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
prob = bench.make_problem(name)
comm = stat_fem.COMM_SELF
_lib.DEVICE_CACHE[0] = False
for _ in range(2):
    r = bench.gpu_step(prob, comm, False, {})
    r.pop("ls"); r.pop("G")
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
r = bench.gpu_step(prob, comm, False, {})
r.pop("ls"); r.pop("G")
del r
torch.cuda.synchronize()
pr.disable()
print("step wall %.3f s" % (time.perf_counter() - t0))
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
st.sort_stats("tottime").print_stats(25)
