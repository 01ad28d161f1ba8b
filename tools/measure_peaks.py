"""One-off: measure the FP64 denominators MEASURED_PEAKS.json lacks (cuBLAS DGEMM, f64 copy/triad).
Library calls here are only denominators for roofline fractions, never the product path."""
import json, torch, time
dev = torch.device("cuda:0")
out = {"gpu": torch.cuda.get_device_name(0)}
def ev_time(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) * 1e-3)
    return best
for n in (4096, 8192, 16384):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    t = ev_time(lambda: torch.matmul(a, b))
    out[f"dgemm_{n}_tflops"] = 2 * n**3 / t / 1e12
    t = ev_time(lambda: torch.matmul(a.t(), b))
    out[f"dgemm_tn_{n}_tflops"] = 2 * n**3 / t / 1e12
    del a, b
# sustained dgemm 3 s
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
torch.cuda.synchronize(); t0 = time.time(); cnt = 0
s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True); s.record()
while time.time() - t0 < 3.0:
    torch.matmul(a, b); cnt += 1
    if cnt % 8 == 0: torch.cuda.synchronize()
e.record(); torch.cuda.synchronize()
out["dgemm_8192_tflops_sustained"] = cnt * 2 * n**3 / (s.elapsed_time(e) * 1e-3) / 1e12
del a, b
# tall-skinny TN: C[4096x4096] = W^T Z, K = 262144
K = 262144; m = 4096
w = torch.randn(K, m, dtype=torch.float64, device=dev); z = torch.randn(K, m, dtype=torch.float64, device=dev)
t = ev_time(lambda: torch.matmul(w.t(), z), reps=3, warm=1)
out["dgemm_tn_tallskinny_tflops"] = 2 * m * m * K / t / 1e12
del w, z
x = torch.empty(1 << 29, dtype=torch.float64, device=dev).normal_(); y = torch.empty_like(x)
t = ev_time(lambda: y.copy_(x), reps=10)
out["f64_copy_gbs"] = 2 * x.numel() * 8 / t / 1e9
t = ev_time(lambda: torch.add(x, y, alpha=2.0, out=y), reps=10)
out["f64_axpy_gbs"] = 3 * x.numel() * 8 / t / 1e9
# potrf via cusolver for reference
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device=dev); spd = a @ a.t() + n * torch.eye(n, dtype=torch.float64, device=dev)
t = ev_time(lambda: torch.linalg.cholesky(spd), reps=3, warm=1)
out["cusolver_potrf_8192_tflops"] = n**3 / 3 / t / 1e12
print(json.dumps(out, indent=1))
open("gpurun_out/fp64_peaks.json", "w").write(json.dumps(out, indent=1))
