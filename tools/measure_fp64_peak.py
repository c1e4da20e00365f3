"""Measures the practical FP64 GEMM peak of the box (cuBLAS DGEMM 8192^3 through torch) — the roofline denominator
for the FP64 tensor-core contractions (SURVEY 6 asked for it; MEASURED_PEAKS.json only has bf16 and HBM)."""
import json, sys, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(3): c = a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); c = a @ b; e.record(); torch.cuda.synchronize()
    best = min(best, s.elapsed_time(e))
print(json.dumps({"fp64_dgemm_tflops": 2 * n**3 / best / 1e9, "ms": best, "how": "torch.matmul float64 8192^3 (cuBLAS), best of 5, CUDA events", "gpu": torch.cuda.get_device_name(0)}))
