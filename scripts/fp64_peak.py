"""fp64 tensor-core ceiling of this GPU, independent of the bench: cuBLAS DGEMM (torch.matmul, float64) at 8192^3 and 16384^3,
burst (best of 5) and sustained (back to back for ~4 s).  MEASURED_PEAKS.json holds only bf16 and HBM."""
import json, time, torch
dev = torch.device("cuda", 0)
out = {"gpu": torch.cuda.get_device_name(0)}
for n in (8192, 16384):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    reps = max(3, int(4000 / best))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b)
    e1.record(); torch.cuda.synchronize()
    out[f"dgemm_{n}"] = {"burst_tflops": round(2.0 * n ** 3 / (best * 1e-3) / 1e12, 2), "sustained_tflops": round(2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12, 2), "reps": reps}
    del a, b
print(json.dumps(out))
