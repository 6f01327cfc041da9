"""cuBLAS DGEMM peak on this box (torch.matmul fp64), burst and ~2 s sustained. Measurement helper only."""
import json, time, torch
torch.backends.cuda.matmul.allow_tf32 = False
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2):
    c = a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * n**3 / best * 1e-9
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
k = 0
t0 = time.time()
while time.time() - t0 < 2.0:
    c = a @ b; k += 1
    if k % 4 == 0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sus = 2 * n**3 * k / e0.elapsed_time(e1) * 1e-9
print(json.dumps({"dgemm_burst_tflops": burst, "dgemm_sustained_tflops": sus, "n": n, "gpu": torch.cuda.get_device_name(0)}))
