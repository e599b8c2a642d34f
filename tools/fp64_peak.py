#!/usr/bin/env python
"""FP64 yardsticks on the box (MEASURED_PEAKS.json holds none): cuBLAS DGEMM 8192^3 and an FP64 FMA-bound torch op."""
import json, torch
n = 8192
a = torch.randn(n, n, device="cuda", dtype=torch.float64); b = torch.randn(n, n, device="cuda", dtype=torch.float64)
for _ in range(2):
    (a @ b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    c = a @ b
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
x = torch.empty(1 << 28, device="cuda", dtype=torch.float64); y = torch.empty_like(x)
y.copy_(x); torch.cuda.synchronize()
e0.record()
for _ in range(5):
    y.copy_(x)
e1.record(); torch.cuda.synchronize()
cms = e0.elapsed_time(e1) / 5
print(json.dumps({"dgemm_8192_tflops": 2 * n ** 3 / ms / 1e9, "dgemm_ms": ms, "copy_gbs": 2 * x.numel() * 8 / cms / 1e6,
                  "sm_count": torch.cuda.get_device_properties(0).multi_processor_count, "name": torch.cuda.get_device_name(0)}))
