"""cuBLAS TF32 / FP32 matmul peaks measured the way the driver measured bf16 (MEASURED_PEAKS.json `how`):
torch.matmul 8192^3, best of 10 (burst) and back to back for 4 s (sustained).  Writes gpurun_out/peaks_tf32.json;
the committed copy is profiles/r2_peaks_tf32.json (bench.py reads it for the tensor roofline denominator)."""
import json
import os
import time

import torch


def measure(n, allow_tf32, seconds=4.0):
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    torch.backends.cudnn.allow_tf32 = allow_tf32
    a = torch.randn((n, n), device='cuda')
    b = torch.randn((n, n), device='cuda')
    c = torch.empty((n, n), device='cuda')
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    k = 0
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(10):
            torch.matmul(a, b, out=c)
        k += 10
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / k
    fl = 2.0 * n ** 3
    return fl / best / 1e9, fl / sus / 1e9


if __name__ == '__main__':
    out = {'gpu_name': torch.cuda.get_device_name(0), 'torch': torch.__version__,
           'how': 'torch.matmul fp32 inputs 8192^3 (2*N^3): best of 10 (burst) and back to back for 4 s (sustained); '
                  'allow_tf32 True = cuBLAS TF32 tensor-core path, False = cuBLAS SGEMM'}
    out['tf32_tflops'], out['tf32_tflops_sustained'] = measure(8192, True)
    out['fp32_tflops'], out['fp32_tflops_sustained'] = measure(8192, False)
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(out, open('gpurun_out/peaks_tf32.json', 'w'), indent=1)
    print(json.dumps(out))
