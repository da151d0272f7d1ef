"""cuBLAS FP64 GEMM throughput on this GPU (the FP64-tensor roofline denominator:
MEASURED_PEAKS.json has no FP64 entry).  Writes gpurun_out/fp64_peaks.json."""
import json
import os
import sys

import torch


def bench(fn, flop, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return flop / best * 1e-9


def main():
    dev = torch.device('cuda:0')
    out = {'gpu': torch.cuda.get_device_name(0)}
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    out['dgemm_8192_tflops'] = bench(lambda: torch.matmul(a, b), 2.0 * n ** 3)
    # sustained: back to back for ~3 s
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 60
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b)
    e1.record()
    torch.cuda.synchronize()
    out['dgemm_8192_tflops_sustained'] = 2.0 * n ** 3 * reps / e0.elapsed_time(e1) * 1e-9
    del a, b
    n = 4096
    a = torch.randn(n, n, dtype=torch.complex128, device=dev)
    b = torch.randn(n, n, dtype=torch.complex128, device=dev)
    out['zgemm_4096_tflops'] = bench(lambda: torch.matmul(a, b), 8.0 * n ** 3)
    del a, b
    # c4-like shapes: real L [11664 x 500] times x [500 x 16384]
    a = torch.randn(11664, 500, dtype=torch.float64, device=dev)
    b = torch.randn(500, 16384, dtype=torch.float64, device=dev)
    out['dgemm_vhs_c4_tflops'] = bench(lambda: torch.matmul(a, b), 2.0 * 11664 * 500 * 16384)
    a = torch.randn(500, 4536, dtype=torch.float64, device=dev)
    b = torch.randn(4536, 16384, dtype=torch.float64, device=dev)
    out['dgemm_fb_c4_tflops'] = bench(lambda: torch.matmul(a, b), 2.0 * 500 * 4536 * 16384)
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/fp64_peaks.json', 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == '__main__':
    sys.exit(main())
