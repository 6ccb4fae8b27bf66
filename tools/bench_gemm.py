"""Times the K5 GEMM on the encoder's shapes (CUDA events, L2-cold via rotating buffers)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from turbo_whisper_workspace_b200 import ops

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
M = B * 1500
res = []
for (N, K, act, name) in [(3840, 1280, 0, "qkv"), (1280, 1280, 0, "out"), (5120, 1280, 1, "fc1"), (1280, 5120, 0, "fc2")]:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    for _ in range(3):
        ops.gemm(a, w, rows=M, bias=bias, act=act, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for _ in range(iters):
        ops.gemm(a, w, rows=M, bias=bias, act=act, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS for comparison (library baseline)
    for _ in range(3):
        torch.nn.functional.linear(a, w)
    e0.record()
    for _ in range(iters):
        torch.nn.functional.linear(a, w)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    res.append(dict(name=name, M=M, N=N, K=K, ms=round(ms, 4), tflops=round(tf, 1), cublas_ms=round(ms2, 4),
                    cublas_tflops=round(2.0 * M * N * K / ms2 / 1e9, 1)))
    print(res[-1], flush=True)
