"""Micro-benchmarks of the encoder kernels on the large-v3-turbo shapes (CUDA events; operands far larger than L2
or rotated).  Usage: python tools/bench_kernels.py [B]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from turbo_whisper_workspace_b200 import ops

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
M = B * 1500


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


only = sys.argv[2] if len(sys.argv) > 2 else ""
for (N, K, act, resid, name) in [] if only == "attention" else [(3840, 1280, 0, 0, "qkv"), (1280, 1280, 0, 1, "out+res"), (5120, 1280, 1, 0, "fc1+gelu"),
                                 (1280, 5120, 0, 1, "fc2+res"), (10240, 1280, 0, 0, "cross_kv")]:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    if resid:
        out = torch.zeros(M, N, dtype=torch.float32, device=dev)
        fn = lambda: ops.gemm(a, w, rows=M, bias=bias, act=act, resid=out, resid_ld=N, out=out)
    else:
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        fn = lambda: ops.gemm(a, w, rows=M, bias=bias, act=act, out=out)
    ms = timed(fn)
    ms2 = timed(lambda: torch.nn.functional.linear(a, w))
    print(json.dumps(dict(kernel="gemm", name=name, M=M, N=N, K=K, ms=round(ms, 4), tflops=round(2.0 * M * N * K / ms / 1e9, 1),
                          cublas_ms=round(ms2, 4), cublas_tflops=round(2.0 * M * N * K / ms2 / 1e9, 1))), flush=True)
    del a, w, out

qkv = torch.randn(M, 3840, device=dev).to(torch.bfloat16)
qkv[:, :1280] *= 0.35
out = torch.empty(M, 1280, dtype=torch.bfloat16, device=dev)
ms = timed(lambda: ops.attention_enc(qkv, B, 1500, 20, out=out))
fl = 4.0 * 1500 * 1500 * 1280 * B
q, k, v = [t.view(B, 1500, 20, 64).transpose(1, 2) for t in qkv.split(1280, dim=1)]
ms2 = timed(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=1.0))
# accuracy against fp32 softmax(QK^T)V of the same bf16 inputs (TF32 off), first two windows
torch.backends.cuda.matmul.allow_tf32 = False
ref = torch.nn.functional.scaled_dot_product_attention(q[:2].float(), k[:2].float(), v[:2].float(), scale=1.0)
got = out.view(B, 1500, 20, 64).transpose(1, 2)[:2].float()
sd = torch.nn.functional.scaled_dot_product_attention(q[:2], k[:2], v[:2], scale=1.0).float()
rms = float(ref.pow(2).mean().sqrt())
print(json.dumps(dict(kernel="attention_enc", lib=os.environ.get("TWB200_LIB", "default"), B=B, ms=round(ms, 4),
                      tflops=round(fl / ms / 1e9, 1), sdpa_ms=round(ms2, 4), sdpa_tflops=round(fl / ms2 / 1e9, 1),
                      max_err_over_rms=round(float((got - ref).abs().max()) / rms, 5),
                      mean_err_over_rms=round(float((got - ref).abs().mean()) / rms, 6),
                      sdpa_bf16_max_err_over_rms=round(float((sd - ref).abs().max()) / rms, 5),
                      sdpa_bf16_mean_err_over_rms=round(float((sd - ref).abs().mean()) / rms, 6))), flush=True)
if only == "attention":
    sys.exit(0)
x = torch.randn(M, 1280, device=dev)
g = torch.ones(1280, device=dev)
xo = torch.empty(M, 1280, dtype=torch.bfloat16, device=dev)
ms = timed(lambda: ops.layernorm(x, g, g, out=xo))
print(json.dumps(dict(kernel="layernorm", ms=round(ms, 4), gbs=round(M * 1280 * 6 / ms / 1e6, 1))), flush=True)
pcm = torch.randn(B, 480000, device=dev) * 0.1
lm = ops.LogMel(dev, B)
of = torch.empty(B, 128, 3000, device=dev)
ms = timed(lambda: lm(pcm, None, out_f32=of))
print(json.dumps(dict(kernel="logmel", ms=round(ms, 4), gbs=round(B * 3.456e6 / ms / 1e6, 1))), flush=True)
