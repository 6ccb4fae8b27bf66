"""One encoder GEMM shape for `ncu --set full`: python tools/ncu_gemm.py [name]  (qkv|out|fc1|fc2)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from turbo_whisper_workspace_b200 import ops
dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "qkv"
N, K, act, resid = {"qkv": (3840, 1280, 0, 0), "out": (1280, 1280, 0, 1), "fc1": (5120, 1280, 1, 0), "fc2": (1280, 5120, 0, 1)}[name]
M = 24 * 1500
a = torch.randn(M, K, device=dev).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(N, device=dev)
out = torch.zeros(M, N, dtype=torch.float32 if resid else torch.bfloat16, device=dev)
for _ in range(5):
    ops.gemm(a, w, rows=M, bias=bias, act=act, resid=out if resid else None, resid_ld=N, out=out)
torch.cuda.synchronize()
print("ok")
