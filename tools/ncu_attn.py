import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from turbo_whisper_workspace_b200 import ops
dev = torch.device("cuda:0")
B = 24; M = B * 1500
qkv = torch.randn(M, 3840, device=dev).to(torch.bfloat16); qkv[:, :1280] *= 0.35
out = torch.empty(M, 1280, dtype=torch.bfloat16, device=dev)
for _ in range(4):
    ops.attention_enc(qkv, B, 1500, 20, out=out)
torch.cuda.synchronize(); print("ok")
