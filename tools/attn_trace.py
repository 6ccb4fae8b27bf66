import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from turbo_whisper_workspace_b200 import ops, _lib
dev = torch.device("cuda:0")
B = 24; M = B * 1500
qkv = torch.randn(M, 3840, device=dev).to(torch.bfloat16); qkv[:, :1280] *= 0.35
out = torch.empty(M, 1280, dtype=torch.bfloat16, device=dev)
for _ in range(2): ops.attention_enc(qkv, B, 1500, 20, out=out)
buf = torch.zeros(192, dtype=torch.int64, device=dev)
lib = C.CDLL(_lib.LIB_PATH)
lib.tw_attention_enc_set_trace(C.c_void_p(buf.data_ptr()))
ops.attention_enc(qkv, B, 1500, 20, out=out); torch.cuda.synchronize()
t = buf.cpu().tolist()
t0 = min(x for x in t if x > 0)
mma = [x - t0 for x in t[:64] if x > 0]; sm = [x - t0 for x in t[64:128] if x > 0]
print("MMA thread stamps (start, after q/k0, then [p_full_A, p_full_B] per j):"); print(mma[:60])
print("softmax A row0 (keys 0-63) stamps per kv tile: [S ready, S loaded, max agreed + PV(j-1) done, P stored]")
for j in range(0, len(sm) - 3, 4):
    a = sm[j:j + 4]
    print(j // 4, a, "ld", a[1] - a[0], "max+o_wait", a[2] - a[1], "exp+st", a[3] - a[2], "period", (a[0] - sm[j - 4]) if j else None)
