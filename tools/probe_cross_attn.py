"""Probe: decoder cross-attention launch time by decode rows (streaming kernel vs TWB200_CROSS_ATTN=scalar)."""
import sys, os, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from turbo_whisper_workspace_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
H, S, D, L = 20, 1500, 1280, 4
p = lambda t: C.c_void_p(t.data_ptr())
cap = int(os.environ.get("TWB200_CROSS_SPLITS", 12))
ROWS = [int(os.environ['PROBE_ROWS'])] if os.environ.get('PROBE_ROWS') else [96, 72, 48, 24, 16, 12, 5, 1]
for B in ROWS:
    ckv = torch.randn(L * 2 * H, B, S, 64, device=dev).to(torch.bfloat16)
    q = torch.randn(B, D, device=dev).to(torch.bfloat16)
    out = torch.empty(B, D, dtype=torch.bfloat16, device=dev)
    blk = B * S * 64
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    part = torch.zeros(B, H, cap, 66, device=dev)
    cnt = torch.zeros(B, H, dtype=torch.int32, device=dev)
    def launch(i):
        k = C.c_void_p(ckv.data_ptr() + ((i * 2 + 0) * H) * blk * 2)
        v = C.c_void_p(ckv.data_ptr() + ((i * 2 + 1) * H) * blk * 2)
        _lib.check(lib.tw_dec_cross_attn(p(q), p(out), k, v, 64, S * 64, blk, None, S, B, H, cap, p(part), p(cnt), st), "x")
    for i in range(L): launch(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 40 if B <= 24 else 16
    e0.record()
    for it in range(n): launch(it % L)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    gb = B * 2.0 * S * D * 2 / 1e9
    print(json.dumps({"mode": os.environ.get("TWB200_CROSS_ATTN", "stream"), "split_cap": cap, "rows": B, "us": round(ms * 1e3, 1),
                      "GBps": round(gb / ms * 1e3, 1), "frac_hbm": round(gb / ms * 1e3 / 6458.7, 3)}), flush=True)
    del ckv
