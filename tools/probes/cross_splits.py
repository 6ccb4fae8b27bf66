"""Probe: decode cross-attention launch time vs number of key splits (wave quantisation of the 1920-CTA grid)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from turbo_whisper_workspace_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
H, S, D, L = 20, 1500, 1280, 4
p = lambda t: C.c_void_p(t.data_ptr())
for B in (24, 16, 8):
    ckv = torch.randn(L * 2 * H, B, S, 64, device=dev).to(torch.bfloat16)
    q = torch.randn(B, D, device=dev).to(torch.bfloat16)
    out = torch.empty(B, D, dtype=torch.bfloat16, device=dev)
    blk = B * S * 64
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for splits in (3, 4, 5, 6, 7, 8, 10, 12, 15):
        part = torch.zeros(B, H, splits, 66, device=dev)
        cnt = torch.zeros(B, H, dtype=torch.int32, device=dev)
        def launch(i):
            k = C.c_void_p(ckv.data_ptr() + ((i * 2 + 0) * H) * blk * 2)
            v = C.c_void_p(ckv.data_ptr() + ((i * 2 + 1) * H) * blk * 2)
            _lib.check(lib.tw_dec_cross_attn(p(q), p(out), k, v, 64, S * 64, blk, None, S, B, H, splits, p(part), p(cnt), st), "x")
        for i in range(L): launch(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(40): launch(it % L)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 40
        gb = B * 2.0 * S * D * 2 / 1e9
        print(f"B={B} splits={splits:2d} grid={splits*H*B:5d}: {ms*1e3:6.1f} us  {gb/ms*1e3:7.1f} GB/s  frac {gb/ms*1e3/6458.7:.3f}", flush=True)
