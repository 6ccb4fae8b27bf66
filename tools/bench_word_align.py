"""Micro-benchmark of the word-timestamp kernels on the large-v3-turbo shapes (CUDA events; synthetic operands):
the alignment tap of one decode step (6 alignment heads, B rows), the normalise / median / head-mean stage over a
full 445-token batch, and the host DTW.  Usage: python tools/bench_word_align.py [B]"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from turbo_whisper_workspace_b200 import _lib
from turbo_whisper_workspace_b200._lib import check

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
S, H, D, T, NS = 1500, 20, 1280, 448, 6
lib = _lib.load()
p = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timed(fn, iters=20, warm=3):
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


# head-major K of 4 decoder layers x (k|v) so that successive calls rotate through > L2 worth of data
kv = torch.randn(8 * H, B, S, 64, device=dev).to(torch.bfloat16)
q = (torch.randn(B, D, device=dev) * 0.3).to(torch.bfloat16)
state = torch.zeros(B, 8, dtype=torch.int32, device=dev)
state[:, 0] = 7
enc_row = torch.arange(B, dtype=torch.int32, device=dev)
heads = torch.tensor([3, 6, 11, 14, 4, 9], dtype=torch.int32, device=dev)
probs = torch.zeros(B, NS, T, S, dtype=torch.float32, device=dev)
blk = B * S * 64
layer = [0]


def tap():
    layer[0] = (layer[0] + 1) % 8
    k = C.c_void_p(kv.data_ptr() + layer[0] * H * blk * 2)
    check(lib.tw_dec_align_tap(p(q), D, k, 64, S * 64, blk, p(enc_row), p(state), p(heads), NS, 0, NS, T, S, B, p(probs), st))


ms = timed(tap)
byt = B * NS * (S * 64 * 2 + S * 4)
print(json.dumps(dict(kernel="align_tap", B=B, heads=NS, ms=round(ms, 4), gbs=round(byt / ms / 1e6, 1))), flush=True)

probs.uniform_(0, 1)
probs /= probs.sum(-1, keepdim=True)
n_tok = 444
stats = torch.zeros(B, NS, S, 2, device=dev)
matrix = torch.zeros(B, T, S, device=dev)
nf = torch.full((B,), S, dtype=torch.int32, device=dev)
ms = timed(lambda: check(lib.tw_align_matrix(p(probs), p(nf), B, NS, T, S, 3, n_tok, 7, p(stats), p(matrix), st)), iters=5)
byt = B * NS * n_tok * S * 4 * 3 + B * n_tok * S * 4     # probabilities streamed for mean, variance and filter + matrix out
print(json.dumps(dict(kernel="align_matrix", B=B, n_tok=n_tok, ms=round(ms, 3), gbs=round(byt / ms / 1e6, 1))), flush=True)

host = matrix[:, :n_tok].cpu().contiguous()
out = np.empty(n_tok, dtype=np.int32)
t0 = time.perf_counter()
for b in range(B):
    check(lib.tw_dtw_token_frames(C.c_void_p(host[b].data_ptr()), S, n_tok, S, out.ctypes.data_as(C.c_void_p)))
print(json.dumps(dict(kernel="dtw_host", B=B, ms_per_window=round((time.perf_counter() - t0) * 1e3 / B, 2))), flush=True)
