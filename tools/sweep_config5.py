"""BASELINE.json configs[4]: log-mel front end + encoder-only sweep, batch 1..256 x 30 s windows on one GPU.

Per batch size: log-mel time -> achieved HBM GB/s on the algorithmic bytes (480000*4 in + 128*3000*2 bf16 out
= 2.688 MB per window) and encoder time (conv stem + 32 layers + final LN + cross-K/V GEMM) ->
TFLOP/s on the algorithmic FLOPs (2.2738e12 + L_dec*4*1500*1280^2 per window), both against MEASURED_PEAKS.json.
CUDA events on the launching stream, 3 warm-up runs, inputs larger than L2 from B >= 8 (noted per line)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers
from turbo_whisper_workspace_b200.config import WhisperDims
from turbo_whisper_workspace_b200.engine import WhisperEngine

batches = [int(b) for b in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 2, 4, 8, 16, 24, 32, 64, 128, 256]
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm_peak = float(peaks.get("hbm_gbs", 6458.7))
tc_burst, tc_sustained = float(peaks.get("bf16_tflops", 1678.6)), float(peaks.get("bf16_tflops_sustained", 1427.9))

dims = WhisperDims.large_v3_turbo()
eng = WhisperEngine(dims, helpers.random_state_dict(dims, 0, "hf"), device="cuda:0", max_batch=24, max_enc_batch=max(batches))
base = [helpers.synth_clip(i) for i in range(8)]


def timed(fn, iters, warm=3):
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


rows = []
for B in batches:
    eng.load_pcm([base[i % 8] for i in range(B)])
    torch.cuda.synchronize()
    mel_ms = timed(lambda: eng.features(B), iters=20 if B <= 32 else 5)
    enc_ms = timed(lambda: eng.encode(B), iters=3 if B <= 32 else 2)
    mel_bytes = B * (480000 * 4 + 128 * 3000 * 2)
    enc_flops = B * (2.2738e12 + dims.dec_layers * 4 * 1500 * 1280 ** 2)
    row = {"batch": B, "logmel_ms": round(mel_ms, 4), "logmel_GBps": round(mel_bytes / mel_ms / 1e6, 1),
           "logmel_frac_hbm": round(mel_bytes / mel_ms / 1e6 / hbm_peak, 3),
           "encoder_ms": round(enc_ms, 3), "encoder_TFLOPs": round(enc_flops / enc_ms / 1e9, 1),
           "encoder_frac_burst": round(enc_flops / enc_ms / 1e9 / tc_burst, 3),
           "encoder_frac_sustained": round(enc_flops / enc_ms / 1e9 / tc_sustained, 3),
           "windows_per_s": round(B / (enc_ms + mel_ms) * 1e3, 1),
           "l2": "inputs+activations exceed the 126 MB L2" if B >= 8 else "activations partly L2-resident"}
    rows.append(row)
    print(json.dumps(row), flush=True)
