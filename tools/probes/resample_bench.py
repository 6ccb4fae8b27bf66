"""Timing of the ingest kernel: 1 h of 44.1 kHz stereo PCM16 and 48 kHz mono fp32 -> 16 kHz mono fp32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from turbo_whisper_workspace_b200 import ops
dev = torch.device("cuda:0")
for (sr, ch, dt) in ((44100, 2, torch.int16), (48000, 1, torch.float32)):
    n = sr * 3600
    x = (torch.randn(n, ch, device=dev) * 3000).to(dt) if dt == torch.int16 else torch.randn(n, device=dev) * 0.1
    rs = ops.Resampler(sr, 16000, dev)
    out = torch.empty(rs.out_len(n), dtype=torch.float32, device=dev)
    for _ in range(3): rs(x, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): rs(x, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    by = x.numel() * x.element_size() + out.numel() * 4
    print(f"{sr} Hz x{ch} {dt}: 1 h -> {ms:.3f} ms, {by/ms/1e6:.0f} GB/s algorithmic ({by/1e6:.0f} MB), taps/phase {rs.taps}, RTFx {3600/(ms/1e3):.3g}")
