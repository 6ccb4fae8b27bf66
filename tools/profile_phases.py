"""Phase timing of the full-size engine (CUDA events): log-mel, encoder, decode step."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers
from turbo_whisper_workspace_b200.config import WhisperDims
from turbo_whisper_workspace_b200.engine import WhisperEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
model = sys.argv[2] if len(sys.argv) > 2 else "turbo"
dims = WhisperDims.large_v3_turbo() if model == "turbo" else WhisperDims.large_v3()
t0 = time.time()
sd = helpers.random_state_dict(dims, 0, "hf")
print("state dict", round(time.time() - t0, 1), "s", flush=True)
eng = WhisperEngine(dims, sd, device="cuda:0", max_batch=B)
del sd
clips = [helpers.synth_clip(i) for i in range(B)]
eng.load_pcm(clips)
torch.cuda.synchronize()

def timed(fn, iters=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

res = {}
res["logmel_ms"] = timed(lambda: eng.features(B), iters=10)
res["encode_ms"] = timed(lambda: eng.encode(B), iters=3)
enc_flops = 2.2738e12 * B + (dims.dec_layers * 4 * 1500 * 1280 ** 2) * B
res["encode_tflops"] = enc_flops / res["encode_ms"] / 1e9
# decode: run a full decode (graph), time per step
prompts = torch.tensor([[50258, -1, 50360]] * B, dtype=torch.int32)
eng.finish_check_every = 0
eng.decode(B, prompts, n_steps=8)  # builds graph
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.decode(B, prompts, n_steps=447); e1.record(); torch.cuda.synchronize()
res["decode_447_ms"] = e0.elapsed_time(e1)
res["decode_step_us"] = res["decode_447_ms"] / 447 * 1e3
# eager per-step for comparison
eng.use_graphs = False
e0.record(); eng.decode(B, prompts, n_steps=50); e1.record(); torch.cuda.synchronize()
res["decode_step_eager_us"] = e0.elapsed_time(e1) / 50 * 1e3
eng.use_graphs = True
t0 = time.time()
out = eng.generate(B)
torch.cuda.synchronize()
res["generate_s"] = time.time() - t0
res["generate_rtfx"] = 30.0 * B / res["generate_s"]
res["tokens_per_row"] = [len(o) for o in out][:6]
res["stats"] = eng.stats
print(json.dumps(res, indent=1))
