import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, helpers
from turbo_whisper_workspace_b200.config import WhisperDims
from turbo_whisper_workspace_b200.engine import WhisperEngine
B = 24
dims = WhisperDims.large_v3_turbo()
eng = WhisperEngine(dims, helpers.random_state_dict(dims, 0, "hf"), device="cuda:0", max_batch=B)
eng.load_pcm([helpers.synth_clip(i) for i in range(B)]); eng.features(B); eng.encode(B)
eng.finish_check_every = 0
eng.step_timing = torch.zeros(128, dtype=torch.int64, device="cuda:0")
prompts = torch.tensor([[50258, -1, 50360]] * B, dtype=torch.int32)
eng.decode(B, prompts, n_steps=200)
torch.cuda.synchronize()
t = eng.step_timing.cpu().tolist()
names = ["embed"] + [f"L{l}.{n}" for l in range(4) for n in ("qkv", "self", "out", "ln2", "cq", "cross", "cout", "ln3", "fc1", "fc2", "ln")] + ["lmhead", "final(start)"]
prev = t[0]
print("phase durations (us) at step 200 (time between barrier exits of CTA 0):")
acc = {}
for i in range(1, 48):
    if t[i] == 0: break
    d = (t[i] - prev) / 1e3; prev = t[i]
    nm = names[i - 1] if i - 1 < len(names) else str(i)
    key = nm.split(".")[-1]
    acc[key] = acc.get(key, 0) + d
    if i <= 12 or i >= 45: print(f"  {nm:12s} {d:7.2f}")
print("sum by phase type:", {k: round(v, 1) for k, v in acc.items()}, "total", round((prev - t[0]) / 1e3, 1))
