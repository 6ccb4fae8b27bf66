import sys, os
ROOT = "/root/repo" if os.path.exists("/root/repo/tests") else os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, helpers
from oracle import logmel_ref as L
from oracle import whisper_ref as R
from turbo_whisper_workspace_b200.config import WhisperDims
from turbo_whisper_workspace_b200.engine import WhisperEngine
clips = [helpers.synth_clip(0), helpers.synth_clip(1, kind="mod"), helpers.synth_clip(2, seconds=11.3, kind="mod")]
feats = torch.stack([torch.from_numpy(L.log_mel(c)) for c in clips])
rd = R.WhisperDims(**helpers.TINY)
sd = helpers.variant_state_dict(rd, "decisive")
ref = R.WhisperRef(rd, sd)
eng = WhisperEngine(WhisperDims(**helpers.TINY), sd, device="cuda:0", max_batch=4)
trace = {}
want = ref.generate(feats.to(torch.bfloat16).float(), trace=trace)
got = eng.generate_from_pcm(clips)
for b in range(3):
    n = next((i for i, (x, y) in enumerate(zip(want[b], got[b])) if x != y), None)
    print("row", b, "len", len(want[b]), len(got[b]), "first diff", n, "want", want[b][:6] if n is None else want[b][max(0, n-2):n+4], "got", got[b][:6] if n is None else got[b][max(0,n-2):n+4])
for k, it in enumerate(trace["iterations"]):
    rec = it["record"]
    print("iteration", k, "rows", it.get("rows"), "steps recorded", len(rec))
    for g in range(min(5, len(rec))):
        print("  step", g, "margins", [round(float(m), 4) for m in rec[g]["margin"]], "rule_gap", [round(float(m), 4) for m in rec[g]["rule_gap"]])
