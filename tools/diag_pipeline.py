import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers
from oracle import whisper_ref as R, logmel_ref as L
from turbo_whisper_workspace_b200.config import WhisperDims
from turbo_whisper_workspace_b200.engine import WhisperEngine
from turbo_whisper_workspace_b200 import pipeline as P
pcm = helpers.quantize_pcm16(np.concatenate([helpers.synth_clip(10 + i, kind="mod" if i % 2 else "noise") for i in range(3)])[:70 * 16000])
rd = R.WhisperDims(**helpers.TINY)
sd = helpers.variant_state_dict(rd, "decisive")
ref = R.WhisperRef(rd, sd)
eng = WhisperEngine(WhisperDims(**helpers.TINY), sd, device="cuda:0", max_batch=4)
for (cl, st) in ((30, 5),):
    wins = P.chunk_windows(len(pcm), cl * 16000, st * 16000, st * 16000)
    clips = [pcm[s:e][:480000] for (s, e, _, _) in wins]
    feats = torch.stack([torch.from_numpy(L.log_mel(c)) for c in clips]).to(torch.bfloat16).float()
    tr = {}
    want = ref.generate(feats, trace=tr)
    etr = {}
    B = eng.load_pcm(clips); eng.features(B)
    got = eng.generate(B, trace=etr)
    print("config", cl, st, "windows", len(wins), "iters", len(tr["iterations"]), len(etr["iterations"]))
    for b in range(B):
        print(" row", b, "equal", got[b] == want[b], len(got[b]), len(want[b]))
    for k, (a, e) in enumerate(zip(tr["iterations"], etr["iterations"])):
        print(" iter", k, "rows", a["rows"], e["rows"], "seek", a["seek"], e["seek"])
        for i, b in enumerate(a["rows"]):
            if b not in e["rows"]: continue
            j = e["rows"].index(b)
            wt = a["tokens"][i].tolist(); gt = e["tokens"][j][3:3 + len(wt)]
            for g, (x, y) in enumerate(zip(wt, gt)):
                if x != y:
                    rec = a["record"][g]
                    print("   row", b, "first mismatch at", g, "want", x, "got", y, "margin", float(rec["margin"][i]), "rule_gap", float(rec["rule_gap"][i]))
                    break
