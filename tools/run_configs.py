"""BASELINE.json configs 3 and 4 on one GPU: (3) 1 h synthetic long-form audio, 30 s chunks with stride 5 (180 windows,
chunk-level timestamps) and the reference's literal 60/5 call (72 truncated windows); (4) whisper-large-v3
(32 decoder layers), batch 16, decoder-heavy."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers
from turbo_whisper_workspace_b200.config import WhisperDims
from turbo_whisper_workspace_b200.pipeline import B200WhisperPipeline
from turbo_whisper_workspace_b200.engine import WhisperEngine

which = sys.argv[1] if len(sys.argv) > 1 else "3"
tok = helpers.build_tokenizer()
if which == "3":
    dims = WhisperDims.large_v3_turbo()
    pipe = B200WhisperPipeline(helpers.random_state_dict(dims, 0, "hf"), dims, tok, devices=["cuda:0"], max_batch=24)
    audio = np.concatenate([helpers.synth_clip(100 + i) for i in range(120)])  # 3600 s
    for (cl, st, bs) in ((30, 5, 24), (60, 5, 512)):
        pipe(audio[:16000 * 200], chunk_length_s=cl, stride_length_s=st, batch_size=bs, return_timestamps=True)  # warm-up
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = pipe(audio, chunk_length_s=cl, stride_length_s=st, batch_size=bs, generate_kwargs={"task": "transcribe"},
                 return_timestamps=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        ts = [c["timestamp"] for c in r["chunks"]]
        print(json.dumps({"config": f"1 h audio, chunk {cl}/{st}, batch_size {bs}", "windows": pipe.last_stats["windows"],
                          "seconds": round(dt, 3), "rtfx": round(3600 / dt, 1), "chunks": len(ts), "first": ts[:2], "last": ts[-1:]}))
else:
    dims = WhisperDims.large_v3()
    B = 16
    eng = WhisperEngine(dims, helpers.random_state_dict(dims, 0, "hf"), device="cuda:0", max_batch=B)
    clips = [helpers.synth_clip(i) for i in range(B)]
    eng.generate_from_pcm(clips)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = eng.generate_from_pcm(clips)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps({"config": "large-v3 (32 decoder layers), batch 16", "seconds": round(dt, 3), "rtfx": round(30 * B / dt, 1),
                      "dec_steps": eng.stats["dec_steps"] // 2, "ms_per_decode_step": None, "tokens": [len(o) for o in out][:4]}))
