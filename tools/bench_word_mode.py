"""End-to-end throughput of return_timestamps="word" next to the segment-timestamp call on the bench workload
(large-v3-turbo shape, random-init weights, 24 x 30 s windows per call, host PCM in, dict out).
The alignment heads are those of openai/whisper-large-v3-turbo's generation config (from memory of the hub file:
[[2,4],[2,11],[3,3],[3,6],[3,11],[3,14]] — six heads in the last two decoder layers).
Usage: python tools/bench_word_mode.py [calls]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import helpers
from turbo_whisper_workspace_b200.config import ALIGNMENT_HEADS, GenerationSettings, WhisperDims

from turbo_whisper_workspace_b200.pipeline import B200WhisperPipeline

K = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = 24
dims = WhisperDims.large_v3_turbo()
gen = GenerationSettings(alignment_heads=ALIGNMENT_HEADS["large-v3-turbo"])
pipe = B200WhisperPipeline(helpers.random_state_dict(dims, 0, "hf"), dims, helpers.build_tokenizer(), gen,
                           devices=["cuda:0"], max_batch=B, contexts_per_device=4)
audio = np.concatenate([helpers.synth_clip(i) for i in range(B)] * K)
kw = dict(chunk_length_s=30, stride_length_s=0, batch_size=B, generate_kwargs={"task": "transcribe"})
for mode in (True, "word", True, "word"):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = pipe(audio, return_timestamps=mode, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"return_timestamps": mode, "audio_s": len(audio) / 16000, "wall_s": round(dt, 3),
                      "rtfx": round(len(audio) / 16000 / dt, 1), "chunks": len(r["chunks"]),
                      "dec_steps": sum(e.stats["dec_steps"] for e in pipe.scheduler.flat_engines)}), flush=True)
pipe.close()
