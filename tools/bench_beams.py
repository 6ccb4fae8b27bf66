"""num_beams=5 (the reference's literal decoding mode under transformers >= 4.53) on the bench workload: large-v3-turbo
dims, 24 x 30 s windows through B200WhisperPipeline.__call__ with generate_kwargs={"num_beams": 5} next to greedy."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers
from turbo_whisper_workspace_b200.config import WhisperDims
from turbo_whisper_workspace_b200.pipeline import B200WhisperPipeline

dims = WhisperDims.large_v3_turbo()
pipe = B200WhisperPipeline(helpers.random_state_dict(dims, 0, "hf"), dims, helpers.build_tokenizer(), devices=["cuda:0"],
                           max_batch=30, contexts_per_device=4)
B = 24
audio = np.concatenate([helpers.synth_clip(i) for i in range(B)])
for beams in (1, 5):
    kw = dict(chunk_length_s=30, stride_length_s=0, batch_size=B, return_timestamps=True,
              generate_kwargs={"task": "transcribe", "num_beams": beams})
    pipe(audio, **kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = pipe(audio, **kw)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    eng = pipe.scheduler.flat_engines[0]
    print(json.dumps({"num_beams": beams, "windows": B, "seconds": round(dt, 3), "rtfx": round(B * 30 / dt, 1),
                      "chunks": len(r["chunks"]), "microbatches": [b - a for a, b in pipe.scheduler.last_stats["microbatches"]],
                      "beam_steps_last": getattr(eng, "last_beam", {}).get("steps")}), flush=True)
