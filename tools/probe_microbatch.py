"""How should a rank's share of ONE long file be cut into micro-batches?  (config 3 on 8 GPUs: 22-23 windows per rank.)
Times pipe(audio of n windows) on one GPU for several micro-batch floors; prints the host phases as well."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers
from turbo_whisper_workspace_b200 import scheduler as S
from turbo_whisper_workspace_b200.config import WhisperDims
from turbo_whisper_workspace_b200.pipeline import B200WhisperPipeline

dims = WhisperDims.large_v3_turbo()
pipe = B200WhisperPipeline(helpers.random_state_dict(dims, 0, "hf"), dims, helpers.build_tokenizer(), devices=["cuda:0"],
                           max_batch=24, contexts_per_device=4)
for n in (23, 45, 90):
    audio = np.concatenate([helpers.synth_clip(100 + i) for i in range(n)])
    for floor in (6, 8, 12, 24):
        S.MIN_MICROBATCH = floor
        kw = dict(chunk_length_s=30, stride_length_s=0, batch_size=24, return_timestamps=True)
        pipe(audio, **kw)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        pipe(audio, **kw)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(json.dumps({"windows": n, "floor": floor, "microbatches": [b - a for a, b in pipe.scheduler.last_stats["microbatches"]],
                          "seconds": round(dt, 4), "rtfx": round(n * 30 / dt, 1),
                          "host": {k: round(v, 4) for k, v in pipe.last_stats["host_seconds"].items()}}), flush=True)
