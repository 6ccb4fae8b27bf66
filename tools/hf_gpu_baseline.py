"""On-box LIBRARY baseline: the unmodified transformers ASR pipeline — the object the reference's
`load_transcription_model` builds (ref:vocalis/core/audio_pipeline.py:195-200) — on cuda:0 in bf16 with sdpa attention,
greedy, on the bench workload (large-v3-turbo shape, random-init weights, 24 x 30 s windows, host PCM in, dict out).
Not part of bench.py's contract (its reference arm is the CPU path); printed as one JSON line for profiles/.
Usage: python tools/hf_gpu_baseline.py [windows]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("HF_HUB_OFFLINE", "1")
import numpy as np
import torch
import helpers
import bench
from transformers import WhisperFeatureExtractor, pipeline

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
t0 = time.perf_counter()
model = bench.build_hf_turbo(0).to(torch.bfloat16).to("cuda:0")
pipe = pipeline("automatic-speech-recognition", model=model, tokenizer=helpers.build_tokenizer(),
                feature_extractor=WhisperFeatureExtractor(feature_size=128), device="cuda:0", dtype=torch.bfloat16)
pipe.generation_config.num_beams = 1
t_build = time.perf_counter() - t0
kw = dict(chunk_length_s=30, stride_length_s=0, batch_size=B, generate_kwargs={"task": "transcribe"},
          return_timestamps=True, ignore_warning=True)
warm = np.concatenate([helpers.synth_clip(100 + i) for i in range(2)])
pipe(warm.copy(), **{**kw, "batch_size": 2})
audio = np.concatenate([helpers.synth_clip(i) for i in range(B)])
torch.cuda.synchronize()
t0 = time.perf_counter()
r = pipe(audio.copy(), **kw)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(json.dumps({"impl": "transformers pipeline on cuda:0 (library baseline)", "dtype": "bf16",
                  "attn": getattr(model.config, "_attn_implementation", None), "windows": B,
                  "audio_s": len(audio) / 16000, "wall_s": round(dt, 3), "rtfx": round(len(audio) / 16000 / dt, 1),
                  "chunks": len(r["chunks"]), "build_s": round(t_build, 1), "torch": torch.__version__,
                  "gpu": torch.cuda.get_device_name(0)}), flush=True)
