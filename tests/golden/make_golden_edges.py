"""Golden fixtures for edge-case inputs of the ASR pipeline call (transformers 5.5.0, "varied" fixture model, greedy):
tests/golden/edges_tiny.json — empty audio without chunking (one zero-padded window), 100 samples with chunking, a 40 s
clip whose last window ends exactly at the end of the audio (chunk 30 / stride 5), each with return_timestamps True and
omitted; plus the exception HF raises for empty audio WITH chunk_length_s (StopIteration: no window is ever yielded).
Run in the build container (CPU):   python tests/golden/make_golden_edges.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402
import helpers  # noqa: E402
from transformers import WhisperFeatureExtractor, pipeline  # noqa: E402


def cases():
    return [("empty_plain", np.zeros(0, np.float32), {}),
            ("tiny_chunked", np.zeros(100, np.float32), dict(chunk_length_s=30, stride_length_s=5)),
            ("exact_multiple", helpers.synth_clip(5, seconds=40.0), dict(chunk_length_s=30, stride_length_s=5)),
            ("empty_chunked", np.zeros(0, np.float32), dict(chunk_length_s=30, stride_length_s=5))]


def main():
    model, _ = G.hf_model("varied")
    pipe = pipeline("automatic-speech-recognition", model=model, tokenizer=helpers.build_tokenizer(),
                    feature_extractor=WhisperFeatureExtractor(feature_size=128), device="cpu", dtype=torch.float32)
    pipe.generation_config.num_beams = 1
    res = {}
    for name, x, kw in cases():
        for rt in (True, None):
            key = f"{name}_{'ts' if rt else 'nots'}"
            try:
                r = pipe(x.copy(), batch_size=4, return_timestamps=rt, generate_kwargs={"task": "transcribe"}, **kw)
                res[key] = {"text": r["text"], "keys": sorted(r.keys()),
                            "chunks": [{"text": c["text"], "timestamp": list(c["timestamp"])} for c in r.get("chunks", [])]}
            except Exception as e:  # noqa: BLE001
                res[key] = {"raises": type(e).__name__}
    with open(os.path.join(HERE, "edges_tiny.json"), "w") as f:
        json.dump(res, f, ensure_ascii=False)
    print({k: (v.get("raises") or len(v["chunks"])) for k, v in res.items()})


if __name__ == "__main__":
    main()
