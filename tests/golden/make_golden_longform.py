"""Golden fixture for UN-CHUNKED long-form input (transformers 5.5.0, greedy): tests/golden/longform_tiny.json.
A 75.3 s clip goes through the ASR pipeline without chunk_length_s — the pipeline then extracts features of the whole
clip (truncation=False) and `generate` runs its seek loop over all 7530 frames
($TF/pipelines/automatic_speech_recognition.py:446-454, $TF/models/whisper/generation_whisper.py:654-658) — for both
fixture models, with return_timestamps=True (tokens of the generate call + the pipeline dict), with num_beams=3 on the
"varied" model, and the ValueError HF raises without timestamps.
Run in the build container (CPU):   python tests/golden/make_golden_longform.py
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


def clip():
    return np.concatenate([helpers.synth_clip(0), helpers.synth_clip(1, kind="mod"),
                           helpers.synth_clip(2, seconds=15.3, kind="mod")])


def main():
    x = clip()
    fe = WhisperFeatureExtractor(feature_size=128)
    res = {"samples": int(x.shape[0])}
    for variant in ("decisive", "varied"):
        model, _ = G.hf_model(variant)
        pipe = pipeline("automatic-speech-recognition", model=model, tokenizer=helpers.build_tokenizer(),
                        feature_extractor=fe, device="cpu", dtype=torch.float32)
        pipe.generation_config.num_beams = 1
        feats = fe(x, sampling_rate=16000, truncation=False, padding="longest", return_tensors="pt",
                   return_attention_mask=True)
        toks = model.generate(input_features=feats["input_features"], attention_mask=feats["attention_mask"],
                              return_timestamps=True, task="transcribe", num_beams=1)
        r = pipe(x.copy(), return_timestamps=True, generate_kwargs={"task": "transcribe"})
        res[variant] = {"tokens": toks[0].tolist(), "text": r["text"],
                        "chunks": [{"text": c["text"], "timestamp": list(c["timestamp"])} for c in r["chunks"]]}
        if variant == "varied":
            rb = pipe(x.copy(), return_timestamps=True, generate_kwargs={"task": "transcribe", "num_beams": 3})
            res["varied_beams3"] = {"text": rb["text"],
                                    "chunks": [{"text": c["text"], "timestamp": list(c["timestamp"])} for c in rb["chunks"]]}
        try:
            pipe(x.copy(), generate_kwargs={"task": "transcribe"})
            res[variant + "_nots"] = {"raises": None}
        except Exception as e:  # noqa: BLE001
            res[variant + "_nots"] = {"raises": type(e).__name__, "message": str(e)}
    with open(os.path.join(HERE, "longform_tiny.json"), "w") as f:
        json.dump(res, f, ensure_ascii=False)
    print({k: (len(v["chunks"]) if isinstance(v, dict) and "chunks" in v else v) for k, v in res.items()})


if __name__ == "__main__":
    main()
