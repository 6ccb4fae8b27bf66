"""Golden fixtures for token timestamps under beam search (the mode the HF pipeline's default num_beams=5 runs when a
caller asks for return_timestamps="word"), from the installed transformers 5.5.0:  tests/golden/word_beams_tiny.json
  * WhisperGenerationMixin.generate(num_beams=5, return_timestamps=True, return_token_timestamps=True,
    return_segments=True, attention_mask=...) on the three fixture clips, both fixture models: the per-segment tokens
    and token timestamps (cross-attention rows gathered by `beam_indices`) and the padded `token_timestamps`.
Run in the build container (CPU):   python tests/golden/make_golden_word_beams.py
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
from make_golden_word import ALIGNMENT_HEADS  # noqa: E402
from transformers import WhisperFeatureExtractor  # noqa: E402


def main():
    fe = WhisperFeatureExtractor(feature_size=128)
    clips = [helpers.synth_clip(0), helpers.synth_clip(1, kind="mod"), helpers.synth_clip(2, seconds=11.3, kind="mod")]
    f = [fe(c, sampling_rate=16000, return_tensors="np", return_attention_mask=True) for c in clips]
    feats = torch.from_numpy(np.stack([x["input_features"][0] for x in f])).to(torch.bfloat16).float()
    mask = torch.from_numpy(np.stack([x["attention_mask"][0] for x in f]))
    res = {"alignment_heads": ALIGNMENT_HEADS, "num_frames": mask.sum(-1).tolist()}
    for variant in ("decisive", "varied"):
        model, _ = G.hf_model(variant)
        model.generation_config.alignment_heads = ALIGNMENT_HEADS
        with torch.no_grad():
            out = model.generate(input_features=feats, attention_mask=mask, return_timestamps=True, task="transcribe",
                                 num_beams=5, do_sample=False, return_token_timestamps=True, return_segments=True)
        res[f"{variant}_generate_beams5"] = {
            "token_timestamps": out["token_timestamps"].double().numpy().tolist(),
            "segment_token_timestamps": [torch.cat([s["token_timestamps"] for s in segs]).double().numpy().tolist()
                                         for segs in out["segments"]],
            "segment_tokens": [torch.cat([s["tokens"] for s in segs]).numpy().astype(int).tolist()
                               for segs in out["segments"]],
        }
    with open(os.path.join(HERE, "word_beams_tiny.json"), "w") as fjson:
        json.dump(res, fjson)
    print("written", {k: (list(v.keys()) if isinstance(v, dict) else v) for k, v in res.items()})


if __name__ == "__main__":
    main()
