"""Golden fixture for the 5-beam search mode (SURVEY.md §8f rank 1 — what the reference's literal pipeline call runs
under transformers >= 4.53, whose ASR pipeline defaults to num_beams=5) from the installed transformers 5.5.0:
tests/golden/beams_tiny.json = WhisperGenerationMixin.generate(return_timestamps=True, num_beams=5) token ids on
the three fixture clips for both fixture models.  Run (CPU):  python tests/golden/make_golden_beams.py
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
from transformers import WhisperFeatureExtractor  # noqa: E402


def main():
    fe = WhisperFeatureExtractor(feature_size=128)
    clips = [helpers.synth_clip(0), helpers.synth_clip(1, kind="mod"), helpers.synth_clip(2, seconds=11.3, kind="mod")]
    feats = torch.from_numpy(np.stack([fe(c, sampling_rate=16000, return_tensors="np")["input_features"][0] for c in clips]))
    feats_bf = feats.to(torch.bfloat16).float()
    res = {}
    for variant in ("decisive", "varied"):
        model, _ = G.hf_model(variant)
        with torch.no_grad():
            gen = model.generate(input_features=feats_bf, return_timestamps=True, task="transcribe", num_beams=5,
                                 do_sample=False)
        res[f"{variant}_generate_beams5"] = gen.numpy().astype(int).tolist()
    with open(os.path.join(HERE, "beams_tiny.json"), "w") as f:
        json.dump(res, f)
    print({k: [len(r) for r in v] for k, v in res.items()})


if __name__ == "__main__":
    main()
