"""Golden fixtures for the no-timestamp generation mode (return_timestamps falsy — the HF pipeline's default),
from the installed transformers 5.5.0:  tests/golden/notimestamps_tiny.json
  * WhisperGenerationMixin.generate(return_timestamps=False) token ids on the three fixture clips, both models;
  * the ASR pipeline's {"text"} for the 71.3 s file, chunk 30 / stride 5 and chunk 30 / stride 0.
Run in the build container (CPU):   python tests/golden/make_golden_notimestamps.py
"""
import io
import json
import os
import sys
import wave

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (hf_model, paths, offline switches)
import helpers  # noqa: E402
from transformers import WhisperFeatureExtractor, pipeline  # noqa: E402
import transformers.pipelines.automatic_speech_recognition as asr  # noqa: E402


def main():
    fe = WhisperFeatureExtractor(feature_size=128)
    clips = [helpers.synth_clip(0), helpers.synth_clip(1, kind="mod"), helpers.synth_clip(2, seconds=11.3, kind="mod")]
    feats = torch.from_numpy(np.stack([fe(c, sampling_rate=16000, return_tensors="np")["input_features"][0] for c in clips]))
    feats_bf = feats.to(torch.bfloat16).float()
    tok = helpers.build_tokenizer()
    asr.ffmpeg_read = lambda b, sr: (np.frombuffer(wave.open(io.BytesIO(b)).readframes(10 ** 9), np.int16)
                                     .astype(np.float32) / 32768.0)
    wav_path = "/tmp/golden_71s.wav"
    helpers.write_wav16(wav_path, np.concatenate(clips))
    res = {}
    for variant in ("decisive", "varied"):
        model, _ = G.hf_model(variant)
        with torch.no_grad():
            gen = model.generate(input_features=feats_bf, return_timestamps=False, task="transcribe", num_beams=1,
                                 do_sample=False)
        res[f"{variant}_generate"] = gen.numpy().astype(int).tolist()
        pipe = pipeline("automatic-speech-recognition", model=model, tokenizer=tok, feature_extractor=fe, device="cpu",
                        dtype=torch.float32)
        pipe.generation_config.num_beams = 1
        for (cl, st, bs) in ((30, 5, 24), (30, 0, 24)):
            r = pipe(wav_path, chunk_length_s=cl, stride_length_s=st, batch_size=bs, generate_kwargs={"task": "transcribe"})
            res[f"{variant}_{cl}_{st}_{bs}"] = {"keys": sorted(r.keys()), "text": r["text"]}
    with open(os.path.join(HERE, "notimestamps_tiny.json"), "w") as f:
        json.dump(res, f, ensure_ascii=False)
    print("written", {k: (len(v) if isinstance(v, list) else v["keys"]) for k, v in res.items()})


if __name__ == "__main__":
    main()
