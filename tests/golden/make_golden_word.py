"""Golden fixtures for word-level timestamps (return_timestamps="word"), from the installed transformers 5.5.0:
tests/golden/word_tiny.json
  * WhisperGenerationMixin.generate(return_timestamps=True, return_token_timestamps=True, return_segments=True,
    attention_mask=...) on the three fixture clips, both fixture models: sequences, the padded `token_timestamps`
    and the per-segment token timestamps the ASR pipeline reads;
  * the ASR pipeline's {"text", "chunks"} with return_timestamps="word" for the 71.3 s file (chunk 30 / stride 5)
    and for a single 11.3 s clip without chunking.
Alignment heads of the fixture models (2 decoder layers x 4 heads): [[0, 1], [1, 0], [1, 3]]; median_filter_width 7.
Run in the build container (CPU):   python tests/golden/make_golden_word.py
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

ALIGNMENT_HEADS = [[0, 1], [1, 0], [1, 3]]


def main():
    fe = WhisperFeatureExtractor(feature_size=128)
    clips = [helpers.synth_clip(0), helpers.synth_clip(1, kind="mod"), helpers.synth_clip(2, seconds=11.3, kind="mod")]
    f = [fe(c, sampling_rate=16000, return_tensors="np", return_attention_mask=True) for c in clips]
    feats = torch.from_numpy(np.stack([x["input_features"][0] for x in f]))
    mask = torch.from_numpy(np.stack([x["attention_mask"][0] for x in f]))
    feats_bf = feats.to(torch.bfloat16).float()
    tok = helpers.build_tokenizer()
    asr.ffmpeg_read = lambda b, sr: (np.frombuffer(wave.open(io.BytesIO(b)).readframes(10 ** 9), np.int16)
                                     .astype(np.float32) / 32768.0)
    wav_path = "/tmp/golden_71s.wav"
    helpers.write_wav16(wav_path, np.concatenate(clips))
    res = {"alignment_heads": ALIGNMENT_HEADS, "num_frames": mask.sum(-1).tolist()}
    for variant in ("decisive", "varied"):
        model, _ = G.hf_model(variant)
        model.generation_config.alignment_heads = ALIGNMENT_HEADS
        assert model.config.median_filter_width == 7
        with torch.no_grad():
            out = model.generate(input_features=feats_bf, attention_mask=mask, return_timestamps=True, task="transcribe",
                                 num_beams=1, do_sample=False, return_token_timestamps=True, return_segments=True)
        res[f"{variant}_generate"] = {
            "sequences": out["sequences"].numpy().astype(int).tolist(),
            "token_timestamps": out["token_timestamps"].double().numpy().tolist(),
            "segment_token_timestamps": [torch.cat([s["token_timestamps"] for s in segs]).double().numpy().tolist()
                                         for segs in out["segments"]],
            "segment_tokens": [torch.cat([s["tokens"] for s in segs]).numpy().astype(int).tolist()
                               for segs in out["segments"]],
        }
        pipe = pipeline("automatic-speech-recognition", model=model, tokenizer=tok, feature_extractor=fe, device="cpu",
                        dtype=torch.float32)
        pipe.generation_config.num_beams = 1
        pipe.generation_config.alignment_heads = ALIGNMENT_HEADS
        r = pipe(wav_path, chunk_length_s=30, stride_length_s=5, batch_size=24, generate_kwargs={"task": "transcribe"},
                 return_timestamps="word")
        res[f"{variant}_30_5_24"] = {"text": r["text"], "chunks": [{"text": c["text"], "timestamp": list(c["timestamp"])}
                                                                 for c in r["chunks"]]}
        if variant == "varied":
            # the reference's literal chunking (chunk_length_s=60: every window is truncated to its first 30 s by the
            # feature extractor while the stride bookkeeping keeps 60 s) with word timestamps
            r = pipe(wav_path, chunk_length_s=60, stride_length_s=5, batch_size=32, generate_kwargs={"task": "transcribe"},
                     return_timestamps="word")
            res[f"{variant}_60_5_32"] = {"text": r["text"], "chunks": [{"text": c["text"], "timestamp": list(c["timestamp"])}
                                                                     for c in r["chunks"]]}
        r = pipe(clips[2].copy(), generate_kwargs={"task": "transcribe"}, return_timestamps="word")
        res[f"{variant}_single"] = {"text": r["text"], "chunks": [{"text": c["text"], "timestamp": list(c["timestamp"])}
                                                                for c in r["chunks"]]}
    with open(os.path.join(HERE, "word_tiny.json"), "w") as fjson:
        json.dump(res, fjson, ensure_ascii=False)
    print("written", {k: (list(v.keys()) if isinstance(v, dict) else v) for k, v in res.items()})


if __name__ == "__main__":
    main()
