"""Generates the committed golden fixtures from the INSTALLED third-party implementation
(transformers 5.5.0 — the library the reference delegates the whole hot path to) and, when
/root/reference is present, through the reference's own entry point
``vocalis.core.audio_pipeline.AudioProcessingPipeline.process_audio`` (recipe: SURVEY.md §8c / A.2).

Run in the build container (CPU):   python tests/golden/make_golden.py
Outputs (small, committed):  tests/golden/logmel_kat.npz, tests/golden/model_tiny.npz,
                             tests/golden/pipeline_tiny.json
The fixtures pin oracle/ (tests/test_oracle_*.py) and the CUDA path (tests/test_gpu_*.py).
"""
import io
import json
import os
import sys
import types
import wave

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["HF_HUB_OFFLINE"] = "1"

import helpers  # noqa: E402
from transformers import (GenerationConfig, WhisperConfig, WhisperFeatureExtractor,  # noqa: E402
                          WhisperForConditionalGeneration, pipeline)
import transformers.pipelines.automatic_speech_recognition as asr  # noqa: E402
from transformers.models.whisper.tokenization_whisper import LANGUAGES  # noqa: E402

from oracle import whisper_ref as R  # noqa: E402  (only for the token-id constants / suppress list)


def hf_model(variant: str):
    cfg = WhisperConfig(vocab_size=51866, num_mel_bins=128, d_model=helpers.TINY["d_model"],
                        encoder_layers=helpers.TINY["enc_layers"], decoder_layers=helpers.TINY["dec_layers"],
                        encoder_attention_heads=helpers.TINY["heads"], decoder_attention_heads=helpers.TINY["heads"],
                        encoder_ffn_dim=helpers.TINY["ffn"], decoder_ffn_dim=helpers.TINY["ffn"],
                        max_source_positions=1500, max_target_positions=448, pad_token_id=50257, bos_token_id=50257,
                        eos_token_id=50257, decoder_start_token_id=50258)
    model = WhisperForConditionalGeneration(cfg).eval()
    sd = helpers.variant_state_dict(R.WhisperDims(**helpers.TINY), variant)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("proj_out" in m for m in missing), (missing, unexpected)
    model.tie_weights()
    model.generation_config = GenerationConfig(
        begin_suppress_tokens=list(R.BEGIN_SUPPRESS_TOKENS), suppress_tokens=list(R.SUPPRESS_TOKENS),
        max_initial_timestamp_index=50, max_length=448, is_multilingual=True, no_timestamps_token_id=50364,
        lang_to_id={f"<|{l}|>": 50259 + i for i, l in enumerate(LANGUAGES)},
        task_to_id={"transcribe": 50360, "translate": 50359}, return_timestamps=False, pad_token_id=50257,
        bos_token_id=50257, eos_token_id=50257, decoder_start_token_id=50258)
    return model, sd


def main():
    fe = WhisperFeatureExtractor(feature_size=128)
    if "--pipeline-only" not in sys.argv:
        part12(fe)
    part3(fe)
    print("golden fixtures written")


def part12(fe):
    # ---------------------------------------------------------------- 1. log-mel known answers
    A = (0.1 * np.random.default_rng(0).standard_normal(480000)).astype(np.float32)
    t = np.arange(480000) / 16000.0
    sine = (0.5 * np.sin(2 * np.pi * 440.0 * t)).astype(np.float32)
    kat = {}
    for name, pcm in (("noise30", A), ("noise10", A[:160000]), ("sine440", sine), ("short", A[:12345]),
                      ("mod", helpers.synth_clip(1, kind="mod"))):
        f = fe(pcm, sampling_rate=16000, return_tensors="np", return_attention_mask=True)
        x = f["input_features"][0]
        kat[name + "_sub"] = x[::8, ::25].copy()       # [16, 120] sub-sampled grid
        kat[name + "_stats"] = np.array([x.mean(), x.min(), x.max()], dtype=np.float64)
        kat[name + "_mask_sum"] = np.array(int(f["attention_mask"][0].sum()))
    np.savez_compressed(os.path.join(HERE, "logmel_kat.npz"), **kat)

    # ---------------------------------------------------------------- 2. model: encoder / logits / generate
    clips = [helpers.synth_clip(0), helpers.synth_clip(1, kind="mod"), helpers.synth_clip(2, seconds=11.3, kind="mod")]
    feats = torch.from_numpy(np.stack([fe(c, sampling_rate=16000, return_tensors="np")["input_features"][0] for c in clips]))
    feats_bf = feats.to(torch.bfloat16).float()   # what a bf16 pipeline feeds the model
    out = {}
    for variant in ("decisive", "varied"):
        model, _ = hf_model(variant)
        with torch.no_grad():
            enc = model.model.encoder(feats_bf).last_hidden_state
            out[f"{variant}_enc_sub"] = enc[:, ::50, ::8].numpy().copy()
            dec_ids = torch.tensor([[50258, 50259, 50360, 50365, 400, 401, 50400, 50400, 402]] * 3)
            logits = model(input_features=feats_bf, decoder_input_ids=dec_ids).logits
            out[f"{variant}_logits_sub"] = logits[:, :, ::97].numpy().copy()
            out[f"{variant}_logits_argmax"] = logits.argmax(-1).numpy().copy()
            gen = model.generate(input_features=feats_bf, return_timestamps=True, task="transcribe", num_beams=1,
                                 do_sample=False)
            out[f"{variant}_generate"] = gen.numpy().astype(np.int32)
            lang = model.detect_language(input_features=feats_bf)
            out[f"{variant}_langs"] = lang.numpy().astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "model_tiny.npz"), **out)



def part3(fe):
    # ---------------------------------------------------------------- 3. pipeline end to end
    tok = helpers.build_tokenizer()
    asr.ffmpeg_read = lambda b, sr: (np.frombuffer(wave.open(io.BytesIO(b)).readframes(10 ** 9), np.int16)
                                     .astype(np.float32) / 32768.0)
    # 71.3 s: with stride 0 the 30 s windows are exactly the three clips of the model fixtures (on which the
    # "decisive" model's greedy picks all have margins far above the bf16 tolerance)
    long_pcm = np.concatenate([helpers.synth_clip(0), helpers.synth_clip(1, kind="mod"),
                               helpers.synth_clip(2, seconds=11.3, kind="mod")])
    wav_path = "/tmp/golden_71s.wav"
    helpers.write_wav16(wav_path, long_pcm)
    results = {}
    # decisiveness of the stride-0 windows under PCM16 quantisation, measured with the oracle
    from oracle import logmel_ref as L
    q = helpers.quantize_pcm16(long_pcm)
    wfeats = torch.stack([torch.from_numpy(L.log_mel(q[i * 480000:(i + 1) * 480000])) for i in range(3)])
    tr = {}
    R.WhisperRef(R.WhisperDims(**helpers.TINY), helpers.variant_state_dict(R.WhisperDims(**helpers.TINY), "decisive")
                 ).generate(wfeats.to(torch.bfloat16).float(), trace=tr)
    results["decisive_30_0_min_margin"] = float(min(min(float(r["margin"].min()), float(r["rule_gap"].min()))
                                                    for it in tr["iterations"] for r in it["record"]))
    for variant in ("decisive", "varied"):
        model, _ = hf_model(variant)
        pipe = pipeline("automatic-speech-recognition", model=model, tokenizer=tok, feature_extractor=fe, device="cpu",
                        dtype=torch.float32)
        pipe.generation_config.num_beams = 1      # greedy oracle
        for (cl, st, bs) in ((30, 0, 24), (30, 5, 24), (60, 5, 32)):
            r = pipe(wav_path, chunk_length_s=cl, stride_length_s=st, batch_size=bs, generate_kwargs={"task": "transcribe"},
                     return_timestamps=True)
            results[f"{variant}_{cl}_{st}_{bs}"] = {"text": r["text"], "chunks": [
                {"timestamp": list(c["timestamp"]), "text": c["text"]} for c in r["chunks"]]}
        # through the reference's own entry point, when the reference tree is mounted
        if variant == "varied" and os.path.isdir("/root/reference/vocalis"):
            for m in ("librosa", "soundfile", "sherpa_onnx", "pydub"):
                sys.modules.setdefault(m, types.ModuleType(m))
            sys.modules["pydub"].AudioSegment = type("AudioSegment", (), {})
            sys.path.insert(0, "/root/reference")
            import vocalis.core.audio_pipeline as ap
            ap.LLM_AVAILABLE = False
            p = ap.AudioProcessingPipeline()
            p.transcription_model = pipe
            p.diarize = lambda *a, **k: []
            res = p.process_audio(wav_path, task="transcribe")
            results["reference_process_audio_varied"] = {
                "text": res["text"], "segments": [{"timestamp": list(c["timestamp"]), "text": c["text"]} for c in res["segments"]],
                "keys": sorted(res.keys())}
    with open(os.path.join(HERE, "pipeline_tiny.json"), "w") as f:
        json.dump(results, f, indent=1, ensure_ascii=False)


if __name__ == "__main__":
    main()
