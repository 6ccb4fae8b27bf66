"""-m gpu end-to-end tests through the reference-facing callable (HF ASR pipeline signature) against the
golden outputs of the real transformers pipeline / the reference's process_audio (tests/golden/)."""
import json
import os

import numpy as np
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return json.load(open(os.path.join(GOLD, "pipeline_tiny.json")))


@pytest.fixture(scope="module")
def long_wav(tmp_path_factory):
    pcm = np.concatenate([helpers.synth_clip(0), helpers.synth_clip(1, kind="mod"),
                          helpers.synth_clip(2, seconds=11.3, kind="mod")])
    p = tmp_path_factory.mktemp("audio") / "golden_71s.wav"
    helpers.write_wav16(p, pcm)
    return str(p), helpers.quantize_pcm16(pcm)


@pytest.fixture(scope="module")
def pipes(cuda_device):
    from oracle import whisper_ref as R
    from turbo_whisper_workspace_b200.config import WhisperDims
    from turbo_whisper_workspace_b200.pipeline import B200WhisperPipeline
    tok = helpers.build_tokenizer()
    out = {}
    for variant in ("decisive", "varied"):
        sd = helpers.variant_state_dict(R.WhisperDims(**helpers.TINY), variant)
        out[variant] = B200WhisperPipeline(sd, WhisperDims(**helpers.TINY), tok, devices=[cuda_device], max_batch=4)
    return out


def _norm(r):
    return {"text": r["text"], "chunks": [{"timestamp": list(c["timestamp"]), "text": c["text"]} for c in r["chunks"]]}


def _common_prefix(a, b):
    n = 0
    for x, y in zip(a, b):
        if x != y:
            break
        n += 1
    return n


@pytest.mark.parametrize("cl,st,bs", [(30, 0, 24), (30, 5, 24), (60, 5, 32)])
def test_pipeline_vs_hf_golden_decisive(pipes, gold, long_wav, cl, st, bs):
    """The reference's literal call (chunk_length_s=60, stride 5; ref:vocalis/core/audio_pipeline.py:351-358), the
    legacy 30 s call and a stride-0 call on the "decisive" fixture model.  The host logic is pinned exactly on CPU
    (tests/test_pipeline_host_golden.py); the GPU token ids are margin-checked in tests/test_gpu_engine.py.  The
    fp32 pipeline has a few near-tie picks on these windows (minimum margin recorded in the golden file), after
    which greedy paths may legitimately diverge, so here: output contract, first chunk identical, and a long common
    prefix of chunks with the transformers output."""
    path, _ = long_wav
    r = _norm(pipes["decisive"](path, chunk_length_s=cl, stride_length_s=st, batch_size=bs,
                                generate_kwargs={"task": "transcribe"}, return_timestamps=True))
    g = gold[f"decisive_{cl}_{st}_{bs}"]
    assert set(r) == {"text", "chunks"} and all(set(c) == {"timestamp", "text"} for c in r["chunks"])
    assert r["chunks"][0]["timestamp"] == g["chunks"][0]["timestamp"]
    n = _common_prefix(r["text"], g["text"])
    assert n >= min(len(g["text"]), 200), f"text diverges from the transformers output after {n} characters"


def test_pipeline_input_kinds_agree(pipes, long_wav):
    path, pcm = long_wav
    p = pipes["decisive"]
    kw = dict(chunk_length_s=30, stride_length_s=5, batch_size=24, generate_kwargs={"task": "transcribe"},
              return_timestamps=True)
    a = p(path, **kw)
    b = p(open(path, "rb").read(), **kw)
    c = p(pcm, **kw)
    d = p({"raw": pcm, "sampling_rate": 16000}, **kw)
    assert _norm(a) == _norm(b) == _norm(c) == _norm(d)


def test_pipeline_varied_structure(pipes, gold, long_wav):
    """Low-margin fixture: bf16 may legitimately flip near-tie picks, so check the contract, the first chunk
    and that most chunk boundaries agree with the fp32 pipeline."""
    path, _ = long_wav
    r = _norm(pipes["varied"](path, chunk_length_s=60, stride_length_s=5, batch_size=32,
                              generate_kwargs={"task": "transcribe"}, return_timestamps=True))
    g = gold["reference_process_audio_varied"]
    assert set(r) == {"text", "chunks"} and all(set(c) == {"timestamp", "text"} for c in r["chunks"])
    assert r["chunks"][0]["timestamp"][0] == g["segments"][0]["timestamp"][0]
    assert _common_prefix(r["text"], g["text"]) >= 20


def test_pipeline_short_clip_and_errors(pipes):
    p = pipes["decisive"]
    pcm = helpers.synth_clip(4, seconds=3.7)
    r = p(pcm, return_timestamps=True)
    assert isinstance(r["text"], str) and len(r["chunks"]) >= 1
    with pytest.raises(ValueError):
        p(pcm, chunk_length_s=10, stride_length_s=6, return_timestamps=True)
    with pytest.raises(ValueError):
        p(pcm, return_timestamps=True, generate_kwargs={"task": "summarize"})
    with pytest.raises(ValueError, match="alignment_heads"):     # HF's error for a generation config without them
        p(pcm, return_timestamps="word")
    with pytest.raises(ValueError, match="alignment_heads"):
        p(pcm, return_timestamps="word", generate_kwargs={"num_beams": 2})
    r2 = p(pcm)                      # HF default: no timestamps -> {"text"} only
    assert set(r2) == {"text"} and isinstance(r2["text"], str)
    # beam search through the pipeline: every window occupies num_beams decode rows (4-row engine: 2 windows x 2 beams)
    r3 = p(np.concatenate([pcm, pcm, pcm]), chunk_length_s=3, stride_length_s=0, batch_size=8, return_timestamps=True,
           generate_kwargs={"num_beams": 2})
    assert set(r3) == {"text", "chunks"} and len(r3["chunks"]) >= 1
    with pytest.raises(ValueError):
        p(pcm, return_timestamps=True, generate_kwargs={"num_beams": 5})   # 5 rows do not fit the 4-row fixture engine
    # explicit language / translate task use the forced-prompt path (no language detection step)
    gen = p.generation
    gen.lang_to_id = {"<|en|>": 50259, "<|fr|>": 50265}
    eng = p.scheduler.flat_engines[0]
    B = eng.load_pcm([pcm])
    eng.features(B)
    tr = {}
    eng.generate(B, task="translate", language="fr", trace=tr)
    assert tr["iterations"][0]["tokens"][0][:3] == [50258, 50265, 50359]
    with pytest.raises(ValueError):
        eng.generate(B, language="xx")


def test_multi_gpu_single_process_sharding(gold, long_wav):
    """WindowScheduler over every visible GPU (one host thread + engine contexts per device): same tokens as one GPU."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import whisper_ref as R
    from turbo_whisper_workspace_b200.config import WhisperDims
    from turbo_whisper_workspace_b200.scheduler import WindowScheduler
    from turbo_whisper_workspace_b200.config import GenerationSettings
    sd = helpers.variant_state_dict(R.WhisperDims(**helpers.TINY), "decisive")
    dims = WhisperDims(**helpers.TINY)
    clips = [helpers.synth_clip(50 + i, kind="mod" if i % 2 else "noise", seconds=30 if i % 3 else 12.5) for i in range(7)]
    one = WindowScheduler(sd, dims, GenerationSettings(), devices=["cuda:0"], max_batch=2, contexts_per_device=1)
    many = WindowScheduler(sd, dims, GenerationSettings(), devices=[f"cuda:{i}" for i in range(torch.cuda.device_count())],
                           max_batch=2, contexts_per_device=2)
    assert many.run(clips) == one.run(clips)
    assert many.last_stats["workers"] == torch.cuda.device_count()


def test_pipeline_ingests_other_sample_rates(pipes, tmp_path):
    """A 44.1 kHz stereo PCM16 WAV and a {"raw", "sampling_rate": 8000} dict go through the GPU ingest kernel and
    give the PCM torchaudio's resample (the reference's branch) produces; the transcription call runs on it."""
    import wave
    import torchaudio.functional as AF
    from turbo_whisper_workspace_b200 import pipeline as P
    pipe = pipes["decisive"]
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((44100 * 4, 2)) * 0.1).clip(-1, 1)
    xi = (x * 32767).round().astype("<i2")
    path = tmp_path / "stereo44k.wav"
    with wave.open(str(path), "wb") as wf:
        wf.setnchannels(2); wf.setsampwidth(2); wf.setframerate(44100); wf.writeframes(xi.tobytes())
    audio, _ = P.load_audio(str(path), 16000, pipe.ingest_device)
    want = AF.resample(torch.from_numpy(xi.astype(np.float32) / 32768.0).mean(dim=1), 44100, 16000).numpy()
    assert audio.shape == want.shape and float(np.abs(audio - want).max()) < 1e-5
    y = (rng.standard_normal(8000 * 3) * 0.1).astype(np.float32)
    audio2, _ = P.load_audio({"raw": y, "sampling_rate": 8000}, 16000, pipe.ingest_device)
    want2 = AF.resample(torch.from_numpy(y), 8000, 16000).numpy()
    assert audio2.shape == want2.shape and float(np.abs(audio2 - want2).max()) < 1e-5
    r = pipe(str(path), chunk_length_s=30, batch_size=4, return_timestamps=True)
    assert set(r) == {"text", "chunks"} and len(r["chunks"]) >= 1
