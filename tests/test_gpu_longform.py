"""-m gpu tests of UN-CHUNKED long-form input (> 30 s without chunk_length_s; not the reference's call, which always
chunks — SURVEY.md section 8f): whole-clip features from tw_logmel_long against the oracle's numpy restatement of
WhisperFeatureExtractor(truncation=False), the long-form seek loop against the oracle's (pinned token-exact to
transformers, tests/golden/longform_tiny.json) under the margin rule of tests/test_gpu_engine.py, and the pipeline call."""
import json
import os

import numpy as np
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu
MARGIN_TOL = 0.30
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _clip():
    return np.concatenate([helpers.synth_clip(0), helpers.synth_clip(1, kind="mod"),
                           helpers.synth_clip(2, seconds=15.3, kind="mod")])


@pytest.mark.parametrize("n", [None, 2997 * 160 + 480000 + 333, 480000 + 161])
def test_longform_logmel_matches_oracle(cuda_device, n):
    """75.3 s fixture clip; a clip whose third window owns only the last few frames; one frame more than 30 s."""
    from oracle import logmel_ref as L
    from turbo_whisper_workspace_b200 import ops
    if n is None:
        audio = _clip()
    else:
        rng = np.random.default_rng(n)
        t = np.arange(n) / 16000.0
        audio = (0.3 * np.sin(2 * np.pi * (200 + 40 * t) * t) * (0.2 + np.abs(np.sin(0.7 * t)))
                 + 0.01 * rng.standard_normal(n)).astype(np.float32)
        audio[n // 3: n // 3 + 40000] = 0.0          # a silent stretch: values at the clamp floor
    got = ops.LogMel(cuda_device, 1).long(audio).float().cpu().numpy().T
    want = L.log_mel_long(audio)
    assert got.shape == want.shape == (128, audio.shape[0] // 160)
    # the output is bf16 (2^-9 relative rounding) of values the fp32 kernel computes within 1e-4
    err = np.abs(got - want)
    assert (err <= 1e-4 + np.abs(want) * 2.0 ** -8).all(), float(err.max())
    assert got.min() >= want.min() - 1e-2 and abs(float(got.max()) - float(want.max())) <= 1e-2


@pytest.mark.parametrize("variant", ["decisive", "varied"])
def test_longform_generate_matches_oracle(cuda_device, variant):
    from oracle import logmel_ref as L
    from oracle import whisper_ref as R
    from turbo_whisper_workspace_b200.config import WhisperDims
    from turbo_whisper_workspace_b200.engine import WhisperEngine
    gold = json.load(open(os.path.join(GOLD, "longform_tiny.json")))
    audio = _clip()
    rd = R.WhisperDims(**helpers.TINY)
    sd = helpers.variant_state_dict(rd, variant)
    ref = R.WhisperRef(rd, sd)
    eng = WhisperEngine(WhisperDims(**helpers.TINY), sd, device=cuda_device, max_batch=4)
    feats = torch.from_numpy(L.log_mel_long(audio))[None].to(torch.bfloat16).float()
    trace, etrace = {}, {}
    want = ref.generate(feats, trace=trace)[0]
    got = eng.generate_long_from_pcm(audio, trace=etrace)
    assert len(trace["iterations"]) >= 3, "the fixture is meant to walk several 30 s segments"
    agreed, identical_rows, first_diffs = helpers.compare_generate_traces(trace, etrace, MARGIN_TOL)
    assert agreed >= 20, f"free-running agreement is implausibly short: {first_diffs}"
    if identical_rows == [0]:
        assert got == want
        if feats.shape[-1] == 7530 and want == gold[variant]["tokens"]:
            assert got == gold[variant]["tokens"]
    # the short-form path of the same engine still works afterwards (other feature buffer, same graphs)
    assert len(eng.generate_from_pcm([audio[:480000]])[0]) > 0


def test_longform_pipeline_call(cuda_device):
    """The pipeline callable without chunk_length_s: dict with chunk timestamps that run past 30 s, HF's ValueError
    without timestamps, and a list call that mixes a long and a short input."""
    from oracle import whisper_ref as R
    from turbo_whisper_workspace_b200.config import WhisperDims
    from turbo_whisper_workspace_b200.pipeline import B200WhisperPipeline
    gold = json.load(open(os.path.join(GOLD, "longform_tiny.json")))["decisive"]
    sd = helpers.variant_state_dict(R.WhisperDims(**helpers.TINY), "decisive")
    pipe = B200WhisperPipeline(sd, WhisperDims(**helpers.TINY), helpers.build_tokenizer(), devices=[cuda_device], max_batch=4)
    audio = _clip()
    r = pipe(audio, return_timestamps=True, generate_kwargs={"task": "transcribe"})
    assert set(r) == {"text", "chunks"} and r["chunks"]
    assert list(r["chunks"][0]["timestamp"]) == gold["chunks"][0]["timestamp"]
    assert max(c["timestamp"][1] for c in r["chunks"] if c["timestamp"][1] is not None) > 30.0
    with pytest.raises(ValueError, match="long-form generation"):
        pipe(audio, generate_kwargs={"task": "transcribe"})
    both = pipe([audio[:16000 * 12], audio], return_timestamps=True)
    assert both[1]["text"] == r["text"] and both[0]["chunks"]
    pipe.close()
