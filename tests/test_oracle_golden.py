"""CPU tests: the oracle restatements against the committed golden fixtures generated from the installed
third-party implementation (tests/golden/make_golden.py) and the known-answer values of SURVEY.md §8c."""
import json
import os

import numpy as np
import pytest
import torch

import helpers
from oracle import logmel_ref as L
from oracle import whisper_ref as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def kat():
    return np.load(os.path.join(GOLD, "logmel_kat.npz"))


@pytest.fixture(scope="module")
def model_gold():
    return np.load(os.path.join(GOLD, "model_tiny.npz"))


def _kat_inputs():
    A = (0.1 * np.random.default_rng(0).standard_normal(480000)).astype(np.float32)
    t = np.arange(480000) / 16000.0
    return {"noise30": A, "noise10": A[:160000], "sine440": (0.5 * np.sin(2 * np.pi * 440.0 * t)).astype(np.float32),
            "short": A[:12345], "mod": helpers.synth_clip(1, kind="mod")}


@pytest.mark.parametrize("name", ["noise30", "noise10", "sine440", "short", "mod"])
def test_logmel_oracle_vs_hf_golden(kat, name):
    pcm = _kat_inputs()[name]
    x = L.log_mel(pcm)
    np.testing.assert_allclose(x[::8, ::25], kat[name + "_sub"], atol=2e-5, rtol=0)
    np.testing.assert_allclose([x.mean(), x.min(), x.max()], kat[name + "_stats"], atol=2e-5)
    assert int(L.attention_mask(len(pcm)).sum()) == int(kat[name + "_mask_sum"])


def test_logmel_survey_known_answers():
    """Values measured from the HF torch path during the survey."""
    A = _kat_inputs()["noise30"]
    x = L.log_mel(A)
    assert x.shape == (128, 3000)
    np.testing.assert_allclose(x[0, :4], [0.45651639, 0.56622541, 0.60694182, 0.39072639], atol=1e-5)
    np.testing.assert_allclose(x[127, -4:], [0.74227071, 0.59086835, 0.66143847, 0.55686784], atol=1e-5)
    assert abs(x.mean() - 0.597388) < 1e-5 and abs(x.min() + 0.501018) < 1e-5 and abs(x.max() - 0.942317) < 1e-5
    x10 = L.log_mel(A[:160000])
    assert int(L.attention_mask(160000).sum()) == 1000
    np.testing.assert_allclose(x10[0, 999:1002], [0.65147579, 0.54963934, -0.22125149], atol=1e-5)
    assert abs(x10[0, -1] + 1.05846596) < 1e-5
    s = L.log_mel(_kat_inputs()["sine440"])
    assert int(s[:, 1500].argmax()) == 18 and abs(s.max() - 1.48535407) < 1e-5 and abs(s.min() + 0.51464593) < 1e-5
    w = L.hann_periodic()
    assert w[0] == 0 and abs(w[200] - 1) < 1e-12 and abs(w[399] - 6.169e-5) < 1e-7


def test_mel_filterbank_structure():
    fb = L.mel_filter_bank()
    nz = fb != 0
    assert fb.shape == (201, 128) and int(nz.sum()) == 394 and int(nz.sum(0).max()) == 9 and int(nz.sum(1).max()) == 2


def _clips_feats():
    clips = [helpers.synth_clip(0), helpers.synth_clip(1, kind="mod"), helpers.synth_clip(2, seconds=11.3, kind="mod")]
    feats = torch.stack([torch.from_numpy(L.log_mel(c)) for c in clips])
    return clips, feats.to(torch.bfloat16).float()


@pytest.mark.parametrize("variant", ["decisive", "varied"])
def test_model_oracle_vs_hf_golden(model_gold, variant):
    _, fb = _clips_feats()
    dims = R.WhisperDims(**helpers.TINY)
    ref = R.WhisperRef(dims, helpers.variant_state_dict(dims, variant))
    enc = ref.encode(fb)
    np.testing.assert_allclose(enc[:, ::50, ::8].numpy(), model_gold[f"{variant}_enc_sub"], atol=2e-4, rtol=1e-4)
    dec_ids = torch.tensor([[50258, 50259, 50360, 50365, 400, 401, 50400, 50400, 402]] * 3)
    logits = ref.decode(dec_ids, enc)
    np.testing.assert_allclose(logits[:, :, ::97].numpy(), model_gold[f"{variant}_logits_sub"], atol=2e-3, rtol=1e-4)
    assert ref.detect_language(enc, R.GenConfig()) == model_gold[f"{variant}_langs"].tolist()


@pytest.mark.parametrize("variant", ["decisive", "varied"])
def test_generate_oracle_vs_hf_golden(model_gold, variant):
    """Token-exact agreement of the oracle's seek loop / logits processors with WhisperGenerationMixin.generate."""
    _, fb = _clips_feats()
    dims = R.WhisperDims(**helpers.TINY)
    ref = R.WhisperRef(dims, helpers.variant_state_dict(dims, variant))
    got = ref.generate(fb)
    gold = model_gold[f"{variant}_generate"]
    for b, row in enumerate(got):
        want = gold[b].tolist()
        while want and want[-1] == 50257:
            want.pop()
        assert row == want, f"row {b}"


@pytest.mark.parametrize("variant", ["decisive", "varied"])
def test_longform_generate_oracle_vs_hf_golden(variant):
    """Un-chunked long-form input (75.3 s = 7530 frames in ONE generate call): the oracle's whole-clip log-mel equals the
    HF extractor's (truncation=False) and its seek loop over all frames is token-exact with transformers
    (tests/golden/make_golden_longform.py)."""
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "longform_tiny.json")))
    pcm = np.concatenate([helpers.synth_clip(0), helpers.synth_clip(1, kind="mod"),
                          helpers.synth_clip(2, seconds=15.3, kind="mod")])
    feats = L.log_mel_long(pcm)
    assert feats.shape == (128, pcm.shape[0] // 160)
    if variant == "varied":
        from transformers import WhisperFeatureExtractor
        hf = WhisperFeatureExtractor(feature_size=128)(pcm, sampling_rate=16000, truncation=False, padding="longest",
                                                      return_tensors="np")["input_features"][0]
        assert np.abs(hf - feats).max() <= 1e-4
        # the first 2998 frames do not see the clip-global clamp or the far end: same as the 30 s extractor's
        # unless the global maximum lifts the floor
        short = L.log_mel(pcm[:480000])
        assert np.abs(np.maximum(short[:, :2998], feats.min()) - feats[:, :2998]).max() <= 1e-4
    dims = R.WhisperDims(**helpers.TINY)
    ref = R.WhisperRef(dims, helpers.variant_state_dict(dims, variant))
    assert ref.generate(torch.from_numpy(feats)[None])[0] == gold[variant]["tokens"]
    with pytest.raises(ValueError):
        ref.generate(torch.from_numpy(feats)[None], return_timestamps=False)


@pytest.mark.parametrize("variant", ["decisive", "varied"])
def test_generate_without_timestamps_oracle_vs_hf_golden(variant):
    """generate(return_timestamps=False): <|notimestamps|> in the prompt, suppress lists only — token-exact with
    transformers, including the extra seek iterations the (unmasked) timestamp ids of these random models cause."""
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "notimestamps_tiny.json")))
    _, fb = _clips_feats()
    dims = R.WhisperDims(**helpers.TINY)
    ref = R.WhisperRef(dims, helpers.variant_state_dict(dims, variant))
    got = ref.generate(fb, return_timestamps=False)
    for b, row in enumerate(got):
        want = list(gold[f"{variant}_generate"][b])
        while want and want[-1] == 50257:
            want.pop()
        assert row == want, f"row {b}"


@pytest.mark.parametrize("variant", ["decisive", "varied"])
def test_beam_search_oracle_vs_hf_golden(variant):
    """generate(num_beams=5) — SURVEY.md §8f rank 1, the mode the reference's literal pipeline call runs under
    transformers >= 4.53: the oracle's restatement of GenerationMixin._beam_search (log-softmax before the processors,
    2*num_beams candidates, length-penalised finished slots, early-stop heuristic, cache re-gather) is token-exact with
    transformers on both fixture models, seek loop included.  Groundwork for the GPU beam path of the next round."""
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "beams_tiny.json")))
    _, fb = _clips_feats()
    dims = R.WhisperDims(**helpers.TINY)
    ref = R.WhisperRef(dims, helpers.variant_state_dict(dims, variant))
    got = ref.generate(fb, num_beams=5)
    for b, row in enumerate(got):
        want = list(gold[f"{variant}_generate_beams5"][b])
        while want and want[-1] == 50257:
            want.pop()
        assert row == want, f"row {b}"


@pytest.mark.parametrize("variant", ["decisive", "varied"])
def test_token_timestamps_oracle_vs_hf_golden(variant):
    """generate(return_token_timestamps=True, return_segments=True, attention_mask=...) — SURVEY.md §8f rank 4: the
    oracle's eager cross-attention tap, per-row crop to (num_frames - seek) // 2 frames (python slice rules, negative
    counts included), population-std normalisation, median filter, head mean, float32-cost DTW and jump extraction
    give the SAME token ids and the same fp32 time for every token as transformers, for both the padded
    `token_timestamps` output and the per-segment values the ASR pipeline reads, across seek iterations."""
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "word_tiny.json")))
    _, fb = _clips_feats()
    dims = R.WhisperDims(**helpers.TINY)
    ref = R.WhisperRef(dims, helpers.variant_state_dict(dims, variant))
    ts = {}
    got = ref.generate(fb, alignment_heads=gold["alignment_heads"], num_frames=gold["num_frames"], token_ts=ts)
    g = gold[f"{variant}_generate"]
    for b, row in enumerate(got):
        assert row == g["segment_tokens"][b], f"row {b}"
        np.testing.assert_array_equal(np.asarray(ts["segments"][b], dtype=np.float32),
                                      np.asarray(g["segment_token_timestamps"][b], dtype=np.float32))
        np.testing.assert_array_equal(np.asarray(ts["sequences"][b], dtype=np.float32),
                                      np.asarray(g["token_timestamps"][b][:len(row)], dtype=np.float32))


@pytest.mark.parametrize("variant", ["varied"])
def test_token_timestamps_under_beam_search_oracle_vs_hf_golden(variant):
    """generate(num_beams=5, return_token_timestamps=True): the oracle's beam search also tracks HF's `beam_indices`
    (same gathers as the sequences, batch-offset parent row per generated position, -1 beyond the hypothesis) and the
    token-timestamp stage gathers every position's alignment rows from the beam that produced it — tokens and fp32
    times identical to transformers.  Groundwork for word timestamps with beam search on the GPU path (the engine
    raises NotImplementedError for that combination today)."""
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "word_beams_tiny.json")))
    _, fb = _clips_feats()
    dims = R.WhisperDims(**helpers.TINY)
    ref = R.WhisperRef(dims, helpers.variant_state_dict(dims, variant))
    ts = {}
    got = ref.generate(fb, num_beams=5, alignment_heads=gold["alignment_heads"], num_frames=gold["num_frames"], token_ts=ts)
    g = gold[f"{variant}_generate_beams5"]
    for b, row in enumerate(got):
        assert row == g["segment_tokens"][b], f"row {b}"
        np.testing.assert_array_equal(np.asarray(ts["segments"][b], dtype=np.float32),
                                      np.asarray(g["segment_token_timestamps"][b], dtype=np.float32))
        np.testing.assert_array_equal(np.asarray(ts["sequences"][b], dtype=np.float32),
                                      np.asarray(g["token_timestamps"][b][:len(row)], dtype=np.float32))


def test_dtw_oracle_and_native_match_transformers_function():
    """The oracle's anti-diagonal DTW and the C library's tw_dtw_token_frames against transformers'
    _dynamic_time_warping + jump extraction on random matrices, tie-heavy integer matrices included."""
    import ctypes as C
    from transformers.models.whisper.generation_whisper import _dynamic_time_warping
    from turbo_whisper_workspace_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for trial in range(32):
        n, m = int(rng.integers(1, 32)), int(rng.integers(1, 64))
        mat = (rng.integers(-2, 3, size=(n, m)) if trial % 4 == 0 else rng.standard_normal((n, m))).astype(np.float32)
        ld = m + int(rng.integers(0, 5))
        buf = np.zeros((n, ld), dtype=np.float32)
        buf[:, :m] = mat
        out = np.empty(n, dtype=np.int32)
        assert lib.tw_dtw_token_frames(buf.ctypes.data_as(C.c_void_p), ld, n, m, out.ctypes.data_as(C.c_void_p)) == 0
        ti, tj = _dynamic_time_warping(-mat.astype(np.float64))
        want = tj[np.pad(np.diff(ti), (1, 0), constant_values=1).astype(bool)].tolist()
        assert out.tolist() == want and R.WhisperRef.dtw_token_frames(-mat.astype(np.float64)) == want, trial
    assert lib.tw_dtw_token_frames(None, 4, 2, 4, None) != 0 and b"tw_dtw_token_frames" in lib.tw_last_error()
    # the threaded batch entry point: same frames as the single-window call, -1 everywhere for an empty frame axis
    B, T, S, n = 5, 12, 40, 9
    mats = rng.standard_normal((B, T, S)).astype(np.float32)
    nf = np.array([40, 0, 17, 1, 40], dtype=np.int32)
    got = np.empty((B, n), dtype=np.int32)
    assert lib.tw_dtw_token_frames_batch(mats.ctypes.data_as(C.c_void_p), T * S, S, B, n, nf.ctypes.data_as(C.c_void_p),
                                         got.ctypes.data_as(C.c_void_p), 3) == 0
    for b in range(B):
        want = [-1] * n if nf[b] == 0 else R.WhisperRef.dtw_token_frames(-mats[b, :n, :nf[b]].astype(np.float64))
        assert got[b].tolist() == want, b


def test_retrieve_segment_cases():
    TB = R.TIMESTAMP_BEGIN
    f = R.WhisperRef.retrieve_segment
    # single closing timestamp: everything consumed
    segs, adv = f([TB, 5, 6, TB + 100, TB + 100, 7, TB + 200], 3000, TB)
    assert adv == 3000 and segs == [[TB, 5, 6, TB + 100], [TB + 100, 7, TB + 200]]
    # pair but no closing timestamp: keep up to the last pair, seek to it
    segs, adv = f([TB, 5, TB + 100, TB + 100, 7, 8], 3000, TB)
    assert adv == 200 and segs == [[TB, 5, TB + 100, TB + 100]]
    # no consecutive pair: one segment, window consumed
    segs, adv = f([TB, 5, 6, 7], 1234, TB)
    assert adv == 1234 and segs == [[TB, 5, 6, 7]]


def test_pipeline_golden_is_consistent_with_reference_entry_point():
    """The fixture produced through the reference's own process_audio equals the HF pipeline call with the
    reference's literal arguments (chunk_length_s=60, stride_length_s=5, batch_size=32 on CPU)."""
    r = json.load(open(os.path.join(GOLD, "pipeline_tiny.json")))
    ref = r["reference_process_audio_varied"]
    assert ref["text"] == r["varied_60_5_32"]["text"]
    assert ref["segments"] == r["varied_60_5_32"]["chunks"]
    assert {"text", "segments", "merged_segments", "duration", "processing_times"} <= set(ref["keys"])
