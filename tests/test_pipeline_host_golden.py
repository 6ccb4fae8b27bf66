"""CPU test: the host side of B200WhisperPipeline (audio ingest, windowing, stride plumbing, batch padding,
_decode_asr call) against the golden dicts produced by the real transformers pipeline and by the reference's own
process_audio.  The GPU engines are replaced by a scheduler that returns the ORACLE's token ids (which
tests/test_oracle_golden.py shows to be token-exact with WhisperGenerationMixin.generate), so the comparison is exact
and isolates the host logic from bf16 numerics."""
import json
import os

import numpy as np
import pytest
import torch

import helpers
from oracle import logmel_ref as L
from oracle import whisper_ref as R
from turbo_whisper_workspace_b200.config import WhisperDims
from turbo_whisper_workspace_b200.pipeline import B200WhisperPipeline

WORD_HEADS = [[0, 1], [1, 0], [1, 3]]      # tests/golden/make_golden_word.py
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class OracleScheduler:
    def __init__(self, variant):
        rd = R.WhisperDims(**helpers.TINY)
        self.ref = R.WhisperRef(rd, helpers.variant_state_dict(rd, variant))
        self.last_stats = {}

    def run(self, clips, task="transcribe", language=None, return_timestamps=True, token_timestamps=False, group=None,
            num_beams=1):
        feats = torch.stack([torch.from_numpy(L.log_mel(c)) for c in clips])
        if not token_timestamps:
            return self.ref.generate(feats, task=task, return_timestamps=return_timestamps, num_beams=num_beams)
        # word timestamps: one generate call per HF batch of `group` consecutive windows; num_frames = the feature
        # extractor's attention-mask sum
        rows = []
        for g0 in range(0, len(clips), group):
            sub = clips[g0:g0 + group]
            nf = [min(3000, -(-len(c) // 160)) for c in sub]
            ts = {}
            ids = self.ref.generate(feats[g0:g0 + group], task=task, alignment_heads=WORD_HEADS, num_frames=nf, token_ts=ts,
                                    num_beams=num_beams)
            rows += [(ids[b], ts["segments"][b]) for b in range(len(sub))]
        return rows

    def run_long(self, audio, task="transcribe", language=None, num_beams=1):
        feats = torch.from_numpy(L.log_mel_long(audio))[None]
        return self.ref.generate(feats, task=task, num_beams=num_beams)[0]


@pytest.fixture(scope="module")
def wav(tmp_path_factory):
    pcm = np.concatenate([helpers.synth_clip(0), helpers.synth_clip(1, kind="mod"),
                          helpers.synth_clip(2, seconds=11.3, kind="mod")])
    p = tmp_path_factory.mktemp("audio") / "golden_71s.wav"
    helpers.write_wav16(p, pcm)
    return str(p)


def _norm(r):
    return {"text": r["text"], "chunks": [{"timestamp": list(c["timestamp"]), "text": c["text"]} for c in r["chunks"]]}


@pytest.mark.parametrize("variant,cl,st,bs", [("varied", 30, 0, 24), ("varied", 30, 5, 24), ("varied", 60, 5, 32),
                                              ("decisive", 30, 5, 24)])
def test_host_pipeline_matches_hf_golden(wav, variant, cl, st, bs):
    gold = json.load(open(os.path.join(GOLD, "pipeline_tiny.json")))
    pipe = B200WhisperPipeline(None, WhisperDims(**helpers.TINY), helpers.build_tokenizer(),
                               scheduler=OracleScheduler(variant))
    # NOTE: HF feeds fp32 features to the fp32 model here, so the oracle scheduler does too
    r = pipe(wav, chunk_length_s=cl, stride_length_s=st, batch_size=bs, generate_kwargs={"task": "transcribe"},
             return_timestamps=True)
    assert _norm(r) == gold[f"{variant}_{cl}_{st}_{bs}"]
    if (variant, cl, st) == ("varied", 60, 5):
        ref = gold["reference_process_audio_varied"]   # produced through vocalis...process_audio
        assert r["text"] == ref["text"] and _norm(r)["chunks"] == ref["segments"]


@pytest.mark.parametrize("variant,cl,st,bs", [("varied", 30, 5, 24), ("decisive", 30, 0, 24)])
def test_host_pipeline_without_timestamps_matches_hf_golden(wav, variant, cl, st, bs):
    """return_timestamps omitted (the HF default): <|notimestamps|> generation + text-only merge of the windows;
    the oracle's no-timestamp generate is pinned token-exact on the same golden file (test_oracle_golden.py)."""
    gold = json.load(open(os.path.join(GOLD, "notimestamps_tiny.json")))
    pipe = B200WhisperPipeline(None, WhisperDims(**helpers.TINY), helpers.build_tokenizer(),
                               scheduler=OracleScheduler(variant))
    r = pipe(wav, chunk_length_s=cl, stride_length_s=st, batch_size=bs, generate_kwargs={"task": "transcribe"})
    g = gold[f"{variant}_{cl}_{st}_{bs}"]
    assert sorted(r.keys()) == g["keys"] and r["text"] == g["text"]


@pytest.mark.parametrize("variant", ["varied", "decisive"])
def test_host_pipeline_word_timestamps_match_hf_golden(wav, variant):
    """return_timestamps="word": per-token times of the oracle (pinned time-exact to generate(return_token_timestamps=
    True) in tests/test_oracle_golden.py) through the pipeline's stride plumbing and the word-level _decode_asr
    restatement -> the {"text", "chunks"} of the transformers pipeline, chunked (30 / 5) and for a single short clip."""
    gold = json.load(open(os.path.join(GOLD, "word_tiny.json")))
    pipe = B200WhisperPipeline(None, WhisperDims(**helpers.TINY), helpers.build_tokenizer(),
                               scheduler=OracleScheduler(variant))
    r = pipe(wav, chunk_length_s=30, stride_length_s=5, batch_size=24, generate_kwargs={"task": "transcribe"},
             return_timestamps="word")
    assert _norm(r) == gold[f"{variant}_30_5_24"]
    r = pipe(helpers.synth_clip(2, seconds=11.3, kind="mod"), generate_kwargs={"task": "transcribe"},
             return_timestamps="word")
    assert _norm(r) == gold[f"{variant}_single"]
    if variant == "varied":
        # the reference's literal chunking (60 / 5, batch 32; windows truncated to 30 s) with word timestamps
        r = pipe(wav, chunk_length_s=60, stride_length_s=5, batch_size=32, generate_kwargs={"task": "transcribe"},
                 return_timestamps="word")
        assert _norm(r) == gold["varied_60_5_32"]
    if variant == "decisive":
        # word timestamps under beam search (the reference's literal decoding mode with transformers >= 4.53) go through
        # the same host plumbing: one (start, end) per word, monotone within a window
        r = pipe(helpers.synth_clip(2, seconds=11.3, kind="mod"), generate_kwargs={"num_beams": 3}, return_timestamps="word")
        assert r["chunks"] and all(len(c["timestamp"]) == 2 and isinstance(c["text"], str) for c in r["chunks"])


def test_host_pipeline_edge_inputs_match_hf_golden():
    """Edge cases of the call the reference makes (tests/golden/make_golden_edges.py): empty audio without chunking is
    one zero-padded 30 s window, 100 samples with chunking likewise, a last window that ends exactly at the end of the
    audio, and the StopIteration transformers raises for empty audio with chunk_length_s."""
    gold = json.load(open(os.path.join(GOLD, "edges_tiny.json")))
    pipe = B200WhisperPipeline(None, WhisperDims(**helpers.TINY), helpers.build_tokenizer(),
                               scheduler=OracleScheduler("varied"))
    cases = [("empty_plain", np.zeros(0, np.float32), {}),
             ("tiny_chunked", np.zeros(100, np.float32), dict(chunk_length_s=30, stride_length_s=5)),
             ("exact_multiple", helpers.synth_clip(5, seconds=40.0), dict(chunk_length_s=30, stride_length_s=5)),
             ("empty_chunked", np.zeros(0, np.float32), dict(chunk_length_s=30, stride_length_s=5))]
    # both modes are in the golden file; each input runs one of them here to keep the CPU suite short
    modes = {"empty_plain": (True,), "tiny_chunked": (None,), "exact_multiple": (True,), "empty_chunked": (True, None)}
    for name, x, kw in cases:
        for rt in modes[name]:
            g = gold[f"{name}_{'ts' if rt else 'nots'}"]
            call = lambda: pipe(x.copy(), batch_size=4, return_timestamps=rt, generate_kwargs={"task": "transcribe"}, **kw)
            if "raises" in g:
                assert g["raises"] == "StopIteration"
                with pytest.raises(StopIteration):
                    call()
                continue
            r = call()
            assert sorted(r.keys()) == g["keys"] and r["text"] == g["text"], (name, rt)
            if rt:
                assert _norm(r)["chunks"] == g["chunks"], (name, rt)



@pytest.mark.parametrize("variant", ["varied", "varied_beams3"])     # "decisive" (1259 tokens): tests/test_oracle_golden.py
def test_host_pipeline_unchunked_longform_matches_hf_golden(variant):
    """A 75.3 s clip WITHOUT chunk_length_s: HF extracts features of the whole clip and generate's seek loop walks all 7530
    frames (tests/golden/make_golden_longform.py).  The host path (windowing decision, the one-row model output without
    stride, native _decode_asr over timestamps that restart in every 30 s segment) must give HF's dict; without
    timestamps HF's ValueError; the oracle's long-form generate is pinned token-exact in tests/test_oracle_golden.py."""
    gold = json.load(open(os.path.join(GOLD, "longform_tiny.json")))
    pcm = np.concatenate([helpers.synth_clip(0), helpers.synth_clip(1, kind="mod"),
                          helpers.synth_clip(2, seconds=15.3, kind="mod")])
    assert pcm.shape[0] == gold["samples"]
    model = variant.split("_")[0]
    pipe = B200WhisperPipeline(None, WhisperDims(**helpers.TINY), helpers.build_tokenizer(), scheduler=OracleScheduler(model))
    kw = {"task": "transcribe"}
    if variant.endswith("beams3"):
        kw["num_beams"] = 3
    r = pipe(pcm, return_timestamps=True, generate_kwargs=kw)
    assert _norm(r) == {"text": gold[variant]["text"], "chunks": gold[variant]["chunks"]}
    if variant == "varied":
        with pytest.raises(ValueError, match="long-form generation which requires the model to predict timestamp tokens"):
            pipe(pcm, generate_kwargs={"task": "transcribe"})
        with pytest.raises(NotImplementedError):
            pipe(pcm, return_timestamps="word")
        # a list call mixing a long clip and a short one keeps the order of the inputs
        both = pipe([pcm[:16000 * 20], pcm], return_timestamps=True, generate_kwargs=kw)
        assert _norm(both[1]) == _norm(r) and both[0]["chunks"]
