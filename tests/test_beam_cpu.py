"""CPU test of the product's beam-search step (csrc/beam.cu): `tw_beam_step_host` runs the SAME `__host__ __device__`
bookkeeping code the kernels run (window plan, hypothesis records, grammar state), with the per-row selection done
serially.  Driven with the fp32 oracle's decoder logits it must reproduce the oracle's beam search — which
tests/test_oracle_golden.py pins token-exact to transformers' generate(num_beams=5) — including HF's `beam_indices`
and the beam-of-origin rows the KV cache is re-gathered by.  No compute call touches a GPU."""
import ctypes as C

import numpy as np
import pytest
import torch

import helpers
from oracle import logmel_ref as L
from oracle import whisper_ref as R
from turbo_whisper_workspace_b200 import _lib
from turbo_whisper_workspace_b200.engine import _bitmap


class HostBeam:
    """Host-memory twin of the device state the engine keeps (engine.BeamWorkspace)."""

    def __init__(self, gc, n, K, P, prompt, vocab, timestamps=True, track=True):
        self.lib = _lib.load()
        c = _lib.BeamConfig()
        c.num_beams, c.vocab, c.max_length, c.prompt_len = K, vocab, gc.max_length, P
        c.eos, c.pad, c.no_timestamps = gc.eos_token_id, gc.pad_token_id, gc.no_timestamps_token_id
        c.max_initial_ts = -1 if gc.max_initial_timestamp_index is None else gc.max_initial_timestamp_index
        c.timestamps, c.track_indices, c.length_penalty = int(timestamps), int(track), 1.0
        self.cfg, self.n, self.K, self.P = c, n, K, P
        W = self.W = int(self.lib.tw_beam_record_width(C.byref(c)))
        L_ = gc.max_length
        self.hist = np.full((2, n, K, W), gc.pad_token_id, dtype=np.int32)
        self.fin = np.full((2, n, K, W), gc.pad_token_id, dtype=np.int32)
        if track:
            self.hist[..., L_:] = -1
            self.fin[..., L_:] = -1
        self.hist[0, :, :, :P] = prompt[:, None, :]
        self.run_score = np.full((n, K), -1.0e9, dtype=np.float32)
        self.run_score[:, 0] = 0
        self.fin_score = np.full((n, K), -1.0e9, dtype=np.float32)
        self.fin_flag = np.zeros((n, K), dtype=np.int32)
        self.fin_len = np.zeros((n, K), dtype=np.int32)
        self.gram = np.tile(np.array([0, 1, -1, 0], dtype=np.int32), (n, K, 1))
        self.improvable = np.ones(n, dtype=np.int32)
        self.hits_all = np.zeros(n, dtype=np.int32)
        self.ctrl = np.zeros(8, dtype=np.int32)
        s = _lib.BeamState()
        for name in ("hist", "fin", "run_score", "fin_score", "fin_flag", "fin_len", "gram", "improvable", "hits_all", "ctrl"):
            setattr(s, name, getattr(self, name).ctypes.data)
        self.state = s
        self.sup = _bitmap(gc.suppress_tokens, vocab)
        self.bsup = _bitmap(gc.begin_suppress_tokens, vocab)
        self.cur = P

    def step(self, logits: np.ndarray):
        R_ = self.n * self.K
        nxt = np.zeros(R_, dtype=np.int32)
        origin = np.zeros(R_, dtype=np.int32)
        lg = np.ascontiguousarray(logits, dtype=np.float32)
        _lib.check(self.lib.tw_beam_step_host(C.byref(self.cfg), C.byref(self.state), lg.ctypes.data, self.cur,
                                              self.sup.ctypes.data, self.bsup.ctypes.data, nxt.ctypes.data,
                                              origin.ctypes.data, self.n), "tw_beam_step_host")
        self.cur += 1
        return nxt, origin

    @property
    def done(self):
        return bool(self.ctrl[1])

    def result(self, gc):
        par = int(self.ctrl[0])
        m = int(self.fin_len[:, 0].max())
        toks = self.fin[par, :, 0, self.P:self.P + m]
        idx = self.fin[par, :, 0, gc.max_length:gc.max_length + m]
        return torch.from_numpy(toks.astype(np.int64)), torch.from_numpy(idx.astype(np.int64))


def _drive(ref, enc, prompt, gc, K, timestamps=True):
    n, P = prompt.shape
    hb = HostBeam(gc, n, K, P, prompt.numpy().astype(np.int32), ref.dims.vocab, timestamps=timestamps)
    encK = enc.repeat_interleave(K, dim=0)
    cache = ref.new_cache()
    rows = prompt.repeat_interleave(K, dim=0)
    logits = ref.decode(rows, encK, cache, 0)[:, -1]
    origins = []
    while True:
        nxt, origin = hb.step(logits.numpy())
        if hb.done:
            break
        origins.append(origin.copy())
        sel = torch.from_numpy(origin.astype(np.int64))
        for layer in cache:
            for kind in ("self", "cross"):
                for name in ("k", "v"):
                    if name in layer[kind]:
                        layer[kind][name] = layer[kind][name].index_select(0, sel)
        logits = ref.decode(torch.from_numpy(nxt.astype(np.int64))[:, None], encK, cache, hb.cur - 1)[:, -1]
    return hb, origins


@pytest.mark.parametrize("variant,K", [("decisive", 5), ("varied", 3)])
def test_native_beam_step_matches_oracle(variant, K):
    clips = [helpers.synth_clip(0), helpers.synth_clip(2, seconds=11.3, kind="mod")]
    feats = torch.stack([torch.from_numpy(L.log_mel(c)) for c in clips]).to(torch.bfloat16).float()
    dims = R.WhisperDims(**helpers.TINY)
    ref = R.WhisperRef(dims, helpers.variant_state_dict(dims, variant))
    gc = R.GenConfig()
    enc = ref.encode(feats)
    langs = ref.detect_language(enc, gc)
    prompt = torch.tensor([[gc.decoder_start_token_id, langs[b], gc.task_to_id["transcribe"]] for b in range(2)])
    aux = {}
    want = ref.beam_search(enc, prompt, gc, num_beams=K, aux=aux)
    hb, origins = _drive(ref, enc, prompt, gc, K)
    got, idx = hb.result(gc)
    assert got.shape == want.shape and torch.equal(got, want)
    # HF's `beam_indices` of the returned hypotheses (the oracle's are pinned through the token timestamps they select,
    # tests/test_oracle_golden.py::test_token_timestamps_under_beam_search_oracle_vs_hf_golden)
    bi = aux["beam_indices"]
    assert torch.equal(idx[:, :bi.shape[1]], bi)
    # every origin stays inside its window (the KV re-gather never crosses windows)
    for o in origins:
        assert all(o[r] // K == r // K for r in range(2 * K))


def test_native_beam_step_without_timestamps_matches_oracle():
    clips = [helpers.synth_clip(1, kind="mod")]
    feats = torch.stack([torch.from_numpy(L.log_mel(c)) for c in clips]).to(torch.bfloat16).float()
    dims = R.WhisperDims(**helpers.TINY)
    ref = R.WhisperRef(dims, helpers.variant_state_dict(dims, "varied"))
    gc = R.GenConfig()
    enc = ref.encode(feats)
    langs = ref.detect_language(enc, gc)
    prompt = torch.tensor([[gc.decoder_start_token_id, langs[0], gc.task_to_id["transcribe"], gc.no_timestamps_token_id]])
    want = ref.beam_search(enc, prompt, gc, num_beams=3, timestamps=False)
    hb, _ = _drive(ref, enc, prompt, gc, 3, timestamps=False)
    got, _ = hb.result(gc)
    assert got.shape == want.shape and torch.equal(got, want)


def test_native_beam_masks_match_oracle_processors():
    """One host step on random logits and random histories: the ids the native step can pick (finite candidates) are
    exactly the ids the oracle's processors leave finite, for every grammar situation (begin, after a closed pair,
    after an opening timestamp, monotone timestamps, the timestamp-probability rule)."""
    gc = R.GenConfig()
    g = torch.Generator().manual_seed(0)
    V, TB = 51866, gc.no_timestamps_token_id + 1
    lib = _lib.load()
    for trial in range(30):
        glen = [0, 1, 2, 5, 9][trial % 5]
        gen = torch.randint(0, 50257, (1, glen), generator=g)
        if glen:
            ts_mask = torch.rand(1, glen, generator=g) < 0.4
            ts_vals = torch.sort(torch.randint(TB, TB + 1500, (1, glen), generator=g), dim=1).values
            gen = torch.where(ts_mask, ts_vals, gen)
        logits = torch.randn(1, V, generator=g) * 3
        if trial % 3 == 0:
            logits[:, TB:] += 4.0          # make the timestamp mass win the probability rule sometimes
        logp = torch.log_softmax(logits, -1)
        want = R.WhisperRef.process_logits(logp, [gen[0].tolist()], gc)[0]
        K, P = 4, 3
        hb = HostBeam(gc, 1, K, P, np.array([[gc.decoder_start_token_id, 50259, 50360]], dtype=np.int32), V, track=False)
        # put the history's grammar state into beam 0 (the only live beam at a first step)
        hist = gen[0].tolist()
        is_ts = [t >= TB for t in hist]
        last_ts = max([t for t in hist if t >= TB][-1:], default=-1)
        hb.gram[0, 0] = [int(is_ts[-1]) if hist else 0, int(is_ts[-2]) if len(hist) >= 2 else 1, last_ts, 0]
        hb.cur = P + glen
        nxt, _ = hb.step(logits.repeat(K, 1).numpy())
        par = int(hb.ctrl[0])
        picked = sorted(int(t) for t in hb.hist[par, 0, :, hb.cur - 1])
        finite = torch.nonzero(~torch.isinf(want))[:, 0]
        top = sorted(int(t) for t in finite[torch.topk(want[finite], K).indices])
        assert picked == top, (trial, glen, picked, top)
