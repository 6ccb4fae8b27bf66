"""CPU test of the product's beam-search bookkeeping (turbo-whisper-workspace_b200/beam.py): driven with the fp32
oracle's decoder logits it must reproduce the oracle's beam search — which tests/test_oracle_golden.py pins token-exact
to transformers' generate(num_beams=5) — including the beam-of-origin indices used to re-gather the KV cache."""
import pytest
import torch

import helpers
from oracle import logmel_ref as L
from oracle import whisper_ref as R


def _drive(ref, enc, prompt, gc, K):
    from turbo_whisper_workspace_b200.beam import BeamConfig, BeamSearch
    cfg = BeamConfig(num_beams=K, vocab=ref.dims.vocab, max_length=gc.max_length, eos_id=gc.eos_token_id,
                     pad_id=gc.pad_token_id, no_timestamps_id=gc.no_timestamps_token_id, suppress=gc.suppress_tokens,
                     begin_suppress=gc.begin_suppress_tokens, max_initial_timestamp_index=gc.max_initial_timestamp_index)
    bs = BeamSearch(cfg, prompt, track_indices=True)
    B, P = prompt.shape
    encK = enc.repeat_interleave(K, dim=0)
    cache = ref.new_cache()
    logits = ref.decode(bs.rows(), encK, cache, 0)[:, -1]
    while True:
        origin = bs.step(logits)
        if bs.done:
            break
        for layer in cache:
            for kind in ("self", "cross"):
                for name in ("k", "v"):
                    if name in layer[kind]:
                        layer[kind][name] = layer[kind][name].index_select(0, origin)
        logits = ref.decode(bs.rows()[:, -1:], encK, cache, bs.cur - 1)[:, -1]
    return bs.result(), bs.beam_indices()


@pytest.mark.parametrize("variant", ["decisive", "varied"])
def test_beam_bookkeeping_matches_oracle(variant):
    clips = [helpers.synth_clip(0), helpers.synth_clip(2, seconds=11.3, kind="mod")]
    feats = torch.stack([torch.from_numpy(L.log_mel(c)) for c in clips]).to(torch.bfloat16).float()
    dims = R.WhisperDims(**helpers.TINY)
    ref = R.WhisperRef(dims, helpers.variant_state_dict(dims, variant))
    gc = R.GenConfig()
    enc = ref.encode(feats)
    langs = ref.detect_language(enc, gc)
    prompt = torch.tensor([[gc.decoder_start_token_id, langs[b], gc.task_to_id["transcribe"]] for b in range(2)])
    aux = {}
    want = ref.beam_search(enc, prompt, gc, num_beams=5, aux=aux)
    got, idx = _drive(ref, enc, prompt, gc, 5)
    assert got.shape == want.shape and torch.equal(got, want)
    # HF's `beam_indices` of the returned hypotheses (the oracle's are pinned through the token timestamps they select,
    # tests/test_oracle_golden.py::test_token_timestamps_under_beam_search_oracle_vs_hf_golden)
    assert torch.equal(idx, aux["beam_indices"])


def test_process_scores_matches_oracle_processors():
    """The batched processors against the oracle's per-row loop on random histories (timestamps, pairs, begin)."""
    from turbo_whisper_workspace_b200.beam import process_scores
    gc = R.GenConfig()
    g = torch.Generator().manual_seed(0)
    V, TB = 51866, gc.no_timestamps_token_id + 1
    for trial in range(40):
        glen = [0, 1, 2, 5, 9][trial % 5]
        rows = 6
        gen = torch.randint(0, 50257, (rows, glen), generator=g)
        if glen:
            ts_mask = torch.rand(rows, glen, generator=g) < 0.4
            ts_vals = torch.sort(torch.randint(TB, TB + 1500, (rows, glen), generator=g), dim=1).values
            gen = torch.where(ts_mask, ts_vals, gen)
        scores = torch.randn(rows, V, generator=g) * 3
        want = R.WhisperRef.process_logits(scores, [r.tolist() for r in gen], gc)
        got = process_scores(scores, gen, suppress=torch.tensor(gc.suppress_tokens), begin_suppress=torch.tensor(gc.begin_suppress_tokens),
                             no_timestamps_id=gc.no_timestamps_token_id, eos_id=gc.eos_token_id,
                             max_initial_timestamp_index=gc.max_initial_timestamp_index)
        assert torch.equal(torch.isinf(got), torch.isinf(want)), trial
        assert torch.equal(got[~torch.isinf(got)], want[~torch.isinf(want)]), trial
