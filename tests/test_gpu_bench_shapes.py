"""-m gpu parity tests of the INSTANTIATIONS bench.py and BASELINE.json configs 2 / 4 actually run.

The decode kernels are templated on NB = ceil(rows / 8) (csrc/decode.cu: skinny_gemm_kernel<NB,...>, lmhead_kernel<NB>);
tests/test_gpu_engine.py only reaches NB = 1 / 2.  Here, against `oracle/` (tiny dims: the CPU oracle; full dims: the
oracle's fp32 modules run on the GPU with TF32 off, fed the engine's encoder states):
  (a) tiny dims, 24 and 32 windows in ONE batch (NB = 3, 4): teacher-forced raw logits + every decisive pick, and the
      free-running seek loop under the margin rule;
  (b) large-v3-turbo dims, teacher-forced logits at B = 24 (bench.py's batch) + the free-running 24-row batch equal
      to the same windows decoded 6 at a time (the NB = 1 instantiation the other full-size tests pin to the oracle);
  (c) large-v3 dims (32 decoder layers, BASELINE.json config 4), B = 16, teacher-forced logits;
  (d) four engine contexts sharing one weight copy (bench.py's mode, `contexts_per_device=4`) give exactly the tokens
      of a single context, through the scheduler and through the pipeline callable;
  (e) WIDE generate calls (round 2: bench.py packs up to 96 windows into one call so that the decoder weights are
      streamed once per decode step): 40 / 72 / 96 rows on tiny dims vs the oracle (row chunks of 24 in every
      projection, `lmhead_kernel<5|6>`, two LM-head launches), and a full-size 96-row call == the same windows six at
      a time.
The file name sorts before the other GPU files so that `pytest -x` reaches it first.
"""
import numpy as np
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu

LOGIT_REL_TOL = 2e-2      # same stated bf16 tolerances as tests/test_gpu_engine.py
LOGIT_ABS_FLOOR = 2e-2
MARGIN_TOL = 0.30


def _clips(n, seed0=100):
    """n windows with varied content and lengths (full, short, modulated) so rows retire at different steps."""
    out = []
    for i in range(n):
        kind = "mod" if i % 3 else "noise"
        seconds = (30.0, 30.0, 11.3, 30.0, 23.7, 30.0, 4.2, 30.0)[i % 8]
        out.append(helpers.synth_clip(seed0 + i, seconds=seconds, kind=kind))
    return out


def oracle_first_iteration(ref, feats):
    """The first seek iteration of WhisperRef.generate (encode, language ids, one greedy pass with the per-step
    decisiveness record) — the later iterations repeat the same kernels on fewer rows and are covered at NB = 1 / 2 by
    tests/test_gpu_engine.py; one pass over 32 rows keeps the CPU oracle to well under a minute."""
    from oracle import whisper_ref as R
    gc = R.GenConfig()
    enc = ref.encode(feats)
    langs = ref.detect_language(enc, gc)
    init = torch.tensor([[gc.decoder_start_token_id, l, gc.task_to_id["transcribe"]] for l in langs], dtype=torch.long)
    rec = []
    toks = ref.greedy(enc, init, gc, record=rec)
    return {"langs": langs, "tokens": toks, "record": rec}


@pytest.fixture(scope="module")
def tiny(cuda_device):
    from oracle import logmel_ref as L
    from oracle import whisper_ref as R
    from turbo_whisper_workspace_b200.config import WhisperDims
    from turbo_whisper_workspace_b200.engine import WhisperEngine
    clips = _clips(96)
    feats = torch.stack([torch.from_numpy(L.log_mel(c)) for c in clips]).to(torch.bfloat16).float()
    rd = R.WhisperDims(**helpers.TINY)
    sd = helpers.variant_state_dict(rd, "decisive")
    ref = R.WhisperRef(rd, sd)
    eng = WhisperEngine(WhisperDims(**helpers.TINY), sd, device=cuda_device, max_batch=96)
    eng.enable_taps()
    # windows are independent units for the oracle as well: rows 0..B-1 of ONE 96-row oracle pass serve every B below
    it0 = oracle_first_iteration(ref, feats)
    return clips, feats, ref, eng, it0


@pytest.mark.parametrize("B", [24, 32, 40, 72, 96])
def test_teacher_forced_logits_tiny_nb3_nb4(tiny, B):
    """B = 24 -> NB = 3, B = 32 -> NB = 4 of every decode GEMM and of the LM head; B = 40 / 72 / 96 -> row chunks of 24 in
    the projections (one pass over the weights), `lmhead_kernel<5>`, `<6>` + `<3>`, two `<6>` launches: raw logits at the
    oracle's recorded steps and every decisive un-forced pick of the first seek iteration, language ids included."""
    from oracle import whisper_ref as R
    clips, feats, ref, eng, it0 = tiny
    toks = it0["tokens"][:B]
    n_gen = toks.shape[1]
    eng.load_pcm(clips[:B])
    eng.features(B)
    eng.encode(B)
    forced = torch.full((B, eng.max_len), -1, dtype=torch.int32)
    forced[:, 3:3 + n_gen] = toks.to(torch.int32)
    prompts = torch.tensor([[R.SOT, -1, R.TRANSCRIBE]] * B, dtype=torch.int32)
    logits_at = {}

    def on_step(s):
        g = s - 2
        if g in R.RECORD_LOGIT_STEPS and g < n_gen:
            logits_at[g] = eng.logits[:B].cpu().clone()
    eng.decode(B, prompts, n_steps=2 + n_gen, forced=forced, on_step=on_step)
    assert eng.tokens[:B, 1].cpu().tolist() == it0["langs"][:B], "language detection differs"
    picks = eng.choices[:B].cpu()
    assert len(logits_at) >= 4
    for g, lg in logits_at.items():
        want = it0["record"][g]["logits"][:B]
        err, bound = float((lg - want).abs().max()), LOGIT_REL_TOL * float(want.abs().max()) + LOGIT_ABS_FLOOR
        assert err < bound, f"B={B} step {g}: max |logit - oracle| = {err} > {bound}"
    checked = disagree = 0
    for g in range(n_gen):
        rec = it0["record"][g]
        dec = (rec["margin"][:B] > MARGIN_TOL) & (rec["rule_gap"][:B] > MARGIN_TOL)
        same = picks[:, 3 + g].long() == toks[:, g].long()
        checked += int(dec.sum())
        disagree += int((dec & ~same).sum())
    assert checked > 0.5 * n_gen * B, "tolerance leaves too few decisive steps to be a meaningful check"
    assert disagree == 0, f"B={B}: {disagree} of {checked} decisive picks differ from the oracle"


@pytest.mark.parametrize("B", [24, 32, 96])
def test_free_running_generate_tiny_nb3_nb4(tiny, B):
    """Free-running seek loop at 24 / 32 rows (row retirement changes NB between iterations): the first iteration of
    every row is identical to the oracle's up to the row's first NON-decisive oracle step (north-star rule); the
    later iterations must complete and be reproducible."""
    clips, feats, ref, eng, it0 = tiny
    etrace = {}
    eng.load_pcm(clips[:B])
    eng.features(B)
    got = eng.generate(B, trace=etrace)
    assert etrace["langs"][:B] == it0["langs"][:B]
    e0 = etrace["iterations"][0]
    assert e0["rows"] == list(range(B))
    n_gen = it0["tokens"].shape[1]
    agreed = total = full_rows = 0
    for b in range(B):
        want_row = it0["tokens"][b].tolist()
        got_row = e0["tokens"][b][3:3 + n_gen]
        total += n_gen
        ok = True
        for g, t in enumerate(want_row):
            if got_row[g] != t:
                rec = it0["record"][g]
                assert not (float(rec["margin"][b]) > MARGIN_TOL and float(rec["rule_gap"][b]) > MARGIN_TOL), \
                    f"B={B}: row {b} diverges at DECISIVE step {g}"
                ok = False
                break
            agreed += 1
        full_rows += ok
    print(f"\n[bench-shapes] B={B}: {agreed}/{total} first-iteration tokens identical to the oracle, "
          f"{full_rows}/{B} rows identical end to end")
    assert agreed >= 0.5 * total and full_rows >= B // 4
    eng.load_pcm(clips[:B])
    eng.features(B)
    assert eng.generate(B) == got and all(len(r) > 0 for r in got)


def test_four_contexts_equal_single_context(tiny, cuda_device):
    """bench.py's mode: 4 engine contexts (own streams / host threads / KV pools) over ONE weight copy.  Windows are
    independent units and every kernel reduces a row in a batch-independent order, so the rows must be IDENTICAL to the
    single-context ones — through the scheduler and through the pipeline callable."""
    from turbo_whisper_workspace_b200.config import WhisperDims
    from turbo_whisper_workspace_b200.pipeline import B200WhisperPipeline
    from turbo_whisper_workspace_b200.scheduler import WindowScheduler
    from oracle import whisper_ref as R
    clips, feats, ref, eng, it0 = tiny
    dims = WhisperDims(**helpers.TINY)
    sd = helpers.variant_state_dict(R.WhisperDims(**helpers.TINY), "decisive")
    single = eng.generate_from_pcm(clips[:24])
    sch = WindowScheduler(sd, dims, None, devices=[cuda_device], max_batch=6, contexts_per_device=4)
    assert len({id(e.w) for e in sch.flat_engines}) == 1, "the contexts must share one copy of the weights"
    for _ in range(2):
        assert sch.run(clips[:24]) == single
    pipe = B200WhisperPipeline(None, dims, helpers.build_tokenizer(), scheduler=sch)
    audio = np.concatenate([np.pad(c, (0, 480000 - len(c))) for c in clips[:12]])
    r4 = pipe(audio, chunk_length_s=30, stride_length_s=0, batch_size=24, return_timestamps=True)
    one = WindowScheduler(sd, dims, None, devices=[cuda_device], max_batch=12, contexts_per_device=1)
    r1 = B200WhisperPipeline(None, dims, helpers.build_tokenizer(), scheduler=one)(
        audio, chunk_length_s=30, stride_length_s=0, batch_size=24, return_timestamps=True)
    assert r4 == r1


# ---------------------------------------------------------------------------------------------------------------
# full sizes
# ---------------------------------------------------------------------------------------------------------------
def _teacher_forced_vs_fp32(eng, ref, clips, B, cuda_device, n_free_check=True):
    from oracle import whisper_ref as R
    B = eng.load_pcm(clips[:B])
    eng.features(B)
    enc = eng.encode(B).float().view(B, 1500, -1).clone()
    base = [R.SOT, 50259, R.TRANSCRIBE, 50365 + 7, 1000, 2000, 50365 + 40, 50365 + 40, 3000]
    rows = []
    for b in range(B):     # different histories per row
        r = list(base)
        r[1] = 50259 + (b % 100)
        r[4], r[5], r[8] = 1000 + 37 * b, 2000 + 11 * b, 3000 + 5 * b
        rows.append(r)
    toks = torch.tensor(rows)
    want = ref.decode(toks.to(cuda_device), enc)            # [B, 9, V] fp32, same encoder states
    forced = torch.full((B, eng.max_len), -1, dtype=torch.int32)
    forced[:, :toks.shape[1]] = toks.to(torch.int32)
    got = []
    eng.decode(B, toks[:, :3].to(torch.int32), n_steps=toks.shape[1], forced=forced,
               on_step=lambda s: got.append(eng.logits[:B].clone()))
    n_dec = 0
    for s in range(toks.shape[1]):
        w = want[:, s]
        err = float((got[s] - w).abs().max())
        bound = LOGIT_REL_TOL * float(w.abs().max()) + LOGIT_ABS_FLOOR
        assert err < bound, f"step {s}: {err} > {bound}"
        top2 = w.topk(2, dim=-1).values
        decisive = (top2[:, 0] - top2[:, 1]) > 2 * bound
        n_dec += int(decisive.sum())
        assert torch.equal(got[s].argmax(-1)[decisive], w.argmax(-1)[decisive])
    return n_dec


@pytest.fixture(scope="module")
def turbo24(cuda_device):
    from oracle import whisper_ref as R
    from turbo_whisper_workspace_b200.config import WhisperDims
    from turbo_whisper_workspace_b200.engine import WhisperEngine
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    rd = R.WhisperDims()
    sd = helpers.variant_state_dict(rd, "varied", seed=3)
    eng = WhisperEngine(WhisperDims.large_v3_turbo(), sd, device=cuda_device, max_batch=96)
    eng.enable_taps()
    ref = R.WhisperRef(rd, {k: v.to(cuda_device) for k, v in sd.items()})
    return eng, ref, _clips(96, seed0=300)


@pytest.mark.parametrize("B", [24, 96])
def test_decoder_logits_turbo_batch24_vs_fp32_reference(turbo24, cuda_device, B):
    """BASELINE.json config 2 / bench.py: large-v3-turbo dims, 24 rows (skinny_gemm_kernel<3,...>, lmhead_kernel<3>,
    cross-attention split combine at 24 rows) and the 96-row generate call the bench packs its windows into."""
    eng, ref, clips = turbo24
    _teacher_forced_vs_fp32(eng, ref, clips, B, cuda_device)


def test_generate_turbo_batch24_equals_batches_of_six(turbo24):
    """Free-running full-size generate at the bench's batch: the 24-row batch (NB = 3, later NB = 2 / 1 as rows retire)
    must give exactly the rows of the same windows decoded six at a time (NB = 1, the instantiation
    tests/test_gpu_fullsize.py pins to the fp32 reference)."""
    eng, ref, clips = turbo24
    six = []
    for i in range(0, 96, 6):
        six += eng.generate_from_pcm(clips[i:i + 6])
    assert eng.generate_from_pcm(clips[:24]) == six[:24]
    assert all(len(r) > 0 for r in six)
    # the bench's wide calls: 96 windows (4 row chunks per projection, 2 LM-head launches), 72 and 40 (lmhead<6>+<3>, <5>)
    assert eng.generate_from_pcm(clips) == six
    assert eng.generate_from_pcm(clips[:72]) == six[:72]
    assert eng.generate_from_pcm(clips[:40]) == six[:40]


def test_decoder_logits_large_v3_batch16_vs_fp32_reference(cuda_device):
    """BASELINE.json config 4: large-v3 dims (32 DECODER layers; $TF/models/whisper/modeling_whisper.py:417-506),
    16 rows (NB = 2), teacher-forced logits against the oracle's fp32 decoder on the engine's encoder states."""
    from oracle import whisper_ref as R
    from turbo_whisper_workspace_b200.config import WhisperDims
    from turbo_whisper_workspace_b200.engine import WhisperEngine
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dims = WhisperDims.large_v3()
    assert dims.dec_layers == 32
    rd = R.WhisperDims(dec_layers=32)
    sd = helpers.variant_state_dict(rd, "varied", seed=5)
    eng = WhisperEngine(dims, sd, device=cuda_device, max_batch=16)
    eng.enable_taps()
    ref = R.WhisperRef(rd, {k: v.to(cuda_device) for k, v in sd.items()})
    del sd
    clips = _clips(16, seed0=500)
    _teacher_forced_vs_fp32(eng, ref, clips, 16, cuda_device)
    # a free-running batch runs to completion and is deterministic
    a = eng.generate_from_pcm(clips[:16])
    b = eng.generate_from_pcm(clips[:16])
    assert a == b and all(len(r) > 0 for r in a)
    solo = eng.generate_from_pcm(clips[:2])
    assert solo == a[:2]
