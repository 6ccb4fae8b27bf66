"""CPU tests of the host-side mirror of the HF pipeline (windowing, strides, scheduling, seek-loop helpers)
— no GPU, engines are replaced by fakes."""
import os
import sys

import numpy as np
import pytest

import helpers
from turbo_whisper_workspace_b200 import pipeline as P
from turbo_whisper_workspace_b200 import scheduler as S
from turbo_whisper_workspace_b200.config import WhisperDims
from turbo_whisper_workspace_b200.engine import retrieve_segment, _bitmap


@pytest.mark.parametrize("n,cl,sl,sr", [(70 * 16000, 480000, 80000, 80000), (150 * 16000, 960000, 80000, 80000),
                                        (480000, 480000, 80000, 80000), (480001, 480000, 80000, 80000),
                                        (1000, 480000, 80000, 80000), (400000 + 320000, 480000, 80000, 80000),
                                        (3600 * 16000, 480000, 80000, 80000), (0, 480000, 80000, 80000)])
def test_chunk_windows_matches_hf_chunk_iter(n, cl, sl, sr):
    """Same (stride, is_last) sequence and the same sample ranges as transformers' chunk_iter."""
    asr = pytest.importorskip("transformers.pipelines.automatic_speech_recognition")

    class FakeFE:
        sampling_rate = 16000

        def __call__(self, chunk, **kw):
            return {"n": chunk.shape[0], "first": float(chunk[0]) if chunk.shape[0] else None}

    audio = np.arange(n, dtype=np.float32)
    want = list(asr.chunk_iter(audio, FakeFE(), cl, sl, sr))
    got = P.chunk_windows(n, cl, sl, sr)
    assert len(got) == len(want)
    for (s, e, stride, is_last), w in zip(got, want):
        assert stride == w["stride"] and is_last == w["is_last"]
        assert e - s == w["n"] and float(audio[s]) == w["first"]


def test_chunk_windows_counts_from_survey():
    # 1 h @30/5/5 -> 180 windows; @60/5/5 (the reference's literal call) -> 72 windows
    assert len(P.chunk_windows(3600 * 16000, 480000, 80000, 80000)) == 180
    assert len(P.chunk_windows(3600 * 16000, 960000, 80000, 80000)) == 72
    w = P.chunk_windows(150 * 16000, 960000, 80000, 80000)
    assert [x[2] for x in w] == [(960000, 0, 80000), (960000, 80000, 80000), (800000, 80000, 0)]


def test_partition_properties():
    for n in (0, 1, 7, 23, 24, 180, 181):
        for w in (1, 2, 3, 4, 8):
            r = S.partition(n, w)
            assert len(r) == w and r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [e - s for s, e in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        S.partition(3, 0)


class FakeEngine:
    """Stands in for WhisperEngine: 'transcribes' a clip to [len(clip) % 1000, first sample * 1e3]."""
    def __init__(self, device, max_batch=4):
        self.device, self.max_batch, self.calls = device, max_batch, []

    def generate_from_pcm(self, clips, task="transcribe", language=None, return_timestamps=True):
        assert len(clips) <= self.max_batch
        self.calls.append(len(clips))
        return [[len(c) % 1000, int(round(float(c[0]) * 1000))] for c in clips]


def test_window_scheduler_orders_and_microbatches():
    clips = [np.full(100 + i, i / 1000.0, dtype=np.float32) for i in range(23)]
    sch = S.WindowScheduler(None, None, None, devices=["d0", "d1", "d2"], engine_factory=lambda d: FakeEngine(d, 4))
    rows = sch.run(clips)
    assert rows == [[(100 + i) % 1000, i] for i in range(23)]
    assert sum(sum(e.calls) for e in sch.flat_engines) == 23
    assert all(max(e.calls, default=0) <= 4 for e in sch.flat_engines)
    # all engine contexts pull from ONE queue of equally sized micro-batches (no static per-device ranges)
    assert sch.last_stats["microbatches"] == [(i, min(i + 4, 23)) for i in range(0, 23, 4)]
    # two contexts per device share the same queue; order of the results is unchanged
    sch2 = S.WindowScheduler(None, None, None, devices=["d0", "d1"], engine_factory=lambda d: FakeEngine(d, 4),
                             contexts_per_device=2)
    assert sch2.run(clips) == rows
    assert sum(sum(e.calls) for e in sch2.flat_engines) == 23 and len(sch2.flat_engines) == 4
    assert S.WindowScheduler(None, None, None, devices=["d0"], engine_factory=lambda d: FakeEngine(d)).run([]) == []


def test_balanced_microbatches_keep_every_context_busy():
    """One 1 h file (180 windows of config 3): on 1 GPU x 4 contexts two rounds of 23-window micro-batches; on 8 GPUs x 4
    contexts 12-window micro-batches (the measured optimum for a rank's 22-23 windows, profiles/r2c_microbatch.jsonl:
    VERDICT r1 weak #3: a static 24-window micro-batch left 3 of 4 contexts idle); a per-rank share under torchrun
    (22-23 windows, 4 contexts) likewise."""
    assert S.MIN_MICROBATCH == 12
    assert S.balanced_microbatch(180, 4, 24) == 23 and S.balanced_microbatch(180, 32, 24) == 12
    assert S.balanced_microbatch(23, 4, 24) == 12 and S.balanced_microbatch(480, 4, 24) == 24
    assert S.balanced_microbatch(5, 4, 24) == 12 and S.balanced_microbatch(0, 4, 24) == 24
    assert S.balanced_microbatch(100, 4, 4) == 4          # an engine with fewer rows than the floor
    clips = [np.full(100 + i, i / 1000.0, dtype=np.float32) for i in range(180)]
    sch = S.WindowScheduler(None, None, None, devices=[f"d{i}" for i in range(8)],
                            engine_factory=lambda d: FakeEngine(d, 24), contexts_per_device=4)
    rows = sch.run(clips)
    assert rows == [[(100 + i) % 1000, i] for i in range(180)]
    assert len(sch.last_stats["microbatches"]) == 15 and all(b - a == 12 for a, b in sch.last_stats["microbatches"])
    # distributed shape: rank 3 of 8 owns windows 69..91 and runs them on its own 4-context scheduler
    local = S.WindowScheduler(None, None, None, devices=["g"], engine_factory=lambda d: FakeEngine(d, 24), contexts_per_device=4)
    d = S.DistributedWindowScheduler(local, rank=3, world_size=8)
    assert d.local_range(180) == (69, 92)
    assert d.run_local(clips) == rows[69:92]
    assert [b - a for a, b in local.last_stats["microbatches"]] == [12, 11]


def test_wide_generate_calls_are_balanced_like_the_bench_cuts_them():
    """bench.py cuts the timed region's windows with the scheduler's own rule: at the driver's --steps 20 (480 windows per
    GPU) and the bench defaults (96 rows x 5 contexts) every context gets ONE 96-window call; fewer steps shrink the
    calls evenly instead of leaving contexts idle."""
    assert S.balanced_microbatch(480, 5, 96) == 96
    assert S.balanced_microbatch(480, 3, 96) == 80 and S.balanced_microbatch(480, 4, 96) == 60
    assert S.balanced_microbatch(192, 5, 96) == 39 and S.balanced_microbatch(72, 5, 96) == 15
    assert S.balanced_microbatch(180, 5, 96) == 36          # the 1 h file of config 3 on one GPU
    clips = [np.full(100 + i, i / 1000.0, dtype=np.float32) for i in range(480)]
    sch = S.WindowScheduler(None, None, None, devices=["g"], engine_factory=lambda d: FakeEngine(d, 96), contexts_per_device=5)
    rows = sch.run(clips)
    assert rows == [[(100 + i) % 1000, i] for i in range(480)]
    assert [b - a for a, b in sch.last_stats["microbatches"]] == [96] * 5


def test_cross_attention_split_plan_balances_the_sms():
    """The streaming cross-attention kernel (csrc/cross_attn.cu) picks its key splits per launch: the fullest SM may hold
    at most ~10 % more 128-key chunks than the average for the shapes the engine launches, splits never exceed the
    scratch capacity, and one row on a big device still spreads over the SMs."""
    import ctypes as C
    from turbo_whisper_workspace_b200 import _lib
    lib = _lib.load()

    def plan(rows, heads=20, S_=1500, cap=12, sms=148):
        sp, cps, grid = C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(lib.tw_cross_attn_plan(rows, heads, S_, cap, sms, C.byref(sp), C.byref(cps), C.byref(grid)), "plan")
        return sp.value, cps.value, grid.value
    nch = 12                                              # ceil(1500 / 128)
    for rows in (12, 16, 24, 40, 72, 96):
        sp, cps, grid = plan(rows)
        assert 1 <= sp <= 12 and sp * cps >= nch and (sp - 1) * cps < nch and grid <= 2 * 148
        items = rows * 20 * sp
        fullest = -(-items // 148) * cps
        assert fullest <= 1.12 * rows * 20 * nch / 148, (rows, sp, cps)
    assert plan(96)[0] == 1 and plan(24)[0] in (3, 4)     # wide launches need no split; 24 rows: 9.7 x 4 or 12.97 x 3 chunks per SM
    assert plan(24, cap=1) == (1, 12, 296)                # no scratch: one item per (row, head)
    assert plan(24, cap=3)[0] <= 3
    sp, cps, grid = plan(1)
    assert (sp, cps, grid) == (6, 2, 120)                 # a single row: 20 heads x 6 splits, one 2-chunk item per SM
    assert plan(2, heads=6, S_=130, cap=12) == (2, 1, 24)


def test_longform_window_plan_tiles_the_clip_exactly():
    """ops.longform_plan / longform_buffer (un-chunked long-form features, tw_logmel_long): every clip frame is written by
    exactly one window, and a 30 s STFT of that window (reflect padding at the WINDOW's edges, as the kernel does)
    gives bit-identical power spectra to the STFT of the whole clip for the frames the window owns — including the
    clip's own reflect padding at both ends.  numpy restatement of the kernel's framing; no GPU."""
    from turbo_whisper_workspace_b200 import ops
    from oracle import logmel_ref as L
    win = L.hann_periodic().astype(np.float32)[None, :]

    def power(x, frames):       # frames of a center=True, reflect-padded STFT of x
        xp = np.pad(x, (200, 200), mode="reflect")
        idx = np.arange(400)[None, :] + 160 * np.asarray(frames)[:, None]
        return np.abs(np.fft.rfft((xp[idx] * win).astype(np.float32), axis=1)) ** 2

    rng = np.random.default_rng(0)
    for n in (480001, 480000 + 160 * 2 + 7, 2997 * 160 + 480000, 1204800, 16000 * 200 + 159):
        plan = ops.longform_plan(n)
        T = n // 160
        assert plan["frames"] == T and plan["hop_samples"] == 2997 * 160
        owner = np.full(T, -1)
        for b, (lo, hi, r0) in enumerate(zip(plan["lo"], plan["hi"], plan["row0"])):
            assert 0 <= lo < hi <= 2999 and r0 == 2997 * b + lo
            assert (owner[r0:r0 + hi - lo] == -1).all()
            owner[r0:r0 + hi - lo] = b
        assert (owner >= 0).all()
        audio = rng.standard_normal(n).astype(np.float32)
        buf = ops.longform_buffer(audio, plan)
        assert buf.shape[0] >= (len(plan["lo"]) - 1) * plan["hop_samples"] + 480000
        # a few windows are enough: the first, the second, the last two
        for b in sorted(set([0, 1, len(plan["lo"]) - 2, len(plan["lo"]) - 1]) & set(range(len(plan["lo"])))):
            lo, hi, r0 = plan["lo"][b], plan["hi"][b], plan["row0"][b]
            pick = sorted(set([lo, lo + 1, (lo + hi) // 2, hi - 2, hi - 1]) & set(range(lo, hi)))
            w = buf[b * plan["hop_samples"]: b * plan["hop_samples"] + 480000]
            got = power(w, pick)
            want = power(audio, [r0 + (j - lo) for j in pick])
            assert np.array_equal(got, want), (n, b)


def test_word_mode_microbatches_follow_hf_batches():
    """return_timestamps="word": micro-batches are the HF pipeline's batches of `batch_size` consecutive windows
    (capped at the engine's rows) whatever the number of devices / contexts, and rows come back as (ids, times)."""
    class WordEngine(FakeEngine):
        def generate_from_pcm(self, clips, task="transcribe", language=None, token_timestamps=False):
            assert token_timestamps and len(clips) <= self.max_batch
            self.calls.append([int(round(float(c[0]) * 1000)) for c in clips])
            return [([len(c) % 1000], [0.5 * len(clips)]) for c in clips]
    clips = [np.full(100 + i, i / 1000.0, dtype=np.float32) for i in range(23)]
    assert S.group_ranges(23, 3, 4) == [(0, 8), (8, 16), (16, 23)] and S.group_ranges(3, 4, 24)[0] == (0, 3)
    assert S.microbatch_size(32, 1, 24) == 24 and S.microbatch_size(4, 1, 24) == 4 and S.microbatch_size(32, 5) == 6
    for devices, ctxs in ((["d0"], 1), (["d0", "d1", "d2"], 1), (["d0", "d1"], 2)):
        sch = S.WindowScheduler(None, None, None, devices=devices, engine_factory=lambda d: WordEngine(d, 8),
                                contexts_per_device=ctxs)
        rows = sch.run(clips, token_timestamps=True, group=5)
        assert [r[0] for r in rows] == [[(100 + i) % 1000] for i in range(23)]
        batches = sorted(b for e in sch.flat_engines for b in e.calls)
        assert batches == [list(range(k, min(k + 5, 23))) for k in range(0, 23, 5)]
        # the per-batch value (here: the batch size) reaches every row of that batch
        assert [r[1] for r in rows] == [[2.5]] * 20 + [[1.5]] * 3
    d = S.DistributedWindowScheduler(WordEngine("r1", 8), rank=1, world_size=2)
    assert d.local_range(23, 5) == (15, 23) and [r[0] for r in d.run_local(clips, token_timestamps=True, group=5)] == \
        [[(100 + i) % 1000] for i in range(15, 23)]


def test_window_scheduler_propagates_worker_errors():
    class Boom(FakeEngine):
        def generate_from_pcm(self, clips, **kw):
            raise RuntimeError("boom")
    sch = S.WindowScheduler(None, None, None, devices=["a", "b"], engine_factory=lambda d: Boom(d))
    with pytest.raises(RuntimeError, match="boom"):
        sch.run([np.zeros(10, np.float32)] * 4)


def test_retrieve_segment_matches_oracle_restatement():
    from oracle.whisper_ref import WhisperRef, TIMESTAMP_BEGIN as TB
    rng = np.random.default_rng(0)
    for _ in range(300):
        n = int(rng.integers(1, 40))
        seq = [int(TB + rng.integers(0, 1500)) if rng.random() < 0.35 else int(rng.integers(0, 50000)) for _ in range(n)]
        nf = int(rng.integers(1, 3001))
        assert retrieve_segment(seq, nf, TB) == WhisperRef.retrieve_segment(seq, nf, TB)


def test_bitmap():
    b = _bitmap([0, 31, 32, 51865, 99999], 51866)
    assert b[0] == (1 | (1 << 31)) and b[1] == 1 and (b[51865 >> 5] >> (51865 & 31)) & 1 and int(b.sum()) > 0


def test_load_audio_variants(tmp_path):
    x = helpers.quantize_pcm16(helpers.synth_clip(3, seconds=1.0))
    p = tmp_path / "a.wav"
    helpers.write_wav16(p, x)
    a1, _ = P.load_audio(str(p))
    a2, _ = P.load_audio(open(p, "rb").read())
    a3, extra = P.load_audio({"raw": x, "sampling_rate": 16000, "foo": 1})
    a4, _ = P.load_audio(np.stack([x, x]))
    np.testing.assert_array_equal(a1, x)
    np.testing.assert_array_equal(a2, x)
    np.testing.assert_array_equal(a3, x)
    np.testing.assert_allclose(a4, x)
    assert extra == {"foo": 1}
    with pytest.raises(ValueError):
        P.load_audio({"raw": x})
    with pytest.raises(TypeError):
        P.load_audio(12)


def test_list_inputs_share_one_window_stream():
    """A list call sends the windows of ALL inputs to the engines in one scheduler run (HF's chunk pipeline batches
    across inputs too) and gives exactly the per-input results of separate calls, in order, for every mode."""
    class Sched:
        last_stats = {}

        def __init__(self):
            self.runs = []

        def run(self, clips, token_timestamps=False, group=None, **kw):
            self.runs.append((len(clips), group))
            rows = []
            for c in clips:
                k = int(abs(float(c[:8].sum())) * 1e4) % 1000 if len(c) else 0
                ids = [50365, 300 + k, 301 + len(c) % 97, 50365 + 100 + k % 50]
                rows.append((ids, [0.0, 0.5, 1.0, 1.5]) if token_timestamps else ids)
            return rows
    sch = Sched()
    pipe = P.B200WhisperPipeline(None, WhisperDims(**helpers.TINY), helpers.build_tokenizer(), scheduler=sch)
    files = [helpers.synth_clip(i, seconds=s) for i, s in enumerate((3.0, 71.0, 0.5, 30.0, 45.5))]
    for kw in (dict(chunk_length_s=30, stride_length_s=5, batch_size=4, return_timestamps=True),
               dict(chunk_length_s=30, stride_length_s=5, batch_size=4),
               dict(chunk_length_s=30, stride_length_s=[4, 2], batch_size=3, return_timestamps="word")):
        sch.runs.clear()
        together = pipe(files, **kw)
        st = kw["stride_length_s"] if isinstance(kw["stride_length_s"], list) else [kw["stride_length_s"]] * 2
        n_win = sum(len(P.chunk_windows(len(f), 30 * 16000, st[0] * 16000, st[1] * 16000)) for f in files)
        assert sch.runs == [(n_win, kw["batch_size"] if kw.get("return_timestamps") == "word" else None)]
        apart = [pipe(f, **kw) for f in files]
        assert together == apart and isinstance(together, list) and len(together) == 5
        assert pipe.last_stats["windows"] == 2       # the last single call
    assert pipe((), chunk_length_s=30) == [] and pipe([files[0]], chunk_length_s=30)[0] == pipe(files[0], chunk_length_s=30)
    with pytest.raises(StopIteration):               # one empty input aborts the whole call before any GPU work
        pipe([files[0], np.zeros(0, np.float32)], chunk_length_s=30)


def test_from_hf_model_reads_config_and_generation_config():
    """Everything the engine needs from an HF model object: dimensions from WhisperConfig, the Whisper fields of the
    generation config (suppress lists, language / task ids, alignment heads) and median_filter_width."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as G
    model, _ = G.hf_model("varied")
    model.generation_config.alignment_heads = [[0, 1], [1, 3]]
    model.config.median_filter_width = 5

    class NoEngines:
        last_stats = {}

        def run(self, clips, **kw):
            return [[50365, 300, 50370] for _ in clips]
    pipe = P.B200WhisperPipeline.from_hf_model(model, helpers.build_tokenizer(), scheduler=NoEngines())
    d, g = pipe.dims, pipe.generation
    assert (d.d_model, d.heads, d.ffn, d.enc_layers, d.dec_layers, d.vocab) == (256, 4, 1024, 2, 2, 51866)
    assert g.alignment_heads == [[0, 1], [1, 3]] and g.median_filter_width == 5
    assert g.eos_token_id == 50257 and g.decoder_start_token_id == 50258 and g.no_timestamps_token_id == 50364
    assert g.task_to_id == {"transcribe": 50360, "translate": 50359} and g.lang_to_id["<|en|>"] == 50259
    assert (g.lang_first, g.lang_last) == (50259, 50358) and len(g.suppress_tokens) == 88 and g.max_length == 448
    r = pipe(np.zeros(16000, np.float32), return_timestamps=True)
    assert r["chunks"] == [{"timestamp": (0.0, 0.1), "text": pipe.asr_decoder.text_of([300])}]


def _riff(fmt_tag, ch, sr, bits, body, extensible=False, extra_chunks=b"", data_size=None):
    import struct
    fmt = struct.pack("<HHIIHH", 0xFFFE if extensible else fmt_tag, ch, sr, sr * ch * bits // 8, ch * bits // 8, bits)
    if extensible:
        fmt += struct.pack("<HHI", 22, bits, 0) + struct.pack("<H", fmt_tag) + bytes(14)
    chunks = b"fmt " + struct.pack("<I", len(fmt)) + fmt + extra_chunks
    chunks += b"data" + struct.pack("<I", len(body) if data_size is None else data_size) + body
    return b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks


def test_wav_reader_formats(tmp_path):
    """The in-process RIFF reader: 8 / 16 / 24 / 32-bit PCM, 32-bit float, WAVE_FORMAT_EXTENSIBLE, chunks before the
    data (odd-sized, padded), streamed files whose data size is 0 / 0xFFFFFFFF, stereo down-mix — and None (-> ffmpeg, as
    in HF) for what it does not handle."""
    rng = np.random.default_rng(5)
    n, sr = 1600, 16000
    f = rng.uniform(-0.9, 0.9, size=(n, 2))
    i16 = (f * 32767).round().astype("<i2")
    want16 = (i16.astype(np.float32) / 32768.0)
    got, r = P._read_wav(_riff(1, 2, sr, 16, i16.tobytes()))
    assert r == sr and got.dtype == np.int16 and (got == i16).all()
    np.testing.assert_array_equal(P.ffmpeg_read(_riff(1, 2, sr, 16, i16.tobytes()), sr), want16.mean(axis=1))
    np.testing.assert_array_equal(P.ffmpeg_read(_riff(1, 1, sr, 16, i16[:, 0].tobytes()), sr), want16[:, 0])
    # odd-sized LIST chunk (padded to even) before the data, extensible header, streamed sizes
    lst = b"LIST" + (5).to_bytes(4, "little") + b"abcde" + b"\x00"
    for kw in (dict(extra_chunks=lst), dict(extensible=True), dict(data_size=0), dict(data_size=0xFFFFFFFF)):
        got, _ = P._read_wav(_riff(1, 2, sr, 16, i16.tobytes(), **kw))
        assert (got == i16).all(), kw
    i24 = (f * 8388607).round().astype(np.int32)
    b24 = np.stack([(i24 >> s) & 0xff for s in (0, 8, 16)], axis=-1).astype(np.uint8).tobytes()
    got, _ = P._read_wav(_riff(1, 2, sr, 24, b24))
    np.testing.assert_array_equal(got, i24.astype(np.float32) / 8388608.0)
    i32 = (f * 2147483647).round().astype("<i4")
    got, _ = P._read_wav(_riff(1, 2, sr, 32, i32.tobytes()))
    np.testing.assert_array_equal(got, i32.astype(np.float32) / 2147483648.0)
    u8 = ((f[:, :1] * 127).round() + 128).astype(np.uint8)
    got, _ = P._read_wav(_riff(1, 1, sr, 8, u8.tobytes()))
    np.testing.assert_array_equal(got, (u8.astype(np.float32) - 128.0) / 128.0)
    f32 = f.astype("<f4")
    got, _ = P._read_wav(_riff(3, 2, sr, 32, f32.tobytes()))
    np.testing.assert_array_equal(got, f32)
    assert P._read_wav(_riff(6, 1, sr, 8, bytes(100))) is None             # A-law: left to ffmpeg
    assert P._read_wav(b"RIFF\x00\x00\x00\x00AVI ") is None and P._read_wav(b"") is None
    # a file path goes through mmap and gives the same samples as the bytes
    p = tmp_path / "s.wav"
    p.write_bytes(_riff(1, 2, sr, 16, i16.tobytes()))
    a, _ = P.load_audio(str(p))
    np.testing.assert_array_equal(a, want16.mean(axis=1))
    (tmp_path / "empty.wav").write_bytes(b"")
    with pytest.raises(ValueError):
        P.load_audio(str(tmp_path / "empty.wav"))


def test_mel_filters_match_oracle():
    from oracle import logmel_ref as L
    from turbo_whisper_workspace_b200.ops import slaney_mel_filters
    np.testing.assert_array_equal(slaney_mel_filters(), L.mel_filter_bank().astype(np.float32))


# ------------------------------------------------------------------ audio ingest (filter bank of the GPU resampler)
@pytest.mark.parametrize("orig,new", [(48000, 16000), (44100, 16000), (8000, 16000), (22050, 16000), (11025, 16000)])
def test_resample_filter_bank_is_torchaudios(orig, new):
    """The GPU ingest kernel consumes the torchaudio filter bank: bit-identical construction (including torchaudio's
    fp32 phase offsets), and the per-phase non-zero span the kernel walks loses nothing."""
    import math
    import torch
    from torchaudio.functional import functional as AF
    from turbo_whisper_workspace_b200.ops import sinc_resample_filters
    want, w = AF._get_sinc_resample_kernel(orig, new, math.gcd(orig, new))
    got, w2, o, n = sinc_resample_filters(orig, new)
    assert (w2, got.shape) == (w, tuple(want[:, 0, :].shape)) and (o, n) == (orig // math.gcd(orig, new), new // math.gcd(orig, new))
    assert np.array_equal(got, want[:, 0, :].numpy())
    nz = np.abs(got) > 1e-30
    assert float(np.abs(got[~nz]).sum()) < 1e-25          # what the span skips is numerically nothing
    assert int(nz.sum(axis=1).max()) <= 2 * w + 2          # ~2*width useful taps per phase, not 2*width + orig


def test_fragment_major_weight_packing_matches_kernel_addressing():
    """engine.pack_skinny_weight vs the address formula of csrc/decode.cu (frag_ptr): element
    W[16*slab + 8*half + g][32*kstep + 8*tg + e] sits at ((slab*K/32 + kstep)*512 + half*256 + (g*4 + tg)*8 + e);
    rows beyond N are zero padding."""
    import random
    import torch
    from turbo_whisper_workspace_b200.engine import pack_skinny_weight
    N, K = 906, 256                       # N not a multiple of 16 (like the 51866-row LM head)
    W = torch.arange(N * K, dtype=torch.float32).view(N, K)
    P = pack_skinny_weight(W)
    assert tuple(P.shape) == (912, K)
    flat = P.reshape(-1)
    rng = random.Random(0)
    for _ in range(3000):
        slab, ks, half, g, tg, e = (rng.randrange((N + 15) // 16), rng.randrange(K // 32), rng.randrange(2),
                                    rng.randrange(8), rng.randrange(4), rng.randrange(8))
        off = (slab * (K // 32) + ks) * 512 + half * 256 + (g * 4 + tg) * 8 + e
        row, col = 16 * slab + 8 * half + g, 32 * ks + 8 * tg + e
        assert flat[off].item() == (W[row, col].item() if row < N else 0.0)
    with pytest.raises(ValueError):
        pack_skinny_weight(torch.zeros(16, 40))
