"""Chunk scheduler: shards the 30 s windows of a job across the GPUs of one box.

The reference has no multi-GPU code at all (single process, ``cuda:0``); HF only
batches the windows of one file on one device.  Windows are independent units (no conditioning across
windows in the chunked pipeline), so the path is embarrassingly parallel: weights are replicated, the
workers pull micro-batches of consecutive windows, and only token ids travel back — there is no collective on
the data path.

Two deployment shapes share the same partitioning:
  * :class:`WindowScheduler` — one process, one engine (+ one host thread) per local device;
  * :class:`DistributedWindowScheduler` — one process per GPU (torchrun); each rank runs its range on its
    own engine and the ordered token lists are gathered on the host with ``all_gather_object``.
"""
from __future__ import annotations

import threading
import time
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np


def partition(n_items: int, n_workers: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced ranges [(start, end)] — worker i gets items start..end-1.  Contiguity keeps
    each worker's micro-batches full and the gathered order trivial."""
    if n_workers < 1:
        raise ValueError("n_workers must be >= 1")
    base, rem = divmod(n_items, n_workers)
    out, s = [], 0
    for i in range(n_workers):
        e = s + base + (1 if i < rem else 0)
        out.append((s, e))
        s = e
    return out


def _extra(return_timestamps: bool, num_beams: int, token_timestamps: bool = False) -> Dict[str, Any]:
    """Keyword arguments beyond (task, language), passed only when they differ from the defaults so that simple
    engine stand-ins keep working."""
    kw: Dict[str, Any] = {}
    if not return_timestamps:
        kw["return_timestamps"] = False
    if num_beams > 1:
        kw["num_beams"] = int(num_beams)
    if token_timestamps:
        kw["token_timestamps"] = True
    return kw


def microbatch_size(max_batch: int, num_beams: int = 1, group: Optional[int] = None) -> int:
    """Windows per engine call: every window occupies num_beams decode rows; ``group`` (word timestamps) caps it at
    the HF pipeline's batch_size so that a micro-batch is exactly one of HF's generate batches."""
    mb = max(1, max_batch // max(1, num_beams))
    return mb if not group else max(1, min(mb, int(group)))


def group_ranges(n_items: int, n_workers: int, mb: int) -> List[Tuple[int, int]]:
    """:func:`partition` over whole micro-batches of ``mb`` consecutive items (micro-batch k = items k*mb ..), so the
    micro-batch boundaries do not depend on the number of workers."""
    n_groups = (n_items + mb - 1) // mb
    return [(min(n_items, a * mb), min(n_items, b * mb)) for a, b in partition(n_groups, n_workers)]


def run_in_microbatches(engine, clips: Sequence[np.ndarray], task: str, language: Optional[str],
                        return_timestamps: bool = True, num_beams: int = 1, token_timestamps: bool = False,
                        group: Optional[int] = None) -> List[Any]:
    rows: List[Any] = []
    mb = microbatch_size(engine.max_batch, num_beams, group)
    for i in range(0, len(clips), mb):
        rows.extend(engine.generate_from_pcm(clips[i:i + mb], task=task, language=language,
                                             **_extra(return_timestamps, num_beams, token_timestamps)))
    return rows


# Floor of the micro-batch size.  Measured on one B200 (profiles/r2c_microbatch.jsonl: a rank's share of a 1 h file on 8
# GPUs = 23 windows): 4 x 6 windows 0.563 s, 3 x 8 0.529 s, 2 x 12 0.520 s, 1 x 23 0.538 s — the decode loop costs ~890
# latency-bound steps per micro-batch whatever its size, so fewer, larger micro-batches win until contexts go idle.
MIN_MICROBATCH = 12


def balanced_microbatch(n_items: int, workers: int, max_mb: int, min_mb: int = MIN_MICROBATCH) -> int:
    """Micro-batch size for a job of ``n_items`` windows over ``workers`` engine contexts.  A micro-batch costs about
    the same whatever its size (the decode loop is latency-bound: ~2 x 445 steps per micro-batch), so the job should
    use as FEW micro-batches as possible while keeping every context busy in every round:
    rounds = ceil(n / (workers * max_mb)), size = ceil(n / (workers * rounds)) clamped to [min_mb, max_mb].
    One 1 h file (180 windows): 1 GPU x 4 contexts -> 8 micro-batches of 23; 8 GPUs x 4 contexts -> 15 of 12."""
    workers, max_mb = max(1, workers), max(1, max_mb)
    min_mb = max(1, min(min_mb, max_mb))
    if n_items <= 0:
        return max_mb
    rounds = -(-n_items // (workers * max_mb))
    return max(min_mb, min(max_mb, -(-n_items // (workers * rounds))))


class WorkQueue:
    """One shared queue over the windows of a job: ``take`` hands out the next micro-batch of ``size`` consecutive
    windows to whichever engine context asks first (no static per-device ranges)."""

    def __init__(self, n_items: int, size: int):
        self.n, self.pos, self.size = n_items, 0, max(1, size)
        self.lock = threading.Lock()
        self.handed: List[Tuple[int, int]] = []

    def take(self) -> Optional[Tuple[int, int]]:
        with self.lock:
            if self.pos >= self.n:
                return None
            a, b = self.pos, min(self.n, self.pos + self.size)
            self.pos = b
            self.handed.append((a, b))
            return a, b


class WindowScheduler:
    """One process; ``contexts_per_device`` engine contexts per local GPU, each driven by its own host thread and
    CUDA stream and sharing one copy of the weights.  All contexts of all devices pull micro-batches of consecutive
    windows (:func:`balanced_microbatch`) from ONE :class:`WorkQueue` (no static per-device ranges: a device that
    finishes early takes the next micro-batch), so that a context's latency-bound decode overlaps another context's tensor-bound encoder (and
    another decode) on the same GPU, and one file strong-scales over the devices."""

    def __init__(self, state_dict, dims, generation, devices: Sequence[Any] = ("cuda:0",), max_batch: int = 24,
                 engine_factory: Optional[Callable[..., Any]] = None, contexts_per_device: int = 1):
        self.devices = list(devices)
        self.contexts_per_device = max(1, int(contexts_per_device))
        self.engines: List[List[Any]] = []      # [device][context]
        for d in self.devices:
            ctxs = []
            for c in range(self.contexts_per_device):
                if engine_factory is not None:
                    ctxs.append(engine_factory(d))
                else:
                    from .engine import WhisperEngine
                    shared = ctxs[0].w if ctxs else None
                    ctxs.append(WhisperEngine(dims, state_dict if shared is None else None, device=d, gen=generation,
                                              max_batch=max_batch, shared_weights=shared,
                                              own_stream=self.contexts_per_device > 1))
            self.engines.append(ctxs)
        self.last_stats: Dict[str, Any] = {}

    @property
    def flat_engines(self) -> List[Any]:
        return [e for ctxs in self.engines for e in ctxs]

    @property
    def max_batch(self) -> int:
        return self.engines[0][0].max_batch

    def run(self, clips: Sequence[np.ndarray], task: str = "transcribe", language: Optional[str] = None,
            return_timestamps: bool = True, num_beams: int = 1, token_timestamps: bool = False,
            group: Optional[int] = None) -> List[Any]:
        """Token rows per window, in order.  ``token_timestamps``: every row is ``(ids, times)`` (word timestamps);
        ``group`` = the HF pipeline's batch_size: micro-batches then are HF's generate batches (the per-token times of
        a row depend on the longest row of its batch, $TF/models/whisper/generation_whisper.py:241-381)."""
        t0 = time.perf_counter()
        n = len(clips)
        engines = self.flat_engines
        if not engines or n == 0:
            self.last_stats = {"workers": len(self.devices), "contexts_per_device": self.contexts_per_device,
                               "microbatches": [], "seconds": time.perf_counter() - t0}
            return []
        cap = min(e.max_batch for e in engines)
        # word timestamps: exactly the HF pipeline's batches; otherwise as few, equally sized micro-batches as keep
        # every context busy
        mb = microbatch_size(cap, num_beams, group) if group else \
            balanced_microbatch(n, len(engines), microbatch_size(cap, num_beams), min(MIN_MICROBATCH, max(1, cap // max(1, num_beams))))
        queue = WorkQueue(n, mb)
        results: List[Optional[List[int]]] = [None] * n
        errors: List[BaseException] = []
        extra = _extra(return_timestamps, num_beams, token_timestamps)

        def work(engine):
            try:
                while not errors:
                    item = queue.take()
                    if item is None:
                        return
                    a, b = item
                    results[a:b] = engine.generate_from_pcm(clips[a:b], task=task, language=language, **extra)
            except BaseException as ex:  # surfaced on the calling thread
                errors.append(ex)

        # never more host threads than micro-batches the job can produce
        n_threads = min(len(engines), -(-n // mb))
        # contexts of different devices first, so a small job spreads over the GPUs before it stacks contexts
        order = [self.engines[d][c] for c in range(self.contexts_per_device) for d in range(len(self.devices))]
        threads = [threading.Thread(target=work, args=(eng,), daemon=True) for eng in order[:n_threads]]
        if len(threads) == 1:
            threads[0].run()
        else:
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        if errors:
            raise errors[0]
        self.last_stats = {"workers": len(self.devices), "contexts_per_device": self.contexts_per_device,
                           "microbatches": list(queue.handed), "seconds": time.perf_counter() - t0}
        return [r for r in results]

    def run_long(self, audio: np.ndarray, task: str = "transcribe", language: Optional[str] = None,
                 num_beams: int = 1) -> List[int]:
        """Token row of ONE un-chunked long-form clip (> 30 s): HF's long-form generate is a sequential seek loop over the
        clip's frames, so it runs on one engine context."""
        return self.engines[0][0].generate_long_from_pcm(audio, task=task, language=language, num_beams=num_beams)

    def close(self):
        self.engines = []


class DistributedWindowScheduler:
    """One rank per GPU.  ``run`` must be called by every rank with the same ``clips`` (or the same count):
    rank r processes ``partition(n, world)[r]`` on its local engine — or on its local :class:`WindowScheduler`
    (several engine contexts on the rank's GPU pulling from a queue over the rank's range) — and the token
    lists are gathered in rank order on the host (``all_gather_object``; no data-path collective)."""

    def __init__(self, engine, rank: int, world_size: int, group=None):
        self.engine, self.rank, self.world_size, self.group = engine, rank, world_size, group
        self.last_stats: Dict[str, Any] = {}

    @property
    def devices(self):
        return getattr(self.engine, "devices", [getattr(self.engine, "device", None)])

    @property
    def flat_engines(self) -> List[Any]:
        return getattr(self.engine, "flat_engines", [self.engine])

    def local_range(self, n: int, mb: Optional[int] = None) -> Tuple[int, int]:
        return (group_ranges(n, self.world_size, mb) if mb else partition(n, self.world_size))[self.rank]

    def run_local(self, clips: Sequence[np.ndarray], task: str = "transcribe", language: Optional[str] = None,
                  return_timestamps: bool = True, num_beams: int = 1, token_timestamps: bool = False,
                  group: Optional[int] = None) -> List[Any]:
        mb = microbatch_size(self.engine.max_batch, num_beams, group) if group else None
        s, e = self.local_range(len(clips), mb)
        if e <= s:
            return []
        if hasattr(self.engine, "run"):          # a local WindowScheduler (engine contexts of this rank's GPU)
            return self.engine.run(clips[s:e], task=task, language=language, return_timestamps=return_timestamps,
                                   num_beams=num_beams, token_timestamps=token_timestamps, group=group)
        return run_in_microbatches(self.engine, clips[s:e], task, language, return_timestamps, num_beams,
                                   token_timestamps, group)

    def gather(self, local_rows: List[List[int]]) -> List[List[int]]:
        if self.world_size == 1:
            return local_rows
        import torch.distributed as dist
        buckets: List[Any] = [None] * self.world_size
        dist.all_gather_object(buckets, local_rows, group=self.group)
        rows: List[List[int]] = []
        for b in buckets:
            rows.extend(b)
        return rows

    def run(self, clips: Sequence[np.ndarray], task: str = "transcribe", language: Optional[str] = None,
            return_timestamps: bool = True, num_beams: int = 1, token_timestamps: bool = False,
            group: Optional[int] = None) -> List[Any]:
        t0 = time.perf_counter()
        s, e = self.local_range(len(clips), microbatch_size(self.engine.max_batch, num_beams, group) if group else None)
        rows = self.gather(self.run_local(clips, task, language, return_timestamps, num_beams, token_timestamps, group))
        self.last_stats = {"workers": self.world_size, "rank": self.rank, "local_range": (s, e),
                           "seconds": time.perf_counter() - t0}
        return rows

    def run_long(self, audio: np.ndarray, task: str = "transcribe", language: Optional[str] = None,
                 num_beams: int = 1) -> List[int]:
        """Un-chunked long-form input does not shard (one sequential seek loop): every rank computes the same row."""
        if hasattr(self.engine, "run_long"):
            return self.engine.run_long(audio, task=task, language=language, num_beams=num_beams)
        return self.engine.generate_long_from_pcm(audio, task=task, language=language, num_beams=num_beams)

    def close(self):
        if hasattr(self.engine, "close"):
            self.engine.close()
