"""Chunk scheduler: shards the 30 s windows of a job across the GPUs of one box.

The reference has no multi-GPU code at all (single process, ``cuda:0``); HF only
batches the windows of one file on one device.  Windows are independent units (no conditioning across
windows in the chunked pipeline), so the path is embarrassingly parallel: weights are replicated, every
worker gets a contiguous range of windows, and only token ids travel back — there is no collective on
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


class WindowScheduler:
    """One process; ``contexts_per_device`` engine contexts per local GPU, each driven by its own host thread and
    CUDA stream and sharing one copy of the weights.  Work is handed out as micro-batches of ``max_batch``
    consecutive windows from one queue per device, so that a context's latency-bound decode overlaps another
    context's tensor-bound encoder (and another decode) on the same GPU."""

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

    def run(self, clips: Sequence[np.ndarray], task: str = "transcribe", language: Optional[str] = None,
            return_timestamps: bool = True, num_beams: int = 1, token_timestamps: bool = False,
            group: Optional[int] = None) -> List[Any]:
        """Token rows per window, in order.  ``token_timestamps``: every row is ``(ids, times)`` (word timestamps);
        ``group`` = the HF pipeline's batch_size: micro-batches then are HF's generate batches (the per-token times of
        a row depend on the longest row of its batch, $TF/models/whisper/generation_whisper.py:241-381)."""
        t0 = time.perf_counter()
        n = len(clips)
        mb0 = microbatch_size(self.engines[0][0].max_batch, num_beams, group) if (group and self.engines) else None
        ranges = group_ranges(n, len(self.devices), mb0) if mb0 else partition(n, len(self.devices))
        results: List[Optional[List[int]]] = [None] * n
        errors: List[BaseException] = []
        lock = threading.Lock()
        threads = []
        for di, (s, e) in enumerate(ranges):
            ctxs = self.engines[di]
            mb = mb0 or microbatch_size(ctxs[0].max_batch, num_beams)
            queue = [(i, min(i + mb, e)) for i in range(s, e, mb)]   # micro-batches of this device, in order

            def work(engine, queue=queue):
                try:
                    while True:
                        with lock:
                            if not queue or errors:
                                return
                            a, b = queue.pop(0)
                        rows = engine.generate_from_pcm(clips[a:b], task=task, language=language,
                                                        **_extra(return_timestamps, num_beams, token_timestamps))
                        results[a:b] = rows
                except BaseException as ex:  # surfaced on the calling thread
                    with lock:
                        errors.append(ex)

            for eng in ctxs:
                threads.append(threading.Thread(target=work, args=(eng,), daemon=True))
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
                           "ranges": ranges, "seconds": time.perf_counter() - t0}
        return [r for r in results]

    def close(self):
        self.engines = []


class DistributedWindowScheduler:
    """One rank per GPU.  ``run`` must be called by every rank with the same ``clips`` (or the same count):
    rank r processes ``partition(n, world)[r]`` on its engine; the token lists are gathered in rank order."""

    def __init__(self, engine, rank: int, world_size: int, group=None):
        self.engine, self.rank, self.world_size, self.group = engine, rank, world_size, group
        self.last_stats: Dict[str, Any] = {}

    def local_range(self, n: int, mb: Optional[int] = None) -> Tuple[int, int]:
        return (group_ranges(n, self.world_size, mb) if mb else partition(n, self.world_size))[self.rank]

    def run_local(self, clips: Sequence[np.ndarray], task: str = "transcribe", language: Optional[str] = None,
                  return_timestamps: bool = True, num_beams: int = 1, token_timestamps: bool = False,
                  group: Optional[int] = None) -> List[Any]:
        mb = microbatch_size(self.engine.max_batch, num_beams, group) if group else None
        s, e = self.local_range(len(clips), mb)
        return run_in_microbatches(self.engine, clips[s:e], task, language, return_timestamps, num_beams,
                                   token_timestamps, group) if e > s else []

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
        return self.gather(self.run_local(clips, task, language, return_timestamps, num_beams, token_timestamps, group))
