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


def run_in_microbatches(engine, clips: Sequence[np.ndarray], task: str, language: Optional[str]) -> List[List[int]]:
    rows: List[List[int]] = []
    mb = engine.max_batch
    for i in range(0, len(clips), mb):
        rows.extend(engine.generate_from_pcm(clips[i:i + mb], task=task, language=language))
    return rows


class WindowScheduler:
    def __init__(self, state_dict, dims, generation, devices: Sequence[Any] = ("cuda:0",), max_batch: int = 24,
                 engine_factory: Optional[Callable[..., Any]] = None):
        if engine_factory is None:
            from .engine import WhisperEngine

            def engine_factory(device):
                return WhisperEngine(dims, state_dict, device=device, gen=generation, max_batch=max_batch)
        self.devices = list(devices)
        self.engines = [engine_factory(d) for d in self.devices]
        self.last_stats: Dict[str, Any] = {}

    def run(self, clips: Sequence[np.ndarray], task: str = "transcribe", language: Optional[str] = None
            ) -> List[List[int]]:
        t0 = time.perf_counter()
        n = len(clips)
        ranges = partition(n, len(self.engines))
        results: List[Optional[List[List[int]]]] = [None] * len(self.engines)
        errors: List[Optional[BaseException]] = [None] * len(self.engines)

        def work(i):
            s, e = ranges[i]
            try:
                results[i] = run_in_microbatches(self.engines[i], clips[s:e], task, language) if e > s else []
            except BaseException as ex:  # surfaced on the calling thread
                errors[i] = ex

        if len(self.engines) == 1:
            work(0)
        else:
            threads = [threading.Thread(target=work, args=(i,), daemon=True) for i in range(len(self.engines))]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        for ex in errors:
            if ex is not None:
                raise ex
        rows: List[List[int]] = []
        for r in results:
            rows.extend(r or [])
        self.last_stats = {"workers": len(self.engines), "ranges": ranges, "seconds": time.perf_counter() - t0}
        return rows

    def close(self):
        self.engines = []


class DistributedWindowScheduler:
    """One rank per GPU.  ``run`` must be called by every rank with the same ``clips`` (or the same count):
    rank r processes ``partition(n, world)[r]`` on its engine; the token lists are gathered in rank order."""

    def __init__(self, engine, rank: int, world_size: int, group=None):
        self.engine, self.rank, self.world_size, self.group = engine, rank, world_size, group
        self.last_stats: Dict[str, Any] = {}

    def local_range(self, n: int) -> Tuple[int, int]:
        return partition(n, self.world_size)[self.rank]

    def run_local(self, clips: Sequence[np.ndarray], task: str = "transcribe", language: Optional[str] = None
                  ) -> List[List[int]]:
        s, e = self.local_range(len(clips))
        return run_in_microbatches(self.engine, clips[s:e], task, language) if e > s else []

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

    def run(self, clips: Sequence[np.ndarray], task: str = "transcribe", language: Optional[str] = None
            ) -> List[List[int]]:
        return self.gather(self.run_local(clips, task, language))
