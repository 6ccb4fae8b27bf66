"""Beam-search bookkeeping for the GPU decoder (``num_beams=5`` is what the reference's literal
pipeline call runs under transformers >= 4.53).

Restates ``GenerationMixin._beam_search`` ($TF/generation/utils.py:3076-3400; helpers :2876-3075) and the three Whisper
logits processors ($TF/generation/logits_process.py:1812-2043) as batched tensor operations that run on whatever
device the logits live on: the decode kernels produce the raw fp32 logits of all ``windows x beams`` rows, this module
turns them into the next tokens, the beam of origin of every surviving row (for the KV-cache re-gather) and the
finished hypotheses.  early_stopping=False, do_sample=False, one returned sequence per window — the pipeline's mode.
The arithmetic order follows the original (fp32 log-softmax over the whole vocabulary, THEN the processors; scores
accumulated in fp32; 2*num_beams candidates; finished slots ranked by score / generated_length**length_penalty).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import torch

NEG = -1.0e9


def process_scores(scores: torch.Tensor, generated: torch.Tensor, *, suppress: torch.Tensor, begin_suppress: torch.Tensor,
                   no_timestamps_id: int, eos_id: int, max_initial_timestamp_index: Optional[int],
                   timestamps: bool = True) -> torch.Tensor:
    """SuppressTokens -> SuppressTokensAtBegin -> WhisperTimeStamp on fp32 ``scores`` [R, V].
    ``generated`` [R, g] = the tokens of every row after the decoder prompt (same g for all rows).
    ``suppress`` / ``begin_suppress``: index tensors on the scores' device."""
    s = scores.clone()
    R, V = s.shape
    ninf = float("-inf")
    TB = no_timestamps_id + 1
    s[:, suppress] = ninf
    g = generated.shape[1]
    if g == 0:
        s[:, begin_suppress] = ninf
    if not timestamps:
        return s
    s[:, no_timestamps_id] = ninf
    ids = torch.arange(V, device=s.device)[None, :]
    if g >= 1:
        is_ts = generated >= TB
        last_ts = is_ts[:, -1]
        pen_ts = is_ts[:, -2] if g >= 2 else torch.ones(R, dtype=torch.bool, device=s.device)
        # after a timestamp: a closed pair forbids timestamps, an opening one forbids text below eos
        s = s.masked_fill((last_ts & pen_ts)[:, None] & (ids >= TB), ninf)
        s = s.masked_fill((last_ts & ~pen_ts)[:, None] & (ids < eos_id), ninf)
        # timestamps must not decrease: below the last one (or up to and including it unless it closes a pair)
        any_ts = is_ts.any(dim=1)
        pos = torch.arange(g, device=s.device)[None, :].expand(R, g)
        last_idx = torch.where(is_ts, pos, torch.full_like(pos, -1)).max(dim=1).values.clamp_min(0)
        last_val = generated.gather(1, last_idx[:, None])[:, 0]
        bound = torch.where(last_ts & ~pen_ts, last_val, last_val + 1)
        s = s.masked_fill(any_ts[:, None] & (ids >= TB) & (ids < bound[:, None]), ninf)
    else:
        s[:, :TB] = ninf
        if max_initial_timestamp_index is not None:
            s[:, TB + max_initial_timestamp_index + 1:] = ninf
    logp = torch.log_softmax(s, dim=-1)
    ts_heavier = logp[:, TB:].logsumexp(dim=-1) > logp[:, :TB].max(dim=-1).values
    s = s.masked_fill(ts_heavier[:, None] & (ids < TB), ninf)
    return s


@dataclass
class BeamConfig:
    num_beams: int
    vocab: int
    max_length: int
    eos_id: int
    pad_id: int
    no_timestamps_id: int
    suppress: Sequence[int]
    begin_suppress: Sequence[int]
    max_initial_timestamp_index: Optional[int] = 50
    length_penalty: float = 1.0
    timestamps: bool = True


class BeamSearch:
    """State of one beam search over ``n`` windows x ``num_beams`` rows (row = window * num_beams + beam)."""

    def __init__(self, cfg: BeamConfig, prompt: torch.Tensor, track_indices: bool = False):
        """prompt: int64 [n, P] on the device the logits will live on.  ``track_indices``: also keep HF's
        `beam_indices` ($TF/generation/utils.py:2984-2995, 3065-3070) — for every generated position the row whose
        forward produced it — which word timestamps under beam search gather the cross-attention rows by."""
        self.cfg = cfg
        dev = prompt.device
        n, P = prompt.shape
        K, L = cfg.num_beams, cfg.max_length
        self.n, self.P, self.cur = n, P, P
        self.running = torch.full((n, K, L), cfg.pad_id, dtype=torch.long, device=dev)
        self.running[:, :, :P] = prompt[:, None, :]
        self.finished_seq = self.running.clone()
        self.running_scores = torch.zeros(n, K, device=dev)
        self.running_scores[:, 1:] = NEG
        self.beam_scores = torch.full((n, K), NEG, device=dev)
        self.is_finished = torch.zeros(n, K, dtype=torch.bool, device=dev)
        self.gen_len = torch.zeros(n, K, dtype=torch.long, device=dev)
        self.improvable = torch.ones(n, 1, dtype=torch.bool, device=dev)
        self.top_mask = torch.cat([torch.ones(K, dtype=torch.bool), torch.zeros(K, dtype=torch.bool)]).to(dev)
        self.suppress = torch.as_tensor(list(cfg.suppress), dtype=torch.long, device=dev)
        self.begin_suppress = torch.as_tensor(list(cfg.begin_suppress), dtype=torch.long, device=dev)
        self.batch_off = (torch.arange(n, device=dev) * K)[:, None]
        self.done = False
        self.run_idx = self.fin_idx = None
        if track_indices:
            self.run_idx = torch.full((n, K, L - P), -1, dtype=torch.long, device=dev)
            self.fin_idx = self.run_idx.clone()

    def beam_indices(self) -> torch.Tensor:
        """HF's `beam_indices` of the returned hypotheses: int64 [n, longest generated], -1 beyond a hypothesis."""
        bi = self.fin_idx[:, 0, :]
        return bi[:, :int(((bi + 1).bool()).sum(dim=1).max())]

    def rows(self) -> torch.Tensor:
        """Current token matrix of the running rows, [n * num_beams, cur]."""
        return self.running[:, :, :self.cur].reshape(self.n * self.cfg.num_beams, self.cur)

    def step(self, logits: torch.Tensor) -> torch.Tensor:
        """logits: fp32 [n * num_beams, V] for the next position of every running row.  Advances the search by one
        token and returns ``origin`` int64 [n * num_beams]: the previous row whose history (KV cache) new row r
        continues.  ``self.done`` tells the caller to stop; ``self.rows()[:, -1]`` are the tokens to feed next."""
        c = self.cfg
        n, K, V, L, P, cur = self.n, c.num_beams, c.vocab, c.max_length, self.P, self.cur
        flat = self.rows()
        logp = torch.log_softmax(logits.float(), dim=-1)
        logp = process_scores(logp, flat[:, P:], suppress=self.suppress, begin_suppress=self.begin_suppress,
                              no_timestamps_id=c.no_timestamps_id, eos_id=c.eos_id,
                              max_initial_timestamp_index=c.max_initial_timestamp_index, timestamps=c.timestamps)
        acc = (logp.view(n, K, V) + self.running_scores[:, :, None]).reshape(n, K * V)
        top_lp, top_idx = torch.topk(acc, k=2 * K)
        origin = top_idx // V
        cand = torch.gather(self.running, 1, origin[:, :, None].expand(-1, -1, L)).clone()
        cand[:, :, cur] = top_idx % V
        hits = (cand[:, :, cur] == c.eos_id) | (cur + 1 >= L)
        run_lp = top_lp + hits.float() * NEG
        nxt = torch.topk(run_lp, k=K)[1]
        self.running = torch.gather(cand, 1, nxt[:, :, None].expand(-1, -1, L))
        self.running_scores = torch.gather(run_lp, 1, nxt)
        next_origin = torch.gather(origin, 1, nxt)
        just = hits & self.top_mask[None, :]
        fin_lp = top_lp / float((cur + 1 - P) ** c.length_penalty)
        fin_lp = fin_lp + (~self.improvable).float() * NEG
        fin_lp = fin_lp + (~just).float() * NEG
        m_seq = torch.cat([self.finished_seq, cand], dim=1)
        m_lp = torch.cat([self.beam_scores, fin_lp], dim=1)
        m_fin = torch.cat([self.is_finished, just], dim=1)
        m_len = torch.cat([self.gen_len, torch.full((n, 2 * K), cur + 1 - P, dtype=torch.long, device=logits.device)], dim=1)
        keep = torch.topk(m_lp, k=K)[1]
        self.finished_seq = torch.gather(m_seq, 1, keep[:, :, None].expand(-1, -1, L))
        if self.run_idx is not None:
            cand_idx = torch.gather(self.run_idx, 1, origin[:, :, None].expand(-1, -1, L - P)).clone()
            cand_idx[:, :, cur - P] = origin + self.batch_off
            self.run_idx = torch.gather(cand_idx, 1, nxt[:, :, None].expand(-1, -1, L - P))
            self.fin_idx = torch.gather(torch.cat([self.fin_idx, cand_idx], dim=1), 1, keep[:, :, None].expand(-1, -1, L - P))
        self.beam_scores = torch.gather(m_lp, 1, keep)
        self.is_finished = torch.gather(m_fin, 1, keep)
        self.gen_len = torch.gather(m_len, 1, keep)
        self.cur = cur + 1
        best_possible = self.running_scores[:, :1] / float((self.cur - P) ** c.length_penalty)
        worst_finished = torch.where(self.is_finished, self.beam_scores.min(dim=1, keepdim=True)[0],
                                     torch.full_like(self.beam_scores, NEG))
        self.improvable = self.improvable & (best_possible > worst_finished).any(dim=-1, keepdim=True)
        self.done = not (bool(self.improvable.any()) and not bool(hits.all()))
        return (next_origin + self.batch_off).reshape(-1)

    def result(self) -> torch.Tensor:
        """Best finished hypothesis per window without the prompt, right-padded with pad_id: [n, max generated]."""
        m = int(self.gen_len[:, 0].max())
        return self.finished_seq[:, 0, self.P:self.P + m]
