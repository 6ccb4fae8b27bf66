"""Oracle: Whisper encoder / decoder / greedy generate, plain PyTorch fp32 on CPU
(test infrastructure, see oracle/__init__.py).

Restates, for the configuration the reference's call reaches (greedy, timestamps on, no
prompt, no temperature fallback, `condition_on_prev_tokens` off) and the §8(f) "next" modes built
on it (no timestamps, beam search, token timestamps), the following functions of
transformers 5.5.0 (`$TF/`):
  * WhisperEncoder.forward / WhisperEncoderLayer      $TF/models/whisper/modeling_whisper.py:593-647, 380-414
  * WhisperAttention.forward                          $TF/models/whisper/modeling_whisper.py:284-357
  * WhisperDecoder.forward / WhisperDecoderLayer      $TF/models/whisper/modeling_whisper.py:691-796, 449-506
  * proj_out (tied to embed_tokens)                   $TF/models/whisper/modeling_whisper.py:964-1090
  * sinusoids                                         $TF/models/whisper/modeling_whisper.py:55-64
  * WhisperGenerationMixin.generate (seek loop)       $TF/models/whisper/generation_whisper.py:383-968
  * detect_language / _retrieve_init_tokens           $TF/models/whisper/generation_whisper.py:1610-1673, 1455-1608
  * _retrieve_segment                                 $TF/models/whisper/generation_whisper.py:1976-2073
  * GenerationMixin._sample (greedy)                  $TF/generation/utils.py:2658-2841
  * Suppress*/WhisperTimeStamp logits processors      $TF/generation/logits_process.py:1812-2043
  * GenerationMixin._beam_search                      $TF/generation/utils.py:3076-3400
  * _extract_token_timestamps (word timestamps)       $TF/models/whisper/generation_whisper.py:241-381
    with _median_filter :43-61, _dynamic_time_warping :64-112, the num_frames plumbing of
    _postprocess_outputs :1133-1151 and the per-segment slices of _retrieve_segment :1976-2073
The weights are read from an HF-layout ``state_dict`` (names as in SURVEY.md appendix B).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

# large-v3 special-token layout
EOS = 50257
SOT = 50258
LANG_FIRST, LANG_LAST = 50259, 50358
TRANSLATE, TRANSCRIBE = 50359, 50360
NO_TIMESTAMPS = 50364
TIMESTAMP_BEGIN = 50365
VOCAB = 51866

SUPPRESS_TOKENS = [
    1, 2, 7, 8, 9, 10, 14, 25, 26, 27, 28, 29, 31, 58, 59, 60, 61, 62, 63, 90, 91, 92, 93, 359, 503, 522, 542, 873,
    893, 902, 918, 922, 931, 1350, 1853, 1982, 2460, 2627, 3246, 3253, 3268, 3536, 3846, 3961, 4183, 4667, 6585, 6647,
    7273, 9061, 9383, 10428, 10929, 11938, 12033, 12331, 12562, 13793, 14157, 14635, 15265, 15618, 16553, 16604, 18362,
    18956, 20075, 21675, 22520, 26130, 26161, 26435, 28279, 29464, 31650, 32302, 32470, 36865, 42863, 47425, 49870,
    50254, 50258, 50359, 50360, 50361, 50362, 50363,
]
BEGIN_SUPPRESS_TOKENS = [220, 50257]


# generated-token steps whose raw logits WhisperRef.greedy keeps when recording
RECORD_LOGIT_STEPS = (0, 1, 2, 3, 7, 20, 100, 300, 444)


@dataclass
class WhisperDims:
    d_model: int = 1280
    heads: int = 20
    ffn: int = 5120
    enc_layers: int = 32
    dec_layers: int = 4
    n_mels: int = 128
    max_source_positions: int = 1500
    max_target_positions: int = 448
    vocab: int = VOCAB

    @property
    def head_dim(self) -> int:
        return self.d_model // self.heads


@dataclass
class GenConfig:
    """The generation_config.json fields of the openai/whisper-large-v3(-turbo) checkpoints that the
    path reads."""
    eos_token_id: int = EOS
    pad_token_id: int = EOS
    decoder_start_token_id: int = SOT
    no_timestamps_token_id: int = NO_TIMESTAMPS
    max_length: int = 448
    max_initial_timestamp_index: int = 50
    suppress_tokens: List[int] = field(default_factory=lambda: list(SUPPRESS_TOKENS))
    begin_suppress_tokens: List[int] = field(default_factory=lambda: list(BEGIN_SUPPRESS_TOKENS))
    lang_ids: List[int] = field(default_factory=lambda: list(range(LANG_FIRST, LANG_LAST + 1)))
    task_to_id: Dict[str, int] = field(default_factory=lambda: {"transcribe": TRANSCRIBE, "translate": TRANSLATE})


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    t = torch.arange(length).view(-1, 1) * inv.view(1, -1)
    return torch.cat([t.sin(), t.cos()], dim=1)


class WhisperRef:
    def __init__(self, dims: WhisperDims, state_dict: Dict[str, torch.Tensor]):
        self.dims = dims
        self.sd = {k: v.detach().to(torch.float32) for k, v in state_dict.items()}

    # -------------------------------------------------------------------------------- attention
    def _attn(self, pfx: str, x: torch.Tensor, kv_src: Optional[torch.Tensor] = None, causal: bool = False,
              cache: Optional[dict] = None, probs_out: Optional[list] = None) -> torch.Tensor:
        """Multi-head attention; q is scaled by head_dim**-0.5 after the bias ($TF ...:310)."""
        sd, H, dh = self.sd, self.dims.heads, self.dims.head_dim
        B, T, _ = x.shape
        q = F.linear(x, sd[pfx + "q_proj.weight"], sd[pfx + "q_proj.bias"]) * (dh ** -0.5)
        if cache is not None and "k" in cache and kv_src is not None:
            k, v = cache["k"], cache["v"]  # cross-attention K/V computed once
        else:
            src = x if kv_src is None else kv_src
            k = F.linear(src, sd[pfx + "k_proj.weight"])  # no bias on k
            v = F.linear(src, sd[pfx + "v_proj.weight"], sd[pfx + "v_proj.bias"])
            k = k.view(B, -1, H, dh).transpose(1, 2)
            v = v.view(B, -1, H, dh).transpose(1, 2)
            if cache is not None:
                if kv_src is None and "k" in cache:  # growing self-attention cache
                    k = torch.cat([cache["k"], k], dim=2)
                    v = torch.cat([cache["v"], v], dim=2)
                cache["k"], cache["v"] = k, v
        q = q.view(B, T, H, dh).transpose(1, 2)
        s = q @ k.transpose(-1, -2)
        if causal and T > 1:
            S = k.shape[2]
            mask = torch.ones(T, S, dtype=torch.bool, device=s.device).tril(diagonal=S - T)
            s = s.masked_fill(~mask, float("-inf"))
        p = torch.softmax(s, dim=-1)
        if probs_out is not None:
            probs_out.append(p)     # eager-attention `attn_weights` [B, H, T, S] (what output_attentions returns)
        o = (p @ v).transpose(1, 2).reshape(B, T, H * dh)
        return F.linear(o, sd[pfx + "out_proj.weight"], sd[pfx + "out_proj.bias"])

    def _ln(self, pfx: str, x: torch.Tensor) -> torch.Tensor:
        return F.layer_norm(x, (x.shape[-1],), self.sd[pfx + "weight"], self.sd[pfx + "bias"], 1e-5)

    def _mlp(self, pfx: str, x: torch.Tensor) -> torch.Tensor:
        sd = self.sd
        h = F.gelu(F.linear(x, sd[pfx + "fc1.weight"], sd[pfx + "fc1.bias"]))
        return F.linear(h, sd[pfx + "fc2.weight"], sd[pfx + "fc2.bias"])

    # -------------------------------------------------------------------------------- encoder
    def conv_stem(self, feats: torch.Tensor) -> torch.Tensor:
        sd = self.sd
        x = F.gelu(F.conv1d(feats, sd["model.encoder.conv1.weight"], sd["model.encoder.conv1.bias"], padding=1))
        x = F.gelu(F.conv1d(x, sd["model.encoder.conv2.weight"], sd["model.encoder.conv2.bias"], stride=2, padding=1))
        return x.permute(0, 2, 1) + sd["model.encoder.embed_positions.weight"]

    def encode(self, feats: torch.Tensor, return_layers: bool = False):
        """feats fp32 [B, n_mels, 3000] -> [B, 1500, d_model]; attention_mask is ignored by HF (:608-611)."""
        x = self.conv_stem(feats.to(torch.float32))
        layers = [x]
        for i in range(self.dims.enc_layers):
            p = f"model.encoder.layers.{i}."
            x = x + self._attn(p + "self_attn.", self._ln(p + "self_attn_layer_norm.", x))
            x = x + self._mlp(p, self._ln(p + "final_layer_norm.", x))
            layers.append(x)
        out = self._ln("model.encoder.layer_norm.", x)
        return (out, layers) if return_layers else out

    # -------------------------------------------------------------------------------- decoder
    def new_cache(self) -> List[dict]:
        return [{"self": {}, "cross": {}} for _ in range(self.dims.dec_layers)]

    def decode(self, tokens: torch.Tensor, enc_out: torch.Tensor, cache: Optional[List[dict]] = None,
               past_len: int = 0, cross: Optional[List[list]] = None) -> torch.Tensor:
        """tokens [B, T] (new positions only when a cache is given) -> fp32 logits [B, T, vocab].
        ``cross``: one list per decoder layer; receives that layer's cross-attention weights [B, H, T, S]."""
        sd = self.sd
        B, T = tokens.shape
        x = sd["model.decoder.embed_tokens.weight"][tokens] + sd["model.decoder.embed_positions.weight"][past_len:past_len + T]
        for i in range(self.dims.dec_layers):
            p = f"model.decoder.layers.{i}."
            c = cache[i] if cache is not None else {"self": None, "cross": None}
            x = x + self._attn(p + "self_attn.", self._ln(p + "self_attn_layer_norm.", x), causal=True, cache=c["self"])
            x = x + self._attn(p + "encoder_attn.", self._ln(p + "encoder_attn_layer_norm.", x), kv_src=enc_out,
                               cache=c["cross"], probs_out=None if cross is None else cross[i])
            x = x + self._mlp(p, self._ln(p + "final_layer_norm.", x))
        x = self._ln("model.decoder.layer_norm.", x)
        return F.linear(x, sd["model.decoder.embed_tokens.weight"])  # tied proj_out, no bias

    # -------------------------------------------------------------------------------- processors
    @staticmethod
    def process_logits(scores: torch.Tensor, generated: Sequence[Sequence[int]], gc: GenConfig,
                       apply_rule: bool = True, timestamps: bool = True) -> torch.Tensor:
        """SuppressTokens -> SuppressTokensAtBegin -> WhisperTimeStamp, on fp32 scores [B, V].
        ``generated[k]`` = tokens of row k after the decoder prompt (begin_index).  ``timestamps=False``: the
        processor list of generate(return_timestamps=False) — no WhisperTimeStampLogitsProcessor
        ($TF/models/whisper/generation_whisper.py, _retrieve_logit_processors)."""
        s = scores.clone().float()
        NEG = float("-inf")
        TB = gc.no_timestamps_token_id + 1
        s[:, gc.suppress_tokens] = NEG
        g = len(generated[0])
        if g == 0:
            s[:, gc.begin_suppress_tokens] = NEG
        if not timestamps:
            return s
        s[:, gc.no_timestamps_token_id] = NEG
        for k, seq in enumerate(generated):
            last_ts = len(seq) >= 1 and seq[-1] >= TB
            pen_ts = len(seq) < 2 or seq[-2] >= TB
            if last_ts:
                if pen_ts:
                    s[k, TB:] = NEG
                else:
                    s[k, :gc.eos_token_id] = NEG
            ts = [t for t in seq if t >= TB]
            if ts:
                ts_last = ts[-1] if (last_ts and not pen_ts) else ts[-1] + 1
                s[k, TB:ts_last] = NEG
        if g == 0:
            s[:, :TB] = NEG
            if gc.max_initial_timestamp_index is not None:
                s[:, TB + gc.max_initial_timestamp_index + 1:] = NEG
        if not apply_rule:
            return s
        logp = torch.log_softmax(s, dim=-1)
        for k in range(s.shape[0]):
            if logp[k, TB:].logsumexp(dim=-1) > logp[k, :TB].max():
                s[k, :TB] = NEG
        return s

    @staticmethod
    def _pre_rule_scores(scores, generated, gc):
        return WhisperRef.process_logits(scores, generated, gc, apply_rule=False)

    # -------------------------------------------------------------------------------- generation
    def detect_language(self, enc_out: torch.Tensor, gc: GenConfig) -> List[int]:
        B = enc_out.shape[0]
        logits = self.decode(torch.full((B, 1), gc.decoder_start_token_id, dtype=torch.long), enc_out)[:, -1]
        mask = torch.ones(logits.shape[-1], dtype=torch.bool)
        mask[gc.lang_ids] = False
        logits = logits.masked_fill(mask[None], float("-inf"))
        return logits.argmax(-1).tolist()

    def greedy(self, enc_out: torch.Tensor, prompt: torch.Tensor, gc: GenConfig, record: Optional[list] = None,
               timestamps: bool = True, cross: Optional[List[list]] = None):
        """GenerationMixin._sample, greedy: returns generated tokens per row, [B, <=max_length-len(prompt)];
        finished rows keep emitting pad (= eos).  ``record`` collects (raw fp32 logits, processed scores)."""
        B, P = prompt.shape
        cache = self.new_cache()
        tokens = prompt.clone()
        finished = torch.zeros(B, dtype=torch.bool)
        logits = self.decode(prompt, enc_out, cache, 0, cross=cross)[:, -1]
        while True:
            gen = [tokens[k, P:].tolist() for k in range(B)]
            scores = self.process_logits(logits, gen, gc, timestamps=timestamps)
            if record is not None and not timestamps:
                top2 = scores.topk(2, dim=-1).values
                record.append({"margin": (top2[:, 0] - top2[:, 1]), "rule_gap": torch.full((B,), float("inf")),
                               "logits": None})
            elif record is not None:
                # decisiveness of this step: top-1 margin of the processed scores and the gap of the
                # timestamp-probability rule; raw logits are kept only for the sampled steps
                TB = gc.no_timestamps_token_id + 1
                top2 = scores.topk(2, dim=-1).values
                raw = logits.float().clone()
                pre = self._pre_rule_scores(raw, gen, gc)
                rule_gap = (pre[:, TB:].logsumexp(-1) - pre[:, :TB].max(-1).values).abs()
                step = tokens.shape[1] - P
                record.append({"margin": (top2[:, 0] - top2[:, 1]), "rule_gap": rule_gap,
                               "logits": raw if step in RECORD_LOGIT_STEPS else None})
            nxt = scores.argmax(-1)
            nxt = torch.where(finished, torch.full_like(nxt, gc.pad_token_id), nxt)
            tokens = torch.cat([tokens, nxt[:, None]], dim=1)
            finished = finished | (nxt == gc.eos_token_id)
            if bool(finished.all()) or tokens.shape[1] >= gc.max_length:
                break
            logits = self.decode(nxt[:, None], enc_out, cache, tokens.shape[1] - 1, cross=cross)[:, -1]
        return tokens[:, P:]

    # -------------------------------------------------------------------------------- token timestamps
    @staticmethod
    def median_filter(x: torch.Tensor, width: int) -> torch.Tensor:
        """_median_filter ($TF/models/whisper/generation_whisper.py:43-61): reflect-padded median along the last axis."""
        pad = width // 2
        if x.shape[-1] <= pad:
            return x
        x = F.pad(x, (pad, pad, 0, 0), mode="reflect")
        return x.unfold(-1, width, 1).sort()[0][..., pad]

    @staticmethod
    def dtw_token_frames(matrix) -> List[int]:
        """_dynamic_time_warping (:64-112) on ``matrix`` (float64 numpy [tokens, frames], already negated) followed by
        the jump extraction of _extract_token_timestamps (:367-369): first frame of every token's run on the path.
        Cost table in float32, each cell the float64 sum rounded to float32; anti-diagonal sweep (cell (i, j) only
        needs diagonals i+j-1 and i+j-2), same values as HF's double loop."""
        import numpy as np
        n, m = matrix.shape
        cost = np.full((n + 1, m + 1), np.inf, dtype=np.float32)
        trace = -np.ones((n + 1, m + 1), dtype=np.int8)
        cost[0, 0] = 0
        for d in range(2, n + m + 1):
            i = np.arange(max(1, d - m), min(n, d - 1) + 1)
            j = d - i
            c0, c1, c2 = cost[i - 1, j - 1], cost[i - 1, j], cost[i, j - 1]
            t = np.where((c0 < c1) & (c0 < c2), 0, np.where((c1 < c0) & (c1 < c2), 1, 2)).astype(np.int8)
            c = np.where(t == 0, c0, np.where(t == 1, c1, c2))
            cost[i, j] = (matrix[i - 1, j - 1] + c.astype(np.float64)).astype(np.float32)
            trace[i, j] = t
        trace[0, :] = 2
        trace[:, 0] = 1
        i, j = n, m
        text, time = [], []
        while i > 0 or j > 0:
            text.append(i - 1)
            time.append(j - 1)
            t = trace[i, j]
            if t == 0:
                i, j = i - 1, j - 1
            elif t == 1:
                i -= 1
            else:
                j -= 1
        text, time = np.array(text)[::-1], np.array(time)[::-1]
        jumps = np.pad(np.diff(text), (1, 0), constant_values=1).astype(bool)
        return time[jumps].tolist()

    @staticmethod
    def beam_weights(cross: List[list], alignment_heads, beam_indices: torch.Tensor, n_prompt: int) -> torch.Tensor:
        """The beam-search branch of _extract_token_timestamps (:265-303): for every output position take the
        alignment heads' cross-attention row of the beam that produced it (the first index is repeated for the prompt
        positions of the prefill; -1 = past the end of the hypothesis -> row 0).  -> [B, heads, length, S]."""
        w = torch.stack([torch.cat(cross[l], dim=2)[:, h] for l, h in alignment_heads]).permute(1, 0, 2, 3)
        wl = int((beam_indices != -1).sum(-1).max())
        bi = beam_indices[:, :wl]
        if n_prompt > 1:
            bi = torch.cat([bi[:, :1].expand(-1, n_prompt - 1), bi], dim=-1)
        bi = bi.masked_fill(bi == -1, 0)
        return torch.stack([torch.index_select(w[:, :, i, :], 0, bi[:, i]) for i in range(bi.shape[1])], dim=2)

    def alignment_matrix(self, cross, alignment_heads, row: int, n_prompt: int, num_frames: int,
                         median_width: int = 7, pre_crop: Optional[int] = None) -> torch.Tensor:
        """The per-row branch of _extract_token_timestamps (:352-365): stack the alignment (layer, head) pairs, crop to
        num_frames // 2 encoder positions, drop the prompt positions, standardise over the token axis (population
        std), median-filter along frames, average the heads.  -> fp32 [tokens, frames]."""
        if torch.is_tensor(cross):            # already gathered weights [B, heads, T, S] (beam search)
            w = cross[row]
        else:
            w = torch.stack([torch.cat(cross[l], dim=2)[row, h] for l, h in alignment_heads])      # [heads, T, S]
        if pre_crop is not None:
            w = w[..., : pre_crop // 2]      # the whole-batch crop taken when every row has the same num_frames (:316-323)
        w = w[..., : num_frames // 2][:, n_prompt:, :]
        std = torch.std(w, dim=-2, keepdim=True, unbiased=False)
        mean = torch.mean(w, dim=-2, keepdim=True)
        w = self.median_filter((w - mean) / std, median_width)
        return w.mean(dim=0)

    def token_timestamps(self, cross: List[list], alignment_heads, n_rows: int, n_prompt: int, num_frames: Sequence[int],
                         time_precision: float = 0.02, median_width: int = 7) -> torch.Tensor:
        """_extract_token_timestamps for a greedy batch: fp32 [rows, prompt + generated]: 0 for the prompt positions,
        the DTW jump time of every generated token, the last one repeated for the final token (its cross-attention
        is never computed)."""
        T = cross.shape[2] if torch.is_tensor(cross) else sum(p.shape[2] for p in cross[0])
        out = torch.zeros(n_rows, T + 1, dtype=torch.float32)
        if T - n_prompt <= 0:
            return out
        # num_frames is a tensor here: when all rows agree the batch is cropped once up front AND per row again
        # (:322-323, :354) — the same thing for a non-negative count, a double crop from the end for a negative one
        pre = int(num_frames[0]) if len(set(int(f) for f in num_frames)) == 1 else None
        for b in range(n_rows):
            m = self.alignment_matrix(cross, alignment_heads, b, n_prompt, int(num_frames[b]), median_width, pre)
            frames = self.dtw_token_frames(-m.double().numpy())
            jt = torch.tensor([f * time_precision for f in frames], dtype=torch.float64)
            out[b] = torch.cat([torch.zeros(n_prompt, dtype=torch.float64), jt, jt[-1:]]).to(torch.float32)
        return out


    def beam_search(self, enc_out: torch.Tensor, prompt: torch.Tensor, gc: GenConfig, num_beams: int = 5,
                    length_penalty: float = 1.0, timestamps: bool = True, cross: Optional[List[list]] = None,
                    aux: Optional[dict] = None) -> torch.Tensor:
        """GenerationMixin._beam_search ($TF/generation/utils.py:3076-3400 and its helpers :2876-3075) for the
        Whisper decoder, early_stopping=False, do_sample=False, one returned sequence per row.

        Per step: fp32 log-softmax of the logits over the whole vocabulary, THEN the logits processors (they mask
        log-probabilities here, not logits), accumulated onto the running beam scores; the best 2*num_beams of the
        num_beams*vocab continuations are kept, those that hit a stopping criterion (eos, max_length) compete —
        divided by (generated length)**length_penalty — for the num_beams finished slots, the best num_beams
        others continue (the self-attention cache rows are re-gathered by their beam of origin).  The loop ends when
        the best running score over (current generated length)**length_penalty cannot beat the worst finished one.
        Returns the best finished sequence per row without the prompt, right-padded with pad_token_id.
        ``cross`` collects the cross-attention weights of every forward (rows = batch x beams in their order at that
        step); ``aux["beam_indices"]`` receives HF's `beam_indices` [B, generated]: for every generated position the
        row (batch-offset beam) whose forward produced it, -1 beyond the hypothesis (:2984-2995, :3377-3386)."""
        B, P = prompt.shape
        K, V, L = num_beams, self.dims.vocab, gc.max_length
        NEG = -1.0e9
        enc = enc_out.repeat_interleave(K, dim=0)
        running = torch.full((B, K, L), gc.pad_token_id, dtype=torch.long)
        running[:, :, :P] = prompt[:, None, :]
        finished_seq = running.clone()
        running_scores = torch.zeros(B, K)
        running_scores[:, 1:] = NEG
        beam_scores = torch.full((B, K), NEG)
        is_finished = torch.zeros(B, K, dtype=torch.bool)
        gen_len = torch.zeros(B, K, dtype=torch.long)          # generated length of the finished slots
        improvable = torch.ones(B, 1, dtype=torch.bool)
        top_mask = torch.cat([torch.ones(K, dtype=torch.bool), torch.zeros(K, dtype=torch.bool)])
        cache = self.new_cache()
        cur = P
        logits = self.decode(running[:, :, :P].reshape(B * K, P), enc, cache, 0, cross=cross)[:, -1]
        batch_off = (torch.arange(B) * K)[:, None]
        run_idx = torch.full((B, K, L - P), -1, dtype=torch.long)
        fin_idx = run_idx.clone()
        while True:
            flat = running[:, :, :cur].reshape(B * K, cur)
            logp = torch.log_softmax(logits.float(), dim=-1)
            logp = self.process_logits(logp, [flat[r, P:].tolist() for r in range(B * K)], gc, timestamps=timestamps)
            acc = (logp.view(B, K, V) + running_scores[:, :, None]).reshape(B, K * V)
            top_lp, top_idx = torch.topk(acc, k=2 * K)
            origin = top_idx // V
            cand = torch.gather(running, 1, origin[:, :, None].expand(-1, -1, L)).clone()
            cand[:, :, cur] = top_idx % V
            cand_idx = torch.gather(run_idx, 1, origin[:, :, None].expand(-1, -1, L - P)).clone()
            cand_idx[:, :, cur - P] = origin + batch_off
            hits = (cand[:, :, cur] == gc.eos_token_id) | (cur + 1 >= L)
            # beams that continue
            run_lp = top_lp + hits.float() * NEG
            nxt = torch.topk(run_lp, k=K)[1]
            running = torch.gather(cand, 1, nxt[:, :, None].expand(-1, -1, L))
            run_idx = torch.gather(cand_idx, 1, nxt[:, :, None].expand(-1, -1, L - P))
            running_scores = torch.gather(run_lp, 1, nxt)
            next_origin = torch.gather(origin, 1, nxt)
            # finished slots
            just = hits & top_mask[None, :]
            fin_lp = top_lp / float((cur + 1 - P) ** length_penalty)
            fin_lp = fin_lp + (~improvable).float() * NEG
            fin_lp = fin_lp + (~just).float() * NEG
            m_seq = torch.cat([finished_seq, cand], dim=1)
            m_lp = torch.cat([beam_scores, fin_lp], dim=1)
            m_fin = torch.cat([is_finished, just], dim=1)
            m_len = torch.cat([gen_len, torch.full((B, 2 * K), cur + 1 - P, dtype=torch.long)], dim=1)
            keep = torch.topk(m_lp, k=K)[1]
            finished_seq = torch.gather(m_seq, 1, keep[:, :, None].expand(-1, -1, L))
            fin_idx = torch.gather(torch.cat([fin_idx, cand_idx], dim=1), 1, keep[:, :, None].expand(-1, -1, L - P))
            beam_scores = torch.gather(m_lp, 1, keep)
            is_finished = torch.gather(m_fin, 1, keep)
            gen_len = torch.gather(m_len, 1, keep)
            # re-gather the self-attention cache rows of the surviving beams
            rows = (next_origin + batch_off).reshape(-1)
            for layer in cache:
                for kind in ("self", "cross"):
                    for name in ("k", "v"):
                        if name in layer[kind]:
                            layer[kind][name] = layer[kind][name].index_select(0, rows)
            cur += 1
            best_possible = running_scores[:, :1] / float((cur - P) ** length_penalty)
            worst_finished = torch.where(is_finished, beam_scores.min(dim=1, keepdim=True)[0], torch.tensor(NEG))
            improvable = improvable & (best_possible > worst_finished).any(dim=-1, keepdim=True)
            if not (bool(improvable.any()) and not bool(hits.all())):
                break
            logits = self.decode(running[:, :, cur - 1:cur].reshape(B * K, 1), enc, cache, cur - 1, cross=cross)[:, -1]
        n = int(gen_len[:, 0].max())
        if aux is not None:
            bi = fin_idx[:, 0, :]
            aux["beam_indices"] = bi[:, :int(((bi + 1).bool()).sum(dim=1).max())]
        return finished_seq[:, 0, P:P + n]

    @staticmethod
    def retrieve_segment(seq: List[int], seek_num_frames: int, TB: int):
        """_retrieve_segment: returns (list of token lists, seek advance in mel frames)."""
        is_ts = [t >= TB for t in seq]
        single_ending = is_ts[-2:] == [False, True]
        idxs = [i + 1 for i in range(len(seq) - 1) if is_ts[i] and is_ts[i + 1]]
        if idxs:
            slices = list(idxs)
            if single_ending:
                slices.append(len(seq))
            else:
                slices[-1] += 1
            segs, last = [], 0
            for cur in slices:
                segs.append(seq[last:cur])
                last = cur
            if single_ending:
                return segs, seek_num_frames
            return segs, (seq[last - 2] - TB) * 2
        return [list(seq)], seek_num_frames

    def generate(self, feats: torch.Tensor, task: str = "transcribe", gc: Optional[GenConfig] = None,
                 trace: Optional[dict] = None, return_timestamps: bool = True, num_beams: int = 1,
                 alignment_heads=None, num_frames: Optional[Sequence[int]] = None, median_width: int = 7,
                 token_ts: Optional[dict] = None) -> List[List[int]]:
        """WhisperGenerationMixin.generate for a batch of <=30 s windows (short-form), greedy, timestamps on.
        feats [B, n_mels, 3000] fp32.  Returns the generated ids per row (segments concatenated, no padding).
        feats wider than 3000 frames = HF's long-form mode ($TF/models/whisper/generation_whisper.py:654-658, the same
        seek loop over `total_input_frames`; the language is detected on the first 3000 frames :1649; timestamps are
        mandatory :1388-1394) — one row at a time, as the ASR pipeline calls it for an un-chunked long input.

        ``alignment_heads`` + ``num_frames`` (attention_mask.sum(-1) per row) + ``token_ts`` = the
        return_token_timestamps=True path (greedy only): token_ts["segments"][b] receives the per-token times of the
        returned ids as the `segments` carry them (seek offset added, _retrieve_segment :2036-2039, what the ASR
        pipeline reads) and token_ts["sequences"][b] the ones of the padded `token_timestamps` output (no offset,
        _pad_to_max_length :188-193)."""
        gc = gc or GenConfig()
        TB = gc.no_timestamps_token_id + 1
        B = feats.shape[0]
        feats = feats.to(torch.float32)
        if feats.shape[-1] > 3000 and not return_timestamps:
            raise ValueError("You have passed more than 3000 mel input features (> 30 seconds) which automatically enables "
                             "long-form generation which requires the model to predict timestamp tokens.")
        enc0 = self.encode(feats[..., :3000])
        langs = self.detect_language(enc0, gc)
        tail = [] if return_timestamps else [gc.no_timestamps_token_id]   # <|notimestamps|> joins the prompt
        init = torch.tensor([[gc.decoder_start_token_id, langs[b], gc.task_to_id[task]] + tail for b in range(B)],
                            dtype=torch.long)
        seek = [0] * B
        max_frames = [feats.shape[-1]] * B
        out: List[List[int]] = [[] for _ in range(B)]
        if trace is not None:
            trace.update({"langs": langs, "iterations": []})
        guard = 0
        while any(s < m for s, m in zip(seek, max_frames)):
            guard += 1
            if guard > 64:
                raise RuntimeError("seek loop does not advance")
            rows = [b for b in range(B) if seek[b] < max_frames[b]]
            nfr = {b: min(max_frames[b] - seek[b], 3000) for b in rows}
            seg = torch.zeros(len(rows), feats.shape[1], 3000)
            for i, b in enumerate(rows):
                seg[i, :, :nfr[b]] = feats[b, :, seek[b]:seek[b] + nfr[b]]
            enc = enc0 if (guard == 1) else self.encode(seg)
            rec = [] if trace is not None else None
            cross = [[] for _ in range(self.dims.dec_layers)] if alignment_heads is not None else None
            if num_beams > 1:
                aux = {}
                toks = self.beam_search(enc, init[rows], gc, num_beams=num_beams, timestamps=return_timestamps,
                                        cross=cross, aux=aux)
                if cross is not None:
                    cross = self.beam_weights(cross, alignment_heads, aux["beam_indices"], init.shape[1])
            else:
                toks = self.greedy(enc, init[rows], gc, record=rec, timestamps=return_timestamps, cross=cross)
            if cross is not None:
                # num_frames - seek per active row (_postprocess_outputs :1146-1151); negative values crop from the
                # end exactly as the tensor slice `[..., : n // 2]` does
                P = init.shape[1]
                tt = self.token_timestamps(cross, alignment_heads, len(rows), P,
                                           [int(num_frames[b]) - seek[b] for b in rows], median_width=median_width)
                if token_ts is not None:
                    token_ts.setdefault("segments", [[] for _ in range(B)])
                    token_ts.setdefault("sequences", [[] for _ in range(B)])
            if trace is not None:
                trace["iterations"].append({"rows": rows, "seek": [seek[b] for b in rows], "tokens": toks.clone(),
                                            "record": rec, "enc": enc})
            for i, b in enumerate(rows):
                s = toks[i].tolist()
                if s[-1] == gc.pad_token_id:  # strip padding, then the eos itself
                    n_pad = sum(1 for t in s if t == gc.pad_token_id)
                    if gc.pad_token_id == gc.eos_token_id:
                        n_pad -= 1
                    if n_pad:
                        s = s[:-n_pad]
                if s[-1] == gc.eos_token_id:
                    s = s[:-1]
                segs, adv = self.retrieve_segment(s, nfr[b], TB)
                if cross is not None and token_ts is not None:
                    n_kept = sum(len(sg) for sg in segs)
                    raw = tt[i, P:P + n_kept]
                    off = torch.tensor(seek[b], dtype=torch.float64) * 0.02 / 2       # time_offset (:800-802)
                    token_ts["sequences"][b].extend(raw.tolist())
                    token_ts["segments"][b].extend((raw + off).tolist())
                seek[b] += adv
                for sg in segs:
                    out[b].extend(sg)
        return out


def dims_from_hf_config(cfg) -> WhisperDims:
    return WhisperDims(d_model=cfg.d_model, heads=cfg.encoder_attention_heads, ffn=cfg.encoder_ffn_dim,
                       enc_layers=cfg.encoder_layers, dec_layers=cfg.decoder_layers, n_mels=cfg.num_mel_bins,
                       max_source_positions=cfg.max_source_positions, max_target_positions=cfg.max_target_positions,
                       vocab=cfg.vocab_size)
