"""Shared test helpers: seeded random Whisper weights in the HF state_dict layout and synthetic audio.
No dependency on transformers (the GPU box tests must not need /root/reference or HF model classes)."""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

TINY = dict(d_model=256, heads=4, ffn=1024, enc_layers=2, dec_layers=2)


def sinusoids(length: int, channels: int) -> torch.Tensor:
    inc = math.log(10000.0) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    t = torch.arange(length).view(-1, 1) * inv.view(1, -1)
    return torch.cat([t.sin(), t.cos()], dim=1)


def random_state_dict(dims, seed: int = 0, scheme: str = "unit") -> Dict[str, torch.Tensor]:
    """HF-layout fp32 state_dict.  scheme="hf": N(0, 0.02) like WhisperPreTrainedModel._init_weights;
    scheme="unit": fan-in scaled weights and non-trivial biases / LayerNorm parameters, so attention
    is peaky and every bias / gain path is exercised (a uniform softmax would hide indexing bugs)."""
    g = torch.Generator().manual_seed(seed)
    D, F, V = dims.d_model, dims.ffn, dims.vocab

    def lin(out_f, in_f, gain=1.0):
        std = 0.02 if scheme == "hf" else gain / math.sqrt(in_f)
        return torch.randn(out_f, in_f, generator=g) * std

    def vec(n, std=0.1):
        return torch.zeros(n) if scheme == "hf" else torch.randn(n, generator=g) * std

    def ln(prefix, sd):
        sd[prefix + ".weight"] = torch.ones(D) if scheme == "hf" else 1.0 + 0.1 * torch.randn(D, generator=g)
        sd[prefix + ".bias"] = vec(D)

    def attn(prefix, sd):
        sd[prefix + "q_proj.weight"] = lin(D, D, 2.0)
        sd[prefix + "q_proj.bias"] = vec(D)
        sd[prefix + "k_proj.weight"] = lin(D, D, 2.0)
        sd[prefix + "v_proj.weight"] = lin(D, D)
        sd[prefix + "v_proj.bias"] = vec(D)
        sd[prefix + "out_proj.weight"] = lin(D, D, 0.5)
        sd[prefix + "out_proj.bias"] = vec(D)

    def mlp(prefix, sd):
        sd[prefix + "fc1.weight"] = lin(F, D)
        sd[prefix + "fc1.bias"] = vec(F)
        sd[prefix + "fc2.weight"] = lin(D, F, 0.5)
        sd[prefix + "fc2.bias"] = vec(D)

    sd: Dict[str, torch.Tensor] = {}
    e = "model.encoder."
    s1 = 0.02 if scheme == "hf" else 1.0 / math.sqrt(3 * dims.n_mels)
    s2 = 0.02 if scheme == "hf" else 1.0 / math.sqrt(3 * D)
    sd[e + "conv1.weight"] = torch.randn(D, dims.n_mels, 3, generator=g) * s1
    sd[e + "conv1.bias"] = vec(D)
    sd[e + "conv2.weight"] = torch.randn(D, D, 3, generator=g) * s2
    sd[e + "conv2.bias"] = vec(D)
    sd[e + "embed_positions.weight"] = sinusoids(dims.max_source_positions, D)
    for i in range(dims.enc_layers):
        p = f"{e}layers.{i}."
        attn(p + "self_attn.", sd)
        ln(p + "self_attn_layer_norm", sd)
        mlp(p, sd)
        ln(p + "final_layer_norm", sd)
    ln(e + "layer_norm", sd)
    d = "model.decoder."
    sd[d + "embed_tokens.weight"] = torch.randn(V, D, generator=g) * (0.02 if scheme == "hf" else 0.5)
    sd[d + "embed_positions.weight"] = torch.randn(dims.max_target_positions, D, generator=g) * (0.02 if scheme == "hf" else 0.3)
    for i in range(dims.dec_layers):
        p = f"{d}layers.{i}."
        attn(p + "self_attn.", sd)
        ln(p + "self_attn_layer_norm", sd)
        attn(p + "encoder_attn.", sd)
        ln(p + "encoder_attn_layer_norm", sd)
        mlp(p, sd)
        ln(p + "final_layer_norm", sd)
    ln(d + "layer_norm", sd)
    sd["proj_out.weight"] = sd[d + "embed_tokens.weight"]
    return sd


def synth_clip(seed: int, seconds: float = 30.0, kind: str = "noise") -> np.ndarray:
    """Synthetic 16 kHz fp32 audio: 0.1*N(0,1) (BASELINE.json config 1/2) or amplitude-modulated noise +
    a sine sweep so that the log-mel clamp and the timestamp grammar see structure."""
    n = int(round(seconds * 16000))
    rng = np.random.default_rng(seed)
    x = 0.1 * rng.standard_normal(n)
    if kind == "mod":
        t = np.arange(n) / 16000.0
        env = 0.5 * (1 + np.sin(2 * np.pi * 0.3 * t + seed))
        x = x * env + 0.2 * np.sin(2 * np.pi * (150 + 40 * t) * t) * (env > 0.5)
    return x.astype(np.float32)


def variant_state_dict(dims, variant: str, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Named weight variants used by the parity tests (chosen by running the oracle):
    "decisive": large top-1 margins everywhere, a two-iteration seek loop in which one row retires;
    "varied":   diverse tokens, ~13 timestamps per window, eos before max_length, two seek iterations
                with different per-row seeks; margins down to 1e-2 (margin-aware comparison needed)."""
    emb_scale, ts_boost = {"decisive": (0.2, 0.6), "varied": (0.1, 0.3)}[variant]
    sd = random_state_dict(dims, seed, "unit")
    E = sd["model.decoder.embed_tokens.weight"]
    E.mul_(emb_scale / 0.5)
    g = torch.Generator().manual_seed(seed + 1)
    d = torch.randn(dims.d_model, generator=g)
    d /= d.norm()
    E[50365:] += ts_boost * d
    sd["model.decoder.layer_norm.bias"] = sd["model.decoder.layer_norm.bias"] + 2.0 * d
    return sd


def build_tokenizer():
    """Synthetic WhisperTokenizer with the exact large-v3 id layout (51866 ids):
    ids 0..255 raw bytes, filler words up to 50256, <|endoftext|> 50257, the 107 specials, 1501 timestamps.
    No tokenizer files exist offline, so tests construct it from the transformers library class."""
    from transformers import WhisperTokenizer
    from transformers.models.whisper.tokenization_whisper import LANGUAGES

    bs = list(range(33, 127)) + list(range(161, 173)) + list(range(174, 256))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    b2u = dict(zip(bs, map(chr, cs)))
    vocab = {}
    for b in range(256):
        vocab[b2u[b]] = len(vocab)
    i = 0
    while len(vocab) < 50257:
        s = "".join(b2u[ord(c)] for c in f"w{i}")
        vocab.setdefault(s, len(vocab))
        i += 1
    tok = WhisperTokenizer(vocab=vocab, merges=[])
    tok.add_special_tokens({"additional_special_tokens": ["<|startoftranscript|>"] + [f"<|{l}|>" for l in LANGUAGES] +
                            ["<|translate|>", "<|transcribe|>", "<|startoflm|>", "<|startofprev|>", "<|nospeech|>",
                             "<|notimestamps|>"]})
    tok.add_tokens(["<|%.2f|>" % (k * 0.02) for k in range(1501)])
    tok.pad_token = "<|endoftext|>"
    assert len(tok) == 51866 and tok.convert_tokens_to_ids("<|notimestamps|>") == 50364
    return tok


def write_wav16(path, pcm: np.ndarray, sr: int = 16000) -> None:
    import wave
    x = np.clip(np.asarray(pcm, dtype=np.float64) * 32768.0, -32768, 32767).astype("<i2")
    with wave.open(str(path), "wb") as wf:
        wf.setnchannels(1)
        wf.setsampwidth(2)
        wf.setframerate(sr)
        wf.writeframes(x.tobytes())


def quantize_pcm16(pcm: np.ndarray) -> np.ndarray:
    """What a PCM16 WAV round trip does to the samples (so array inputs and file inputs agree)."""
    return (np.clip(np.asarray(pcm, dtype=np.float64) * 32768.0, -32768, 32767).astype("<i2").astype(np.float32)
            / 32768.0)


def compare_generate_traces(ref_trace, eng_trace, margin_tol: float):
    """Margin-aware comparison of two free-running seek loops (oracle trace vs engine trace).

    North-star rule: greedy ids must be identical wherever the oracle's top-1 margin (and the timestamp-rule
    gap) exceeds the stated tolerance.  Rows are walked iteration by iteration; a row stops being compared at
    its first differing token, which must be a NON-decisive oracle step (afterwards the two loops legitimately
    see different contexts / seeks).  Returns (tokens_compared_equal, rows_fully_identical, first_diffs)."""
    assert eng_trace["langs"] == ref_trace["langs"], "language detection differs"
    alive = None
    agreed, first_diffs = 0, {}
    for k, it in enumerate(ref_trace["iterations"]):
        rows = list(it["rows"])
        if alive is None:
            alive = set(rows)
        if not any(b in alive for b in rows):
            break
        assert k < len(eng_trace["iterations"]), f"engine ran {len(eng_trace['iterations'])} seek iterations, oracle more"
        e = eng_trace["iterations"][k]
        for i, b in enumerate(rows):
            if b not in alive:
                continue
            assert b in e["rows"], f"row {b} retired early in the engine (iteration {k})"
            ei = e["rows"].index(b)
            assert e["seek"][ei] == it["seek"][i], f"row {b}: seek differs at iteration {k}"
            want_row = [int(t) for t in it["tokens"][i].tolist()]
            got_row = [int(t) for t in e["tokens"][ei][3:3 + len(want_row)]]
            for g, t in enumerate(want_row):
                if g >= len(it["record"]):
                    break
                if got_row[g] != t:
                    rec = it["record"][g]
                    decisive = float(rec["margin"][i]) > margin_tol and float(rec["rule_gap"][i]) > margin_tol
                    assert not decisive, (f"row {b} diverges at DECISIVE step {g} of iteration {k}: margin "
                                          f"{float(rec['margin'][i]):.3f}, rule gap {float(rec['rule_gap'][i]):.3f}")
                    first_diffs[b] = (k, g, float(rec["margin"][i]))
                    alive.discard(b)
                    break
                agreed += 1
    return agreed, sorted(alive or []), first_diffs
