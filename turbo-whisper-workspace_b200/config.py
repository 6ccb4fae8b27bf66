"""Model dimensions and generation settings of the Whisper checkpoints the path serves.

Mirrors the fields of ``WhisperConfig`` / ``generation_config.json`` that the hot path reads
(large-v3-turbo = 32 encoder / 4 decoder layers, large-v3 = 32 / 32)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

N_SAMPLES = 480000   # 30 s at 16 kHz
N_FRAMES = 3000      # mel frames per window
SAMPLING_RATE = 16000

# generation_config.json of openai/whisper-large-v3(-turbo)
_SUPPRESS = [
    1, 2, 7, 8, 9, 10, 14, 25, 26, 27, 28, 29, 31, 58, 59, 60, 61, 62, 63, 90, 91, 92, 93, 359, 503, 522, 542, 873,
    893, 902, 918, 922, 931, 1350, 1853, 1982, 2460, 2627, 3246, 3253, 3268, 3536, 3846, 3961, 4183, 4667, 6585, 6647,
    7273, 9061, 9383, 10428, 10929, 11938, 12033, 12331, 12562, 13793, 14157, 14635, 15265, 15618, 16553, 16604, 18362,
    18956, 20075, 21675, 22520, 26130, 26161, 26435, 28279, 29464, 31650, 32302, 32470, 36865, 42863, 47425, 49870,
    50254, 50258, 50359, 50360, 50361, 50362, 50363,
]


# generation_config.alignment_heads of the two checkpoints (from memory of the hub files, as the suppress list above —
# verify when a checkpoint is available; GenerationSettings.from_hf copies the real field).  Used by tools/ only.
ALIGNMENT_HEADS = {
    "large-v3-turbo": [[2, 4], [2, 11], [3, 3], [3, 6], [3, 11], [3, 14]],
    "large-v3": [[7, 0], [10, 17], [12, 18], [13, 12], [16, 1], [17, 14], [19, 11], [21, 4], [24, 1], [25, 6]],
}


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
    vocab: int = 51866

    def validate(self) -> None:
        if self.d_model != self.heads * 64:
            raise ValueError("turbo-whisper-workspace_b200 kernels are built for head_dim 64")
        if self.d_model % 256 or self.ffn % 256:
            raise ValueError("d_model and ffn must be multiples of 256")
        if self.n_mels != 128 or self.max_source_positions != 1500:
            raise ValueError("the log-mel / conv-stem kernels are built for 128 mel bins and 1500 source positions")
        if self.max_target_positions > 448:
            raise ValueError("max_target_positions > 448 is not supported")

    @classmethod
    def large_v3_turbo(cls) -> "WhisperDims":
        return cls()

    @classmethod
    def large_v3(cls) -> "WhisperDims":
        return cls(dec_layers=32)

    @classmethod
    def from_hf_config(cls, cfg) -> "WhisperDims":
        return cls(d_model=cfg.d_model, heads=cfg.encoder_attention_heads, ffn=cfg.encoder_ffn_dim,
                   enc_layers=cfg.encoder_layers, dec_layers=cfg.decoder_layers, n_mels=cfg.num_mel_bins,
                   max_source_positions=cfg.max_source_positions, max_target_positions=cfg.max_target_positions,
                   vocab=cfg.vocab_size)


@dataclass
class GenerationSettings:
    eos_token_id: int = 50257
    pad_token_id: int = 50257
    decoder_start_token_id: int = 50258
    no_timestamps_token_id: int = 50364
    max_length: int = 448
    max_initial_timestamp_index: int = 50
    suppress_tokens: List[int] = field(default_factory=lambda: list(_SUPPRESS))
    begin_suppress_tokens: List[int] = field(default_factory=lambda: [220, 50257])
    lang_first: int = 50259
    lang_last: int = 50358
    task_to_id: Dict[str, int] = field(default_factory=lambda: {"transcribe": 50360, "translate": 50359})
    lang_to_id: Dict[str, int] = field(default_factory=dict)  # "<|en|>" -> id, optional (explicit language)
    # word-level timestamps: (decoder layer, head) pairs of generation_config.alignment_heads (None: not available, as
    # for a generation config without the field) and WhisperConfig.median_filter_width
    alignment_heads: Optional[List[List[int]]] = None
    median_filter_width: int = 7

    @property
    def timestamp_begin(self) -> int:
        return self.no_timestamps_token_id + 1

    @classmethod
    def from_hf(cls, gc) -> "GenerationSettings":
        """From a transformers GenerationConfig carrying the Whisper fields."""
        s = cls()
        for name in ("eos_token_id", "pad_token_id", "decoder_start_token_id", "no_timestamps_token_id", "max_length",
                     "max_initial_timestamp_index"):
            v = getattr(gc, name, None)
            if v is not None:
                setattr(s, name, int(v[0] if isinstance(v, (list, tuple)) else v))
        if getattr(gc, "suppress_tokens", None) is not None:
            s.suppress_tokens = [int(t) for t in gc.suppress_tokens]
        if getattr(gc, "begin_suppress_tokens", None) is not None:
            s.begin_suppress_tokens = [int(t) for t in gc.begin_suppress_tokens]
        if getattr(gc, "task_to_id", None):
            s.task_to_id = {k: int(v) for k, v in gc.task_to_id.items()}
        if getattr(gc, "lang_to_id", None):
            s.lang_to_id = {k: int(v) for k, v in gc.lang_to_id.items()}
            ids = sorted(s.lang_to_id.values())
            if ids != list(range(ids[0], ids[-1] + 1)):
                raise ValueError("language ids must be contiguous")
            s.lang_first, s.lang_last = ids[0], ids[-1]
        if getattr(gc, "alignment_heads", None) is not None:
            s.alignment_heads = [[int(l), int(h)] for l, h in gc.alignment_heads]
        return s
