"""Native token -> text/chunk stitching: the step right after the GPU path.

``AsrDecoder(tokenizer)(model_outputs, return_timestamps=..., return_language=..., time_precision=...)`` has the
contract of ``tokenizer._decode_asr`` ($TF/models/whisper/tokenization_whisper.py:901-1150): it takes the list of
``{"tokens": [[ids]], "stride": (chunk_len, stride_left, stride_right)}`` records the pipeline's postprocess builds
($TF/pipelines/automatic_speech_recognition.py:562-656) and returns ``(text, {"chunks": [...]})``.

The timestamp / stride state machine and the overlap merge run in the C library (``tw_decode_asr``,
csrc/decode_asr.cpp); the tokenizer is only read once, at construction, for its vocabulary: every id is turned into
its byte string (byte-level BPE: each vocabulary character stands for one byte, the GPT-2 table), so a chunk's text is
``b"".join(...)`` decoded as UTF-8 with replacement — no tokenizer call per chunk.
``return_timestamps="word"`` is not covered (it needs cross-attention weights the engine does not export yet).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib

# Whisper's language codes in vocabulary order (<|en|> = <|startoftranscript|> + 1, ...) and their names
# (public model metadata; tests/test_decode_asr.py checks the table against transformers' LANGUAGES).
LANGUAGE_NAMES: Dict[str, str] = {
    "en": "english", "zh": "chinese", "de": "german", "es": "spanish", "ru": "russian", "ko": "korean", "fr": "french",
    "ja": "japanese", "pt": "portuguese", "tr": "turkish", "pl": "polish", "ca": "catalan", "nl": "dutch", "ar": "arabic",
    "sv": "swedish", "it": "italian", "id": "indonesian", "hi": "hindi", "fi": "finnish", "vi": "vietnamese",
    "he": "hebrew", "uk": "ukrainian", "el": "greek", "ms": "malay", "cs": "czech", "ro": "romanian", "da": "danish",
    "hu": "hungarian", "ta": "tamil", "no": "norwegian", "th": "thai", "ur": "urdu", "hr": "croatian", "bg": "bulgarian",
    "lt": "lithuanian", "la": "latin", "mi": "maori", "ml": "malayalam", "cy": "welsh", "sk": "slovak", "te": "telugu",
    "fa": "persian", "lv": "latvian", "bn": "bengali", "sr": "serbian", "az": "azerbaijani", "sl": "slovenian",
    "kn": "kannada", "et": "estonian", "mk": "macedonian", "br": "breton", "eu": "basque", "is": "icelandic",
    "hy": "armenian", "ne": "nepali", "mn": "mongolian", "bs": "bosnian", "kk": "kazakh", "sq": "albanian",
    "sw": "swahili", "gl": "galician", "mr": "marathi", "pa": "punjabi", "si": "sinhala", "km": "khmer", "sn": "shona",
    "yo": "yoruba", "so": "somali", "af": "afrikaans", "oc": "occitan", "ka": "georgian", "be": "belarusian",
    "tg": "tajik", "sd": "sindhi", "gu": "gujarati", "am": "amharic", "yi": "yiddish", "lo": "lao", "uz": "uzbek",
    "fo": "faroese", "ht": "haitian creole", "ps": "pashto", "tk": "turkmen", "nn": "nynorsk", "mt": "maltese",
    "sa": "sanskrit", "lb": "luxembourgish", "my": "myanmar", "bo": "tibetan", "tl": "tagalog", "mg": "malagasy",
    "as": "assamese", "tt": "tatar", "haw": "hawaiian", "ln": "lingala", "ha": "hausa", "ba": "bashkir",
    "jw": "javanese", "su": "sundanese", "yue": "cantonese",
}


def _unicode_to_byte() -> Dict[str, int]:
    """Inverse of the byte-level BPE alphabet: printable bytes stand for themselves, the other 68 are mapped to
    code points 256, 257, ... in byte order."""
    keep = list(range(33, 127)) + list(range(161, 173)) + list(range(174, 256))
    table, extra = {}, 0
    for b in range(256):
        if b in keep:
            table[chr(b)] = b
        else:
            table[chr(256 + extra)] = b
            extra += 1
    return table


class AsrWindow(C.Structure):
    _fields_ = [("tokens", C.POINTER(C.c_int32)), ("n_tokens", C.c_int32), ("has_stride", C.c_int32),
                ("chunk_len", C.c_double), ("stride_left", C.c_double), ("stride_right", C.c_double)]


class AsrConfig(C.Structure):
    _fields_ = [("timestamp_begin", C.c_int32), ("prompt_token_id", C.c_int32), ("decoder_start_token_id", C.c_int32),
                ("return_timestamps", C.c_int32), ("segment_size", C.c_int32), ("n_special", C.c_int32),
                ("special_ids", C.POINTER(C.c_int32)), ("special_lang", C.POINTER(C.c_int32)),
                ("time_precision", C.c_double)]


class AsrDecoder:
    def __init__(self, tokenizer, segment_size: int = 1500):
        n = len(tokenizer)
        strings = tokenizer.convert_ids_to_tokens(list(range(n)))
        added = set(int(v) for v in tokenizer.get_added_vocab().values())
        u2b = _unicode_to_byte()
        self.id_bytes: List[bytes] = []
        for i, s in enumerate(strings):
            if s is None:
                self.id_bytes.append(b"")
            elif i in added or any(ch not in u2b for ch in s):
                self.id_bytes.append(s.encode("utf-8"))       # added tokens are literal text
            else:
                self.id_bytes.append(bytes(u2b[ch] for ch in s))
        self.timestamp_begin = int(tokenizer.convert_tokens_to_ids("<|notimestamps|>")) + 1
        self.prompt_token_id = int(tokenizer.convert_tokens_to_ids("<|startofprev|>"))
        self.decoder_start_token_id = int(tokenizer.convert_tokens_to_ids("<|startoftranscript|>"))
        self.segment_size = int(segment_size)
        special = sorted(set(int(i) for i in tokenizer.all_special_ids))
        self.languages: List[str] = []       # language index -> name
        lang_idx = []
        for i in special:
            code = strings[i][2:-2] if strings[i] is not None else ""
            name = LANGUAGE_NAMES.get(code)
            if name is None:
                lang_idx.append(-1)
            else:
                lang_idx.append(len(self.languages))
                self.languages.append(name)
        self._special = np.asarray(special, dtype=np.int32)
        self._special_lang = np.asarray(lang_idx, dtype=np.int32)
        self.last_flags = 0

    def text_of(self, ids: Sequence[int]) -> str:
        tb = self.id_bytes
        return b"".join(tb[int(t)] for t in ids).decode("utf-8", errors="replace")

    def __call__(self, model_outputs: Sequence[Dict[str, Any]], *, return_timestamps, return_language=None,
                 time_precision: float) -> Tuple[str, Dict[str, Any]]:
        if return_timestamps == "word":
            raise NotImplementedError("word timestamps are not implemented by the native _decode_asr")
        lib = _lib.load()
        wins = (AsrWindow * max(1, len(model_outputs)))()
        keep = []
        total = 0
        for w, out in enumerate(model_outputs):
            toks = out["tokens"]
            ids = np.ascontiguousarray(np.asarray(toks[0].tolist() if hasattr(toks[0], "tolist") else toks[0], dtype=np.int32))
            keep.append(ids)
            wins[w].tokens = ids.ctypes.data_as(C.POINTER(C.c_int32))
            wins[w].n_tokens = int(ids.size)
            total += int(ids.size)
            if "stride" in out:
                cl, sl, sr = out["stride"]
                wins[w].has_stride, wins[w].chunk_len, wins[w].stride_left, wins[w].stride_right = 1, float(cl), float(sl), float(sr)
        cfg = AsrConfig()
        cfg.timestamp_begin, cfg.prompt_token_id = self.timestamp_begin, self.prompt_token_id
        cfg.decoder_start_token_id = self.decoder_start_token_id
        cfg.return_timestamps = 1 if return_timestamps else 0
        cfg.segment_size = self.segment_size
        cfg.n_special = int(self._special.size)
        cfg.special_ids = self._special.ctypes.data_as(C.POINTER(C.c_int32))
        cfg.special_lang = self._special_lang.ctypes.data_as(C.POINTER(C.c_int32))
        cfg.time_precision = float(time_precision)
        cap = total + 1
        out_tokens = np.empty(cap, dtype=np.int32)
        offsets = np.zeros(cap + 1, dtype=np.int64)
        t0, t1 = np.empty(cap, dtype=np.float64), np.empty(cap, dtype=np.float64)
        lang = np.empty(cap, dtype=np.int32)
        n_chunks, flags = C.c_int32(0), C.c_int32(0)
        p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        _lib.check(lib.tw_decode_asr(wins, len(model_outputs), C.byref(cfg), p(out_tokens, C.c_int32), cap,
                                     p(offsets, C.c_int64), p(t0, C.c_double), p(t1, C.c_double), p(lang, C.c_int32),
                                     cap, C.byref(n_chunks), C.byref(flags)), "tw_decode_asr")
        self.last_flags = int(flags.value)
        chunks = []
        for c in range(n_chunks.value):
            ids = out_tokens[offsets[c]:offsets[c + 1]]
            ch: Dict[str, Any] = {"text": self.text_of(ids)}
            if return_timestamps:
                ch["timestamp"] = (None if math.isnan(t0[c]) else float(t0[c]), None if math.isnan(t1[c]) else float(t1[c]))
            if return_language:
                ch["language"] = self.languages[lang[c]] if lang[c] >= 0 else None
            chunks.append(ch)
        text = "".join(ch["text"] for ch in chunks)
        optional = {"chunks": chunks} if (return_timestamps or return_language) else {}
        return text, optional
