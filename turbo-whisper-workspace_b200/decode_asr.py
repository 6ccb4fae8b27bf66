"""Native token -> text/chunk stitching: the step right after the GPU path.

``AsrDecoder(tokenizer)(model_outputs, return_timestamps=..., return_language=..., time_precision=...)`` has the
contract of ``tokenizer._decode_asr`` ($TF/models/whisper/tokenization_whisper.py:901-1150): it takes the list of
``{"tokens": [[ids]], "stride": (chunk_len, stride_left, stride_right)}`` records the pipeline's postprocess builds
($TF/pipelines/automatic_speech_recognition.py:562-656) and returns ``(text, {"chunks": [...]})``.

The timestamp / stride state machine and the overlap merge run in the C library (``tw_decode_asr``,
csrc/decode_asr.cpp); the tokenizer is only read once, at construction, for its vocabulary: every id is turned into
its byte string (byte-level BPE: each vocabulary character stands for one byte, the GPT-2 table), so a chunk's text is
``b"".join(...)`` decoded as UTF-8 with replacement — no tokenizer call per chunk.
``return_timestamps="word"`` (records also carry ``token_timestamps``, one time per id, from the engine's alignment
path) is host Python below: the same state machine with per-token (start, end) pairs, the timestamp-aware overlap
merge, and the grouping of tokens into words (_collate_word_timestamps / _combine_tokens_into_words / _split_tokens_on_*
/ _merge_punctuations, $TF/models/whisper/tokenization_whisper.py:1153-1408).
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
        self.eos_token_id = int(tokenizer.eos_token_id) if getattr(tokenizer, "eos_token_id", None) is not None else self.decoder_start_token_id - 1
        self.default_language = getattr(tokenizer, "language", None)     # word splitting falls back to it, then to english
        self._special = np.asarray(special, dtype=np.int32)
        self._special_lang = np.asarray(lang_idx, dtype=np.int32)
        self.last_flags = 0

    def text_of(self, ids: Sequence[int]) -> str:
        tb = self.id_bytes
        return b"".join(tb[int(t)] for t in ids).decode("utf-8", errors="replace")

    def __call__(self, model_outputs: Sequence[Dict[str, Any]], *, return_timestamps, return_language=None,
                 time_precision: float) -> Tuple[str, Dict[str, Any]]:
        if return_timestamps == "word":
            return self._decode_words(model_outputs, return_language=return_language, time_precision=time_precision)
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

    # ------------------------------------------------------------------------------------------------ word mode
    _NO_SPACE_LANGUAGES = {"chinese", "japanese", "thai", "lao", "myanmar", "cantonese"}
    _PREPEND = "\"'“¡¿([{-"
    _APPEND = "\"'.。,，!！?？:：”)]}、"
    _PUNCT = "!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~"

    def _split_on_unicode(self, tokens: Sequence[int]):
        """_split_tokens_on_unicode (:1315-1343): cut wherever the bytes so far decode to complete code points."""
        full = self.text_of(tokens)
        rep = "\ufffd"
        words, word_tokens, indices = [], [], []
        cur_t, cur_i, off = [], [], 0
        for i, t in enumerate(tokens):
            cur_t.append(t)
            cur_i.append(i)
            dec = self.text_of(cur_t)
            if rep not in dec or full[off + dec.index(rep)] == rep:
                words.append(dec)
                word_tokens.append(cur_t)
                indices.append(cur_i)
                cur_t, cur_i = [], []
                off += len(dec)
        return words, word_tokens, indices

    def _split_on_spaces(self, tokens: Sequence[int]):
        """_split_tokens_on_spaces (:1346-1367)."""
        sub, sub_t, sub_i = self._split_on_unicode(tokens)
        words, word_tokens, indices = [], [], []
        for w, t, ix in zip(sub, sub_t, sub_i):
            special = t[0] >= self.eos_token_id
            if special or w.startswith(" ") or (w.strip() in self._PUNCT) or not words:
                words.append(w)
                word_tokens.append(t)
                indices.append(ix)
            else:
                words[-1] = words[-1] + w
                word_tokens[-1].extend(t)
                indices[-1].extend(ix)
        return words, word_tokens, indices

    def _merge_punctuations(self, words, tokens, indices):
        """_merge_punctuations (:1370-1405), in place."""
        i, j = len(words) - 2, len(words) - 1
        while i >= 0:
            if words[i].startswith(" ") and words[i].strip() in self._PREPEND:
                words[j] = words[i] + words[j]
                tokens[j] = tokens[i] + tokens[j]
                indices[j] = indices[i] + indices[j]
                words[i], tokens[i], indices[i] = "", [], []
            else:
                j = i
            i -= 1
        i, j = 0, 1
        while j < len(words):
            if not words[i].endswith(" ") and words[j] in self._APPEND:
                words[i] += words[j]
                tokens[i] += tokens[j]
                indices[i] += indices[j]
                words[j], tokens[j], indices[j] = "", [], []
            else:
                i = j
            j += 1
        words[:] = [w for w in words if w]
        tokens[:] = [t for t in tokens if t]
        indices[:] = [ix for ix in indices if ix]

    def _collate_words(self, tokens, token_times, language, return_language):
        """_collate_word_timestamps (:1273-1286)."""
        lang = language if language is not None else (self.default_language or "english")
        split = self._split_on_unicode if lang in self._NO_SPACE_LANGUAGES else self._split_on_spaces
        words, word_tokens, indices = split(tokens)
        self._merge_punctuations(words, word_tokens, indices)
        extra = {"language": language} if return_language else {}
        return [{"text": w, "timestamp": (token_times[ix[0]][0], token_times[ix[-1]][1]), **extra}
                for w, ix in zip(words, indices)]

    @staticmethod
    def _merge_with_times(sequences, time_sequences):
        """_find_longest_common_sequence with token_timestamp_sequences (:1153-1270): a position only counts as a
        match when the ids agree AND the left token's (start, end) pair is <= the right token's."""
        left, left_t = sequences[0], time_sequences[0]
        total, total_t = [], []
        for k, right in enumerate(sequences[1:]):
            right_t = time_sequences[k + 1]
            ll, rl = len(left), len(right)
            best, best_idx = 0.0, (ll, ll, 0, 0)
            for i in range(1, ll + rl):
                ls, le = max(0, ll - i), min(ll, ll + rl - i)
                rs, re_ = max(0, i - ll), min(rl, i)
                matches = 0
                for d in range(le - ls):
                    if left[ls + d] == right[rs + d] and left_t[ls + d] <= right_t[rs + d]:
                        matches += 1
                matching = matches / i + i / 10000.0
                if matches > 1 and matching > best:
                    best, best_idx = matching, (ls, le, rs, re_)
            ls, le, rs, re_ = best_idx
            lm, rm = (le + ls) // 2, (re_ + rs) // 2
            total.extend(left[:lm])
            total_t.extend(left_t[:lm])
            left, left_t = right[rm:], right_t[rm:]
        total.extend(left)
        total_t.extend(left_t)
        return total, total_t

    def _decode_words(self, model_outputs, *, return_language, time_precision):
        """tokenizer._decode_asr(..., return_timestamps="word") (:901-1150)."""
        tb = self.timestamp_begin
        special = {int(i): int(l) for i, l in zip(self._special, self._special_lang)}
        last_language = None
        new_chunk = lambda: {"language": last_language, "timestamp": [None, None], "text": ""}
        chunks, chunk = [], new_chunk()
        time_offset = 0.0
        previous_tokens, previous_times = [], []
        skip = False
        right_stride_start = None
        self.last_flags = 0

        def close(cur_chunk):
            toks, times = self._merge_with_times(previous_tokens, previous_times)
            cur_chunk["text"] = self.text_of(toks)
            cur_chunk["words"] = self._collate_words(toks, times, last_language, return_language)
            chunks.append(cur_chunk)

        for output in model_outputs:
            toks0 = output["tokens"][0]
            token_ids = [int(t) for t in (toks0.tolist() if hasattr(toks0, "tolist") else toks0)]
            if token_ids and token_ids[0] == self.prompt_token_id:      # _strip_prompt (:725-741)
                token_ids = (token_ids[token_ids.index(self.decoder_start_token_id):]
                             if self.decoder_start_token_id in token_ids else [])
            tt0 = output["token_timestamps"][0]
            token_times = [float(x) for x in (tt0.tolist() if hasattr(tt0, "tolist") else tt0)]
            last_timestamp, first_timestamp = None, tb
            cur_max, prev_len, penultimate = 0.0, 0.0, 0.0
            if "stride" in output:
                chunk_len, stride_left, stride_right = output["stride"]
                time_offset -= stride_left
                right_stride_start = chunk_len - stride_right
                if stride_left:
                    first_timestamp = stride_left / time_precision + tb
                if stride_right:
                    for token in reversed(token_ids):
                        if token >= tb:
                            if last_timestamp is not None and (token - tb) * time_precision < right_stride_start:
                                break
                            last_timestamp = token
            current_tokens, current_times = [], []
            for i, token in enumerate(token_ids):
                if token in special:
                    li = special[token]
                    if li >= 0:
                        chunk["language"] = self.languages[li]
                        last_language = self.languages[li]
                elif token >= tb:
                    timestamp = float((token - tb) * time_precision)
                    if timestamp < cur_max:
                        single_ending = i >= 2 and not (token_ids[i - 1] >= tb and token_ids[i - 2] >= tb)
                        if single_ending:
                            prev_len += time_precision * self.segment_size
                        else:
                            cur_max = penultimate
                            prev_len += penultimate
                    penultimate = cur_max
                    cur_max = timestamp
                    time = round((token - tb) * time_precision + time_offset + prev_len, 2)
                    if last_timestamp and token >= last_timestamp:
                        skip = True
                    elif skip or (previous_tokens and token < first_timestamp):
                        skip = False
                    elif chunk["timestamp"][0] is None:
                        chunk["timestamp"][0] = time
                    elif time != chunk["timestamp"][0]:
                        chunk["timestamp"][1] = time
                        previous_tokens.append(current_tokens)
                        previous_times.append(current_times)
                        close(chunk)
                        previous_tokens, current_tokens = [], []
                        previous_times, current_times = [], []
                        chunk = new_chunk()
                else:
                    current_tokens.append(token)
                    start = round(0.0 + time_offset, 2) if i == 0 else round(token_times[i - 1] + time_offset, 2)
                    current_times.append((start, round(token_times[i] + time_offset, 2)))
            if "stride" in output:
                time_offset += chunk_len - stride_right
            if current_tokens:
                previous_tokens.append(current_tokens)
                previous_times.append(current_times)
            elif not any(p for p in previous_tokens):
                chunk = new_chunk()
                previous_tokens, current_tokens = [], []
                previous_times, current_times = [], []
        if previous_tokens:
            self.last_flags |= 1
            close(chunk)
        text = "".join(c["text"] for c in chunks)
        words = []
        for c in chunks:
            words.extend(c["words"])
        return text, {"chunks": words}
