"""Oracle: Whisper log-mel front end, numpy restatement (test infrastructure, see oracle/__init__.py).

Follows ``WhisperFeatureExtractor`` of transformers 5.5.0:
  * ``__init__``                          $TF/models/whisper/feature_extraction_whisper.py:69-103
  * ``_torch_extract_fbank_features``     $TF/models/whisper/feature_extraction_whisper.py:135-164
  * ``_np_extract_fbank_features``        $TF/models/whisper/feature_extraction_whisper.py:105-133
  * ``__call__`` (pad / truncate / mask)  $TF/models/whisper/feature_extraction_whisper.py:189-342
  * ``mel_filter_bank``                   $TF/audio_utils.py:453-544 (+ hertz_to_mel :263-296,
                                          mel_to_hertz :299-332, _create_triangular_filter_bank :356-375)
which is what the reference's pipeline call reaches (ref:vocalis/core/audio_pipeline.py:351-358 ->
$TF/pipelines/automatic_speech_recognition.py:61-84).
"""
from __future__ import annotations

import numpy as np

SAMPLING_RATE = 16000
N_FFT = 400
HOP = 160
N_SAMPLES = 480000
N_FRAMES = 3000
N_MELS = 128


def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    mels = 3.0 * f / 200.0
    logstep = 27.0 / np.log(6.4)
    with np.errstate(divide="ignore", invalid="ignore"):
        log_part = 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) * logstep
    return np.where(f >= 1000.0, log_part, mels)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    logstep = np.log(6.4) / 27.0
    return np.where(m >= 15.0, 1000.0 * np.exp(logstep * (m - 15.0)), 200.0 * m / 3.0)


def mel_filter_bank(n_bins: int = 201, n_mels: int = N_MELS, fmin: float = 0.0, fmax: float = 8000.0,
                    sr: int = SAMPLING_RATE) -> np.ndarray:
    """[n_bins, n_mels] float64, slaney scale + slaney (area) normalisation."""
    mel_pts = np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2)
    hz_pts = _mel_to_hz_slaney(mel_pts)
    fft_freqs = np.linspace(0, sr // 2, n_bins)
    diff = np.diff(hz_pts)
    slopes = hz_pts[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / diff[:-1]
    up = slopes[:, 2:] / diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    fb *= (2.0 / (hz_pts[2:n_mels + 2] - hz_pts[:n_mels]))[None, :]
    return fb


def hann_periodic(n: int = N_FFT) -> np.ndarray:
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def pad_or_trim(pcm: np.ndarray, n: int = N_SAMPLES):
    """truncation=True, padding='max_length' of the extractor's __call__ (zero padding on the right).
    Returns (padded fp32 [n], number of valid samples)."""
    pcm = np.asarray(pcm, dtype=np.float32).reshape(-1)
    nv = min(len(pcm), n)
    out = np.zeros(n, dtype=np.float32)
    out[:nv] = pcm[:nv]
    return out, nv


def log_mel(pcm: np.ndarray, dtype=np.float32) -> np.ndarray:
    """[128, 3000] features for one clip of <= 30 s (longer clips are truncated, like HF).

    ``dtype=np.float32`` mirrors the torch path the pipeline uses (fp32 throughout);
    ``np.float64`` gives a higher-precision restatement used to bound both implementations."""
    x, _ = pad_or_trim(pcm)
    x = x.astype(dtype)
    xp = np.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect")  # torch.stft(center=True, pad_mode="reflect")
    n_frames = 1 + (len(xp) - N_FFT) // HOP  # 3001
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(n_frames)[:, None]
    frames = xp[idx] * hann_periodic().astype(dtype)[None, :]
    spec = np.fft.rfft(frames.astype(np.float64 if dtype == np.float64 else np.float32), axis=1)
    power = (np.abs(spec) ** 2).astype(dtype)[:-1]  # drop the last frame -> [3000, 201]
    mel = power @ mel_filter_bank().astype(dtype)  # [3000, 128]
    logspec = np.log10(np.maximum(mel, 1e-10)).T  # [128, 3000]
    logspec = np.maximum(logspec, logspec.max() - 8.0)
    return ((logspec + 4.0) / 4.0).astype(np.float32)


def log_mel_long(pcm: np.ndarray, dtype=np.float32) -> np.ndarray:
    """[128, n // 160] features of a clip LONGER than 30 s, as the ASR pipeline builds them for un-chunked long-form
    input ($TF/pipelines/automatic_speech_recognition.py:446-454: truncation=False, padding="longest"): the whole
    waveform goes through the same STFT (reflect padding only at the two ends of the clip), the last frame is
    dropped, and the `max - 8` clamp uses the maximum of the WHOLE clip
    ($TF/models/whisper/feature_extraction_whisper.py:135-164)."""
    x = np.asarray(pcm, dtype=np.float32).reshape(-1).astype(dtype)
    xp = np.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect")
    n_frames = 1 + (len(xp) - N_FFT) // HOP
    win = hann_periodic().astype(dtype)[None, :]
    fb = mel_filter_bank().astype(dtype)
    out = np.empty((N_MELS, n_frames - 1), dtype=dtype)
    for f0 in range(0, n_frames - 1, 4096):      # blocks of frames: a 1 h clip would not fit as one index matrix
        f1 = min(n_frames - 1, f0 + 4096)
        idx = np.arange(N_FFT)[None, :] + HOP * np.arange(f0, f1)[:, None]
        frames = xp[idx] * win
        spec = np.fft.rfft(frames.astype(np.float64 if dtype == np.float64 else np.float32), axis=1)
        power = (np.abs(spec) ** 2).astype(dtype)
        out[:, f0:f1] = np.log10(np.maximum(power @ fb, 1e-10)).T
    out = np.maximum(out, out.max() - 8.0)
    return ((out + 4.0) / 4.0).astype(np.float32)


def attention_mask(n_valid: int) -> np.ndarray:
    """Frame mask of the extractor: sample mask [::160] ($TF/...feature_extraction_whisper.py:331-339)."""
    m = np.zeros(N_SAMPLES, dtype=np.int32)
    m[:min(n_valid, N_SAMPLES)] = 1
    return m[::HOP]
