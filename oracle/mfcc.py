"""Oracle: the reference's MFCC front end, restated without librosa.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

PARITY UNPINNED: the arithmetic of ``src/loe_speech_recognition/mfcc.py:31-40``
lives in ``librosa`` (bare ``"librosa"`` in ``pyproject.toml:17-21``, i.e. no
pinned version; ``numpy==2.2.1`` in ``requirements.txt`` implies >= 0.10.2),
which is neither vendored under ``/root/reference`` nor installable here, and
the reference has no test or golden vector for it.  This file restates the
published librosa >= 0.10 algorithm for exactly the calls the reference makes,
built from the same scipy primitives librosa itself calls (cross-checked stage by stage against the
librosa-compatible paths of transformers.audio_utils and torchaudio in
tests/test_oracle_golden.py::test_mfcc_oracle_against_independent_librosa_restatements):

  melspectrogram(y, sr, n_mels=40, n_fft=320, hop_length=160,
                 fmin=133.33, fmax=6855.4976)          mfcc.py:31-34
  power_to_db(S, ref=np.max)                           mfcc.py:35
  feature.mfcc(S=log_mel, n_mfcc=13)                   mfcc.py:36
  feature.delta(m), feature.delta(m, order=2)          mfcc.py:39-40
  MFCC.normalize_mfccs (per-frame, over coefficients)  mfcc.py:50-69
  concatenate((norm, d1, d2), axis=0)                  mfcc.py:43
"""
from __future__ import annotations

import numpy as np
import scipy.fft
import scipy.signal

N_FFT = 320
HOP = 160
N_MELS = 40
N_MFCC = 13
FMIN = 133.33
FMAX = 6855.4976
DELTA_WIDTH = 9


def _hz_to_mel(f):
    """librosa.hz_to_mel(htk=False): Slaney scale."""
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if f.ndim:
        log_t = f >= min_log_hz
        mels[log_t] = min_log_mel + np.log(f[log_t] / min_log_hz) / logstep
    elif f >= min_log_hz:
        mels = min_log_mel + np.log(f / min_log_hz) / logstep
    return mels


def _mel_to_hz(m):
    """librosa.mel_to_hz(htk=False)."""
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if m.ndim:
        log_t = m >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (m[log_t] - min_log_mel))
    elif m >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (m - min_log_mel))
    return freqs


def mel_basis(sr=16000, n_fft=N_FFT, n_mels=N_MELS, fmin=FMIN, fmax=FMAX):
    """librosa.filters.mel(htk=False, norm='slaney', dtype=float32) -> (n_mels, 1+n_fft/2)."""
    n_bins = 1 + n_fft // 2
    weights = np.zeros((n_mels, n_bins), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def stft_power(y, n_fft=N_FFT, hop=HOP, window="hann", win_length=None):
    """|STFT|^2 as librosa computes it: float64 windowed rfft stored as complex64,
    centre-padded with zeros, periodic window, power in the real dtype (float32)."""
    y = np.asarray(y)
    if win_length is None:
        win_length = n_fft
    win = scipy.signal.get_window(window, win_length, fftbins=True)
    if win_length < n_fft:                       # librosa.util.pad_center
        lpad = (n_fft - win_length) // 2
        win = np.pad(win, (lpad, n_fft - win_length - lpad))
    pad = n_fft // 2
    yp = np.pad(y, (pad, pad), mode="constant")
    n_frames = 1 + (len(yp) - n_fft) // hop
    idx = np.arange(n_fft)[:, None] + hop * np.arange(n_frames)[None, :]
    frames = yp[idx]                              # (n_fft, T), dtype of y
    cdtype = np.complex64 if y.dtype == np.float32 else np.complex128
    spec = scipy.fft.rfft(win[:, None] * frames, axis=0).astype(cdtype)
    return np.abs(spec) ** 2.0                    # float32 for float32 input


def power_to_db(S, amin=1e-10, top_db=80.0):
    """librosa.power_to_db(S, ref=np.max): both maxima run over the whole array."""
    ref_value = np.max(S)
    log_spec = 10.0 * np.log10(np.maximum(amin, S))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    return np.maximum(log_spec, log_spec.max() - top_db)


def delta(m, order):
    """librosa.feature.delta(width=9, order=order, axis=-1, mode='interp')."""
    return scipy.signal.savgol_filter(m, DELTA_WIDTH, deriv=order, axis=-1,
                                      mode="interp", polyorder=order)


def normalize_mfccs(m):
    """mfcc.py:62-66: statistics over axis 0 = over the coefficients of each frame."""
    mean = np.mean(m, axis=0, keepdims=True)
    std = np.std(m, axis=0, keepdims=True)
    return (m - mean) / (std + 1e-8)


def log_mel(y, sr=16000):
    """(40, T) float32 dB mel spectrogram (stages 1-3 of the recipe)."""
    power = stft_power(y)
    mel = np.einsum("ft,mf->mt", power, mel_basis(sr), optimize=True)
    return power_to_db(mel)


def mfcc_feature_vector(y, sr=16000, n_mfcc=N_MFCC):
    """Equivalent of ``MFCC(signal, sr).feature_vector`` -> (3*n_mfcc, T) float32."""
    if not isinstance(y, np.ndarray):
        raise TypeError("Input signal must be a numpy array.")
    if y.ndim != 1:
        raise ValueError("Input signal must be 1-dimensional.")
    lm = log_mel(y, sr)
    ceps = scipy.fft.dct(lm, axis=-2, type=2, norm="ortho")[:n_mfcc, :]
    d1 = delta(ceps, 1)
    d2 = delta(ceps, 2)
    return np.concatenate((normalize_mfccs(ceps), d1, d2), axis=0)


def mfcc_batch(signals, sr=16000):
    """Equivalent of ``MFCC.batch``: list of (T, 39) transposed views."""
    return [mfcc_feature_vector(s, sr).T for s in signals]


# --------------------------------------------------------------------------------------
# Parameterised front end (BASELINE.json configs[3] "spec" set; SURVEY.md §8c last row)
# --------------------------------------------------------------------------------------
# PARITY UNPINNED BY CONSTRUCTION: the live reference has exactly one parameter set (mfcc.py:31-34) and never
# calls pre-emphasis, a Hamming window, a 512-point FFT or cepstral mean normalisation.  What follows is the
# same librosa-shaped pipeline as above with those stages made parameters; with REFERENCE_CONFIG it reduces to
# mfcc_feature_vector (tests/test_oracle_golden.py::test_parameterised_mfcc_reduces_to_reference).
REFERENCE_CONFIG = dict(n_fft=320, win_length=320, hop_length=160, window="hann", n_mels=40, fmin=FMIN, fmax=FMAX,
                        preemphasis=0.0, log="db", n_mfcc=13, norm="frame")
# 25 ms / 10 ms frames at 16 kHz, 512-point FFT, Hamming, pre-emphasis 0.97, natural log, CMN
SPEC_CONFIG = dict(n_fft=512, win_length=400, hop_length=160, window="hamming", n_mels=40, fmin=FMIN, fmax=FMAX,
                   preemphasis=0.97, log="ln", n_mfcc=13, norm="cmn")


def preemphasis(y, coef):
    """y[n] - coef * y[n-1] with y[-1] := y[0] (the first sample keeps (1 - coef) of its value), float32 arithmetic."""
    y = np.asarray(y, dtype=np.float32)
    if coef == 0.0:
        return y
    prev = np.concatenate((y[:1], y[:-1]))
    return (y - np.float32(coef) * prev).astype(np.float32)


def mfcc_feature_vector_ex(y, sr=16000, cfg=None):
    """(3 * n_mfcc, T) float32 features of one signal under ``cfg`` (a dict like REFERENCE_CONFIG)."""
    c = dict(REFERENCE_CONFIG)
    c.update(cfg or {})
    if not isinstance(y, np.ndarray):
        raise TypeError("Input signal must be a numpy array.")
    if y.ndim != 1:
        raise ValueError("Input signal must be 1-dimensional.")
    y = preemphasis(y.astype(np.float32), c["preemphasis"])
    power = stft_power(y, n_fft=c["n_fft"], hop=c["hop_length"], window=c["window"], win_length=c["win_length"])
    basis = mel_basis(sr, n_fft=c["n_fft"], n_mels=c["n_mels"], fmin=c["fmin"], fmax=c["fmax"])
    mel = np.einsum("ft,mf->mt", power, basis, optimize=True)
    if c["log"] == "db":
        lm = power_to_db(mel)
    elif c["log"] == "ln":
        lm = np.log(np.maximum(1e-10, mel))
    else:
        raise ValueError(c["log"])
    ceps = scipy.fft.dct(lm, axis=-2, type=2, norm="ortho")[: c["n_mfcc"], :]
    d1 = delta(ceps, 1)
    d2 = delta(ceps, 2)
    if c["norm"] == "frame":
        static = normalize_mfccs(ceps)
    elif c["norm"] == "cmn":                 # cepstral mean normalisation: per coefficient, over the utterance's frames
        static = ceps - np.mean(ceps, axis=1, keepdims=True)
    elif c["norm"] == "cmvn":
        static = (ceps - np.mean(ceps, axis=1, keepdims=True)) / (np.std(ceps, axis=1, keepdims=True) + 1e-8)
    elif c["norm"] == "none":
        static = ceps
    else:
        raise ValueError(c["norm"])
    return np.concatenate((static, d1, d2), axis=0).astype(np.float32)
