"""MFCC front end -- same public API as the reference's ``mfcc.py`` (MFCC dataclass,
``feature_vector``, ``normalize_mfccs``, ``batch``), computed by the sm_100a kernels in
csrc/mfcc.cu through ``loe_mfcc_dev`` (include/loe_b200.h).

Reference: src/loe_speech_recognition/mfcc.py:12-84.  The reference delegates the arithmetic
to librosa (melspectrogram n_mels=40 n_fft=320 hop=160 fmin=133.33 fmax=6855.4976,
power_to_db(ref=np.max), mfcc(n_mfcc=13), delta, delta order 2); the kernels restate that
pipeline.  Only the filterbank weights are prepared on the host (once per
sample rate).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np
from numpy.typing import NDArray

from . import _native

N_FFT = 320
HOP = 160
N_MELS = 40
FMIN = 133.33
FMAX = 6855.4976


def _slaney_hz_to_mel(f: float) -> float:
    f_sp = 200.0 / 3
    if f >= 1000.0:
        return 1000.0 / f_sp + np.log(f / 1000.0) / (np.log(6.4) / 27.0)
    return f / f_sp


def _slaney_mel_to_hz(m: np.ndarray) -> np.ndarray:
    f_sp = 200.0 / 3
    lin = f_sp * m
    log = 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 1000.0 / f_sp))
    return np.where(m >= 1000.0 / f_sp, log, lin)


def mel_filterbank(sample_rate: float) -> NDArray[np.float32]:
    """Dense (40, 161) slaney-normalised triangular filterbank (librosa.filters.mel semantics)."""
    n_bins = 1 + N_FFT // 2
    freqs = np.arange(n_bins, dtype=np.float64) * (float(sample_rate) / N_FFT)
    edges = _slaney_mel_to_hz(np.linspace(_slaney_hz_to_mel(FMIN), _slaney_hz_to_mel(FMAX), N_MELS + 2))
    width = np.diff(edges)
    w = np.zeros((N_MELS, n_bins), dtype=np.float32)
    for m in range(N_MELS):
        rise = (freqs - edges[m]) / width[m]
        fall = (edges[m + 2] - freqs) / width[m + 1]
        w[m] = np.maximum(0.0, np.minimum(rise, fall))
    w *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return w


def mel_filterbank_sparse(sample_rate: float):
    """(start int32 [40], length int32 [40], weights float32 [LOE_MEL_MAXW*40]) with weight j
    of filter m stored at weights[j*40 + m] (the layout loe_mfcc_dev reads)."""
    dense = mel_filterbank(sample_rate)
    start = np.zeros(N_MELS, dtype=np.int32)
    length = np.zeros(N_MELS, dtype=np.int32)
    w = np.zeros((_native.LOE_MEL_MAXW, N_MELS), dtype=np.float32)
    for m in range(N_MELS):
        nz = np.nonzero(dense[m])[0]
        if nz.size == 0:
            continue
        a, b = int(nz[0]), int(nz[-1]) + 1
        if b - a > _native.LOE_MEL_MAXW:
            raise NotImplementedError(f"mel filter {m} spans {b - a} bins at sample_rate={sample_rate}; "
                                      f"the kernel supports {_native.LOE_MEL_MAXW}")
        start[m], length[m] = a, b - a
        w[: b - a, m] = dense[m, a:b]
    return start, length, np.ascontiguousarray(w.reshape(-1))


@dataclass
class MFCC:
    # Input
    signal: np.ndarray
    sample_rate: int | float

    # Settings
    n_mfcc: int = field(default=13)

    # Internals
    _feature_vector: np.ndarray = field(init=False)

    def __post_init__(self) -> None:
        _validate(self.signal)
        if self.n_mfcc != 13:
            raise NotImplementedError("the B200 MFCC kernel is built for n_mfcc=13 (the value every reference call site uses)")
        self._feature_vector = self.batch([self.signal], self.sample_rate)[0].T

    @property
    def feature_vector(self) -> np.ndarray:
        return self._feature_vector

    @staticmethod
    def normalize_mfccs(mfccs):
        """mfcc.py:50-69: statistics over axis 0, i.e. over the coefficients of each frame
        (host helper kept for API compatibility; the kernel applies the same formula)."""
        mean = np.mean(mfccs, axis=0, keepdims=True)
        std = np.std(mfccs, axis=0, keepdims=True)
        return (mfccs - mean) / (std + 1e-8)

    @classmethod
    def batch(cls, signals: List[NDArray], sample_rate: int) -> List[NDArray[np.float32]]:
        """List of (T, 39) float32 feature matrices (row = frame), one kernel pass for all signals."""
        from ._engine import get_engine

        for s in signals:
            _validate(s)
        if len(signals) == 0:
            return []
        eng = get_engine()
        b = eng.mfcc(signals, sample_rate)
        flat = b.feat.cpu().numpy()
        off = b.frm_off_host
        return [flat[off[i]:off[i + 1]] for i in range(len(signals))]


def _validate(signal) -> None:
    if not isinstance(signal, np.ndarray):
        raise TypeError("Input signal must be a numpy array.")
    if signal.ndim != 1:
        raise ValueError("Input signal must be 1-dimensional.")
