"""Energy-based silence stripper with the reference's API (signal_separation.py:43-164), computed by
``loe_silence_dev`` (csrc/vad.cu) for a whole batch in one launch.

Same constructor fields, ``remove_empty`` / ``remove_empty_batch`` / ``get_all_noises``, the
``FailToProcess`` exception and the reference's list quirks: noise frames collected during a signal
whose speech never ends are carried into the next successful signal's noise clip, and a stripped
signal shorter than 9 frames fails after its noise has been recorded (:88-101).
"""
from __future__ import annotations

import logging
from dataclasses import dataclass, field
from typing import List

import numpy as np
from numpy.typing import NDArray

logger = logging.getLogger(__name__)


@dataclass
class SignalSeparation:
    class FailToProcess(Exception):
        def __init__(self, *args: object) -> None:
            super().__init__(*args)
            logger.error("Failed to process signal")

    sample_rate: int = field(default=16000)
    frame_time: float = field(default=0.01)
    speech_high_threshold: float = field(default=0.08)
    speech_low_threshold: float = field(default=0.01)
    silence_duration_threshold: float = field(default=0.02)

    _noises: List[NDArray[np.float32]] = field(default_factory=list)
    _max_volume: float = field(init=False)
    _result: List[NDArray[np.float32]] = field(default_factory=list)
    _noise: List[NDArray[np.float32]] = field(default_factory=list)

    @property
    def frame_size(self) -> int:
        return int(self.sample_rate * self.frame_time)

    @property
    def maximum_silence_frames(self) -> int:
        return int(self.silence_duration_threshold / self.frame_time)

    def _segment_batch(self, signals):
        from ._engine import get_engine
        return get_engine().silence(signals, self.frame_size, self.speech_high_threshold, self.speech_low_threshold,
                                    self.maximum_silence_frames)

    def _finish(self, signal, noise_mask, seg, max_volume):
        """Host bookkeeping of one signal given the kernel's decisions; returns the stripped signal or
        raises FailToProcess exactly where the reference does."""
        fs = self.frame_size
        done, start, end, n_frames = (int(v) for v in seg)
        self._max_volume = float(max_volume)
        upto = end + 1 if done else n_frames
        for f in np.nonzero(noise_mask[:upto])[0]:
            self._noise.append(signal[f * fs:(f + 1) * fs])
        self._result = [signal[f * fs:(f + 1) * fs] for f in range(start, end)] if done else []
        if not done:
            raise self.FailToProcess
        self._noises.append(np.concatenate(self._noise, dtype=np.float32))
        self._noise = []
        if len(self._result) < 9:                       # threshold based on the MFCC delta window
            raise self.FailToProcess
        return np.ascontiguousarray(signal[start * fs:end * fs], dtype=np.float32)

    def remove_empty(self, signal: NDArray[np.float32]) -> NDArray[np.float32]:
        _, noise, seg, mx, eoff = self._segment_batch([signal])
        return self._finish(signal, noise, seg[0], mx[0])

    def remove_empty_batch(self, signals: List[NDArray[np.float32]]) -> List[NDArray[np.float32]]:
        if len(signals) == 0:
            return []
        _, noise, seg, mx, eoff = self._segment_batch(signals)
        out = []
        for i, signal in enumerate(signals):
            try:
                out.append(self._finish(signal, noise[eoff[i]:eoff[i + 1]], seg[i], mx[i]))
            except self.FailToProcess:
                logger.warning(f"Signal with property: length {signal.shape[0]}, max {np.abs(np.max(signal))} failed")
        return out

    def get_all_noises(self) -> List[NDArray[np.float32]]:
        return self._noises
