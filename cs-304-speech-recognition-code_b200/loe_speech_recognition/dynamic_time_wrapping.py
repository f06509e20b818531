"""Template DTW recogniser with the reference's API (dynamic_time_wrapping.py:13-120): the templates'
and the sample's MFCCs come from the MFCC kernel, the trellis fill from ``loe_dtw_dev`` (csrc/dtw.cu).

``DynamicTimeWarping(sequences, sample, ...).search() -> (index, min_distance)`` as in the reference;
added: ``from_features`` (skip the front end) and ``search_batch`` (many samples against the same
templates in one launch).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .mfcc import MFCC


@dataclass
class DynamicTimeWarping:
    sequences: List[np.ndarray]            # raw template signals (one per word)
    sample: np.ndarray                     # raw sample signal
    sample_rate: int | float = field(default=16000)
    trace_back: bool = field(default=False)
    pruning: bool = field(default=True)
    pruning_factor: float = field(default=4)

    _sequences: np.ndarray = field(init=False)
    _sample: np.ndarray = field(init=False)
    _cost_matrix: Optional[np.ndarray] = field(init=False, default=None)
    _path_matrix: Optional[np.ndarray] = field(init=False, default=None)
    _number_of_words_in_sequences: int = field(init=False)
    _word_length_in_sequences: List[int] = field(init=False)
    _word_starting_positions: List[int] = field(init=False)
    _height: int = field(init=False)
    _length: int = field(init=False)

    def __post_init__(self):
        feats = MFCC.batch(list(self.sequences) + [self.sample], self.sample_rate)   # one pass for all signals
        self._set_features(feats[:-1], feats[-1])

    def _set_features(self, seq_feats: Sequence[np.ndarray], sample_feat: np.ndarray) -> None:
        self._seq_feats = [np.ascontiguousarray(f, dtype=np.float32) for f in seq_feats]
        self._word_length_in_sequences = [f.shape[0] for f in self._seq_feats]
        self._sequences = np.concatenate(self._seq_feats)
        self._sample = np.ascontiguousarray(sample_feat, dtype=np.float32)
        self._number_of_words_in_sequences = len(self._seq_feats)
        self._height, self._length = self._sequences.shape[0], self._sample.shape[0]
        self._word_starting_positions = [0] + np.cumsum(self._word_length_in_sequences)[:-1].astype(int).tolist()

    @classmethod
    def from_features(cls, seq_feats: Sequence[np.ndarray], sample_feat: np.ndarray, trace_back: bool = False,
                      pruning: bool = True, pruning_factor: float = 4) -> "DynamicTimeWarping":
        """Build from (T, D) feature matrices instead of raw signals (added)."""
        obj = cls.__new__(cls)
        obj.sequences, obj.sample, obj.sample_rate = [], None, 16000
        obj.trace_back, obj.pruning, obj.pruning_factor = trace_back, pruning, pruning_factor
        obj._cost_matrix = obj._path_matrix = None
        obj._set_features(seq_feats, sample_feat)
        return obj

    def search(self) -> Tuple[int, float]:
        from ._engine import get_engine
        bi, bd, _, cost, path = get_engine().dtw(self._seq_feats, [self._sample], self.pruning, self.pruning_factor,
                                                 want_matrices=True)
        self._cost_matrix = cost
        self._path_matrix = path.astype(int) if self.trace_back else np.zeros_like(path, dtype=int)
        return int(bi[0]), np.float64(bd[0])

    def search_batch(self, sample_feats: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
        """(index [n], min_distance [n]) for many (T, D) sample feature matrices (added)."""
        from ._engine import get_engine
        bi, bd, _, _, _ = get_engine().dtw(self._seq_feats, list(sample_feats), self.pruning, self.pruning_factor)
        return bi.astype(int), bd

    @staticmethod
    def euclidean_distance(point1, point2):
        return np.sqrt(np.sum((point1 - point2) ** 2))
