"""Alignment containers of the reference API (signal.py:15-91).

``Signal(num_of_state, signal, path)`` stays a plain picklable record.  The bookkeeping the
reference does with these objects in Python loops (bucket frames by state, count transitions)
is done on the device by ``loe_align_dev`` / ``loe_kmeans_dev``; the host methods below are
kept for API compatibility and for callers that build ``Signal`` lists by hand.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np
from numpy.typing import NDArray

from .transition_probability import TransitionProbabilities


@dataclass
class Signal:
    num_of_state: int
    signal: NDArray[np.float32]
    path: NDArray[np.int8]

    @property
    def order_by_state(self) -> List[NDArray | None]:
        """Frames of each state: the contiguous run of state 0, then of state 1, ... (a state
        that is skipped yields None; frames after a decrease are dropped)."""
        path = np.asarray(self.path)
        out: List[NDArray | None] = []
        start = 0
        for s in range(self.num_of_state):
            end = start
            while end < len(path) and path[end] == s:
                end += 1
            out.append(self.signal[start:end] if end > start else None)
            start = end
        return out

    @property
    def order_by_signal(self) -> List[Tuple[NDArray, int]]:
        return list(zip(self.signal, self.path))


@dataclass
class SortedSignals:
    num_of_states: int
    _signals: List[Signal] = field(init=False)

    def __post_init__(self) -> None:
        self._signals = []

    def append(self, signal: Signal) -> None:
        self._signals.append(signal)

    @property
    def order_by_state(self) -> List[List[NDArray]]:
        out: List[List[NDArray]] = [[] for _ in range(self.num_of_states)]
        for sig in self._signals:
            for s, seg in enumerate(sig.order_by_state):
                if seg is not None:
                    out[s].append(seg)
        return out

    @property
    def transition_counts(self) -> NDArray[np.int32]:
        counts = np.zeros((self.num_of_states, self.num_of_states), dtype=np.int32)
        for sig in self._signals:
            p = np.asarray(sig.path).astype(np.int64)
            np.add.at(counts, (p[:-1], p[1:]), 1)
        return counts

    @property
    def transition_probabilities(self) -> TransitionProbabilities:
        counts = self.transition_counts
        with np.errstate(all="ignore"):
            probs = (counts / np.sum(counts, axis=1, keepdims=True)).astype(np.float32)
        return TransitionProbabilities.from_transition_probability(probs)
