"""Word <-> state-range bookkeeping of concatenated word models (reference:
model_boundary.py:11-179).  Host-side only; the kernels receive the same information as the
flat ``flags`` / ``word`` / ``word_lo`` tables (_trellis.py)."""
from __future__ import annotations

import bisect
from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np
from numpy.typing import NDArray


@dataclass
class ModelBoundary:
    _boundaries: List[int] = field(default_factory=list)   # cumulative state counts
    _labels: List[str] = field(default_factory=list)
    _isFrozen: bool = field(default=False)

    @property
    def lower_boundaries(self) -> List[int]:
        self._isFrozen = True
        return [0] + self._boundaries[:-1]

    @property
    def upper_boundaries(self) -> List[int]:
        self._isFrozen = True
        return [b - 1 for b in self._boundaries]

    @property
    def sizes(self) -> List[int]:
        return [b - a for a, b in zip([0] + self._boundaries[:-1], self._boundaries)]

    @property
    def num_of_words(self) -> int:
        return len(self._boundaries)

    def append(self, num_of_states: int) -> None:
        if self._isFrozen:
            raise Exception("ModelBoundary modified after its boundaries were read")
        self._boundaries.append((self._boundaries[-1] if self._boundaries else 0) + num_of_states)

    def _word_index(self, state: int) -> int:
        if not self._boundaries or state < 0 or state >= self._boundaries[-1]:
            raise Exception(f"state {state} outside every word")   # bare Exception, like the reference
        return bisect.bisect_right(self._boundaries, state)

    def find_lower_boundary(self, state: int) -> int:
        if state < 0:
            raise Exception(f"Failed to find lower boundary for state {state}")
        w = min(bisect.bisect_right(self._boundaries, state), len(self._boundaries) - 1)
        return self.lower_boundaries[w]

    def find_upper_boundary(self, state: int) -> int:
        for ub in self.upper_boundaries:
            if state <= ub:
                return ub
        raise Exception(f"Failed to find upper boundary for state {state}")

    def add_model_labels(self, model_labels: List[str]) -> None:
        assert len(model_labels) == self.num_of_words
        self._labels = model_labels

    def get_label(self, state: int) -> str:
        return self._labels[self.lower_boundaries.index(self.find_lower_boundary(state))]

    def get_state_range(self, label: str) -> Tuple[int, int]:
        i = self._labels.index(label)
        return ((0 if i == 0 else self._boundaries[i - 1]), self._boundaries[i])

    def append_to_labels(self, state: int, skip_silence: bool, labels: List[str]) -> None:
        lab = self.get_label(state)
        if not (lab == "S" and skip_silence):
            labels.append(lab)

    def get_labels(self, path: NDArray[np.int8], skip_silence: bool = True) -> List[str]:
        """Run-length compress the state path and read off the word sequence: a word is
        emitted when the path leaves the current word's state range, or re-enters the same
        word's first state from its last state (a repeated word).  "S" is dropped."""
        seq = np.asarray(path).tolist()
        comp = [seq[0]]
        for s in seq[1:]:
            if s != comp[-1]:
                comp.append(s)
        out: List[str] = []
        lo, hi = self.find_lower_boundary(comp[0]), self.find_upper_boundary(comp[0])
        self.append_to_labels(comp[0], skip_silence, out)
        for prev, cur in zip(comp[:-1], comp[1:]):
            if cur < lo or cur > hi:
                lo, hi = self.find_lower_boundary(cur), self.find_upper_boundary(cur)
                self.append_to_labels(cur, skip_silence, out)
            elif prev == hi and cur == lo:
                self.append_to_labels(cur, skip_silence, out)
        return out
