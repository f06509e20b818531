"""Dict-backed transition matrices with the reference's on-disk identity.

Pickled models (``log_trans_probs.pickle``, hidden_markov_model.py:93-115) store an instance
of ``loe_speech_recognition.transition_probability.LogTransitionProbabilities`` with the
attributes ``num_of_states`` and ``_core`` (dict keyed by (from, to)), so these classes keep
that module path, those attribute names and the reference's read semantics
(transition_probability.py:11-82): an absent key reads back as 0.0, zero probabilities are
stored and become -inf after the log.  The kernels never touch these objects; they are
flattened once per model by ``to_dense`` (see _trellis.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Self, Tuple

import numpy as np
from numpy.typing import NDArray


@dataclass
class SparseMatrix:
    num_of_states: int = field(default=0)
    _core: Dict[Tuple[int, ...], float] = field(default_factory=dict)

    def __getitem__(self, key: Tuple[int, ...]) -> float:
        return self._core.get(tuple(key), 0.0)

    def __setitem__(self, key: Tuple[int, ...], value: float) -> None:
        self._core[tuple(key)] = value

    def __str__(self) -> str:
        return f"{self._core}"

    def to_dense(self, missing: float = 0.0) -> NDArray[np.float32]:
        """Dense [S, S] float32 copy; absent keys -> ``missing`` (0.0, like __getitem__)."""
        out = np.full((self.num_of_states, self.num_of_states), np.float32(missing), dtype=np.float32)
        for (i, j), v in self._core.items():
            out[i, j] = v
        return out

    @classmethod
    def from_dense(cls, dense: NDArray) -> Self:
        m = cls(int(dense.shape[0]))
        for i, row in enumerate(dense):
            for j, v in enumerate(row):
                m._core[(i, j)] = v
        return m


@dataclass
class TransitionProbabilities(SparseMatrix):

    @classmethod
    def from_num_of_states(cls, num_of_states: int) -> Self:
        """Initial guess: state i spreads its mass uniformly over itself and every later state."""
        dense = np.zeros((num_of_states, num_of_states), dtype=np.float32)
        for i in range(num_of_states):
            dense[i, i:] = 1 / (num_of_states - i)
        return cls.from_transition_probability(dense)

    @classmethod
    def from_transition_probability(cls, transition_probability: NDArray[np.float32]) -> Self:
        return cls.from_dense(transition_probability)


@dataclass
class LogTransitionProbabilities(SparseMatrix):

    def append(self, ltp: Self) -> None:
        """Block-diagonal concatenation (cross-block keys stay absent, i.e. read as 0.0)."""
        base = self.num_of_states
        self.num_of_states += ltp.num_of_states
        for (i, j), v in ltp._core.items():
            self._core[(i + base, j + base)] = v

    @classmethod
    def from_transition_probability(cls, tp: TransitionProbabilities) -> Self:
        out = cls(tp.num_of_states)
        with np.errstate(divide="ignore", invalid="ignore"):
            for key, v in tp._core.items():
                out._core[key] = np.log(v)
        return out
