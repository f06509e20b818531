"""Isolated-word classifier (reference: model_collection.py:14-40): the label of the word
model with the best Viterbi score, first label in TI_DIGITS_LABELS order on ties.

The reference loops over the 11 word models in Python; here the 11 models sit side by side in
one trellis ("multi", _trellis.py) so a batch of utterances is classified with one emission
launch and one Viterbi launch."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Self, Sequence

import itertools
import os

import numpy as np
from numpy.typing import NDArray

from . import _trellis
from .hidden_markov_model import HiddenMarkovModel, _Fingerprint, _PackCache
from .ti_digits import TI_DIGITS_LABELS


@dataclass
class ModelCollection(_PackCache):
    num_of_states: int = field(default=5)
    dim_of_feature: int = field(default=39)
    _models: List[HiddenMarkovModel] = field(default_factory=list)

    def _packs(self):
        def build():
            from ._engine import get_engine
            eng = get_engine()
            normals = [mn for m in self._models for mn in m._multivariate_normals]
            sizes = [len(m._multivariate_normals) for m in self._models]
            cols = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(int).tolist()
            tr = _trellis.build([m._log_transition_probs.to_dense() for m in self._models], cols,
                                list(range(len(sizes))), "multi")
            return eng.pack_gaussians(normals), eng.pack_trellises([tr])
        # object ids + content (in-place edits of a word model are seen, like HiddenMarkovModel._packs): ONE fingerprint
        # over the Gaussians of all word models, the transition tables hashed per model
        normals = [mn for m in self._models for mn in m._multivariate_normals]
        fp = self.__dict__.get("_fingerprint")
        if fp is None or not fp.same_objects(normals):
            fp = self.__dict__["_fingerprint"] = _Fingerprint(normals)
        tables = [m._log_transition_probs._core for m in self._models]
        values = np.fromiter(itertools.chain.from_iterable(t.values() for t in tables), dtype=np.float64)
        key = (fp.value(), hash(values.tobytes()), tuple(hash(tuple(t)) for t in tables),
               tuple((id(m), id(m._multivariate_normals), id(m._log_transition_probs)) for m in self._models))
        return self._cached(key, build)

    def scores_batch(self, signals: Sequence[NDArray[np.float32]], precision: Optional[str] = None) -> NDArray[np.float32]:
        """[n_utt, n_models] Viterbi scores (column order = self._models)."""
        from ._engine import get_engine
        eng = get_engine()
        gp, tp = self._packs()
        batch = eng.upload_features(signals, gp.dim)
        scores = eng.emission(batch.feat, gp, precision)
        _, end_scores, _, _ = eng.viterbi(scores, batch.frm_off, batch.n_utt, batch.max_frames, batch.total_frames, tp)
        return end_scores.cpu().numpy()

    def predict_batch(self, signals: Sequence[NDArray[np.float32]], precision: Optional[str] = None) -> List[str]:
        sc = self.scores_batch(signals, precision)
        # sorted(..., reverse=True) is stable: the first model wins ties == argmax (first maximum)
        return [str(self._models[int(i)]) for i in np.argmax(sc, axis=1)]

    def predict(self, signal: NDArray[np.float32]) -> str:
        return self.predict_batch([signal])[0]

    @classmethod
    def load_from_files(cls, folder_path: str) -> Self:
        mc = cls()
        for label in TI_DIGITS_LABELS:
            mc._models.append(HiddenMarkovModel.from_folder(os.path.join(folder_path, f"{label}")))
        return mc
