"""HMM decoder and trainers -- the reference's public classes, backed by the sm_100a kernels.

Reference: src/loe_speech_recognition/hidden_markov_model.py (all line numbers below).
Same class names, constructor / classmethod signatures, attribute names, exceptions and
on-disk layout (``<root>/<label>/{log_trans_probs,multivariate_normals}.pickle``); the
per-(frame, state) Python loops are replaced by three launches:

  loe_emission_dev   all Gaussian log-densities of a batch        (:46-48)
  loe_viterbi_dev    trellis walk + backtrace, one CTA/utterance  (:160-208, :481-581, :591)
  loe_align_dev + loe_kmeans_dev   sufficient statistics of the M-step (:320-350, :602-636)

Added (does not alter existing signatures): ``predict_batch`` on the decoders -- the
reference API is one utterance per call, which cannot fill a GPU.

The scipy frozen distributions stay the model's persistent form (the pickles embed them); the
device copy ("pack") is rebuilt lazily whenever the Python-side model objects change.
"""
from __future__ import annotations

import logging
import os
import pickle
import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Self, Sequence, Tuple

import numpy as np
import scipy as sp
import scipy.stats
from numpy.typing import NDArray
from tqdm import tqdm

from . import _trellis
from .model_boundary import ModelBoundary
from .signal import Signal, SortedSignals
from .transition_probability import LogTransitionProbabilities, TransitionProbabilities

logger = logging.getLogger(__name__)


def _engine():
    from ._engine import get_engine
    return get_engine()


@dataclass
class MultivariateNormal:
    """Full-covariance Gaussian; persistent form = scipy frozen distribution (:20-48)."""
    dim_of_features: int = field(init=False)
    _core: "sp.stats._multivariate.multivariate_normal_frozen" = field(init=False)

    @classmethod
    def from_means_covariances(cls, mean: NDArray[np.float32], covariance: NDArray[np.float32]) -> Self:
        mn = cls()
        # raises LinAlgError (singular) / ValueError (non-finite, not PSD) exactly like the reference
        mn._core = sp.stats.multivariate_normal(mean=mean, cov=covariance, allow_singular=False)
        mn.dim_of_features = mean.shape[0]
        return mn

    def log_pdf(self, x: NDArray) -> float:
        """Single-frame log-density through the float64 emission kernel, rounded to np.float32 like :46-48
        (scipy evaluates in float64; one frame costs nothing, so the exact mode is used here)."""
        assert x.shape[0] == self.dim_of_features
        eng = _engine()
        gp = eng.pack_gaussians([self])
        feat = eng._to_dev(np.asarray(x, dtype=np.float32)[None, :])
        return eng.emission(feat, gp, "fp64").cpu().numpy()[0, 0]


# ----------------------------------------------------------------------------------------
# device packs, cached on the Python objects but never pickled
# ----------------------------------------------------------------------------------------
class _PackCache:
    """Mixin: drops device handles when pickled (ProcessPoolExecutor ships models to workers)."""

    _PACK_ATTRS = ("_pack_cache", "_native_decoder", "_fingerprint")

    def __getstate__(self):
        state = dict(self.__dict__)
        for k in self._PACK_ATTRS:
            state.pop(k, None)
        return state

    def _cached(self, key, builder, slot: str = "_pack_cache"):
        cache = self.__dict__.get(slot)
        if cache is None or cache[0] != key or cache[2] != os.getpid():
            cache = (key, builder(), os.getpid())
            self.__dict__[slot] = cache
        return cache[1]


class _Fingerprint:
    """Content fingerprint of a model's Gaussians for the device-pack cache, cheap enough to take before every
    single-utterance call (the reference re-reads its objects on every predict, so an in-place edit must be seen).
    The arrays of the scipy objects are referenced here (kept alive: their addresses stay valid) together with a pointer
    table, and one C call (loe_host_fingerprint, host code) hashes them all; per call the Python side only checks that
    every object still holds the SAME array.  Arrays that are not contiguous in memory are hashed through zlib."""

    __slots__ = ("arrays", "ptrs", "sizes", "odd")

    def __init__(self, normals):
        import ctypes
        self.arrays = [a for n in normals for a in (n._core.mean, n._core.cov_object._LP)]
        flat = [a for a in self.arrays if a.flags.forc]           # scipy's whitening matrices are Fortran-ordered
        self.odd = [a for a in self.arrays if not a.flags.forc]
        self.ptrs = (ctypes.c_void_p * max(len(flat), 1))(*[a.ctypes.data for a in flat])
        self.sizes = (ctypes.c_int64 * max(len(flat), 1))(*[a.nbytes for a in flat])

    def same_objects(self, normals) -> bool:
        arrays = self.arrays
        if 2 * len(normals) != len(arrays):
            return False
        i = 0
        for n in normals:
            core = n._core
            if core.mean is not arrays[i] or core.cov_object._LP is not arrays[i + 1]:
                return False
            i += 2
        return True

    def value(self) -> int:
        from . import _native
        n = len(self.arrays) - len(self.odd)
        h = int(_native.load().loe_host_fingerprint(self.ptrs, self.sizes, n))
        for a in self.odd:
            h = zlib.crc32(np.ascontiguousarray(a).view(np.uint8), h & 0xFFFFFFFF) | (h & ~0xFFFFFFFF)
        return h


def _transitions_hash(ltp):
    """(hash of the keys, hash of the values) of a dict-backed transition table."""
    core = ltp._core
    if not core:
        return (0, 0)
    return (hash(tuple(core)), hash(np.fromiter(core.values(), dtype=np.float64, count=len(core)).tobytes()))


def _model_key(normals, ltp, holder=None):
    """Identity of a model for the device-pack cache: object ids plus a content fingerprint (of the means, the
    whitening matrices and the transition table), so that in-place edits and recycled addresses are seen.  ``holder``
    (the model's __dict__) keeps the _Fingerprint between calls."""
    fp = holder.get("_fingerprint") if holder is not None else None
    if fp is None or not fp.same_objects(normals):
        fp = _Fingerprint(normals)
        if holder is not None:
            holder["_fingerprint"] = fp
    return (id(normals), len(normals), id(ltp), len(ltp._core), id(ltp._core), fp.value()) + _transitions_hash(ltp)


@dataclass
class HiddenMarkovModel(_PackCache):
    label: str
    isMultiProcessing: bool = field(default=True)
    isTqdm: bool = field(default=True)
    _multivariate_normals: List[MultivariateNormal] = field(default_factory=list)
    _log_transition_probs: LogTransitionProbabilities = field(default_factory=LogTransitionProbabilities)

    def __str__(self) -> str:
        return self.label

    @property
    def num_of_states(self) -> int:
        return len(self._multivariate_normals)

    @property
    def dim_of_features(self) -> int:
        return self._multivariate_normals[0].dim_of_features

    # -- device side ---------------------------------------------------------------------
    def _trellis_kind(self) -> str:
        return "word"

    def _host_trellis(self) -> _trellis.HostTrellis:
        return _trellis.build([self._log_transition_probs.to_dense()], [0], [0], "word")

    def _packs(self):
        def build():
            eng = _engine()
            return eng.pack_gaussians(self._multivariate_normals), eng.pack_trellises([self._host_trellis()])
        return self._cached(_model_key(self._multivariate_normals, self._log_transition_probs, self.__dict__), build)

    def predict(self, signal: NDArray[np.float32]) -> Tuple[float, NDArray[np.int8]]:
        assert len(self._multivariate_normals) > 0
        assert self.dim_of_features == signal.shape[1]
        return self._viterbi(signal)

    def _viterbi(self, observation_sequence: NDArray[np.float32]) -> Tuple[float, NDArray[np.int8]]:
        scores, paths = self.predict_batch([observation_sequence])
        return scores[0], paths[0]

    def predict_batch(self, signals: Sequence[NDArray[np.float32]], precision: Optional[str] = None
                      ) -> Tuple[NDArray[np.float32], List[NDArray[np.int8]]]:
        """Viterbi scores and state paths of many utterances in one pass (added entry point)."""
        eng = _engine()
        gp, tp = self._packs()
        batch = eng.upload_features(signals, self.dim_of_features)
        scores = eng.emission(batch.feat, gp, precision)
        path, _, _, best_score = eng.viterbi(scores, batch.frm_off, batch.n_utt, batch.max_frames, batch.total_frames, tp,
                                             want_end_scores=False)
        path_h = path.cpu().numpy()
        off = batch.frm_off_host
        return best_score.cpu().numpy(), [path_h[off[i]:off[i + 1]] for i in range(batch.n_utt)]

    # -- persistence (:93-158) ------------------------------------------------------------
    def save(self, parent_folder_path: str = "./cache") -> None:
        folder = os.path.join(parent_folder_path, f"{self.label}")
        os.makedirs(folder, exist_ok=True)
        with open(os.path.join(folder, "log_trans_probs.pickle"), "wb") as f:
            pickle.dump(self._log_transition_probs, f, pickle.HIGHEST_PROTOCOL)
        with open(os.path.join(folder, "multivariate_normals.pickle"), "wb") as f:
            pickle.dump(self._multivariate_normals, f, pickle.HIGHEST_PROTOCOL)
        logger.info(f"Finish saving all files for {self.label} model")

    @classmethod
    def from_folder(cls, model_folder_path: str) -> Self:
        if not os.path.isdir(model_folder_path):
            raise FileNotFoundError
        model = cls(cls._model_folder_name_parser(model_folder_path))
        with open(os.path.join(model_folder_path, "log_trans_probs.pickle"), "rb") as f:
            model._log_transition_probs = pickle.load(f)
        with open(os.path.join(model_folder_path, "multivariate_normals.pickle"), "rb") as f:
            model._multivariate_normals = pickle.load(f)
        return model

    @staticmethod
    def _model_folder_name_parser(folder_path: str) -> str:
        return str(folder_path.split("/")[-1])

    # -- flat cache format (added) -----------------------------------------
    # The reference's pickles embed scipy's private frozen-distribution classes, which couples a
    # saved model to the scipy version that wrote it.  ``model.npz`` beside them holds only plain
    # arrays (means, covariances, dense log-transitions); loading rebuilds the Gaussians with scipy.
    def save_flat(self, parent_folder_path: str = "./cache") -> str:
        folder = os.path.join(parent_folder_path, f"{self.label}")
        os.makedirs(folder, exist_ok=True)
        path = os.path.join(folder, "model.npz")
        np.savez(path, label=np.array(self.label),
                 means=np.stack([np.asarray(mn._core.mean, dtype=np.float64) for mn in self._multivariate_normals]),
                 covariances=np.stack([np.asarray(mn._core.cov_object.covariance, dtype=np.float64) for mn in self._multivariate_normals]),
                 log_transitions=self._log_transition_probs.to_dense())
        return path

    @classmethod
    def from_flat(cls, model_folder_path: str) -> Self:
        path = os.path.join(model_folder_path, "model.npz")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        z = np.load(path, allow_pickle=False)
        model = cls(str(z["label"]))
        model._multivariate_normals = [MultivariateNormal.from_means_covariances(m, c) for m, c in zip(z["means"], z["covariances"])]
        model._log_transition_probs = LogTransitionProbabilities.from_dense(z["log_transitions"].astype(np.float32))
        return model


# ----------------------------------------------------------------------------------------
# M-step from device statistics (shared by the isolated and the embedded trainer)
# ----------------------------------------------------------------------------------------
def _unpack_stats(stats: np.ndarray, dim: int):
    """stats [S, 1+D+D(D+1)/2] -> (N [S], s1 [S,D], S2 [S,D,D] symmetric)."""
    S = stats.shape[0]
    n = stats[:, 0]
    s1 = stats[:, 1:1 + dim]
    iu = np.triu_indices(dim)
    s2 = np.zeros((S, dim, dim), dtype=np.float64)
    s2[:, iu[0], iu[1]] = stats[:, 1 + dim:]
    s2[:, iu[1], iu[0]] = stats[:, 1 + dim:]
    return n, s1, s2


def _batched_eigh(cov: np.ndarray):
    """float64 symmetric eigendecomposition of [n, D, D] matrices (lower triangle, like scipy's _PSD); LAPACK
    releases the GIL, so the batch is spread over a few host threads."""
    n = cov.shape[0]
    if n < 16:
        return np.linalg.eigh(cov)
    import concurrent.futures as cf
    k = min(8, os.cpu_count() or 1, n // 8)
    parts = np.array_split(np.arange(n), k)
    with cf.ThreadPoolExecutor(max_workers=k) as ex:
        res = list(ex.map(lambda idx: np.linalg.eigh(cov[idx]), parts))
    return np.concatenate([r[0] for r in res]), np.concatenate([r[1] for r in res])


@dataclass
class HiddenMarkovModelTrainable(HiddenMarkovModel):
    class HMMTrainMeanFail(Exception):
        def __init__(self, *args: object) -> None:
            super().__init__(*args)
            logger.warning("Cannot use all the state for the HMM")

    class HMMTrainConverge(Exception):
        def __init__(self, *args: object) -> None:
            super().__init__(*args)
            logger.info("Successfully train the HMM model")

    _means: NDArray[np.float32] = field(init=False)
    _covariances: NDArray[np.float32] = field(init=False)
    _transition_probs: TransitionProbabilities = field(init=False)

    @property
    def num_of_states(self) -> int:
        return self._means.shape[0]

    @classmethod
    def from_data(cls, label: str, mfccs: List[NDArray[np.float32]], num_of_states: int = 5,
                  max_iterations: int = 100, isMultiProcessingTraining: bool = True, isTqdm: bool = True) -> Self:
        """Segmental K-means for one word (:233-281).  ``isMultiProcessingTraining`` is accepted for
        compatibility; the E-step is one batched device pass either way.  Under an initialised
        torch.distributed group the utterances are sharded by rank and the statistics summed with
        one all-reduce per iteration (every rank returns the same model)."""
        from . import _dist

        model = cls(label, isMultiProcessing=isMultiProcessingTraining, isTqdm=isTqdm)
        model._means, model._covariances, model._transition_probs = model._init_parameters(mfccs[0], num_of_states)
        model._update_inference_weights()
        eng = _engine()
        shard = _dist.shard(list(mfccs))
        batch = eng.upload_features(shard, mfccs[0].shape[1]) if len(shard) else None
        bar = tqdm(desc=f"Train {model.label} model", total=max_iterations, position=1, disable=not model.isTqdm)
        for it in range(max_iterations):
            try:
                model._train_device(batch)
                model._update_inference_weights()
            except cls.HMMTrainMeanFail:
                logger.error("Failed to train model")
                raise
            except cls.HMMTrainConverge:
                logger.info(f"Finish training model {str(model)} after {it} iterations")
                break
            bar.update()
        bar.close()
        model._update_inference_weights()
        return model

    @classmethod
    def from_data_batch(cls, labeled_mfccs: Dict[str, List[NDArray[np.float32]]], num_of_states=5,
                        max_iterations: int = 100, device_mstep: Optional[bool] = None, return_info: bool = False):
        """Train several word models at once (added entry point).  The result equals calling
        :meth:`from_data` once per label (each word keeps its own convergence test and stops on its own),
        but every iteration is ONE device pass over all words and nothing but a status word per model
        leaves the device: per-word 3xFP16 tensor-core emission launches into one score matrix, one
        Viterbi launch with a trellis per word, one statistics pass for all states, -- under
        torch.distributed -- one all-reduce, and the M-step itself (``loe_mstep_dev``: means, the
        reference's means-only convergence test, covariances, transition probabilities, and the
        whitening data written straight into the tensor-core image).  The scipy objects of the
        persistent format are built once at the end from the float32 parameters, which are bit-identical
        to the host M-step's.  ``device_mstep=False`` (or LOE_B200_HOST_MSTEP=1) keeps the round-1 host M-step."""
        if device_mstep is None:
            device_mstep = not os.environ.get("LOE_B200_HOST_MSTEP")
        D = int(next(iter(labeled_mfccs.values()))[0].shape[1])
        if device_mstep and D == 39:
            return cls._from_data_batch_device(labeled_mfccs, num_of_states, max_iterations, return_info)
        models = cls._from_data_batch_host(labeled_mfccs, num_of_states, max_iterations)
        return (models, {"iterations": None, "mstep": "host"}) if return_info else models

    @classmethod
    def _from_data_batch_device(cls, labeled_mfccs, num_of_states, max_iterations: int, return_info: bool):
        from . import _dist, _native
        from ._engine import pack_h16_image

        eng = _engine()
        torch = eng.torch
        labels = list(labeled_mfccs)
        n_st = {l: int(num_of_states[l] if isinstance(num_of_states, dict) else num_of_states) for l in labels}
        W, D = len(labels), 39
        sizes = np.array([n_st[l] for l in labels], dtype=np.int32)
        first = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.int32)
        G = int(sizes.sum())
        tiles = (sizes + 5) // 6
        tile0 = np.concatenate(([0], np.cumsum(tiles)[:-1])).astype(np.int32)
        n_tiles = int(tiles.sum())
        tile_halves = eng.lib.loe_emission_h16_tile_bytes() // 2

        # initial parameters exactly as from_data: uniform segmentation of each word's FIRST utterance (:359-389)
        means0 = np.zeros((G, D), np.float32)
        cov0 = np.zeros((G, D, D), np.float32)
        band0 = np.full((G, 3), -np.inf, np.float32)
        trans0 = {}
        for i, l in enumerate(labels):
            m, c, t = cls._init_parameters(labeled_mfccs[l][0], n_st[l])
            means0[first[i]:first[i] + sizes[i]] = m
            cov0[first[i]:first[i] + sizes[i]] = c
            trans0[l] = t
            with np.errstate(divide="ignore", invalid="ignore"):
                band0[first[i]:first[i] + sizes[i]] = _trellis.build([np.log(t.to_dense())], [0], [0], "word").band

        # this rank's utterances, word by word (the frames of a word are contiguous on the device)
        shards = {l: _dist.shard(list(labeled_mfccs[l])) for l in labels}
        feats = [x for l in labels for x in shards[l]]
        utt_word = np.array([i for i, l in enumerate(labels) for _ in shards[l]], dtype=np.int32)
        batch = eng.upload_features(feats, D) if feats else None
        utt_tr = eng._to_dev(utt_word) if feats else None
        frame_range = {}
        if feats:
            cnt = np.cumsum([0] + [len(shards[l]) for l in labels])
            for i, l in enumerate(labels):
                frame_range[l] = (int(batch.frm_off_host[cnt[i]]), int(batch.frm_off_host[cnt[i + 1]]))
            scores = eng.empty((batch.total_frames, G), torch.float32)
        stride = 1 + D + D * (D + 1) // 2

        # device-resident model: float32 means / covariances, trellis bands, 3xFP16 image, active set
        def host_image(idx):
            """3xFP16 image tiles + constants of word ``idx`` from its float32 covariances, scipy's way (eigh + _PSD checks)."""
            a, n = int(first[idx]), int(sizes[idx])
            cov = cov_h[a:a + n].astype(np.float64)
            if not np.all(np.isfinite(cov)):
                raise ValueError("array must not contain infs or NaNs")
            lam, vec = np.linalg.eigh(cov)
            eps = 1e6 * np.finfo(np.float64).eps * np.max(np.abs(lam), axis=1)
            if np.any(lam.min(axis=1) < -eps):
                raise ValueError("The input matrix must be symmetric positive semidefinite.")
            if np.any(lam <= eps[:, None]):
                raise np.linalg.LinAlgError("When `allow_singular is False`, the input matrix must be symmetric positive definite.")
            U = vec * np.sqrt(1.0 / lam)[:, None, :]
            cst = -0.5 * (D * np.log(2 * np.pi) + np.sum(np.log(lam), axis=1))
            img = pack_h16_image(means_h[a:a + n].astype(np.float64), U, cst)
            if img is None:
                raise NotImplementedError("whitening matrix outside the binary16 range: train this model with from_data(…) under "
                                          "LOE_B200_EMISSION=tc / LOE_B200_HOST_MSTEP=1")
            cp = np.zeros(int(tiles[idx]) * 6, np.float32)
            cp[:n] = cst.astype(np.float32)
            return img, cp

        means_h, cov_h = means0, cov0
        img0 = np.zeros(n_tiles * tile_halves, np.float16)
        cst0 = np.zeros(n_tiles * 6, np.float32)
        for i in range(W):
            im, cp = host_image(i)
            img0[tile0[i] * tile_halves:(tile0[i] + tiles[i]) * tile_halves] = im
            cst0[tile0[i] * 6:(tile0[i] + tiles[i]) * 6] = cp
        means_d, cov_d, b_h16, cst_pad = eng._to_dev(means0), eng._to_dev(cov0), eng._to_dev(img0), eng._to_dev(cst0)
        trellises = [_trellis.build([np.zeros((int(n), int(n)), np.float32)], [int(a)], [i], "word") for i, (a, n) in enumerate(zip(first, sizes))]
        tp = eng.pack_trellises(trellises)
        tp.band.copy_(eng._to_dev(band0))
        state_word = eng._to_dev(np.repeat(np.arange(W, dtype=np.int32), sizes))
        word_first, word_n, word_tile = eng._to_dev(first), eng._to_dev(sizes), eng._to_dev(tile0)
        active = torch.ones(W, dtype=torch.int32, device=eng.device)
        updated = torch.zeros(W, dtype=torch.int32, device=eng.device)
        status = torch.zeros(W, dtype=torch.int32, device=eng.device)
        counts_applied = torch.zeros((G, G), dtype=torch.int32, device=eng.device)
        applied_any = torch.zeros(W, dtype=torch.int32, device=eng.device)
        multi_ok = bool(feats) and int(sizes.max()) <= 12 and not os.environ.get("LOE_B200_EMISSION_PER_WORD")
        if multi_ok:
            seg_begin = eng._to_dev(np.array([frame_range[l][0] for l in labels], dtype=np.int64))
            seg_end = eng._to_dev(np.array([frame_range[l][1] for l in labels], dtype=np.int64))
            # the features never change: their pre-split tensor-core operand is built once, per word on tile boundaries
            seg_tiles = np.array([(frame_range[l][1] - frame_range[l][0] + 127) // 128 for l in labels], dtype=np.int64)
            n_img_tiles = int(seg_tiles.sum())
            seg_img_tile = eng._to_dev(np.concatenate(([0], np.cumsum(seg_tiles)[:-1])).astype(np.int32))
            a_img = eng.empty((max(n_img_tiles, 1) * 20480,), torch.uint8)
            inv2 = eng.empty((max(n_img_tiles, 1) * 128,), torch.float32)
            _native.check(eng.lib.loe_h16_image_dev(batch.feat.data_ptr(), D, W, seg_begin.data_ptr(), seg_end.data_ptr(),
                                                    seg_img_tile.data_ptr(), n_img_tiles, a_img.data_ptr(), inv2.data_ptr(), eng._stream()))
            eng.launches += 1
        status_host = [torch.empty(W, dtype=torch.int32).pin_memory() for _ in range(2)]
        events = [None, None]
        maybe_active = np.ones(W, dtype=bool)          # host view of the active set, one iteration behind the device
        iterations = 0
        stream = torch.cuda.current_stream(eng.device)

        def digest(slot):
            """Host reaction to the status words of an iteration (read one iteration late: the device froze converged
            words itself, the E-step of a word that was already done is wasted work, never a different result)."""
            events[slot].synchronize()
            st = status_host[slot].numpy()
            if np.any(st & _native.LOE_MSTEP_MEAN_FAIL):
                raise cls.HMMTrainMeanFail
            bad = np.nonzero(st & _native.LOE_MSTEP_SUSPECT)[0]
            if bad.size:
                nonlocal means_h, cov_h
                means_h, cov_h = means_d.cpu().numpy(), cov_d.cpu().numpy()
                for i in bad.tolist():                 # scipy's own verdict: raises what the reference raises, or repairs the image
                    im, cp = host_image(i)             # (the device parked the word: its parameters are those of the flagged update)
                    b_h16[int(tile0[i]) * tile_halves:int(tile0[i] + tiles[i]) * tile_halves].copy_(eng._to_dev(im))
                    cst_pad[int(tile0[i]) * 6:int(tile0[i] + tiles[i]) * 6].copy_(eng._to_dev(cp))
                    active[i] = 1
            maybe_active[(st & (_native.LOE_MSTEP_CONVERGED)) != 0] = False

        marks = []                                     # CUDA events between the phases of every iteration (return_info only)

        def mark():
            if return_info:
                e = torch.cuda.Event(enable_timing=True)
                e.record(stream)
                marks.append(e)

        for it in range(max_iterations):
            if not maybe_active.any():
                break
            iterations += 1
            mark()
            if feats:
                if multi_ok:            # all word models in one launch; converged ones are skipped on the device
                    _native.check(eng.lib.loe_emission_h16_multi_img_dev(a_img.data_ptr(), inv2.data_ptr(), b_h16.data_ptr(), cst_pad.data_ptr(), W,
                                                                         seg_begin.data_ptr(), seg_end.data_ptr(), seg_img_tile.data_ptr(),
                                                                         word_tile.data_ptr(), word_n.data_ptr(), word_first.data_ptr(),
                                                                         active.data_ptr(), int(sizes.max()), scores.data_ptr(), G, eng._stream()))
                    eng.launches += 1
                else:
                    for i, l in enumerate(labels):
                        if maybe_active[i]:
                            a, b = frame_range[l]
                            eng.emission_h16_into(batch.feat[a:b], b_h16, cst_pad, int(tile0[i]), int(sizes[i]), scores[a:b], int(first[i]))
                mark()
                path, _, _, _ = eng.viterbi(scores, batch.frm_off, batch.n_utt, batch.max_frames, batch.total_frames, tp,
                                            utt_tr=utt_tr, want_end_scores=False)
                mark()
                stats, counts, _ = eng.kmeans_stats(batch.feat, path, batch.frm_off, batch.n_utt, batch.total_frames, tp,
                                                    utt_tr, False, G, means_d)
                mark()
            else:
                stats = torch.zeros((G, stride), dtype=torch.float64, device=eng.device)
                counts = torch.zeros((G, G), dtype=torch.int32, device=eng.device)
                mark(); mark(); mark()
            stats, counts = _dist.allreduce_stats(stats, counts)
            mark()
            _native.check(eng.lib.loe_mstep_dev(stats.data_ptr(), counts.data_ptr(), G, W, state_word.data_ptr(), word_first.data_ptr(),
                                                word_n.data_ptr(), word_tile.data_ptr(), means_d.data_ptr(), cov_d.data_ptr(),
                                                counts_applied.data_ptr(), tp.band.data_ptr(), b_h16.data_ptr(), cst_pad.data_ptr(),
                                                active.data_ptr(), updated.data_ptr(), status.data_ptr(), D, eng._stream()))
            eng.launches += 2
            applied_any += updated
            mark()
            slot = it & 1
            status_host[slot].copy_(status, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            events[slot] = ev
            if it >= 1:
                digest((it - 1) & 1)
        if iterations:
            digest((iterations - 1) & 1)

        # persistent form: float32 parameters back to the host once, scipy objects built like _update_inference_weights
        means_h, cov_h = means_d.cpu().numpy(), cov_d.cpu().numpy()
        counts_h, applied_h = counts_applied.cpu().numpy().astype(np.int64), applied_any.cpu().numpy()
        models = {}
        for i, l in enumerate(labels):
            a, n = int(first[i]), int(sizes[i])
            m = cls(l, isTqdm=False)
            m._means, m._covariances = means_h[a:a + n].copy(), cov_h[a:a + n].copy()
            if applied_h[i]:
                c = counts_h[a:a + n, a:a + n]
                with np.errstate(all="ignore"):
                    probs = (c / np.sum(c, axis=1, keepdims=True)).astype(np.float32)
                m._transition_probs = TransitionProbabilities.from_transition_probability(probs)
            else:
                m._transition_probs = trans0[l]
            m._update_inference_weights()
            models[l] = m
        info = {"iterations": iterations, "mstep": "device", "n_states": G, "frames_this_rank": int(batch.total_frames) if feats else 0,
                "allreduce_bytes": 8 * (G * stride + G * G)}
        if return_info and iterations:
            torch.cuda.synchronize(eng.device)
            ph = np.array([[marks[6 * k + j].elapsed_time(marks[6 * k + j + 1]) for j in range(5)] for k in range(iterations)])
            total = np.array([marks[6 * k].elapsed_time(marks[6 * (k + 1)]) for k in range(iterations - 1)])
            info["phase_ms"] = {n: float(np.median(ph[:, j])) for j, n in enumerate(("emission", "viterbi", "align_stats", "allreduce", "mstep"))}
            info["phase_ms_first_iteration"] = {n: float(ph[0, j]) for j, n in enumerate(("emission", "viterbi", "align_stats", "allreduce", "mstep"))}
            # start-to-start time of consecutive iterations on the device: kernels + host gaps (status check, launches)
            info["ms_per_iteration"] = float(np.median(total)) if len(total) else float(ph[0].sum())
            info["ms_per_iteration_all"] = [float(v) for v in total]
        return (models, info) if return_info else models

    @classmethod
    def _from_data_batch_host(cls, labeled_mfccs: Dict[str, List[NDArray[np.float32]]], num_of_states=5,
                              max_iterations: int = 100) -> Dict[str, Self]:
        """Round-1 form of :meth:`from_data_batch`: device E-step, host M-step (batched float64 eigendecomposition, image
        packing and upload every iteration).  Kept as the cross-check of the device M-step and for dimensions other than 39."""
        from . import _dist
        from ._engine import Batch

        eng = _engine()
        torch = eng.torch
        labels = list(labeled_mfccs)
        n_st = {l: int(num_of_states[l] if isinstance(num_of_states, dict) else num_of_states) for l in labels}
        models = {}
        for l in labels:
            m = cls(l, isTqdm=False)
            m._means, m._covariances, m._transition_probs = m._init_parameters(labeled_mfccs[l][0], n_st[l])
            models[l] = m
        D = int(labeled_mfccs[labels[0]][0].shape[1])
        start = {}
        G = 0
        for l in labels:
            start[l] = G
            G += n_st[l]
        # this rank's utterances, word by word (the frames of a word are contiguous on the device)
        shards = {l: _dist.shard(list(labeled_mfccs[l])) for l in labels}
        feats = [x for l in labels for x in shards[l]]
        utt_word = np.array([i for i, l in enumerate(labels) for _ in shards[l]], dtype=np.int32)
        batch = eng.upload_features(feats, D) if feats else None
        utt_tr = eng._to_dev(utt_word) if feats else None
        frame_range = {}
        if feats:
            cnt = np.cumsum([0] + [len(shards[l]) for l in labels])
            for i, l in enumerate(labels):
                frame_range[l] = (int(batch.frm_off_host[cnt[i]]), int(batch.frm_off_host[cnt[i + 1]]))
            scores = eng.empty((batch.total_frames, G), torch.float32)
        stride = 1 + D + D * (D + 1) // 2
        active = set(labels)
        whiten = {}

        def refresh(which):
            """(U, cst) of the listed words from their covariances: scipy's _PSD in batched form."""
            cov = np.concatenate([models[l]._covariances.astype(np.float64) for l in which])
            if not np.all(np.isfinite(cov)):
                raise ValueError("array must not contain infs or NaNs")
            lam, vec = _batched_eigh(cov)
            eps = 1e6 * np.finfo(np.float64).eps * np.max(np.abs(lam), axis=1)
            if np.any(lam.min(axis=1) < -eps):
                raise ValueError("The input matrix must be symmetric positive semidefinite.")
            if np.any(lam <= eps[:, None]):
                raise np.linalg.LinAlgError("When `allow_singular is False`, the input matrix must be symmetric positive definite.")
            U = vec * np.sqrt(1.0 / lam)[:, None, :]
            cst = -0.5 * (D * np.log(2 * np.pi) + np.sum(np.log(lam), axis=1))
            o = 0
            for l in which:
                whiten[l] = (U[o:o + n_st[l]], cst[o:o + n_st[l]])
                o += n_st[l]

        refresh(labels)
        for it in range(max_iterations):
            if not active:
                break
            if feats:
                b_packed, cst_pad, first = eng.pack_tc_words([(models[l]._means.astype(np.float64), *whiten[l]) for l in labels])
                with np.errstate(divide="ignore", invalid="ignore"):       # log 0 = -inf is the reference's encoding
                    trellises = [_trellis.build([np.log(models[l]._transition_probs.to_dense())], [start[l]], [i], "word")
                                 for i, l in enumerate(labels)]
                tp = eng.pack_trellises(trellises)
                for i, l in enumerate(labels):
                    if l in active:
                        a, b = frame_range[l]
                        eng.emission_tc_into(batch.feat[a:b], b_packed, cst_pad, first[i], n_st[l], scores[a:b], start[l])
                path, _, _, _ = eng.viterbi(scores, batch.frm_off, batch.n_utt, batch.max_frames, batch.total_frames, tp,
                                            utt_tr=utt_tr, want_end_scores=False)
                shift = eng._to_dev(np.concatenate([models[l]._means for l in labels]).astype(np.float32))
                stats, counts, _ = eng.kmeans_stats(batch.feat, path, batch.frm_off, batch.n_utt, batch.total_frames, tp,
                                                    utt_tr, False, G, shift)
            else:
                stats = torch.zeros((G, stride), dtype=torch.float64, device=eng.device)
                counts = torch.zeros((G, G), dtype=torch.int32, device=eng.device)
            stats, counts = _dist.allreduce_stats(stats, counts)
            stats_h, counts_h = stats.cpu().numpy(), counts.cpu().numpy().astype(np.int64)
            updated = []
            for l in labels:
                if l not in active:
                    continue
                a, n = start[l], n_st[l]
                try:
                    models[l]._update_from_statistics(stats_h[a:a + n], counts_h[a:a + n, a:a + n],
                                                      shift=models[l]._means.astype(np.float64))
                    updated.append(l)
                except cls.HMMTrainConverge:
                    active.discard(l)
            if updated:
                refresh(updated)
        for l in labels:
            models[l]._update_inference_weights()
        return models

    def _update_inference_weights(self) -> None:
        self._log_transition_probs = LogTransitionProbabilities.from_transition_probability(self._transition_probs)
        self._multivariate_normals = self.get_multivariate_normals(self._means, self._covariances)

    # -- E-step on the device ---------------------------------------------------------------
    def _device_statistics(self, batch):
        """Align ``batch`` with the current model and return (stats [S, stride], counts [S,S]) as
        host float64 / int64 arrays, summed over ranks when torch.distributed is initialised."""
        from . import _dist

        eng = _engine()
        S, D = self._means.shape
        if batch is not None:
            gp, tp = self._packs()
            scores = eng.emission(batch.feat, gp)
            path, _, _, _ = eng.viterbi(scores, batch.frm_off, batch.n_utt, batch.max_frames, batch.total_frames, tp,
                                        want_end_scores=False)
            shift = eng._to_dev(np.ascontiguousarray(self._means, dtype=np.float32))
            stats, counts, _ = eng.kmeans_stats(batch.feat, path, batch.frm_off, batch.n_utt, batch.total_frames, tp,
                                                None, False, S, shift)
        else:
            stats = eng.torch.zeros((S, 1 + D + D * (D + 1) // 2), dtype=eng.torch.float64, device=eng.device)
            counts = eng.torch.zeros((S, S), dtype=eng.torch.int32, device=eng.device)
        stats, counts = _dist.allreduce_stats(stats, counts)
        return stats.cpu().numpy(), counts.cpu().numpy().astype(np.int64)

    def _train_device(self, batch) -> None:
        stats, counts = self._device_statistics(batch)
        self._update_from_statistics(stats, counts, shift=self._means.astype(np.float64))

    def _train(self, mfccs: List[NDArray[np.float32]]) -> None:
        """One iteration over a list of feature matrices (:294-318)."""
        self._train_device(_engine().upload_features(mfccs, mfccs[0].shape[1]))

    def _update_from_statistics(self, stats: np.ndarray, counts: np.ndarray, shift: np.ndarray) -> None:
        """M-step (:320-350) from sufficient statistics accumulated around ``shift``:
        means, the means-only convergence test BEFORE covariances / transitions are touched,
        np.cov-style covariance (N-1) + 1e-3 I, row-normalised transition counts."""
        S, D = shift.shape
        n, s1, s2 = _unpack_stats(stats, D)
        if np.any(n == 0):
            raise self.HMMTrainMeanFail
        new_means = (shift + s1 / n[:, None]).astype(np.float32)
        if np.allclose(new_means, self._means):
            raise self.HMMTrainConverge
        self._means = new_means
        with np.errstate(all="ignore"):
            centred = s2 - s1[:, :, None] * s1[:, None, :] / n[:, None, None]
            cov = centred / (n - 1.0)[:, None, None]
            self._covariances = (cov + np.eye(D) * 0.001).astype(np.float32)
            probs = (counts / np.sum(counts, axis=1, keepdims=True)).astype(np.float32)
        self._transition_probs = TransitionProbabilities.from_transition_probability(probs)

    def _update_middleware_parameters(self, sorted_signals: SortedSignals) -> None:
        """Reference entry point (:320-350) for hand-built SortedSignals: the frames are shipped to
        the device once and reduced by the statistics kernel."""
        eng = _engine()
        sigs = sorted_signals._signals
        S, D = self._means.shape
        if not sigs:
            raise self.HMMTrainMeanFail
        feats = [np.asarray(s.signal, dtype=np.float32) for s in sigs]
        batch = eng.upload_features(feats, D)
        path = eng._to_dev(np.concatenate([np.asarray(s.path, dtype=np.int8) for s in sigs]))
        tp = eng.pack_trellises([_trellis.build([np.zeros((S, S), np.float32)], [0], [0], "word")])
        shift = eng._to_dev(np.ascontiguousarray(self._means, dtype=np.float32))
        stats, counts, _ = eng.kmeans_stats(batch.feat, path, batch.frm_off, batch.n_utt, batch.total_frames, tp,
                                            None, False, S, shift)
        self._update_from_statistics(stats.cpu().numpy(), counts.cpu().numpy().astype(np.int64),
                                     shift=self._means.astype(np.float64))

    def _train_external(self, signals: List[Signal]) -> None:
        sorted_signals = SortedSignals(self.num_of_states)
        for s in signals:
            sorted_signals.append(s)
        self._update_middleware_parameters(sorted_signals)

    @classmethod
    def _init_parameters(cls, sample_signal: NDArray[np.float32], num_of_states: int):
        """Uniform segmentation of the FIRST utterance, covariance 0.01 I, uniform forward
        transitions (:359-389).  One utterance: host arithmetic."""
        D = sample_signal.shape[1]
        seg = int(sample_signal.shape[0] / num_of_states)
        means = np.array([np.average(sample_signal[i * seg:(i + 1) * seg, :], axis=0) for i in range(num_of_states)],
                         dtype=np.float32)
        return means, cls._init_covariance(D, num_of_states), TransitionProbabilities.from_num_of_states(num_of_states)

    @staticmethod
    def _init_covariance(dim_of_features: int, num_of_states: int) -> NDArray[np.float32]:
        return (np.tile(np.eye(dim_of_features), (num_of_states, 1, 1)) * 0.01).astype(np.float32)

    @staticmethod
    def get_multivariate_normals(means: NDArray[np.float32], covariances: NDArray[np.float32]) -> List[MultivariateNormal]:
        return [MultivariateNormal.from_means_covariances(mean=m, covariance=c) for m, c in zip(means, covariances)]


# ----------------------------------------------------------------------------------------
# digit-loop decoder (:413-581)
# ----------------------------------------------------------------------------------------
def _penalty_args(penalty) -> Tuple[float, bool]:
    """np.float64 penalties (the default np.log(0.005)) make the reference evaluate word-start
    candidates in float64; Python scalars and np.float32 are weak and stay float32."""
    f64 = isinstance(penalty, np.float64) or (isinstance(penalty, np.ndarray) and penalty.dtype == np.float64)
    return float(penalty), bool(f64)


@dataclass
class HiddenMarkovModelInference(_PackCache):
    _multivariate_normals: List[MultivariateNormal] = field(default_factory=list)
    _log_transition_probs: LogTransitionProbabilities = field(init=False)
    _model_boundaries: ModelBoundary = field(init=False)
    _log_transition_probability_between_words: float = field(default=np.log(0.005))

    @classmethod
    def from_folder(cls, folder_path: str, models_to_load: List[str]) -> Self:
        inf = cls()
        ltp = LogTransitionProbabilities()
        normals: List[MultivariateNormal] = []
        labels: List[str] = []
        mb = ModelBoundary()
        for name in sorted(os.listdir(folder_path)):
            path = os.path.join(folder_path, name)
            label = HiddenMarkovModel._model_folder_name_parser(path)
            if label not in models_to_load:
                continue
            hmm = HiddenMarkovModel.from_folder(path)
            ltp.append(hmm._log_transition_probs)
            normals.extend(hmm._multivariate_normals)
            mb.append(hmm.num_of_states)
            labels.append(label)
        inf._log_transition_probs = ltp
        inf._multivariate_normals = normals
        mb.add_model_labels(labels)
        inf._model_boundaries = mb
        return inf

    @classmethod
    def from_models(cls, models: Sequence[HiddenMarkovModel]) -> Self:
        """Same as from_folder for in-memory word models, in the order given (added)."""
        inf = cls()
        ltp = LogTransitionProbabilities()
        mb = ModelBoundary()
        normals: List[MultivariateNormal] = []
        for m in models:
            ltp.append(m._log_transition_probs)
            normals.extend(m._multivariate_normals)
            mb.append(len(m._multivariate_normals))
        mb.add_model_labels([m.label for m in models])
        inf._log_transition_probs, inf._multivariate_normals, inf._model_boundaries = ltp, normals, mb
        return inf

    def _packs(self):
        def build():
            eng = _engine()
            mb = self._model_boundaries
            dense = self._log_transition_probs.to_dense()
            lows, sizes = mb.lower_boundaries, mb.sizes
            blocks = [dense[a:a + n, a:a + n] for a, n in zip(lows, sizes)]
            tr = _trellis.build(blocks, lows, list(range(len(sizes))), "loop")
            return eng.pack_gaussians(self._multivariate_normals), eng.pack_trellises([tr])
        return self._cached(_model_key(self._multivariate_normals, self._log_transition_probs, self.__dict__), build)

    def predict(self, signal: NDArray[np.float32]) -> str:
        """hidden_markov_model.py:458-461.  The word sequence is decoded by the Viterbi launch itself (fused labels); the
        host label routine only sees what the device table cannot hold (T == 1, more than 32 words) and raises there
        exactly like the reference."""
        return self.predict_batch([signal])[0]

    def _viterbi(self, observation_sequence: NDArray[np.float32]) -> Tuple[float, NDArray[np.int8]]:
        scores, paths = self.viterbi_batch([observation_sequence])
        return scores[0], paths[0]

    def viterbi_batch(self, signals: Sequence[NDArray[np.float32]], precision: Optional[str] = None):
        eng = _engine()
        batch = eng.upload_features(signals, self._multivariate_normals[0].dim_of_features)
        best_score, path = self._decode_device(batch, precision)
        path_h = path.cpu().numpy()
        off = batch.frm_off_host
        return best_score.cpu().numpy(), [path_h[off[i]:off[i + 1]] for i in range(batch.n_utt)]

    def _decode_device(self, batch, precision: Optional[str] = None, max_words: Optional[int] = None):
        """features (device) -> (best_score [n], path [F]) on the device: two launches.  With ``max_words``
        the Viterbi launch also decodes the word sequence: (best_score, path, words, count)."""
        eng = _engine()
        gp, tp = self._packs()
        pen, f64 = _penalty_args(self._log_transition_probability_between_words)   # read per call: scripts poke it
        scores = eng.emission(batch.feat, gp, precision)
        labels = None
        if max_words is not None:
            names = self._model_boundaries._labels
            labels = (names.index("S") if "S" in names else -1, max_words)
        out = eng.viterbi(scores, batch.frm_off, batch.n_utt, batch.max_frames, batch.total_frames, tp,
                          loop=True, penalty=pen, penalty_f64=f64, want_end_scores=False, labels=labels)
        if labels is not None:
            return out[3], out[0], out[4], out[5]
        return out[3], out[0]

    def _strings_host(self, words_h, count_h, path_getter, frm_off_host) -> List[str]:
        """Word-id table -> strings.  Utterances whose count overflowed the table or whose path held a
        negative state (T == 1) go through the host routine, which decides (and raises like the reference)."""
        labels = self._model_boundaries._labels
        n, max_words = words_h.shape
        if all(len(l) == 1 and ord(l) < 256 and l != "\n" for l in labels):
            # one C pass writes every utterance's labels + a newline, one split makes the strings (10 000 utterances:
            # ~0.5 ms instead of ~3 ms of NumPy indexing and slicing -- 6 % of an end-to-end decode step)
            from . import _native
            words_c = np.ascontiguousarray(words_h, dtype=np.int8)
            count_c = np.ascontiguousarray(count_h, dtype=np.int32)
            buf = np.empty(n * (max_words + 1), dtype=np.uint8)
            nb = _native.load().loe_labels_text_host(words_c.ctypes.data, count_c.ctypes.data, n, max_words,
                                                     "".join(labels).encode("latin-1"), len(labels), b"\n", buf.ctypes.data)
            out = buf[:nb].tobytes().decode("latin-1").split("\n")[:n]
        else:
            out = ["".join(labels[k] for k in words_h[i, :max(c, 0)]) for i, c in enumerate(count_h.tolist())]
        bad = np.nonzero((count_h < 0) | (count_h > max_words))[0]
        if bad.size:
            path_h = path_getter()
            for i in bad.tolist():
                out[i] = "".join(self._model_boundaries.get_labels(path_h[frm_off_host[i]:frm_off_host[i + 1]]))
        return out

    def _strings_device(self, batch, precision: Optional[str] = None, max_words: int = 32) -> List[str]:
        _, path, words, count = self._decode_device(batch, precision, max_words)
        return self._strings_host(words.cpu().numpy(), count.cpu().numpy(), lambda: path.cpu().numpy(), batch.frm_off_host)

    def predict_batch(self, signals: Sequence[NDArray[np.float32]], precision: Optional[str] = None) -> List[str]:
        """Digit strings of many (T, 39) feature matrices in one pass (added entry point)."""
        eng = _engine()
        batch = eng.upload_features(signals, self._multivariate_normals[0].dim_of_features)
        return self._strings_device(batch, precision)

    def decode_pcm_batch(self, signals: Sequence[NDArray], sample_rate: int = 16000, precision: Optional[str] = None) -> List[str]:
        """Raw PCM -> digit strings, everything between the H2D copy of the samples and the D2H
        copy of the word ids on the device (MFCC -> emission -> Viterbi -> labels; added entry point)."""
        eng = _engine()
        batch = eng.mfcc(signals, sample_rate)
        return self._strings_device(batch, precision)


    def decode_pcm_flat(self, pcm_flat, sample_offsets: NDArray[np.int64], sample_rate: int = 16000,
                        precision: Optional[str] = None, n_chunks: Optional[int] = None) -> List[str]:
        """Batch ingestion form of :meth:`decode_pcm_batch`: ``pcm_flat`` is ONE host buffer (numpy
        array or pinned torch tensor, utterances back to back; float32 as the reference holds them, or
        the raw int16 WAV samples -- half the PCIe bytes, identical results) and ``sample_offsets`` the
        [n+1] sample offsets.  The batch is cut into chunks of whole utterances; the host->device copy
        of chunk c+1 runs on a copy stream while chunk c goes through MFCC, emission, Viterbi and
        labels on the compute stream, and only the word-id tables come back (added entry point)."""
        from ._engine import Batch
        eng = _engine()
        torch = eng.torch
        off = np.asarray(sample_offsets, dtype=np.int64)
        n = len(off) - 1
        if n <= 0:
            return []
        if isinstance(pcm_flat, torch.Tensor):
            src = pcm_flat
        else:
            arr = np.asarray(pcm_flat)
            src = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int16 if arr.dtype == np.int16 else np.float32))
        frames = 1 + np.diff(off) // 160
        if n_chunks is None:
            n_chunks = int(min(8, max(1, (int(off[-1]) * src.element_size()) // (64 << 20))))      # >= 64 MB per chunk
        cuts = np.searchsorted(off, np.linspace(0, int(off[-1]), n_chunks + 1)[1:-1]).tolist()
        bounds = sorted(set([0] + [min(max(int(c), 0), n) for c in cuts] + [n]))
        comp = torch.cuda.current_stream(eng.device)
        copy = eng.copy_stream()
        copy.wait_stream(comp)
        max_words = 32
        words_h = torch.empty((n, max_words), dtype=torch.int8).pin_memory() if n_chunks > 1 else None
        count_h = torch.empty((n,), dtype=torch.int32).pin_memory() if n_chunks > 1 else None
        keep = []
        for a, b in zip(bounds[:-1], bounds[1:]):
            if a == b:
                continue
            s0, s1 = int(off[a]), int(off[b])
            fr = frames[a:b]
            frm_off = np.concatenate(([0], np.cumsum(fr))).astype(np.int64)
            with torch.cuda.stream(copy):
                pcm = src[s0:s1].to(eng.device, non_blocking=True)
                pcm_off = torch.from_numpy(off[a:b + 1] - s0).to(eng.device, non_blocking=True)
                frm_off_dev = torch.from_numpy(frm_off).to(eng.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            comp.wait_event(ev)
            for t in (pcm, pcm_off, frm_off_dev):
                t.record_stream(comp)
            feat = eng.mfcc_device(pcm, pcm_off, frm_off_dev, b - a, int(frm_off[-1]), int(fr.max()), int(fr.min()), sample_rate)
            batch = Batch(feat, frm_off_dev, frm_off, b - a, int(fr.max()))
            _, path, words, count = self._decode_device(batch, precision, max_words)
            if n_chunks > 1:
                words_h[a:b].copy_(words, non_blocking=True)
                count_h[a:b].copy_(count, non_blocking=True)
            keep.append((a, b, path, frm_off, words, count))
        if n_chunks > 1:
            comp.synchronize()
            wh, ch = words_h.numpy(), count_h.numpy()
        else:
            wh, ch = keep[0][4].cpu().numpy(), keep[0][5].cpu().numpy()

        def full_path():
            return np.concatenate([p.cpu().numpy() for _, _, p, _, _, _ in keep])
        frm_all = np.concatenate(([0], np.cumsum(frames))).astype(np.int64)
        return self._strings_host(wh, ch, full_path, frm_all)

    # -- torch-free route: the C host-buffer decoder (include/loe_b200.h, loe_decoder_*) ----------
    def native_decoder(self, sample_rate: float = 16000, device: int = 0):
        """The :class:`_decoder.NativeDecoder` of this model (device tables, streams and workspace owned by
        the C library; rebuilt when a script replaces the Gaussians / transitions)."""
        from ._decoder import NativeDecoder
        from ._engine import host_gauss_arrays

        def build():
            mb = self._model_boundaries
            dense = self._log_transition_probs.to_dense()
            lows, sizes = mb.lower_boundaries, mb.sizes
            blocks = [dense[a:a + n, a:a + n] for a, n in zip(lows, sizes)]
            tr = _trellis.build(blocks, lows, list(range(len(sizes))), "loop")
            return NativeDecoder(*host_gauss_arrays(self._multivariate_normals), tr, sample_rate, device)
        key = ("native", float(sample_rate), int(device)) + tuple(_model_key(self._multivariate_normals, self._log_transition_probs, self.__dict__))
        return self._cached(key, build, slot="_native_decoder")

    def decode_pcm_host(self, pcm_flat, sample_offsets, sample_rate: float = 16000, n_chunks: int = 0,
                        device: int = 0) -> List[str]:
        """:meth:`decode_pcm_flat` through ``loe_decoder_decode_host``: numpy in, strings out, no torch on
        the way (``pcm_flat`` float32 or int16; a :class:`_decoder.PinnedBuffer` array makes the copies
        asynchronous).  Same results as every other decode entry point."""
        dec = self.native_decoder(sample_rate, device)
        pen, f64 = _penalty_args(self._log_transition_probability_between_words)
        names = self._model_boundaries._labels
        skip = names.index("S") if "S" in names else -1
        off = np.asarray(sample_offsets, dtype=np.int64)
        words, count, _, _ = dec.decode(pcm_flat, off, pen, f64, skip, 32, n_chunks)
        if len(off) <= 1:
            return []

        def full_path():
            return dec.decode(pcm_flat, off, pen, f64, skip, 32, n_chunks, want_path=True)[3]
        frm_all = np.concatenate(([0], np.cumsum(1 + np.diff(off) // 160))).astype(np.int64)
        return self._strings_host(words, count, full_path, frm_all)


# ----------------------------------------------------------------------------------------
# embedded ("continuous") training (:584-797)
# ----------------------------------------------------------------------------------------
@dataclass
class HiddenMarkovModelMultiWord(HiddenMarkovModel):
    _model_boundaries: ModelBoundary = field(init=False)

    def _host_trellis(self) -> _trellis.HostTrellis:
        mb = self._model_boundaries
        dense = self._log_transition_probs.to_dense()
        lows, sizes = mb.lower_boundaries, mb.sizes
        blocks = [dense[a:a + n, a:a + n] for a, n in zip(lows, sizes)]
        uniq = {lab: i for i, lab in enumerate(dict.fromkeys(mb._labels))}
        return _trellis.build(blocks, lows, [uniq[l] for l in mb._labels], "chain")

    def get_remuexed_signals(self, mfccs_sequences: List[NDArray[np.float32]]) -> Dict[str, List[Signal]]:
        out: Dict[str, List[Signal]] = {label: [] for label in self._model_boundaries._labels}
        if len(mfccs_sequences) == 0:
            return out
        _, paths = self.predict_batch(mfccs_sequences)
        for x, path in zip(mfccs_sequences, paths):
            for label, sigs in self._remux_path_and_signal(x, path, self._model_boundaries).items():
                out[label].extend(sigs)
        return out

    @staticmethod
    def _remux_path_and_signal(signal, path, model_boundaries: ModelBoundary) -> Dict[str, List[Signal]]:
        """Host version of the cut-per-word step (:602-636), kept for API compatibility; the
        trainer itself uses loe_align_dev(remux=1) and never materialises these objects."""
        out: Dict[str, List[Signal]] = {label: [] for label in model_boundaries._labels}
        labels = [model_boundaries.get_label(int(s)) for s in path]
        start = 0
        for i in range(1, len(path)):
            if labels[i] != labels[start]:
                lo = model_boundaries.find_lower_boundary(int(path[start]))
                hi = model_boundaries.find_upper_boundary(int(path[start]))
                out[labels[start]].append(Signal(num_of_state=hi - lo + 1, signal=signal[start:i], path=path[start:i] - lo))
                start = i
        return out

    @classmethod
    def from_labels(cls, labels: str, trainable_models: Dict[str, HiddenMarkovModelTrainable]) -> Self:
        hmm = cls(labels)
        hmm.isTqdm = False
        ltp = LogTransitionProbabilities()
        normals: List[MultivariateNormal] = []
        mb = ModelBoundary()
        for label in labels:
            ltp.append(trainable_models[label]._log_transition_probs)
            normals.extend(trainable_models[label]._multivariate_normals)
            mb.append(len(trainable_models[label]._multivariate_normals))
        mb.add_model_labels(list(labels))
        hmm._log_transition_probs, hmm._multivariate_normals, hmm._model_boundaries = ltp, normals, mb
        return hmm


@dataclass
class HiddenMarkovModelTrainContinuous:
    isTqdm: bool = field(default=True)
    isMultiProcessing: bool = field(default=True)
    _trainable_models: Dict[str, HiddenMarkovModelTrainable] = field(default_factory=dict)
    _models_loaded: List[str] = field(default_factory=list)
    _num_of_finished_models: int = field(default=0)

    @classmethod
    def from_folder(cls, folder_path: str, models_to_load: List[str]) -> Self:
        """Seeds from saved isolated models; _means = 0, cov = 0.01 I and uniform transitions are
        placeholders, only the loaded Gaussians / log-transitions carry the seed (:679-712)."""
        tc = cls()
        for name in sorted(os.listdir(folder_path)):
            path = os.path.join(folder_path, name)
            label = HiddenMarkovModel._model_folder_name_parser(path)
            if label not in models_to_load:
                continue
            m = HiddenMarkovModelTrainable.from_folder(model_folder_path=path)
            S, D = len(m._multivariate_normals), m._multivariate_normals[0].dim_of_features
            m._means = np.zeros((S, D), dtype=np.float32)
            m._covariances = m._init_covariance(D, S)
            m._transition_probs = TransitionProbabilities.from_num_of_states(S)
            tc._trainable_models[label] = m
        tc._models_loaded = models_to_load
        return tc

    def train(self, labeled_mfccs: Dict[str, List[NDArray[np.float32]]], max_iterations: int = 100) -> None:
        """Embedded re-estimation (:714-731).  The utterances are uploaded once; every iteration is
        emission + chain Viterbi + align/remux + statistics on the device, then the per-word
        M-step in the reference's order.  Sharded by rank under torch.distributed."""
        from . import _dist

        eng = _engine()
        items = [(lab, x) for lab, xs in labeled_mfccs.items() for x in xs]
        items = _dist.shard(items)
        chains = list(dict.fromkeys(self.insert_silence(lab) for lab, _ in items))
        chain_id = {c: i for i, c in enumerate(chains)}
        batch = eng.upload_features([x for _, x in items]) if items else None
        utt_tr = eng._to_dev(np.array([chain_id[self.insert_silence(lab)] for lab, _ in items], dtype=np.int32)) if items else None
        bar = tqdm(total=max_iterations, desc="Training Iteration", disable=not self.isTqdm, position=0)
        for it in range(max_iterations):
            try:
                stats, counts = self._device_statistics(batch, utt_tr, chains)
                self._update_trainable_model_parameters(statistics=(stats, counts))
            except HiddenMarkovModelTrainable.HMMTrainMeanFail:
                logger.error("Failed to train model")
                raise
            except HiddenMarkovModelTrainable.HMMTrainConverge:
                logger.info(f"Finish training model after {it} iterations")
                break
            bar.update()
        bar.close()

    # global emission table: the loaded word models side by side, in _trainable_models order
    def _global_layout(self):
        labels = list(self._trainable_models.keys())
        sizes = [len(self._trainable_models[l]._multivariate_normals) for l in labels]
        starts = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(int)
        return labels, sizes, {l: int(s) for l, s in zip(labels, starts)}, int(sum(sizes))

    def _device_statistics(self, batch, utt_tr, chains):
        from . import _dist

        eng = _engine()
        labels, sizes, start, G = self._global_layout()
        D = self._trainable_models[labels[0]]._multivariate_normals[0].dim_of_features
        stride = 1 + D + D * (D + 1) // 2
        if batch is not None:
            normals = [mn for l in labels for mn in self._trainable_models[l]._multivariate_normals]
            gp = eng.pack_gaussians(normals)
            label_id = {l: i for i, l in enumerate(labels)}
            trellises = []
            for chain in chains:
                dens = [self._trainable_models[c]._log_transition_probs.to_dense() for c in chain]
                trellises.append(_trellis.build(dens, [start[c] for c in chain], [label_id[c] for c in chain], "chain"))
            tp = eng.pack_trellises(trellises)
            scores = eng.emission(batch.feat, gp)
            path, _, _, _ = eng.viterbi(scores, batch.frm_off, batch.n_utt, batch.max_frames, batch.total_frames, tp,
                                        utt_tr=utt_tr, want_end_scores=False)
            shift = eng._to_dev(np.concatenate([np.asarray(self._trainable_models[l]._means, dtype=np.float32) for l in labels]))
            stats, counts, _ = eng.kmeans_stats(batch.feat, path, batch.frm_off, batch.n_utt, batch.total_frames, tp,
                                                utt_tr, True, G, shift)
        else:
            stats = eng.torch.zeros((G, stride), dtype=eng.torch.float64, device=eng.device)
            counts = eng.torch.zeros((G, G), dtype=eng.torch.int32, device=eng.device)
        stats, counts = _dist.allreduce_stats(stats, counts)
        return stats.cpu().numpy(), counts.cpu().numpy().astype(np.int64)

    def _train(self, labeled_mfccs: Dict[str, List[NDArray[np.float32]]]) -> Dict[str, List[Signal]]:
        """Reference-shaped E-step (:733-752): per label string, forced alignment + remux."""
        out: Dict[str, List[Signal]] = {label: [] for label in self._models_loaded}
        for item in labeled_mfccs.items():
            for lab, sigs in self._train_process(item).items():
                out[lab].extend(sigs)
        return out

    def _train_process(self, labels_and_mfccs) -> Dict[str, List[Signal]]:
        labels, mfccs = labels_and_mfccs
        hmm = HiddenMarkovModelMultiWord.from_labels(self.insert_silence(labels), self._trainable_models)
        return hmm.get_remuexed_signals(mfccs)

    def _update_trainable_model_parameters(self, remuxed_signals: Optional[Dict[str, List[Signal]]] = None,
                                           statistics=None) -> None:
        """Per-word M-step in dict order with the reference's stop rule (:754-770): the counter of
        converged models is CUMULATIVE across iterations and the stop fires when it equals the
        number of models; words after the one that fired are not updated in that iteration."""
        if statistics is None:
            for label, signals in remuxed_signals.items():
                self._one_word(label, lambda m, s=signals: m._train_external(s))
            return
        stats, counts = statistics
        labels, sizes, start, _ = self._global_layout()
        for label in (l for l in self._models_loaded if l in self._trainable_models):
            a, n = start[label], len(self._trainable_models[label]._multivariate_normals)
            self._one_word(label, lambda m, a=a, n=n: m._update_from_statistics(
                stats[a:a + n], counts[a:a + n, a:a + n], shift=m._means.astype(np.float64)))

    def _one_word(self, label, update) -> None:
        model = self._trainable_models[label]
        try:
            update(model)
        except HiddenMarkovModelTrainable.HMMTrainConverge:
            self._num_of_finished_models += 1
            if self._num_of_finished_models == len(self._trainable_models):
                raise
        finally:
            model._update_inference_weights()

    def save(self, folder_path: str) -> None:
        os.makedirs(folder_path, exist_ok=True)
        for model in self._trainable_models.values():
            model.save(folder_path)

    @staticmethod
    def insert_silence(labels: str) -> str:
        return "".join(f"S{c}" for c in labels) + "S"
