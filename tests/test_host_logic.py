"""CPU: host-side logic of the drop-in package (no GPU, no compute calls into the library)."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from helpers import LOOP_ORDER, N_STATES, WORDS, oracle_flat, trained_word_model
from oracle import hmm as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_public_names_and_out_of_scope():
    import loe_speech_recognition as L
    for name in ("MFCC", "HiddenMarkovModel", "HiddenMarkovModelTrainable", "HiddenMarkovModelInference",
                 "HiddenMarkovModelTrainContinuous", "Signal", "ModelCollection", "TI_DIGITS_LABELS"):
        assert hasattr(L, name)
    assert list(L.TI_DIGITS_LABELS) == ["1", "2", "3", "4", "5", "6", "7", "8", "9", "O", "Z"]
    for name in ("TIDigits", "DataLoader", "CSVReader", "CSVWriter", "plot_confusion_matrix_from_lists", "plot_line",
                 "SignalSeparation", "DynamicTimeWarping"):
        assert hasattr(L, name)
    with pytest.raises(NotImplementedError):
        L.Segmentation                                    # microphone capture: the one name that is not rebuilt
    from loe_speech_recognition import HiddenMarkovModelInference
    assert isinstance(HiddenMarkovModelInference()._log_transition_probability_between_words, np.float64)


def test_spawn_start_method_and_reference_fallback(tmp_path):
    """Importing the package leaves the process-wide start method alone; it is switched to "forkserver" (torch and the
    package preloaded into the server; LOE_B200_START_METHOD=spawn for plain spawn) only when this process creates its
    CUDA engine (CUDA does not survive fork) and only if the application has not chosen one.
    Out-of-scope host modules can be borrowed from a reference checkout."""
    pkg = os.path.join(ROOT, "cs-304-speech-recognition-code_b200")
    code = ("import sys; sys.path.insert(0, %r); import multiprocessing, loe_speech_recognition as L; "
            "print(multiprocessing.get_start_method(allow_none=True)); "
            "from loe_speech_recognition import _engine; _engine._prefer_spawn(); "
            "print(multiprocessing.get_start_method(allow_none=True))" % pkg)
    assert subprocess.check_output([sys.executable, "-c", code], text=True).split() == ["None", "forkserver"]
    env = dict(os.environ, LOE_B200_START_METHOD="spawn")
    assert subprocess.check_output([sys.executable, "-c", code], text=True, env=env).split() == ["None", "spawn"]
    env = dict(os.environ, LOE_B200_KEEP_START_METHOD="1")
    assert subprocess.check_output([sys.executable, "-c", code], text=True, env=env).split() == ["None", "None"]
    # an implicitly latched "fork" (tqdm's multiprocessing lock, scripts/project3_predict_simple.py:15) is replaced too
    code_fork = code.replace("import multiprocessing, loe", "import multiprocessing; multiprocessing.get_context(); import loe")
    assert subprocess.check_output([sys.executable, "-c", code_fork], text=True).split() == ["fork", "forkserver"]
    code_fs = code.replace("import multiprocessing, loe", "import multiprocessing; multiprocessing.set_start_method('spawn'); import loe")
    assert subprocess.check_output([sys.executable, "-c", code_fs], text=True).split() == ["spawn", "spawn"]
    ref = "/root/reference/src/loe_speech_recognition"
    if os.path.isdir(ref):
        # the reference's segmentation.py needs sounddevice at import: borrow it with a stub of that module
        code3 = ("import sys, types; sys.path.insert(0, %r); sd = types.ModuleType('sounddevice'); "
                 "sd.InputStream = object; sd.CallbackFlags = object; sys.modules['sounddevice'] = sd; "
                 "import loe_speech_recognition as L; print(L.Segmentation.__module__)" % pkg)
        out = subprocess.run([sys.executable, "-c", code3], text=True, capture_output=True,
                             env=dict(os.environ, LOE_REFERENCE_SRC=ref))
        assert out.returncode != 0 or out.stdout.strip() == "loe_speech_recognition.segmentation"


def test_transition_matrices_semantics():
    from loe_speech_recognition.transition_probability import LogTransitionProbabilities, TransitionProbabilities
    tp = TransitionProbabilities.from_num_of_states(5)
    _, _, ref_trans = O.init_parameters(np.zeros((10, 3), np.float32), 5)
    assert np.array_equal(tp.to_dense(), ref_trans)
    assert len(tp._core) == 25 and tp[(4, 0)] == 0.0
    ltp = LogTransitionProbabilities.from_transition_probability(tp)
    assert np.array_equal(ltp.to_dense(), O.log_transitions(ref_trans))
    assert ltp[(3, 1)] == -np.inf and isinstance(ltp[(0, 0)], np.float32)
    both = LogTransitionProbabilities()
    both.append(ltp); both.append(ltp)
    assert both.num_of_states == 10 and both[(4, 5)] == 0.0 and (4, 5) not in both._core   # absent key reads 0.0
    assert np.array_equal(both.to_dense(), O.block_diag_missing_zero([ltp.to_dense(), ltp.to_dense()]))


def test_reference_pickles_load_and_round_trip(tmp_path):
    from loe_speech_recognition import HiddenMarkovModel, HiddenMarkovModelInference
    src = os.path.join(ROOT, "tests", "golden", "golden_models")
    m = HiddenMarkovModel.from_folder(os.path.join(src, "Z"))
    assert m.label == "Z" and m.num_of_states == 5 and m.dim_of_features == 39
    assert type(m._log_transition_probs).__module__ == "loe_speech_recognition.transition_probability"
    assert type(m._multivariate_normals[0]).__module__ == "loe_speech_recognition.hidden_markov_model"
    m.save(str(tmp_path))
    m2 = HiddenMarkovModel.from_folder(str(tmp_path / "Z"))
    assert m2._log_transition_probs._core == m._log_transition_probs._core
    assert np.array_equal(m2._multivariate_normals[2]._core.mean, m._multivariate_normals[2]._core.mean)
    with pytest.raises(FileNotFoundError):
        HiddenMarkovModel.from_folder(str(tmp_path / "nope"))
    inf = HiddenMarkovModelInference.from_folder(src, ["1", "S", "Z", "7"])
    assert inf._model_boundaries._labels == ["1", "S", "Z"]
    assert inf._model_boundaries.lower_boundaries == [0, 5, 8] and inf._model_boundaries.upper_boundaries == [4, 7, 12]
    blob = pickle.dumps(inf)                       # models are shipped to pool workers by the scripts
    assert pickle.loads(blob)._model_boundaries._labels == ["1", "S", "Z"]


def test_flat_model_format_round_trip(tmp_path, golden):
    """model.npz (plain arrays, no scipy pickles): same emission tables and transitions after reload."""
    from loe_speech_recognition import HiddenMarkovModel
    from loe_speech_recognition._engine import host_gauss_arrays
    m = trained_word_model(golden, "4")
    m.save_flat(str(tmp_path))
    m2 = HiddenMarkovModel.from_flat(str(tmp_path / "4"))
    assert m2.label == "4" and m2.num_of_states == 5
    for a, b in zip(host_gauss_arrays(m._multivariate_normals), host_gauss_arrays(m2._multivariate_normals)):
        assert np.array_equal(a, b)
    assert np.array_equal(m2._log_transition_probs.to_dense(), m._log_transition_probs.to_dense())
    assert isinstance(m2._log_transition_probs[(0, 0)], np.float32)
    with pytest.raises(FileNotFoundError):
        HiddenMarkovModel.from_flat(str(tmp_path / "nope"))


def test_trellis_builder_matches_oracle(golden):
    from loe_speech_recognition import _trellis
    from loe_speech_recognition._native import POS_END, POS_INIT, POS_START
    logAs = [golden[f"train_logA_{w}"] for w in LOOP_ORDER]
    sizes = [a.shape[0] for a in logAs]
    lows = np.concatenate(([0], np.cumsum(sizes)[:-1])).tolist()
    for kind, ref in (("loop", O.loop_trellis(logAs)), ("chain", O.chain_trellis(logAs))):
        t = _trellis.build(logAs, lows, list(range(12)), kind)
        assert np.array_equal(t.band, ref.band)
        assert np.array_equal((t.flags & POS_INIT) != 0, ref.init)
        assert np.array_equal(np.nonzero(t.flags & POS_END)[0], ref.ends)
        if kind == "loop":
            assert np.array_equal(np.nonzero(t.flags & POS_START)[0], ref.loop_starts)
    w = _trellis.build([logAs[0]], [0], [0], "word")
    assert np.array_equal(w.band, O.word_trellis(logAs[0]).band)
    off, col, band, flags, word, word_lo, max_pos, max_ends = _trellis.stack([w, _trellis.build(logAs, lows, list(range(12)), "loop")])
    assert off.tolist() == [0, 5, 63] and max_pos == 58 and max_ends == 12 and band.shape == (63, 3)


def test_model_boundary_labels_match_oracle(golden):
    from loe_speech_recognition.model_boundary import ModelBoundary
    sizes = [N_STATES[w] for w in LOOP_ORDER]
    mb = ModelBoundary()
    for n in sizes:
        mb.append(n)
    mb.add_model_labels(list(LOOP_ORDER))
    for name in ("int", "f64"):
        for i in range(10):
            path = golden[f"loop_path_{name}_{i}"]
            assert "".join(mb.get_labels(path)) == str(golden[f"loop_strings_{name}"][i])
            assert mb.get_labels(path, skip_silence=False) == O.get_labels(path, sizes, list(LOOP_ORDER), skip_silence=False)
    rep = np.array([0, 1, 2, 3, 4, 0, 1, 4, 4, 50, 52, 53], dtype=np.int8)      # repeated word via last->first state
    assert mb.get_labels(rep) == O.get_labels(rep, sizes, list(LOOP_ORDER)) == ["1", "1", "Z"]
    with pytest.raises(Exception):
        mb.get_labels(np.array([-1], dtype=np.int8))
    with pytest.raises(Exception):
        mb.append(3)                                  # frozen after the boundaries were read
    assert mb.get_label(52) == "S" and mb.find_lower_boundary(57) == 53 and mb.find_upper_boundary(50) == 52


def test_signal_bookkeeping_matches_oracle():
    from loe_speech_recognition.signal import Signal, SortedSignals
    rng = np.random.default_rng(0)
    paths = [np.array(p, dtype=np.int8) for p in ([0, 0, 1, 1, 2, 2, 2], [0, 2, 2, 1, 1], [1, 1, 2], [2, 2, 0, 0])]
    sigs = [rng.normal(size=(len(p), 4)).astype(np.float32) for p in paths]
    ss = SortedSignals(3)
    for x, p in zip(sigs, paths):
        s = Signal(3, x, p)
        ref = O.order_by_state(x, p, 3)
        for a, b in zip(s.order_by_state, ref):
            assert (a is None and b is None) or np.array_equal(a, b)
        ss.append(s)
    ref = O.mstep(sigs, paths, 3)
    assert np.array_equal(ss.transition_counts, ref["counts"])
    assert np.array_equal(ss.transition_probabilities.to_dense(), ref["trans"], equal_nan=True)


def _numpy_stats(feats, paths, n_states, shift):
    """Host stand-in for loe_align_dev + loe_kmeans_dev (same packed layout)."""
    D = feats[0].shape[1]
    iu = np.triu_indices(D)
    stats = np.zeros((n_states, 1 + D + D * (D + 1) // 2))
    counts = np.zeros((n_states, n_states), dtype=np.int64)
    for x, p in zip(feats, paths):
        for s, seg in enumerate(O.order_by_state(x, p, n_states)):
            if seg is None:
                continue
            d = seg.astype(np.float64) - shift[s]
            stats[s, 0] += len(seg)
            stats[s, 1:1 + D] += d.sum(0)
            stats[s, 1 + D:] += (d.T @ d)[iu]
        np.add.at(counts, (p[:-1].astype(int), p[1:].astype(int)), 1)
    return stats, counts


def test_mstep_from_statistics_matches_oracle(golden):
    from loe_speech_recognition import HiddenMarkovModelTrainable
    w = "5"
    feats = [golden[f"train_feat_{w}_{i}"] for i in range(8)]
    means, Us, lps, logA = oracle_flat(golden, w)
    tr = O.word_trellis(logA)
    paths = [O.viterbi(O.emission_scores(x, means, Us, lps), tr)[2] for x in feats]
    old = golden[f"train_means_{w}"] + 0.5
    ref = O.mstep(feats, paths, 5, old_means=old)
    m = HiddenMarkovModelTrainable(w)
    m._means = old.astype(np.float32)
    m._covariances = m._init_covariance(39, 5)
    stats, counts = _numpy_stats(feats, paths, 5, m._means.astype(np.float64))
    m._update_from_statistics(stats, counts, shift=m._means.astype(np.float64))
    assert np.allclose(m._means, ref["means"], rtol=1e-6, atol=1e-6)
    assert np.allclose(m._covariances, ref["covs"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(m._transition_probs.to_dense(), ref["trans"])
    # convergence is tested on the means only, before covariances / transitions are touched
    cov_before = m._covariances.copy()
    stats, counts = _numpy_stats(feats, paths, 5, m._means.astype(np.float64))
    with pytest.raises(HiddenMarkovModelTrainable.HMMTrainConverge):
        m._update_from_statistics(stats, counts, shift=m._means.astype(np.float64))
    assert np.array_equal(m._covariances, cov_before)
    stats[2, 0] = 0
    with pytest.raises(HiddenMarkovModelTrainable.HMMTrainMeanFail):
        m._update_from_statistics(stats, counts, shift=m._means.astype(np.float64))


def test_init_parameters_match_oracle(golden):
    from loe_speech_recognition import HiddenMarkovModelTrainable
    x = golden["train_feat_2_0"]
    means, covs, tp = HiddenMarkovModelTrainable._init_parameters(x, 5)
    rm, rc, rt = O.init_parameters(x, 5)
    assert np.array_equal(means, rm) and np.array_equal(covs, rc) and np.array_equal(tp.to_dense(), rt)


def test_embedded_stop_rule_is_cumulative():
    """_num_of_finished_models accumulates across iterations (hidden_markov_model.py:754-770)."""
    from loe_speech_recognition import HiddenMarkovModelTrainContinuous, HiddenMarkovModelTrainable

    class Fake:
        def __init__(self, conv):
            self.conv, self.updated = conv, 0
        def _update_inference_weights(self):
            self.updated += 1

    tc = HiddenMarkovModelTrainContinuous()
    tc._trainable_models = {"a": Fake(True), "b": Fake(False)}

    def upd(m):
        if m.conv:
            raise HiddenMarkovModelTrainable.HMMTrainConverge
    tc._one_word("a", upd)
    assert tc._num_of_finished_models == 1
    with pytest.raises(HiddenMarkovModelTrainable.HMMTrainConverge):
        tc._one_word("a", upd)                        # same model converging again reaches len(models) == 2
    assert tc._trainable_models["a"].updated == 2
    assert tc.insert_silence("Z1") == "SZS1S"


def test_penalty_modes():
    from loe_speech_recognition.hidden_markov_model import _penalty_args
    assert _penalty_args(np.log(0.005)) == (float(np.log(0.005)), True)
    assert _penalty_args(-100) == (-100.0, False)
    assert _penalty_args(-37.25) == (-37.25, False)
    assert _penalty_args(np.float32(-3)) == (-3.0, False)


def test_mel_filterbank_matches_oracle():
    from loe_speech_recognition.mfcc import mel_filterbank, mel_lane_tables
    from oracle import mfcc as OM
    assert np.array_equal(mel_filterbank(16000), OM.mel_basis())
    bins, w, na, nb = mel_lane_tables(16000)
    assert (na, nb) == (11, 5)
    bins = bins.reshape(na + nb, 32); w = w.reshape(na + nb, 32)
    dense = np.zeros((40, 161), np.float32)
    for it in range(na):                              # round A: lane = filter
        for lane in range(32):
            dense[lane, bins[it, lane]] += w[it, lane]
    for it in range(na, na + nb):                     # round B: 4 lanes per filter
        for lane in range(32):
            dense[32 + lane // 4, bins[it, lane]] += w[it, lane]
    assert np.array_equal(dense, OM.mel_basis())
    for sr in (22050, 44100):
        mel_lane_tables(sr)


def test_mfcc_input_validation_needs_no_gpu():
    from loe_speech_recognition import MFCC
    with pytest.raises(TypeError):
        MFCC([0.0] * 2000, 16000)
    with pytest.raises(ValueError):
        MFCC(np.zeros((2, 2000), np.float32), 16000)
    x = np.arange(26, dtype=np.float32).reshape(13, 2)
    from oracle import mfcc as OM
    assert np.allclose(MFCC.normalize_mfccs(x), OM.normalize_mfccs(x))


def test_no_cuda_fails_loudly():
    """Without a CUDA device the product must raise, never fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from loe_speech_recognition import MFCC
    from loe_speech_recognition._engine import NoCudaDevice
    with pytest.raises(NoCudaDevice):
        MFCC.batch([np.zeros(16000, np.float32)], 16000)


def test_product_never_imports_oracle():
    code = ("import sys; sys.path.insert(0, %r); import loe_speech_recognition, loe_speech_recognition._engine, "
            "loe_speech_recognition.synthetic; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
            % os.path.join(ROOT, "cs-304-speech-recognition-code_b200"))
    subprocess.check_call([sys.executable, "-c", code], cwd="/tmp")
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cs-304-speech-recognition-code_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_h16_image_is_triangular_and_scores_like_the_whitening_matrix():
    """pack_h16_image (host, no GPU): the 3xFP16 operand of loe_emission_h16_dev replaces U_s by the
    lower-triangular R_s^T of U_s^T = Q R_s.  Rebuilt from the image (hi + lo parts, layout of include/loe_b200.h):
    the same |W^T (x - mean)|^2 as U_s to ~1e-7, feature chunk c empty beyond column 8 (c + 1) (what lets the
    kernel issue narrower MMAs), column 39 and the padding states zero, the second hi copy identical."""
    from loe_speech_recognition._engine import pack_h16_image
    rng = np.random.default_rng(0)
    S, D = 8, 39
    means = rng.normal(size=(S, D))
    us = np.stack([np.linalg.qr(rng.normal(size=(D, D)))[0] @ np.diag(rng.uniform(0.5, 30, D)) for _ in range(S)])
    flat = pack_h16_image(means, us, np.zeros(S))
    assert flat.dtype == np.float16 and flat.size == 2 * 15 * 240 * 8
    img = flat.reshape(2, 15, 240, 8).astype(np.float64)
    assert np.array_equal(img[:, 0:5], img[:, 10:15])
    x = rng.normal(size=(7, D))
    xa = np.concatenate([x, np.ones((7, 1))], axis=1)
    for s in range(12):
        t, sl = divmod(s, 6)
        W = np.zeros((40, 40))
        for c in range(5):
            for j in range(40):
                n = (j // 8) * 48 + sl * 8 + j % 8
                W[8 * c:8 * c + 8, j] = img[t, c, n] + img[t, 5 + c, n]
        if s >= S:
            assert not W.any()                                        # padding states of the last tile
            continue
        assert not np.triu(W[:39, :39], 1).any() and not W[:, 39].any()
        for c in range(4):
            assert not W[8 * c:8 * c + 8, 8 * (c + 1):].any()
        ref = (((x - means[s]) @ us[s]) ** 2).sum(axis=1)
        got = ((xa @ W) ** 2).sum(axis=1)
        assert np.abs(got - ref).max() <= 2e-7 * ref.max()
    assert pack_h16_image(means, us * 1e4, np.zeros(S)) is None       # outside the binary16 range: TF32 image instead
    bad = us.copy()
    bad[3, 5, 7] = np.nan
    assert pack_h16_image(means, bad, np.zeros(S)) is None


def test_mel_lane_tables_exact_and_bank_conflict_free():
    """mel_lane_tables (host): the lane-balanced filterbank the MFCC kernel reads reproduces the dense slaney
    filterbank exactly, every lane walks consecutive bins, and at 16 kHz the window starts are slid to minimise the
    shared-memory wavefronts of the 16-byte power-spectrum reads."""
    from loe_speech_recognition.mfcc import mel_filterbank, mel_lane_tables
    for sr in (16000, 8000, 22050):
        bins, w, na, nb = mel_lane_tables(sr)
        B, Wt = bins.reshape(na + nb, 32), w.reshape(na + nb, 32)
        dense = mel_filterbank(sr)
        rec = np.zeros_like(dense)
        for it in range(na + nb):
            for lane in range(32):
                m = lane if it < na else 32 + lane // 4
                if Wt[it, lane] != 0:
                    rec[m, B[it, lane]] += Wt[it, lane]
        assert np.array_equal(rec, dense)
        assert np.array_equal(B[:na], B[0][None, :] + np.arange(na)[:, None])
        assert np.array_equal(B[na:], B[na][None, :] + 4 * np.arange(nb)[:, None])
        assert B.min() >= 0 and B.max() < 161 + 3 + 32 + 64               # inside the kernel's frame area slack
        if sr == 16000:
            # mfcc_mel_r_kernel reads 16-byte [bin][4 frames] entries, a quarter-warp per wavefront: the window starts
            # are slid so that its loads take at most half the wavefronts of the natural starts (11 x 4 is the floor of round A)
            from loe_speech_recognition.mfcc import _quarter_wavefronts
            assert (na, nb) == (11, 5)
            nz = [np.nonzero(dense[m])[0] for m in range(40)]
            first, last = [int(z[0]) for z in nz], [int(z[-1]) for z in nz]

            def total(starts_a, starts_b):
                ta = sum(_quarter_wavefronts(list(starts_a[q:q + 8]), first[q:q + 8], last[q:q + 8], na, 1) for q in range(0, 32, 8))
                tb = sum(_quarter_wavefronts(list(starts_b[q:q + 2]), first[32 + q:34 + q], last[32 + q:34 + q], 4 * nb, 4)
                         for q in range(0, 8, 2))
                return ta, tb
            got = total([int(b) for b in B[0]], [int(b) for b in B[na][::4]])
            nat = total(first[:32], first[32:])
            assert got[0] <= 44 and got[1] <= 26 and sum(got) * 2 <= sum(nat), (got, nat)
            assert all(B[0][m] <= first[m] and B[0][m] + na > last[m] for m in range(32))
            assert all(B[na][4 * q] <= first[32 + q] and B[na][4 * q] + 4 * nb > last[32 + q] for q in range(8))


def test_model_key_sees_in_place_edits(built_lib):
    """The device-pack cache key (taken before every single-utterance call) changes when a mean, a whitening matrix or the
    transition table is edited IN PLACE, when an array object is replaced, and is stable otherwise; the fingerprint state
    is never pickled.  Host code only (loe_host_fingerprint)."""
    import pickle
    import time
    from loe_speech_recognition.hidden_markov_model import HiddenMarkovModel, HiddenMarkovModelTrainable, _model_key
    from loe_speech_recognition.transition_probability import LogTransitionProbabilities
    rng = np.random.default_rng(3)
    S, D = 58, 39
    means = rng.normal(size=(S, D)).astype(np.float32)
    a = rng.normal(size=(S, D, D))
    covs = (a @ a.transpose(0, 2, 1) / D + np.eye(D)).astype(np.float32)
    m = HiddenMarkovModel("x")
    m._multivariate_normals = HiddenMarkovModelTrainable.get_multivariate_normals(means, covs)
    m._log_transition_probs = LogTransitionProbabilities.from_dense(np.log(np.full((S, S), 1.0 / S, np.float32)))
    key = lambda: _model_key(m._multivariate_normals, m._log_transition_probs, m.__dict__)
    k0 = key()
    assert key() == k0
    t0 = time.perf_counter()
    for _ in range(50):
        key()
    per_call = (time.perf_counter() - t0) / 50
    print(f"_model_key: {per_call * 1e6:.0f} us per call for {S} Gaussians + {S * S} transitions")
    m._multivariate_normals[17]._core.mean[5] += 1e-9                     # a mean, in place
    k1 = key()
    assert k1 != k0
    m._multivariate_normals[40]._core.cov_object._LP[38, 38] *= 1.0000001    # a whitening matrix, last element
    k2 = key()
    assert k2 != k1
    m._log_transition_probs[(3, 4)] = np.float32(-7.0)                     # the dict-backed table
    k3 = key()
    assert k3 != k2
    core = m._multivariate_normals[2]._core
    core.mean = core.mean.copy()                                           # same content, another array object: no change
    assert key() == k3
    core.mean = core.mean + 1.0
    assert key() != k3
    clone = pickle.loads(pickle.dumps(m))
    assert "_fingerprint" not in clone.__dict__ and "_fingerprint" in m.__dict__
    assert _model_key(clone._multivariate_normals, clone._log_transition_probs, clone.__dict__)[5:] == key()[5:]
