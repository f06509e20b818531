"""Generate the committed golden fixtures by running the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the real ``loe_speech_recognition`` from /root/reference/src (with stub modules
for its absent GUI/audio dependencies, oracle/ref_import.py), drives it on the seeded
synthetic corpus and stores inputs + reference outputs as small ``.npz`` files plus one
reference-written model folder (pickles).  While doing so it asserts that the NumPy oracle
(oracle/hmm.py) reproduces every reference output bit-for-bit -- this is what pins the oracle.

MFCC: librosa is absent, so the features fed to the reference HMM code come from the
restated front end (oracle/mfcc.py, parity unpinned); ``golden_mfcc.npz`` is therefore an
oracle-generated regression fixture, not a reference output.
"""
from __future__ import annotations

import importlib.util
import os
import shutil
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import hmm as O            # noqa: E402
from oracle import mfcc as OM          # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402

warnings.filterwarnings("ignore")


def load_synth():
    path = os.path.join(ROOT, "cs-304-speech-recognition-code_b200", "loe_speech_recognition", "synthetic.py")
    spec = importlib.util.spec_from_file_location("loe_synth_for_golden", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


def silence_training_features(S, seed=40, n_strings=6, keep=16):
    """In-context silence: slices of string features (see synthetic.silence_frames)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_strings):
        ds = [S.DIGITS[int(i)] for i in rng.integers(0, 11, size=3)]
        pcm, segs = S.synth_string(rng, ds, return_segments=True)
        out += S.silence_frames(OM.mfcc_feature_vector(pcm).T, segs)
    return out[:keep]


def dense_logA(ltp):
    n = ltp.num_of_states
    out = np.zeros((n, n), dtype=np.float32)
    for (i, j), v in ltp._core.items():
        out[i, j] = v
    return out


def flat_model(hmm):
    means = np.array([mn._core.mean for mn in hmm._multivariate_normals])
    Us = np.array([mn._core.cov_object._LP for mn in hmm._multivariate_normals])
    lp = np.array([mn._core.cov_object._log_pdet for mn in hmm._multivariate_normals])
    covs = np.array([mn._core.cov_object.covariance for mn in hmm._multivariate_normals])
    return means, Us, lp, covs, dense_logA(hmm._log_transition_probs)


def main():
    R = import_reference()
    S = load_synth()
    from loe_speech_recognition.hidden_markov_model import HiddenMarkovModelMultiWord

    out = {}
    # ------------------------------------------------------------------ corpus + MFCC
    train = S.isolated_corpus(seed=0, n_per_word=8)
    test = S.isolated_corpus(seed=10, n_per_word=2, words=S.DIGITS)
    strings, truth = S.string_corpus(seed=20, n_utts=6, n_digits=7)
    short_strings, short_truth = S.string_corpus(seed=21, n_utts=4, n_digits=2)
    feats_train = {w: OM.mfcc_batch(v) for w, v in train.items()}
    feats_train["S"] = silence_training_features(S)
    feats_test = {w: OM.mfcc_batch(v) for w, v in test.items()}
    feats_str = OM.mfcc_batch(strings)
    feats_short = OM.mfcc_batch(short_strings)

    rng = np.random.default_rng(5)
    mf_in = [S.synth_isolated(rng, "3", 1.0), S.synth_string(rng, ["4", "Z"]), S.synth_isolated(rng, "S", 0.11),
             rng.normal(0, 800, size=16000 + 37).round().astype(np.float32)]
    np.savez_compressed(os.path.join(HERE, "golden_mfcc.npz"),
                        **{f"pcm{i}": p for i, p in enumerate(mf_in)},
                        **{f"feat{i}": OM.mfcc_feature_vector(p) for i, p in enumerate(mf_in)},
                        mel_basis=OM.mel_basis())

    # ------------------------------------------------------------------ a5: isolated training with the reference
    models = {}
    train_log = {}
    for w in S.WORDS:
        n_states = S.STATES_PER_WORD[w]
        m = R.HiddenMarkovModelTrainable.from_data(w, feats_train[w], num_of_states=n_states, max_iterations=4,
                                                   isMultiProcessingTraining=False, isTqdm=False)
        models[w] = m
        train_log[w] = m
    model_dir = os.path.join(HERE, "golden_models")
    shutil.rmtree(model_dir, ignore_errors=True)
    for w in ("1", "S", "Z"):
        models[w].save(model_dir)

    # oracle restatement of the same training run (iteration-exact)
    for w in S.WORDS:
        n_states = S.STATES_PER_WORD[w]
        feats = feats_train[w]
        means, covs, trans = O.init_parameters(feats[0], n_states)
        for it in range(4):
            packs = [O.gaussian_pack(means[s], covs[s]) for s in range(n_states)]
            tr = O.word_trellis(O.log_transitions(trans))
            paths = []
            for x in feats:
                sc = O.emission_scores(x, [p[0] for p in packs], [p[1] for p in packs], [p[2] for p in packs])
                paths.append(O.viterbi(sc, tr)[2])
            r = O.mstep(feats, paths, n_states, old_means=means)
            if r["converged"]:
                break
            means, covs, trans = r["means"], r["covs"], r["trans"]
        ref = models[w]
        assert np.array_equal(means, ref._means), w
        assert np.array_equal(covs, ref._covariances), w
        assert np.array_equal(O.log_transitions(trans), dense_logA(ref._log_transition_probs), equal_nan=True), w
        out[f"train_means_{w}"] = ref._means
        out[f"train_covs_{w}"] = ref._covariances
        out[f"train_logA_{w}"] = dense_logA(ref._log_transition_probs)
    print("isolated training: oracle == reference (bit-exact) for", len(S.WORDS), "words")
    for w in S.WORDS:
        for i, x in enumerate(feats_train[w]):
            out[f"train_feat_{w}_{i}"] = np.ascontiguousarray(x)

    # ------------------------------------------------------------------ a2/a3: isolated decode
    flat = {w: flat_model(models[w]) for w in S.WORDS}
    mc = R.ModelCollection()
    mc._models = [models[w] for w in R.TI_DIGITS_LABELS]
    iso_feats, iso_scores, iso_paths, iso_labels, iso_truth = [], [], [], [], []
    for w in S.DIGITS:
        for x in feats_test[w]:
            sc_row, path_row = [], []
            for lab in R.TI_DIGITS_LABELS:
                score, path = models[lab].predict(x)
                means, Us, lp, _, logA = flat[lab]
                em = O.emission_scores(x, means, Us, lp)
                # emission: oracle vs per-frame scipy calls
                ref_em = np.array([[mn.log_pdf(f) for mn in models[lab]._multivariate_normals] for f in x[:5]])
                assert np.array_equal(em[:5], ref_em)
                es, bi, opath = O.viterbi(em, O.word_trellis(logA))
                assert es[0] == score and es[0].dtype == np.float32 and np.array_equal(opath, path), (w, lab)
                sc_row.append(score); path_row.append(path)
            iso_feats.append(np.ascontiguousarray(x)); iso_scores.append(sc_row); iso_paths.append(path_row)
            iso_labels.append(mc.predict(x)); iso_truth.append(w)
    print("isolated decode: oracle == reference;  accuracy", np.mean([a == b for a, b in zip(iso_labels, iso_truth)]))
    for i, x in enumerate(iso_feats):
        out[f"iso_feat_{i}"] = x
        out[f"iso_paths_{i}"] = np.array(iso_paths[i])
    out["iso_scores"] = np.array(iso_scores, dtype=np.float32)
    out["iso_labels"] = np.array(iso_labels)
    out["iso_truth"] = np.array(iso_truth)
    out["iso_model_order"] = np.array(list(R.TI_DIGITS_LABELS))

    # ------------------------------------------------------------------ a4: loop decode (all 12 models in sorted order)
    full_dir = "/tmp/loe_golden_full_models"
    shutil.rmtree(full_dir, ignore_errors=True)
    for w in S.WORDS:
        models[w].save(full_dir)
    inf = R.HiddenMarkovModelInference.from_folder(full_dir, list(S.WORDS))
    order = inf._model_boundaries._labels
    sizes = [S.STATES_PER_WORD[w] for w in order]
    out["loop_order"] = np.array(order)
    tr = O.loop_trellis([flat[w][4] for w in order])
    means = np.concatenate([flat[w][0] for w in order]); Us = np.concatenate([flat[w][1] for w in order])
    lps = np.concatenate([flat[w][2] for w in order])
    penalties = {"int": -100, "f64": np.log(0.005), "pyfloat": -37.25, "zero": 0}
    all_feats = feats_str + feats_short
    out["loop_truth"] = np.array(truth + short_truth)
    for i, x in enumerate(all_feats):
        out[f"loop_feat_{i}"] = np.ascontiguousarray(x)
    for name, pen in penalties.items():
        inf._log_transition_probability_between_words = pen
        strs, paths, scores = [], [], []
        for i, x in enumerate(all_feats):
            score, path = inf._viterbi(x)
            s = inf.predict(x)
            em = O.emission_scores(x, means, Us, lps)
            es, bi, opath = O.viterbi(em, tr, penalty=pen)
            assert es[bi] == score and np.array_equal(opath, path), (name, i)
            assert "".join(O.get_labels(opath, sizes, order)) == s
            strs.append(s); paths.append(path); scores.append(score)
        # batched oracle == per-utterance oracle
        ems = [O.emission_scores(x, means, Us, lps) for x in all_feats]
        bes, bbi, bpaths = O.viterbi_batch(ems, tr, penalty=pen)
        for i in range(len(all_feats)):
            assert np.array_equal(bpaths[i], paths[i]) and bes[i, bbi[i]] == scores[i]
        out[f"loop_strings_{name}"] = np.array(strs)
        out[f"loop_scores_{name}"] = np.array(scores, dtype=np.float32)
        for i, p in enumerate(paths):
            out[f"loop_path_{name}_{i}"] = p
        print(f"loop decode [{name}]: oracle == reference; strings", strs[:3], "truth", (truth + short_truth)[:3])

    # short utterances / edge cases (T = 2, 3, 9) through the loop and word decoders
    for T in (2, 3, 9):
        x = np.ascontiguousarray(all_feats[0][40:40 + T])
        inf._log_transition_probability_between_words = -100
        score, path = inf._viterbi(x)
        em = O.emission_scores(x, means, Us, lps)
        es, bi, opath = O.viterbi(em, tr, penalty=-100)
        assert np.array_equal(opath, path) and (es[bi] == score or (np.isinf(score) and np.isinf(es[bi])))
        out[f"edge_loop_path_T{T}"] = path
        out[f"edge_loop_score_T{T}"] = np.float32(score)
        score, path = models["1"].predict(x)
        m1 = flat["1"]
        es, bi, opath = O.viterbi(O.emission_scores(x, m1[0], m1[1], m1[2]), O.word_trellis(m1[4]))
        assert np.array_equal(opath, path) and (es[0] == score or (np.isinf(score) and np.isinf(es[0])))
        out[f"edge_word_path_T{T}"] = path
        out[f"edge_word_score_T{T}"] = np.float32(score)
    out["edge_feat"] = np.ascontiguousarray(all_feats[0][40:49])
    print("edge cases: oracle == reference")

    # ------------------------------------------------------------------ a6: embedded training (2 iterations)
    emb = R.HiddenMarkovModelTrainContinuous.from_folder(full_dir, list(S.WORDS))
    emb.isMultiProcessing = False
    emb.isTqdm = False
    erng = np.random.default_rng(30)      # every digit must occur, else the reference raises HMMTrainMeanFail
    emb_truth = [S.DIGITS[i] + S.DIGITS[(i + 3) % 11] + S.DIGITS[(i + 7) % 11] for i in range(11)] + \
                [S.DIGITS[i] + S.DIGITS[(i + 5) % 11] for i in range(11)] + ["12", "12"]
    emb_strings = [S.synth_string(erng, list(t)) for t in emb_truth]
    labeled = {}
    for lab, pcm in zip(emb_truth, emb_strings):
        labeled.setdefault(lab, []).append(OM.mfcc_feature_vector(pcm).T)
    # oracle restatement: chain alignment + remux for iteration 1
    wm = {w: flat[w] for w in S.WORDS}
    pooled = {w: [] for w in S.WORDS}
    for lab, xs in labeled.items():
        chain = O.insert_silence(lab)
        assert chain == emb.insert_silence(lab)
        csizes = [S.STATES_PER_WORD[c] for c in chain]
        ctr = O.chain_trellis([wm[c][4] for c in chain])
        cm = np.concatenate([wm[c][0] for c in chain]); cu = np.concatenate([wm[c][1] for c in chain])
        cl = np.concatenate([wm[c][2] for c in chain])
        ref_chain = HiddenMarkovModelMultiWord.from_labels(chain, emb._trainable_models)
        for x in xs:
            _, rpath = ref_chain._viterbi(x)
            es, bi, opath = O.viterbi(O.emission_scores(x, cm, cu, cl), ctr)
            assert np.array_equal(opath, rpath), lab
            ref_mux = ref_chain._remux_path_and_signal(x, rpath, ref_chain._model_boundaries)
            omux = O.remux(x, opath, csizes, list(chain))
            for w in omux:
                assert len(omux[w]) == len(ref_mux[w])
                for (seg, p, n), rs in zip(omux[w], ref_mux[w]):
                    assert np.array_equal(seg, rs.signal) and np.array_equal(p, rs.path) and n == rs.num_of_state
                pooled[w].extend(omux[w])
    emb.train(labeled, max_iterations=1)
    for w in S.WORDS:
        r = O.mstep([s for s, _, _ in pooled[w]], [p for _, p, _ in pooled[w]], S.STATES_PER_WORD[w],
                    old_means=np.zeros_like(flat[w][0], dtype=np.float32))
        tm = emb._trainable_models[w]
        assert np.array_equal(r["means"], tm._means) and np.array_equal(r["covs"], tm._covariances), w
        assert np.array_equal(O.log_transitions(r["trans"]), dense_logA(tm._log_transition_probs), equal_nan=True)
        out[f"emb1_means_{w}"] = tm._means
        out[f"emb1_covs_{w}"] = tm._covariances
        out[f"emb1_logA_{w}"] = dense_logA(tm._log_transition_probs)
    print("embedded training iteration 1: oracle == reference (bit-exact)")
    out["emb_labels"] = np.array(list(labeled.keys()))
    for lab, xs in labeled.items():
        for i, x in enumerate(xs):
            out[f"emb_feat_{lab}_{i}"] = np.ascontiguousarray(x)
        out[f"emb_count_{lab}"] = np.int32(len(xs))

    np.savez_compressed(os.path.join(HERE, "golden_hmm.npz"), **out)
    print("wrote", os.path.join(HERE, "golden_hmm.npz"))


if __name__ == "__main__":
    main()
