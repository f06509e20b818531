"""CPU: the NumPy oracle against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  Bit-exact everywhere: this is what pins the oracle."""
import numpy as np

from helpers import LOOP_ORDER, N_STATES, WORDS, oracle_flat
from oracle import hmm as O

PENALTIES = {"int": -100, "f64": np.log(0.005), "pyfloat": -37.25, "zero": 0}


def _loop_tables(golden):
    flat = [oracle_flat(golden, w) for w in LOOP_ORDER]
    return (np.concatenate([f[0] for f in flat]), np.concatenate([f[1] for f in flat]), np.concatenate([f[2] for f in flat]),
            O.loop_trellis([f[3] for f in flat]), [f[3].shape[0] for f in flat])


def test_isolated_scores_and_paths(golden):
    order = [str(s) for s in golden["iso_model_order"]]
    for i in range(0, 22, 3):
        x = golden[f"iso_feat_{i}"]
        for k, w in enumerate(order):
            means, Us, lps, logA = oracle_flat(golden, w)
            es, bi, path = O.viterbi(O.emission_scores(x, means, Us, lps), O.word_trellis(logA))
            assert es[0] == golden["iso_scores"][i, k]
            assert np.array_equal(path, golden[f"iso_paths_{i}"][k])
    best = np.argmax(golden["iso_scores"], axis=1)
    assert [order[b] for b in best] == [str(s) for s in golden["iso_labels"]]


def test_loop_decode_all_penalty_modes(golden):
    means, Us, lps, tr, sizes = _loop_tables(golden)
    assert [str(s) for s in golden["loop_order"]] == list(LOOP_ORDER)
    ems = [O.emission_scores(golden[f"loop_feat_{i}"], means, Us, lps) for i in range(10)]
    for name, pen in PENALTIES.items():
        assert O.penalty_mode(pen)[1] == (name == "f64")
        for i in (0, 3, 6, 9):
            es, bi, path = O.viterbi(ems[i], tr, penalty=pen)
            assert es[bi] == golden[f"loop_scores_{name}"][i]
            assert np.array_equal(path, golden[f"loop_path_{name}_{i}"])
            assert "".join(O.get_labels(path, sizes, list(LOOP_ORDER))) == str(golden[f"loop_strings_{name}"][i])
        bes, bbi, bpaths = O.viterbi_batch(ems, tr, penalty=pen)
        for i in range(10):
            assert np.array_equal(bpaths[i], golden[f"loop_path_{name}_{i}"])
            assert bes[i, bbi[i]] == golden[f"loop_scores_{name}"][i]


def test_edge_cases(golden):
    means, Us, lps, tr, sizes = _loop_tables(golden)
    m1 = oracle_flat(golden, "1")
    for T in (2, 3, 9):
        x = golden["edge_feat"][:T]
        es, bi, path = O.viterbi(O.emission_scores(x, means, Us, lps), tr, penalty=-100)
        assert np.array_equal(path, golden[f"edge_loop_path_T{T}"])
        es, bi, path = O.viterbi(O.emission_scores(x, m1[0], m1[1], m1[2]), O.word_trellis(m1[3]))
        assert np.array_equal(path, golden[f"edge_word_path_T{T}"])
        ref = golden[f"edge_word_score_T{T}"]
        assert es[0] == ref or (np.isinf(ref) and np.isinf(es[0]))
    es, bi, path = O.viterbi(O.emission_scores(golden["edge_feat"][:1], m1[0], m1[1], m1[2]), O.word_trellis(m1[3]))
    assert path.tolist() == [-1]


def test_isolated_training_trajectory(golden):
    for w in ("1", "S", "Z"):
        # MFCC.batch hands out TRANSPOSED VIEWS (mfcc.py:84), i.e. column-major (T, 39) arrays, and the
        # reference's float32 np.average depends on that layout at the 1-ulp level (pairwise vs row-wise
        # accumulation) -- restore it for a bit-exact replay.
        feats = [np.asfortranarray(golden[f"train_feat_{w}_{i}"]) for i in range(64) if f"train_feat_{w}_{i}" in golden.files]
        n_states = N_STATES[w]
        means, covs, trans = O.init_parameters(feats[0], n_states)
        for it in range(4):
            packs = [O.gaussian_pack(means[s], covs[s]) for s in range(n_states)]
            tr = O.word_trellis(O.log_transitions(trans))
            paths = [O.viterbi(O.emission_scores(x, [p[0] for p in packs], [p[1] for p in packs], [p[2] for p in packs]), tr)[2]
                     for x in feats]
            r = O.mstep(feats, paths, n_states, old_means=means)
            if r["converged"]:
                break
            means, covs, trans = r["means"], r["covs"], r["trans"]
        assert np.array_equal(means, golden[f"train_means_{w}"])
        assert np.array_equal(covs, golden[f"train_covs_{w}"])
        assert np.array_equal(O.log_transitions(trans), golden[f"train_logA_{w}"], equal_nan=True)


def test_embedded_iteration(golden):
    flat = {w: oracle_flat(golden, w) for w in WORDS}
    pooled = {w: [] for w in WORDS}
    for lab in [str(s) for s in golden["emb_labels"]]:
        chain = O.insert_silence(lab)
        sizes = [N_STATES[c] for c in chain]
        tr = O.chain_trellis([flat[c][3] for c in chain])
        cm = np.concatenate([flat[c][0] for c in chain]); cu = np.concatenate([flat[c][1] for c in chain])
        cl = np.concatenate([flat[c][2] for c in chain])
        for i in range(int(golden[f"emb_count_{lab}"])):
            x = np.asfortranarray(golden[f"emb_feat_{lab}_{i}"])
            _, _, path = O.viterbi(O.emission_scores(x, cm, cu, cl), tr)
            for w, segs in O.remux(x, path, sizes, list(chain)).items():
                pooled[w].extend(segs)
    for w in WORDS:
        r = O.mstep([s for s, _, _ in pooled[w]], [p for _, p, _ in pooled[w]], N_STATES[w],
                    old_means=np.zeros((N_STATES[w], 39), np.float32))
        assert np.array_equal(r["means"], golden[f"emb1_means_{w}"])
        assert np.array_equal(r["covs"], golden[f"emb1_covs_{w}"])
        assert np.array_equal(O.log_transitions(r["trans"]), golden[f"emb1_logA_{w}"], equal_nan=True)


def test_mfcc_oracle_regression(golden_mfcc):
    """oracle/mfcc.py is parity-unpinned (librosa absent); this pins it against its own committed
    output and against independent restatements of two stages."""
    from oracle import mfcc as OM
    for i in range(4):
        assert np.allclose(OM.mfcc_feature_vector(golden_mfcc[f"pcm{i}"]), golden_mfcc[f"feat{i}"], rtol=1e-6, atol=1e-6)
    assert np.array_equal(OM.mel_basis(), golden_mfcc["mel_basis"])
    mb = OM.mel_basis()
    assert mb.shape == (40, 161) and np.all((mb > 0).sum(1) >= 2) and np.all((mb > 0).sum(1) <= 18)
    # delta taps and edge rule
    rng = np.random.default_rng(0)
    c = rng.normal(size=(13, 40)).astype(np.float32)
    k = np.arange(-4, 5)
    d1, d2 = OM.delta(c, 1), OM.delta(c, 2)
    for t in (4, 17, 35):
        assert np.allclose(d1[:, t], (c[:, t - 4:t + 5] * k).sum(1) / 60, atol=1e-5)
        assert np.allclose(d2[:, t], (c[:, t - 4:t + 5] * (3 * k * k - 20)).sum(1) / 462, atol=1e-5)
    assert np.allclose(d1[:, 0], d1[:, 4], atol=1e-5) and np.allclose(d2[:, -1], d2[:, -5], atol=1e-5)
    T = 1 + 16000 // 160
    assert golden_mfcc["feat0"].shape == (39, T)


def test_mfcc_oracle_against_independent_librosa_restatements():
    """The MFCC oracle is "parity unpinned" (librosa is absent and un-pinned upstream).  Second opinions that ARE
    installed: transformers.audio_utils and torchaudio both ship restatements of the same librosa recipe
    (slaney mel scale + slaney area normalisation, centred periodic-Hann STFT, power_to_db, ortho DCT-II).  The
    oracle's stages must agree with them on a seeded signal -- this pins the recipe's conventions (filter edges
    and normalisation, window periodicity, zero centre padding, dB reference / floor, DCT scaling), not librosa's
    last-ulp arithmetic."""
    import pytest
    from oracle import mfcc as OM
    AU = pytest.importorskip("transformers.audio_utils")
    rng = np.random.default_rng(4)
    t = np.arange(16000)
    y = (3000 * np.sin(2 * np.pi * 440 * t / 16000) + 1500 * np.sin(2 * np.pi * 2300 * t / 16000) + rng.normal(0, 30, t.size)).astype(np.float32)
    # mel filterbank
    fb = AU.mel_filter_bank(num_frequency_bins=161, num_mel_filters=40, min_frequency=OM.FMIN, max_frequency=OM.FMAX,
                            sampling_rate=16000, norm="slaney", mel_scale="slaney")
    mine = OM.mel_basis(16000)
    assert fb.T.shape == mine.shape and np.abs(fb.T - mine).max() < 1e-7 * np.abs(mine).max()
    # |STFT|^2 -> mel -> dB
    win = AU.window_function(OM.N_FFT, "hann", periodic=True)
    S = AU.spectrogram(y, win, frame_length=OM.N_FFT, hop_length=OM.HOP, fft_length=OM.N_FFT, power=2.0, center=True,
                       pad_mode="constant", mel_filters=fb, mel_floor=0.0)
    mel = np.einsum("ft,mf->mt", OM.stft_power(y), mine)
    assert S.shape == mel.shape == (40, 101) and np.abs(S - mel).max() < 1e-5 * np.abs(mel).max()
    db = AU.power_to_db(S, reference=S.max(), min_value=1e-10, db_range=80.0)
    lm = OM.power_to_db(mel)
    assert np.abs(db - lm).max() < 1e-3 and lm.max() == 0.0 and lm.min() >= -80.0
    # DCT-II (ortho), 13 cepstra: torchaudio's matrix
    torchaudio = pytest.importorskip("torchaudio")
    dct = torchaudio.functional.create_dct(OM.N_MFCC, OM.N_MELS, norm="ortho").numpy().T          # (13, 40)
    import scipy.fft
    ceps = scipy.fft.dct(lm, axis=-2, type=2, norm="ortho")[:OM.N_MFCC]
    assert np.abs(dct @ lm - ceps).max() < 1e-3
    # and torchaudio's own mel spectrogram with the librosa-compatible switches
    import torch
    ms = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=OM.N_FFT, win_length=OM.N_FFT, hop_length=OM.HOP,
                                              f_min=OM.FMIN, f_max=OM.FMAX, n_mels=OM.N_MELS, power=2.0, center=True,
                                              pad_mode="constant", norm="slaney", mel_scale="slaney")(torch.from_numpy(y)).numpy()
    assert ms.shape == mel.shape and np.abs(ms - mel).max() < 1e-4 * np.abs(mel).max()


def test_parameterised_mfcc_reduces_to_reference():
    """oracle.mfcc.mfcc_feature_vector_ex (the configs[3] front end, parity unpinned by construction) with the
    reference's parameter set IS the pinned-as-far-as-possible restatement; the "spec" set changes what it should."""
    from oracle import mfcc as OM
    rng = np.random.default_rng(4)
    y = np.round(rng.normal(0, 2000, size=24000)).astype(np.float32)
    assert np.array_equal(OM.mfcc_feature_vector_ex(y, 16000, OM.REFERENCE_CONFIG), OM.mfcc_feature_vector(y).astype(np.float32))
    spec = OM.mfcc_feature_vector_ex(y, 16000, OM.SPEC_CONFIG)
    assert spec.shape == (39, 151) and spec.dtype == np.float32
    assert np.abs(spec[:13].mean(axis=1)).max() < 1e-4                       # CMN: zero mean over time per coefficient
    e = OM.preemphasis(y, 0.97)
    assert e[0] == np.float32(y[0] - np.float32(0.97) * y[0]) and e[5] == np.float32(y[5] - np.float32(0.97) * y[4])
    # a 400-sample Hamming window centred in a 512-point FFT: frame t only sees samples [160 t - 200, 160 t + 200)
    y2 = y.copy(); y2[160 * 50 + 201:160 * 50 + 250] += 500.0
    p1 = OM.stft_power(y, 512, 160, "hamming", 400); p2 = OM.stft_power(y2, 512, 160, "hamming", 400)
    assert np.array_equal(p1[:, 50], p2[:, 50]) and not np.array_equal(p1[:, 51], p2[:, 51])


def test_mfcc_oracle_against_librosa_fixture():
    """Row a1's pin, wherever it can be had: tests/golden/make_golden_mfcc_librosa.py writes golden_mfcc_librosa.npz on a box
    with librosa (the reference's exact calls, mfcc.py:31-43); the oracle must reproduce it.  Skipped while the file does not
    exist (librosa is not installable in the authoring container: parity of a1 stays 'unpinned' until someone runs it)."""
    import os
    import pytest
    from oracle import mfcc as OM
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_mfcc_librosa.npz")
    if not os.path.exists(path):
        pytest.skip("no librosa fixture (run tests/golden/make_golden_mfcc_librosa.py where librosa is installed)")
    z = np.load(path, allow_pickle=False)
    i = 0
    while f"pcm{i}" in z.files:
        got = OM.mfcc_feature_vector(z[f"pcm{i}"])
        ref = z[f"feat{i}"]
        assert got.shape == ref.shape
        assert np.all(np.abs(got - ref) <= 1e-4 * np.abs(ref) + 1e-4 * np.abs(ref[:13]).max()), (i, np.abs(got - ref).max())
        i += 1
    assert i > 0
