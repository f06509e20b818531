"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden
fixtures produced by the unmodified reference.  Bars (BASELINE.json north_star): features and
log-likelihoods within 1e-4 relative; state paths and digit strings bit-exact given the same
emission scores, and on every utterance whose margin exceeds the float tolerance otherwise."""
import numpy as np
import pytest

from helpers import LOOP_ORDER, N_STATES, WORDS, emission_close, oracle_flat, rel_close, trained_word_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(built_lib):
    from loe_speech_recognition._engine import get_engine
    return get_engine()


# ------------------------------------------------------------------ a1 MFCC
def test_mfcc_matches_oracle_golden(eng, golden_mfcc):
    from loe_speech_recognition import MFCC
    pcms = [golden_mfcc[f"pcm{i}"] for i in range(4)]
    got = MFCC.batch(pcms, sample_rate=16000)
    for i, g in enumerate(got):
        ref = golden_mfcc[f"feat{i}"].T
        assert g.shape == ref.shape and g.dtype == np.float32
        atol = 1e-4 * np.abs(ref[:, :13]).max()          # floor for delta terms near zero
        assert rel_close(g, ref, rtol=1e-4, atol=atol), (i, np.abs(g - ref).max())
    single = MFCC(pcms[0], 16000).feature_vector
    assert single.shape == (39, got[0].shape[0])
    assert np.array_equal(single.T, got[0])


def test_mfcc_ragged_batch_and_errors(eng):
    from loe_speech_recognition import MFCC
    from oracle import mfcc as OM
    rng = np.random.default_rng(3)
    lens = [1440, 1441, 1599, 1600, 5000, 16000, 64000, 12345]
    pcms = [rng.normal(0, 1000, size=n).round().astype(np.float32) for n in lens]
    got = MFCC.batch(pcms, 16000)
    for p, g in zip(pcms, got):
        ref = OM.mfcc_feature_vector(p).T
        assert g.shape == ref.shape
        assert rel_close(g, ref, rtol=1e-4, atol=1e-4 * np.abs(ref[:, :13]).max())
    with pytest.raises(ValueError):
        MFCC.batch([np.zeros(1000, np.float32)], 16000)      # 7 frames < 9: savgol_filter raises in the reference
    with pytest.raises(TypeError):
        MFCC([1.0, 2.0], 16000)
    with pytest.raises(ValueError):
        MFCC(np.zeros((2, 2000), np.float32), 16000)
    assert MFCC.batch([], 16000) == []
    z = MFCC.batch([np.zeros(4000, np.float32)], 16000)[0]   # digital silence: all-equal dB -> zero cepstra
    assert np.all(np.isfinite(z))


def test_mfcc_independent_of_batch_composition(eng):
    """The mel kernel sizes its frame chunks from the batch and picks two-sample loads from the parity of an
    utterance's first sample: the features of an utterance must not depend on either -- bitwise -- nor on whether
    the PCM arrives as float32 or int16, and the big batch must still agree with the oracle."""
    from loe_speech_recognition import MFCC
    from oracle import mfcc as OM
    rng = np.random.default_rng(11)
    t = np.arange(70000)
    base = []
    for L in (1441, 16000, 23457, 46401, 64000, 69999, 30000, 52000):       # odd lengths shift later utterances to odd starts
        f = rng.uniform(200, 4000, 3)
        sig = sum(3000 * np.sin(2 * np.pi * fi * t[:L] / 16000 + rng.uniform(0, 6)) for fi in f) + rng.normal(0, 30, L)
        base.append(np.round(sig).astype(np.int16))
    alone = [MFCC.batch([b.astype(np.float32)], 16000)[0] for b in base]
    for b, a in zip(base[:3], alone[:3]):
        ref = OM.mfcc_feature_vector(b.astype(np.float32)).T
        assert rel_close(a, ref, rtol=1e-4, atol=1e-4 * np.abs(ref[:, :13]).max())
    big = [base[i % len(base)] for i in range(1200)]                            # ~330 k frames: chunks longer than the minimum
    assert sum(1 + len(b) // 160 for b in big) > 12 * 148 * 160
    for dtype in (np.float32, np.int16):
        got = MFCC.batch([b.astype(dtype) for b in big], 16000)
        for i, g in enumerate(got):
            assert np.array_equal(g, alone[i % len(base)]), (dtype, i)


def test_kernels_stay_inside_their_buffers(eng, golden):
    """Guard bands around every output of the decode path (mel workspace, utterance maxima, features, scores)
    keep their sentinel: ragged utterances, a partial last emission tile, float32 and int16 PCM."""
    torch = eng.torch
    from loe_speech_recognition.synthetic import string_corpus
    utts, _ = string_corpus(seed=77, n_utts=9, n_digits=3)
    utts = [u[: len(u) - 53 * i] for i, u in enumerate(utts)] + [utts[0][:1441]]
    inf = _loop_inference(golden)
    gp, _ = inf._packs()
    G = 4096
    for dtype in (np.float32, np.int16):
        pcm, pcm_off, frm_off_dev, frm_off, frames = eng.upload_pcm([np.round(u).astype(dtype) for u in utts])
        n, F = len(utts), int(frm_off[-1])

        def banded(rows, cols):
            size = rows * max(cols, 1)
            t = torch.full((G + size + G,), -12345.0, dtype=torch.float32, device=eng.device)
            return t, (t[G:G + size].view(rows, cols) if cols else t[G:G + size])
        mel_all, mel = banded(F, 40)
        um_all, um = banded(n, 0)
        feat_all, feat = banded(F, 39)
        sc_all, sc = banded(F, gp.n_states)
        eng.mfcc_device(pcm, pcm_off, frm_off_dev, n, F, int(frames.max()), int(frames.min()), 16000, out=feat, mel_ws=mel, utt_max=um)
        for prec in ("h16", "tc", "fp32"):
            eng.emission(feat, gp, prec, out=sc)
            torch.cuda.synchronize()
            for whole, inner in ((mel_all, F * 40), (um_all, n), (feat_all, F * 39), (sc_all, F * gp.n_states)):
                assert bool((whole[:G] == -12345.0).all()) and bool((whole[G + inner:] == -12345.0).all()), (dtype, prec)
        assert bool(torch.isfinite(feat).all()) and bool(torch.isfinite(sc).all()) and bool((feat != -12345.0).any())


def test_parameterised_mfcc_matches_oracle(eng):
    """g2 / BASELINE configs[3]: the parameterised front end (loe_mfcc_ex_dev) against the restated oracle
    (parity unpinned by construction -- no live reference call site) for the "spec" set and variations of every
    parameter; float32 and int16 PCM bit-identical; ragged batch; errors."""
    from dataclasses import asdict, replace
    from loe_speech_recognition import MFCC, MFCCConfig
    from oracle import mfcc as OM
    rng = np.random.default_rng(21)
    t = np.arange(70000)
    sigs = []
    for L in (1600, 1441, 16000, 23457, 64000, 5000, 30001):
        f = rng.uniform(200, 4000, 3)
        sig = sum(3000 * np.sin(2 * np.pi * fi * t[:L] / 16000 + rng.uniform(0, 6)) for fi in f) + rng.normal(0, 30, L)
        sigs.append(np.round(sig).astype(np.int16))
    spec = MFCCConfig.spec()
    configs = [spec, replace(spec, log="db", norm="frame"), replace(spec, norm="cmvn"), replace(spec, norm="none", preemphasis=0.0),
               MFCCConfig(n_fft=256, win_length=256, hop_length=128), MFCCConfig(n_fft=1024, win_length=800, hop_length=320, window="hamming"),
               MFCCConfig(n_fft=512, win_length=512, hop_length=160), MFCCConfig(n_fft=64, win_length=64, hop_length=32, n_mels=20, fmax=7000.0)]
    for cfg in configs:
        use = [s for s in sigs if 1 + len(s) // cfg.hop_length >= 9]
        got = MFCC.batch([s.astype(np.float32) for s in use], 16000, config=cfg)
        got16 = MFCC.batch(use, 16000, config=cfg)
        for s, g, g16 in zip(use, got, got16):
            ref = OM.mfcc_feature_vector_ex(s.astype(np.float32), 16000, asdict(cfg)).T
            assert g.shape == ref.shape and g.dtype == np.float32, (cfg, g.shape, ref.shape)
            atol = 1e-4 * max(np.abs(ref[:, :13]).max(), 1.0)
            assert rel_close(g, ref, rtol=1e-4, atol=atol), (cfg, len(s), np.abs(g - ref).max())
            assert np.array_equal(g, g16), cfg
    # the reference's own parameter set through config= takes the specialised kernels
    a = MFCC.batch([sigs[2].astype(np.float32)], 16000, config=MFCCConfig())[0]
    assert np.array_equal(a, MFCC.batch([sigs[2].astype(np.float32)], 16000)[0])
    with pytest.raises(NotImplementedError):
        MFCC.batch([sigs[2].astype(np.float32)], 16000, config=MFCCConfig(n_fft=400, win_length=400))
    with pytest.raises(ValueError):
        MFCC.batch([np.zeros(1000, np.float32)], 16000, config=spec)          # 7 frames < 9
    # bitwise reproducible (fixed-order CMN sums)
    b1 = MFCC.batch(sigs, 16000, config=spec); b2 = MFCC.batch(sigs, 16000, config=spec)
    assert all(np.array_equal(x, y) for x, y in zip(b1, b2))


# ------------------------------------------------------------------ a2 emission
@pytest.mark.parametrize("precision,rtol", [("fp32", 1e-4), ("fp64", 1e-6), ("tc", 1e-4), ("h16", 1e-4)])
def test_emission_matches_scipy(eng, golden, precision, rtol):
    from oracle import hmm as O
    x = np.concatenate([golden[f"iso_feat_{i}"] for i in range(6)])
    for w in ("1", "S", "Z"):
        m = trained_word_model(golden, w)
        gp = eng.pack_gaussians(m._multivariate_normals)
        got = eng.emission(eng._to_dev(x), gp, precision).cpu().numpy()
        means, Us, lps, _ = oracle_flat(golden, w)
        ref = O.emission_scores(x, means, Us, lps)
        assert got.shape == ref.shape
        if precision == "fp64":
            assert rel_close(got, ref, rtol=rtol), np.abs(got / ref - 1).max()
        else:
            assert emission_close(got, ref, lps), np.abs(got - ref).max()
        if precision == "fp64":
            assert np.mean(got == ref) > 0.999       # float64 arithmetic, one rounding: bit-identical almost everywhere


def test_emission_in_context_features(eng, golden):
    """All 58 states on string features: scores span -3e5 .. +50 and cross zero."""
    from oracle import hmm as O
    flat = [oracle_flat(golden, w) for w in LOOP_ORDER]
    means = np.concatenate([f[0] for f in flat]); Us = np.concatenate([f[1] for f in flat]); lps = np.concatenate([f[2] for f in flat])
    inf = _loop_inference(golden)
    gp, _ = inf._packs()
    for i in (0, 7):
        x = golden[f"loop_feat_{i}"]
        ref = O.emission_scores(x, means, Us, lps)
        assert emission_close(eng.emission(eng._to_dev(x), gp, "fp32").cpu().numpy(), ref, lps)
        assert emission_close(eng.emission(eng._to_dev(x), gp, "tc").cpu().numpy(), ref, lps)
        got64 = eng.emission(eng._to_dev(x), gp, "fp64").cpu().numpy()
        assert rel_close(got64, ref, rtol=1e-6) and np.mean(got64 == ref) > 0.999


def test_emission_ill_conditioned(eng):
    """Covariances with condition number up to 1e6 (only +1e-3 I regularises real ones)."""
    from oracle import hmm as O
    from loe_speech_recognition.hidden_markov_model import MultivariateNormal
    rng = np.random.default_rng(11)
    D = 39
    q, _ = np.linalg.qr(rng.normal(size=(D, D)))
    lam = np.logspace(-3, 3, D)
    cov = ((q * lam) @ q.T).astype(np.float32)
    cov = (cov + cov.T) / 2
    mean = rng.normal(0, 3, size=D).astype(np.float32)
    mn = MultivariateNormal.from_means_covariances(mean, cov)
    x = (mean + rng.normal(size=(300, D)) * np.sqrt(lam).mean()).astype(np.float32)
    gp = eng.pack_gaussians([mn])
    pk = O.gaussian_pack(mean, cov)
    ref = O.emission_scores(x, *[np.array([v]) for v in pk])
    for precision in ("fp32", "fp64", "tc", "h16"):
        got = eng.emission(eng._to_dev(x), gp, precision).cpu().numpy()
        assert emission_close(got, ref, [pk[2]]), (precision, np.abs(got / ref - 1).max())


def test_emission_tc_many_tiles_and_ragged_sizes(eng, golden):
    """Tensor-core path: frame counts around the 128-row tile edge, state counts around the 6-state
    tile edge, more frame tiles than CTAs (pipeline phases wrap), against the fp64 kernel."""
    rng = np.random.default_rng(5)
    inf = _loop_inference(golden)
    normals = inf._multivariate_normals
    base = np.concatenate([golden[f"loop_feat_{i}"] for i in range(10)])
    for n_states in (1, 5, 6, 7, 12, 58):
        gp = eng.pack_gaussians(normals[:n_states])
        lps = [mn._core.cov_object._log_pdet for mn in normals[:n_states]]
        for n_frames in (1, 127, 128, 129, 1000, 40000):
            x = base[rng.integers(0, len(base), size=n_frames)]
            xd = eng._to_dev(x)
            ref = eng.emission(xd, gp, "fp64").cpu().numpy()
            for kind in ("tc", "h16"):
                got = eng.emission(xd, gp, kind).cpu().numpy()
                assert emission_close(got, ref, lps), (kind, n_states, n_frames, np.abs(got - ref).max())
    # a feature buffer that is only 4-byte aligned takes the non-TMA loads; same values
    gp = eng.pack_gaussians(normals)
    x = eng._to_dev(base[:1000])
    shifted = eng.torch.empty(1000 * 39 + 1, dtype=eng.torch.float32, device=eng.device)[1:].view(1000, 39)
    shifted.copy_(x)
    assert shifted.data_ptr() % 16 != 0
    assert eng.torch.equal(eng.emission(shifted, gp, "tc"), eng.emission(x, gp, "tc"))
    assert eng.torch.equal(eng.emission(shifted, gp, "h16"), eng.emission(x, gp, "h16"))


def test_emission_from_presplit_image_is_bit_identical(eng, golden):
    """Decode hand-off: the cepstrum kernel writes the 3xFP16 kernel's A operand (loe_mfcc_img_dev) and the emission kernel
    bulk-copies it (loe_emission_h16_img_dev).  Same split arithmetic, same MMAs: scores bit-identical to the kernel that
    splits the float32 features itself; the features written alongside are unchanged; ragged batch, partial last tile,
    more tiles than CTAs, guard bands, and the image path without the float32 matrix."""
    torch = eng.torch
    from loe_speech_recognition.synthetic import string_corpus
    inf = _loop_inference(golden)
    gp, _ = inf._packs()
    for n_utts, n_digits in ((3, 2), (40, 7), (700, 7)):
        utts, _ = string_corpus(seed=5 + n_utts, n_utts=n_utts, n_digits=n_digits)
        utts = [u[: len(u) - 37 * (i % 5)] for i, u in enumerate(utts)]
        pcm, pcm_off, frm_off_dev, frm_off, frames = eng.upload_pcm(utts)
        n, F = len(utts), int(frm_off[-1])
        feat = eng.mfcc_device(pcm, pcm_off, frm_off_dev, n, F, int(frames.max()), int(frames.min()))
        ref = eng.emission(feat, gp, "h16")
        G = 4096
        img_all = torch.full((G + int(eng.lib.loe_emission_h16_img_bytes(F)) + G,), 0x5A, dtype=torch.uint8, device=eng.device)
        inv_all = torch.full((G + ((F + 127) // 128) * 128 + G,), -3.0, dtype=torch.float32, device=eng.device)
        image = (img_all[G:-G], inv_all[G:-G])
        feat2 = eng.mfcc_device(pcm, pcm_off, frm_off_dev, n, F, int(frames.max()), int(frames.min()), image=image)
        assert torch.equal(feat2, feat)
        got = eng.emission_image(image, F, gp)
        torch.cuda.synchronize()
        assert torch.equal(got, ref), (n_utts, float((got - ref).abs().max()))
        assert bool((img_all[:G] == 0x5A).all()) and bool((img_all[-G:] == 0x5A).all())
        assert bool((inv_all[:G] == -3.0).all()) and bool((inv_all[-G:] == -3.0).all())
        image2 = eng.image_buffers(F)
        assert eng.mfcc_device(pcm, pcm_off, frm_off_dev, n, F, int(frames.max()), int(frames.min()), image=image2, want_feat=False) is None
        assert torch.equal(eng.emission_image(image2, F, gp), ref)
    # fewer states than a tile, and the host-buffer decoder (which uses this hand-off) against the torch-backed decode
    gp5 = eng.pack_gaussians(inf._multivariate_normals[:5])
    assert torch.equal(eng.emission_image(image2, F, gp5), eng.emission(feat, gp5, "h16"))


def test_multi_model_emission_launches_match_per_model(eng, golden):
    """Training scores every word model on its own frames: loe_emission_h16_multi_dev (one launch, float32 features) and
    loe_h16_image_dev + loe_emission_h16_multi_img_dev (one launch, pre-split operand built once) must write exactly what one
    loe_emission_h16_dev launch per model writes; segments whose active flag is not 1 are left untouched."""
    from loe_speech_recognition import _native
    from loe_speech_recognition._engine import host_gauss_arrays, pack_h16_image
    torch = eng.torch
    words = ("1", "S", "Z", "7")
    feats = [np.concatenate([golden[f"train_feat_{w if w != 'S' else '1'}_{i}"] for i in range(6)]) for w in words]
    feats[1] = feats[1][:131]                                             # a segment that is not a multiple of the tile
    begin = np.concatenate(([0], np.cumsum([len(f) for f in feats])[:-1])).astype(np.int64)
    end = (begin + np.array([len(f) for f in feats])).astype(np.int64)
    x = eng._to_dev(np.concatenate(feats).astype(np.float32))
    sizes = np.array([N_STATES[w] for w in words], np.int32)
    first = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.int32)
    G, W = int(sizes.sum()), len(words)
    tile_halves = eng.lib.loe_emission_h16_tile_bytes() // 2
    img = np.zeros(W * tile_halves, np.float16); cst = np.zeros(W * 6, np.float32)
    for i, w in enumerate(words):
        means, us, c = host_gauss_arrays(trained_word_model(golden, w)._multivariate_normals)
        img[i * tile_halves:(i + 1) * tile_halves] = pack_h16_image(means, us, c)
        cst[i * 6:i * 6 + len(c)] = c.astype(np.float32)
    b_h16, cst_pad = eng._to_dev(img), eng._to_dev(cst)
    want = torch.full((int(end[-1]), G), -5.0, dtype=torch.float32, device=eng.device)
    for i in range(W):
        if i != 2:
            eng.emission_h16_into(x[begin[i]:end[i]], b_h16, cst_pad, i, int(sizes[i]), want[begin[i]:end[i]], int(first[i]))
    sb, se = eng._to_dev(begin), eng._to_dev(end)
    st, sn, sc = eng._to_dev(np.arange(W, dtype=np.int32)), eng._to_dev(sizes), eng._to_dev(first)
    active = eng._to_dev(np.array([1, 1, 0, 1], np.int32))                # the third model has converged: skipped
    got = torch.full_like(want, -5.0)
    _native.check(eng.lib.loe_emission_h16_multi_dev(x.data_ptr(), 39, b_h16.data_ptr(), cst_pad.data_ptr(), W, sb.data_ptr(), se.data_ptr(),
                                                     st.data_ptr(), sn.data_ptr(), sc.data_ptr(), active.data_ptr(), 5, got.data_ptr(), G, eng._stream()))
    assert torch.equal(got, want)
    tiles = (end - begin + 127) // 128
    img_tile = eng._to_dev(np.concatenate(([0], np.cumsum(tiles)[:-1])).astype(np.int32))
    n_img = int(tiles.sum())
    a_img = torch.full((n_img * 20480,), 0x77, dtype=torch.uint8, device=eng.device); inv2 = torch.full((n_img * 128,), -1.0, device=eng.device)
    _native.check(eng.lib.loe_h16_image_dev(x.data_ptr(), 39, W, sb.data_ptr(), se.data_ptr(), img_tile.data_ptr(), n_img, a_img.data_ptr(),
                                            inv2.data_ptr(), eng._stream()))
    got2 = torch.full_like(want, -5.0)
    _native.check(eng.lib.loe_emission_h16_multi_img_dev(a_img.data_ptr(), inv2.data_ptr(), b_h16.data_ptr(), cst_pad.data_ptr(), W, sb.data_ptr(),
                                                         se.data_ptr(), img_tile.data_ptr(), st.data_ptr(), sn.data_ptr(), sc.data_ptr(),
                                                         active.data_ptr(), 5, got2.data_ptr(), G, eng._stream()))
    assert torch.equal(got2, want)
    assert bool((inv2 > 0).all())


def test_emission_h16_range_handling(eng, golden):
    """binary16 tops out at 65504: feature rows of any magnitude are rescaled inside the kernel, a model
    whose whitening matrix leaves the range gets no FP16 image and runs on the TF32 kernel."""
    from loe_speech_recognition._engine import pack_h16_image
    rng = np.random.default_rng(11)
    inf = _loop_inference(golden)
    normals = inf._multivariate_normals[:12]
    gp = eng.pack_gaussians(normals)
    assert gp.b_h16 is not None
    lps = [mn._core.cov_object._log_pdet for mn in normals]
    x = np.concatenate([golden[f"loop_feat_{i}"] for i in range(3)])[:600].copy()
    big = rng.integers(0, len(x), size=60)
    x[big] *= np.float32(10.0) ** rng.integers(3, 9, size=60)[:, None].astype(np.float32)     # rows up to ~1e9
    xd = eng._to_dev(x)
    ref = eng.emission(xd, gp, "fp64").cpu().numpy()
    got = eng.emission(xd, gp, "h16").cpu().numpy()
    assert np.all(np.isfinite(got))
    np.testing.assert_allclose(got, ref, rtol=2e-4)      # huge rows: 1e-4 of |score| plus the float32 input cancellation
    # ordinary rows are untouched by the presence of huge ones
    small = np.setdiff1d(np.arange(len(x)), big)
    assert emission_close(got[small], ref[small], lps)
    # NaN rows stay NaN, rows after them are unaffected
    x2 = x[:256].copy(); x2[5, 7] = np.nan
    g2 = eng.emission(eng._to_dev(x2), gp, "h16").cpu().numpy()
    assert np.all(np.isnan(g2[5])) and np.all(np.isfinite(np.delete(g2, 5, axis=0)))
    # out-of-range model: no FP16 image, "h16" and "auto" run the TF32 kernel
    means = np.stack([np.asarray(mn._core.mean, dtype=np.float64) for mn in normals])
    us = np.stack([np.asarray(mn._core.cov_object._LP, dtype=np.float64) for mn in normals])
    cst = np.zeros(len(normals))
    assert pack_h16_image(means, us * 1e4, cst) is None
    assert pack_h16_image(means, us, cst) is not None
    gp_big = eng.pack_gauss_arrays(means, us * 1e4, cst)
    assert gp_big.b_h16 is None
    xs = eng._to_dev(x[small][:200] * np.float32(1e-4))
    assert eng.torch.equal(eng.emission(xs, gp_big, "h16"), eng.emission(xs, gp_big, "tc"))


# ------------------------------------------------------------------ a3 word Viterbi
def test_word_viterbi_bit_exact_given_scores(eng, golden):
    """Feed the ORACLE's float32 emission scores to the Viterbi kernel: score and path must be
    bit-identical to the reference's (golden) output."""
    from loe_speech_recognition import _trellis
    order = [str(s) for s in golden["iso_model_order"]]
    for i in range(22):
        x = golden[f"iso_feat_{i}"]
        for k, w in enumerate(order):
            from oracle import hmm as O
            means, Us, lps, logA = oracle_flat(golden, w)
            em = O.emission_scores(x, means, Us, lps)
            tp = eng.pack_trellises([_trellis.build([logA], [0], [0], "word")])
            T = x.shape[0]
            off = eng._to_dev(np.array([0, T], dtype=np.int64))
            path, _, _, bs = eng.viterbi(eng._to_dev(em), off, 1, T, T, tp, want_end_scores=False)
            assert bs.cpu().numpy()[0] == golden["iso_scores"][i, k]
            assert np.array_equal(path.cpu().numpy(), golden["iso_paths_" + str(i)][k])


def test_isolated_predict_end_to_end(eng, golden):
    """HiddenMarkovModel.predict / ModelCollection.predict with GPU emission (fp32 and fp64)."""
    from loe_speech_recognition import ModelCollection
    order = [str(s) for s in golden["iso_model_order"]]
    models = {w: trained_word_model(golden, w) for w in order}
    mc = ModelCollection()
    mc._models = [models[w] for w in order]
    feats = [golden[f"iso_feat_{i}"] for i in range(22)]
    for precision in ("fp32", "fp64", "tc", "h16"):
        sc = mc.scores_batch(feats, precision)
        assert rel_close(sc, golden["iso_scores"], rtol=1e-4)
        assert mc.predict_batch(feats, precision) == [str(s) for s in golden["iso_labels"]]
        for k, w in enumerate(order[:3]):
            s, paths = models[w].predict_batch(feats, precision)
            assert rel_close(s, golden["iso_scores"][:, k], rtol=1e-4)
            # every path equals the reference's, or is a near-tie by the oracle's own arithmetic (both state sequences
            # re-scored on the oracle's scores, |delta| <= 1e-4 |score|); the float64 kernel gets no excuse at all
            from oracle import hmm as O
            from oracle.adjudicate import adjudicate
            means, Us, lps, logA = oracle_flat(golden, w)
            tr = O.word_trellis(logA)
            excused = 0
            for i, p in enumerate(paths):
                want = golden[f"iso_paths_{i}"][k]
                if np.array_equal(p, want):
                    continue
                assert precision != "fp64", (w, i)
                ok, rel = adjudicate(O.emission_scores(feats[i], means, Us, lps), tr, None, p, len(means) - 1, want, len(means) - 1)
                assert ok, (precision, w, i, rel)
                excused += 1
            print(f"isolated {w} {precision}: {len(paths) - excused}/{len(paths)} paths identical, {excused} margin-excused")
    score, path = models["1"].predict(feats[0])
    assert isinstance(score, np.float32) and path.dtype == np.int8 and path.shape == (feats[0].shape[0],)
    assert mc.predict(feats[0]) == str(golden["iso_labels"][0])
    with pytest.raises(AssertionError):
        models["1"].predict(np.zeros((20, 13), np.float32))


# ------------------------------------------------------------------ a4 loop Viterbi
def _loop_inference(golden):
    from loe_speech_recognition import HiddenMarkovModelInference
    return HiddenMarkovModelInference.from_models([trained_word_model(golden, w) for w in LOOP_ORDER])


PENALTIES = {"int": -100, "f64": np.log(0.005), "pyfloat": -37.25, "zero": 0}


@pytest.mark.parametrize("name", list(PENALTIES))
def test_loop_viterbi_bit_exact_given_scores(eng, golden, name):
    from oracle import hmm as O
    inf = _loop_inference(golden)
    gp, tp = inf._packs()
    flat = [oracle_flat(golden, w) for w in LOOP_ORDER]
    means = np.concatenate([f[0] for f in flat]); Us = np.concatenate([f[1] for f in flat]); lps = np.concatenate([f[2] for f in flat])
    from loe_speech_recognition.hidden_markov_model import _penalty_args
    pen, f64 = _penalty_args(PENALTIES[name])
    assert f64 == (name == "f64")
    for i in range(10):
        x = golden[f"loop_feat_{i}"]
        em = O.emission_scores(x, means, Us, lps)
        T = x.shape[0]
        off = eng._to_dev(np.array([0, T], dtype=np.int64))
        path, _, _, bs = eng.viterbi(eng._to_dev(em), off, 1, T, T, tp, loop=True, penalty=pen, penalty_f64=f64,
                                     want_end_scores=False)
        assert bs.cpu().numpy()[0] == golden[f"loop_scores_{name}"][i], (name, i)
        assert np.array_equal(path.cpu().numpy(), golden[f"loop_path_{name}_{i}"]), (name, i)


@pytest.mark.parametrize("name", list(PENALTIES))
def test_loop_decode_strings(eng, golden, name):
    inf = _loop_inference(golden)
    inf._log_transition_probability_between_words = PENALTIES[name]
    feats = [golden[f"loop_feat_{i}"] for i in range(10)]
    want = [str(s) for s in golden[f"loop_strings_{name}"]]
    from oracle.adjudicate import OracleLoopDecoder, compare_loop_decodes
    assert inf.predict_batch(feats, "fp64") == want
    od = OracleLoopDecoder({w: (golden[f"train_means_{w}"], golden[f"train_covs_{w}"], golden[f"train_logA_{w}"]) for w in LOOP_ORDER},
                           LOOP_ORDER)
    ems = od.emissions(feats)
    for precision in ("fp32", "tc", "h16"):
        got32 = inf.predict_batch(feats, precision)
        scores, paths = inf.viterbi_batch(feats, precision)
        assert rel_close(scores, golden[f"loop_scores_{name}"], rtol=1e-4)
        # float32 / 3xTF32 / 3xFP16 emission: a differing path must be a near-tie under the ORACLE's arithmetic -- both
        # state sequences re-scored on the oracle's scores (margin test); the oracle decode itself is
        # pinned to the reference's golden paths
        v = compare_loop_decodes(ems, od.trellis, PENALTIES[name], od.sizes, LOOP_ORDER, paths, got32)
        print(f"loop {name} {precision}:", v.summary())
        assert not v.failed, (name, precision, v.failed)
        assert v.identical_paths + v.excused == len(feats)
        assert v.identical_strings >= len(feats) - v.excused
        for i, (g, w) in enumerate(zip(got32, want)):
            assert g == w or not np.array_equal(paths[i], golden[f"loop_path_{name}_{i}"]), (name, precision, i)
    assert inf.predict(feats[0]) == want[0]


def test_in_place_model_edits_reach_the_device(eng, golden):
    """The reference reads its model objects afresh on every predict (hidden_markov_model.py:481-531); the drop-in keeps a
    device copy, so an edit IN PLACE -- an entry of the transition dict, a mean inside a frozen scipy object -- must be
    seen by the next call (content fingerprint of the pack cache)."""
    def model():
        m = _loop_inference(golden)
        m._log_transition_probability_between_words = -100
        return m
    inf, fresh = model(), model()
    x = golden["loop_feat_0"]
    s0, p0 = inf._viterbi(x)
    assert inf._viterbi(x)[0] == s0
    st = int(p0[len(p0) // 4])                                   # a state the best path stays in: its self-loop is used
    assert p0[len(p0) // 4 + 1] == st or p0[len(p0) // 4 - 1] == st
    inf._log_transition_probs[(st, st)] = np.float32(inf._log_transition_probs[(st, st)] - 5.0)
    s1, _ = inf._viterbi(x)
    fresh._log_transition_probs[(st, st)] = np.float32(fresh._log_transition_probs[(st, st)] - 5.0)   # before its first pack
    assert s1 != s0 and s1 == fresh._viterbi(x)[0]
    mid = int(p0[len(p0) // 2])
    inf._multivariate_normals[mid]._core.mean[:] += 3.0
    s2, _ = inf._viterbi(x)
    fresh = model()
    fresh._log_transition_probs[(st, st)] = np.float32(fresh._log_transition_probs[(st, st)] - 5.0)
    fresh._multivariate_normals[mid]._core.mean[:] += 3.0
    assert s2 != s1 and s2 == fresh._viterbi(x)[0]
    assert inf.predict(x) == fresh.predict(x)


def test_edge_cases_short_utterances(eng, golden):
    """T = 2, 3, 9 (unreachable end states: -inf scores, back-pointer 0) and T = 1 (path [-1])."""
    inf = _loop_inference(golden)
    inf._log_transition_probability_between_words = -100
    word = trained_word_model(golden, "1")
    x9 = golden["edge_feat"]
    for T in (2, 3, 9):
        x = np.ascontiguousarray(x9[:T])
        s, p = inf._viterbi(x)
        assert np.array_equal(p, golden[f"edge_loop_path_T{T}"])
        ref = golden[f"edge_loop_score_T{T}"]
        assert (np.isinf(ref) and np.isinf(s) and s < 0) or rel_close(s, ref, rtol=1e-4)
        s, p = word.predict(x)
        assert np.array_equal(p, golden[f"edge_word_path_T{T}"])
        ref = golden[f"edge_word_score_T{T}"]
        assert (np.isinf(ref) and np.isinf(s) and s < 0) or rel_close(s, ref, rtol=1e-4)
    s, p = word.predict(np.ascontiguousarray(x9[:1]))
    assert p.tolist() == [-1] and np.isinf(s)
    with pytest.raises(Exception):
        inf.predict(np.ascontiguousarray(x9[:1]))            # reference: bare Exception from ModelBoundary


def test_too_many_states_overflow(eng, golden):
    from loe_speech_recognition import HiddenMarkovModelInference
    words = [trained_word_model(golden, "1") for _ in range(26)]   # 130 states > int8 range
    inf = HiddenMarkovModelInference.from_models(words)
    with pytest.raises(OverflowError):
        inf.predict(golden["iso_feat_0"])


# ------------------------------------------------------------------ a5 segmental K-means
def test_isolated_training_matches_reference(eng, golden):
    from loe_speech_recognition import HiddenMarkovModelTrainable
    for w in WORDS:
        feats = [golden[f"train_feat_{w}_{i}"] for i in range(64) if f"train_feat_{w}_{i}" in golden.files]
        m = HiddenMarkovModelTrainable.from_data(w, feats, num_of_states=N_STATES[w], max_iterations=4,
                                                 isMultiProcessingTraining=False, isTqdm=False)
        assert rel_close(m._means, golden[f"train_means_{w}"], rtol=1e-4, atol=1e-5), w
        assert rel_close(m._covariances, golden[f"train_covs_{w}"], rtol=1e-3, atol=1e-5), w
        got = m._log_transition_probs.to_dense(); ref = golden[f"train_logA_{w}"]
        assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(got[~np.isnan(got)], ref[~np.isnan(ref)]), w


def test_kmeans_statistics_exact(eng, golden):
    """Statistics kernel against the oracle M-step on the oracle's own alignments."""
    from oracle import hmm as O
    from loe_speech_recognition import Signal, HiddenMarkovModelTrainable
    w = "3"
    feats = [golden[f"train_feat_{w}_{i}"] for i in range(8)]
    means, Us, lps, logA = oracle_flat(golden, w)
    tr = O.word_trellis(logA)
    paths = [O.viterbi(O.emission_scores(x, means, Us, lps), tr)[2] for x in feats]
    ref = O.mstep(feats, paths, 5, old_means=None)
    m = HiddenMarkovModelTrainable(w)
    m._means = np.zeros((5, 39), np.float32)
    m._covariances = m._init_covariance(39, 5)
    m._train_external([Signal(5, x, p) for x, p in zip(feats, paths)])
    assert rel_close(m._means, ref["means"], rtol=1e-6, atol=1e-6)
    assert rel_close(m._covariances, ref["covs"], rtol=1e-4, atol=1e-6)
    assert np.array_equal(m._transition_probs.to_dense(), ref["trans"])


def test_training_empty_state_raises(eng, golden):
    from loe_speech_recognition import HiddenMarkovModelTrainable, Signal
    x = golden["train_feat_1_0"][:12]
    m = HiddenMarkovModelTrainable("1")
    m._means = np.zeros((3, 39), np.float32)
    m._covariances = m._init_covariance(39, 3)
    path = np.array([0] * 6 + [2] * 6, dtype=np.int8)          # state 1 never visited
    with pytest.raises(HiddenMarkovModelTrainable.HMMTrainMeanFail):
        m._train_external([Signal(3, x, path)])


# ------------------------------------------------------------------ a6 embedded training
def test_embedded_training_iteration_matches_reference(eng, golden, tmp_path):
    from loe_speech_recognition import HiddenMarkovModelTrainContinuous
    for w in WORDS:
        trained_word_model(golden, w).save(str(tmp_path))
    tc = HiddenMarkovModelTrainContinuous.from_folder(str(tmp_path), list(WORDS))
    tc.isTqdm = False
    assert tc.insert_silence("12") == "S1S2S"
    labeled = {}
    for lab in [str(s) for s in golden["emb_labels"]]:
        labeled[lab] = [golden[f"emb_feat_{lab}_{i}"] for i in range(int(golden[f"emb_count_{lab}"]))]
    tc.train(labeled, max_iterations=1)
    for w in WORDS:
        m = tc._trainable_models[w]
        assert rel_close(m._means, golden[f"emb1_means_{w}"], rtol=1e-4, atol=1e-5), w
        assert rel_close(m._covariances, golden[f"emb1_covs_{w}"], rtol=1e-3, atol=1e-5), w
        got = m._log_transition_probs.to_dense(); ref = golden[f"emb1_logA_{w}"]
        assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(got[~np.isnan(got)], ref[~np.isnan(ref)]), w
    tc.save(str(tmp_path / "out"))
    assert (tmp_path / "out" / "S" / "multivariate_normals.pickle").exists()


# ------------------------------------------------------------------ full pipeline + full-size properties
def test_pcm_to_strings_pipeline(eng, golden):
    from loe_speech_recognition import MFCC
    from loe_speech_recognition.synthetic import string_corpus
    inf = _loop_inference(golden)
    inf._log_transition_probability_between_words = -100
    utts, truth = string_corpus(seed=20, n_utts=6, n_digits=7)
    got = inf.decode_pcm_batch(utts)
    feats = MFCC.batch(utts, 16000)
    assert got == inf.predict_batch(feats)
    # vs the oracle pipeline on the same PCM: identical state paths, or a near-tie by the oracle's arithmetic
    from oracle.adjudicate import OracleLoopDecoder
    od = OracleLoopDecoder({w: (golden[f"train_means_{w}"], golden[f"train_covs_{w}"], golden[f"train_logA_{w}"]) for w in LOOP_ORDER},
                           LOOP_ORDER)
    _, paths = inf.viterbi_batch(feats)
    v = od.compare(utts, -100, paths, got)
    print("pcm -> strings pipeline:", v.summary())
    assert not v.failed, v.failed
    assert v.identical_strings >= len(utts) - v.excused
    # flat host buffer, chunked copy/compute overlap: same strings for any chunking
    flat = np.concatenate(utts)
    off = np.concatenate(([0], np.cumsum([len(u) for u in utts]))).astype(np.int64)
    for n_chunks in (1, 2, 3, 6, 9):
        assert inf.decode_pcm_flat(flat, off, 16000, n_chunks=n_chunks) == got
    pinned = eng.torch.from_numpy(flat).pin_memory()
    assert inf.decode_pcm_flat(pinned, off) == got
    assert inf.decode_pcm_flat(flat[:0], off[:1]) == []
    # raw int16 WAV samples on the wire: bit-identical features and strings
    assert inf.decode_pcm_flat(flat.astype(np.int16), off, n_chunks=2) == got
    f16 = MFCC.batch([u.astype(np.int16) for u in utts], 16000)
    assert all(np.array_equal(a, b) for a, b in zip(f16, feats))


def test_large_batch_properties(eng, golden):
    """Config-2-sized property checks: batch result == per-utterance result, paths valid,
    idempotent, and batched oracle agreement on a 200-utterance sample."""
    from oracle import hmm as O
    from loe_speech_recognition.synthetic import string_corpus
    inf = _loop_inference(golden)
    inf._log_transition_probability_between_words = -100
    utts, _ = string_corpus(seed=77, n_utts=2000, n_digits=7)
    b = eng.mfcc(utts)
    score, path = inf._decode_device(b, "fp32")
    score2, path2 = inf._decode_device(b, "fp32")
    assert eng.torch.equal(path, path2) and eng.torch.equal(score, score2)
    path_h = path.cpu().numpy(); off = b.frm_off_host
    assert path_h.min() >= 0 and path_h.max() < 58
    feat_h = b.feat.cpu().numpy()
    flat = [oracle_flat(golden, w) for w in LOOP_ORDER]
    means = np.concatenate([f[0] for f in flat]); Us = np.concatenate([f[1] for f in flat]); lps = np.concatenate([f[2] for f in flat])
    tr = O.loop_trellis([f[3] for f in flat])
    idx = list(range(0, 2000, 10))
    ems = [O.emission_scores(feat_h[off[i]:off[i + 1]], means, Us, lps) for i in idx]
    from oracle.adjudicate import compare_loop_decodes
    sizes = [len(f[0]) for f in flat]
    gp, tp = inf._packs()
    for precision in ("fp32", "h16"):
        _, p_dev = inf._decode_device(b, precision)
        p_h = p_dev.cpu().numpy()
        # oracle emission + oracle Viterbi on the kernel's own features: identical paths or adjudicated near-ties
        v = compare_loop_decodes(ems, tr, -100, sizes, LOOP_ORDER, [p_h[off[i]:off[i + 1]] for i in idx])
        print(f"large batch {precision}:", v.summary())
        assert not v.failed, (precision, v.failed[:5])
        assert v.excused <= 10, v.summary()
        # exact given the kernel's own scores
        sc = eng.emission(b.feat, gp, precision).cpu().numpy()
        _, _, opaths2 = O.viterbi_batch([sc[off[i]:off[i + 1]] for i in idx], tr, penalty=-100)
        assert all(np.array_equal(op, p_h[off[i]:off[i + 1]]) for i, op in zip(idx, opaths2)), precision


def test_benchmarked_path_h16_1000_utterances_vs_oracle(eng, golden):
    """The configuration bench.py times -- config-2 strings, 3xFP16 tensor-core emission, through the C-ABI host-buffer
    decoder (loe_decoder_decode_host) -- at 1 000 distinct utterances against the full oracle pipeline (oracle MFCC ->
    oracle emission -> oracle Viterbi, hidden_markov_model.py:481-581).  Every utterance must give the oracle's state
    path, or pass the margin test: both state sequences re-scored with the oracle's arithmetic within 1e-4 |score|.
    The counts are printed; a mismatch that is not a near-tie fails the test."""
    from oracle.adjudicate import OracleLoopDecoder
    from loe_speech_recognition.hidden_markov_model import _penalty_args
    from loe_speech_recognition.synthetic import string_corpus
    inf = _loop_inference(golden)
    inf._log_transition_probability_between_words = -100
    utts, _ = string_corpus(seed=4242, n_utts=1000, n_digits=7)
    off = np.concatenate(([0], np.cumsum([len(u) for u in utts]))).astype(np.int64)
    flat = np.concatenate(utts).astype(np.float32)
    dec = inf.native_decoder()
    assert dec.emission == "h16"
    strings = inf.decode_pcm_host(flat, off)
    pen, f64 = _penalty_args(-100)
    _, _, score, path = dec.decode(flat, off, pen, f64, LOOP_ORDER.index("S"), 32, 0, want_scores=True, want_path=True)
    frm = np.concatenate(([0], np.cumsum(1 + np.diff(off) // 160)))
    paths = [path[frm[i]:frm[i + 1]] for i in range(len(utts))]
    od = OracleLoopDecoder({w: (golden[f"train_means_{w}"], golden[f"train_covs_{w}"], golden[f"train_logA_{w}"]) for w in LOOP_ORDER},
                           LOOP_ORDER)
    v = od.compare(utts, -100, paths, strings)
    print("h16 / C-ABI decode of 1000 config-2 utterances vs oracle:", v.summary())
    assert not v.failed, v.failed[:5]
    assert v.identical_paths + v.excused == len(utts)
    assert v.identical_strings >= len(utts) - v.excused
    assert v.excused <= 50, v.summary()                    # near-ties are rare; a flood of them would be a precision bug
    # the same batch through the torch-backed entry points gives the same strings
    assert inf.decode_pcm_flat(flat, off, precision="h16") == strings


def test_training_at_scale_matches_oracle(eng, golden):
    """600 utterances of one word (tiled pool with noise): GPU trainer vs the oracle training loop."""
    from oracle import hmm as O
    from loe_speech_recognition import HiddenMarkovModelTrainable
    rng = np.random.default_rng(9)
    base = [golden[f"train_feat_7_{i}"] for i in range(8)]
    feats = [(base[i % 8] + rng.normal(0, 0.05, size=base[i % 8].shape)).astype(np.float32) for i in range(600)]
    m = HiddenMarkovModelTrainable.from_data("7", feats, num_of_states=5, max_iterations=3, isMultiProcessingTraining=False, isTqdm=False)
    means, covs, trans = O.init_parameters(feats[0], 5)
    for it in range(3):
        packs = [O.gaussian_pack(means[s], covs[s]) for s in range(5)]
        tr = O.word_trellis(O.log_transitions(trans))
        ems = [O.emission_scores(x, [p[0] for p in packs], [p[1] for p in packs], [p[2] for p in packs]) for x in feats]
        _, _, paths = O.viterbi_batch(ems, tr)
        r = O.mstep(feats, paths, 5, old_means=means)
        if r["converged"]:
            break
        means, covs, trans = r["means"], r["covs"], r["trans"]
    assert rel_close(m._means, means, rtol=1e-3, atol=1e-4)
    assert rel_close(m._covariances, covs, rtol=1e-2, atol=1e-4)
    assert np.allclose(m._transition_probs.to_dense(), trans, atol=2e-3)


def test_kmeans_statistics_large_random(eng):
    """Statistics kernels on 20k utterances with random (valid and invalid) alignments vs NumPy."""
    from loe_speech_recognition import _trellis
    rng = np.random.default_rng(2)
    S, D, n = 5, 39, 20000
    lens = rng.integers(9, 60, size=n)
    off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    F = int(off[-1])
    x = rng.normal(0, 2, size=(F, D)).astype(np.float32)
    path = np.zeros(F, dtype=np.int8)
    for i in range(n):
        p = np.sort(rng.integers(0, S, size=lens[i]))
        if i % 50 == 0:
            p = p[::-1].copy()                       # decreasing: only the first run is credited
        path[off[i]:off[i + 1]] = p
    tp = eng.pack_trellises([_trellis.build([np.zeros((S, S), np.float32)], [0], [0], "word")])
    shift = rng.normal(size=(S, D)).astype(np.float32)
    stats, counts, bucket = eng.kmeans_stats(eng._to_dev(x), eng._to_dev(path), eng._to_dev(off), n, F, tp, None, False, S,
                                             eng._to_dev(shift))
    stats = stats.cpu().numpy(); counts = counts.cpu().numpy()
    # NumPy restatement of order_by_state + statistics
    valid = np.zeros(F, dtype=bool)
    ref_counts = np.zeros((S, S), dtype=np.int64)
    for i in range(n):
        p = path[off[i]:off[i + 1]].astype(int)
        ok = np.concatenate(([True], np.cumprod(np.diff(p) >= 0).astype(bool)))
        valid[off[i]:off[i + 1]] = ok
        np.add.at(ref_counts, (p[:-1], p[1:]), 1)
    assert np.array_equal(counts, ref_counts)
    iu = np.triu_indices(D)
    for s in range(S):
        sel = valid & (path == s)
        d = x[sel].astype(np.float64) - shift[s]
        assert stats[s, 0] == sel.sum()
        assert np.allclose(stats[s, 1:1 + D], d.sum(0), rtol=1e-10, atol=1e-8)
        assert np.allclose(stats[s, 1 + D:], (d.T @ d)[iu], rtol=1e-10, atol=1e-8)
    # bitwise reproducible
    stats2, _, _ = eng.kmeans_stats(eng._to_dev(x), eng._to_dev(path), eng._to_dev(off), n, F, tp, None, False, S, eng._to_dev(shift))
    assert np.array_equal(stats, stats2.cpu().numpy())


def test_long_utterance_fallback_paths(eng, golden):
    """Utterances too long for the warp kernel's shared-memory back-pointers (CTA kernel, then its
    global-workspace variant) must give the same bit-exact result."""
    from oracle import hmm as O
    inf = _loop_inference(golden)
    gp, tp = inf._packs()
    flat = [oracle_flat(golden, w) for w in LOOP_ORDER]
    tr = O.loop_trellis([f[3] for f in flat])
    base = np.concatenate([golden[f"loop_feat_{i}"] for i in range(10)])
    for T in (3000, 12000):                      # 3000: CTA kernel, smem back-pointers; 12000: global workspace
        x = np.ascontiguousarray(np.tile(base, (T // len(base) + 1, 1))[:T])
        xd = eng._to_dev(x)
        sc = eng.emission(xd, gp, "fp32")
        off = eng._to_dev(np.array([0, T], dtype=np.int64))
        path, _, _, bs = eng.viterbi(sc, off, 1, T, T, tp, loop=True, penalty=-100.0, penalty_f64=False, want_end_scores=False)
        es, bi, opath = O.viterbi(sc.cpu().numpy(), tr, penalty=-100)
        assert bs.cpu().numpy()[0] == es[bi]
        assert np.array_equal(path.cpu().numpy(), opath)
    assert not eng.lib.loe_viterbi_bp_fits(12000, 58)
    # label decoding on the fallback path (separate labels kernel) and the stand-alone labels entry point
    inf._log_transition_probability_between_words = -100
    long_feat = np.ascontiguousarray(np.tile(base, (2, 1))[:5000])
    got = inf.predict_batch([long_feat, golden["loop_feat_0"]])
    _, paths = inf.viterbi_batch([long_feat, golden["loop_feat_0"]])
    assert got == ["".join(inf._model_boundaries.get_labels(p)) for p in paths]
    assert len(got[0]) > 32                          # more words than the id table holds: host routine takes over
    batch = eng.upload_features([golden["loop_feat_0"], golden["loop_feat_1"]], 39)
    _, path = inf._decode_device(batch)
    words, count = eng.labels(path, batch.frm_off, 2, tp, skip_label=LOOP_ORDER.index("S"), max_words=32)
    for i in range(2):
        ids = words.cpu().numpy()[i, :int(count[i])]
        assert "".join(LOOP_ORDER[k] for k in ids) == str(golden["loop_strings_int"][i])


def test_config1_isolated_digits_200_utterances(eng):
    """BASELINE.json configs[0]: 11 word HMMs x 5 states, 200 synthetic 1 s utterances.  Models are
    trained by the GPU trainer; scores of the batched classifier are checked against the oracle run
    on the same models, labels must be identical."""
    from oracle import hmm as O
    from loe_speech_recognition import HiddenMarkovModelTrainable, MFCC, ModelCollection, TI_DIGITS_LABELS
    from loe_speech_recognition.synthetic import DIGITS, isolated_corpus, synth_isolated
    train = isolated_corpus(seed=1, n_per_word=10, words=DIGITS, seconds=1.0)
    models = {}
    for w in TI_DIGITS_LABELS:
        models[w] = HiddenMarkovModelTrainable.from_data(w, MFCC.batch(train[w], 16000), num_of_states=5, max_iterations=5,
                                                         isMultiProcessingTraining=False, isTqdm=False)
    mc = ModelCollection()
    mc._models = [models[w] for w in TI_DIGITS_LABELS]
    rng = np.random.default_rng(0)
    truth = [DIGITS[int(i)] for i in rng.integers(0, 11, size=200)]
    utts = [synth_isolated(rng, w, 1.0) for w in truth]
    feats = MFCC.batch(utts, 16000)
    assert all(f.shape == (101, 39) for f in feats)
    got = mc.predict_batch(feats)
    assert np.mean([a == b for a, b in zip(got, truth)]) >= 0.95
    sc = mc.scores_batch(feats)
    ref = np.zeros_like(sc)
    for k, w in enumerate(TI_DIGITS_LABELS):
        m = models[w]
        packs = [O.gaussian_pack(m._means[s], m._covariances[s]) for s in range(5)]
        tr = O.word_trellis(m._log_transition_probs.to_dense())
        ems = [O.emission_scores(x, [p[0] for p in packs], [p[1] for p in packs], [p[2] for p in packs]) for x in feats]
        ref[:, k] = O.viterbi_batch(ems, tr)[0][:, 0]
    finite = np.isfinite(ref)
    assert np.array_equal(finite, np.isfinite(sc))
    # totals are sums of 101 per-frame scores of either sign (magnitudes up to 1e3): 1e-4 relative + 0.05 absolute
    assert rel_close(sc[finite], ref[finite], rtol=1e-4, atol=0.05), np.abs(sc[finite] - ref[finite]).max()
    labels = list(TI_DIGITS_LABELS)
    assert got == [labels[int(i)] for i in np.argmax(ref, axis=1)]


@pytest.mark.parametrize("penalty", [-100, np.log(0.005)])
def test_wide_trellis_four_positions_per_lane(eng, golden, penalty):
    """65..128 positions run with 4 positions per lane (SPL = 4): a 25-word loop grammar (121 states),
    and 20 independent words side by side (100 states), against the oracle on the kernel's own scores."""
    from oracle import hmm as O
    from loe_speech_recognition import HiddenMarkovModelInference, ModelCollection
    from loe_speech_recognition.hidden_markov_model import _penalty_args
    words = list(LOOP_ORDER) + list(LOOP_ORDER) + ["1"]
    models = [trained_word_model(golden, w) for w in words]
    inf = HiddenMarkovModelInference.from_models(models)
    inf._log_transition_probability_between_words = penalty
    gp, tp = inf._packs()
    assert tp.max_pos == 121
    tr = O.loop_trellis([golden[f"train_logA_{w}"] for w in words])
    pen, f64 = _penalty_args(penalty)
    feats = [golden[f"loop_feat_{i}"] for i in range(10)]
    batch = eng.upload_features(feats, 39)
    sc = eng.emission(batch.feat, gp, "fp32")
    path, _, _, bs = eng.viterbi(sc, batch.frm_off, batch.n_utt, batch.max_frames, batch.total_frames, tp, loop=True,
                                 penalty=pen, penalty_f64=f64, want_end_scores=False)
    sc_h, path_h, off = sc.cpu().numpy(), path.cpu().numpy(), batch.frm_off_host
    for i in range(10):
        es, bi, opath = O.viterbi(sc_h[off[i]:off[i + 1]], tr, penalty=penalty)
        assert bs.cpu().numpy()[i] == es[bi] and np.array_equal(path_h[off[i]:off[i + 1]], opath), i
    if penalty == -100:
        mc = ModelCollection()
        mc._models = [trained_word_model(golden, w) for w in (list(LOOP_ORDER[:10]) * 2)]
        got = mc.scores_batch(feats, "fp32")
        gp2, tp2 = mc._packs()
        assert tp2.max_pos == 100
        sc2 = eng.emission(batch.feat, gp2, "fp32").cpu().numpy()
        for k, m in enumerate(mc._models):
            trw = O.word_trellis(m._log_transition_probs.to_dense())
            for i in (0, 5):
                es, _, _ = O.viterbi(sc2[off[i]:off[i + 1], 5 * k:5 * k + 5], trw)
                assert got[i, k] == es[0] or (np.isinf(es[0]) and np.isinf(got[i, k]))


def test_batched_training_equals_per_word_training(eng, golden):
    """from_data_batch (all words in one device pass per iteration, batched eigendecomposition) must give
    the models from_data gives word by word."""
    from loe_speech_recognition import HiddenMarkovModelTrainable
    labeled = {w: [golden[f"train_feat_{w}_{i}"] for i in range(64) if f"train_feat_{w}_{i}" in golden.files] for w in WORDS}
    got = HiddenMarkovModelTrainable.from_data_batch(labeled, num_of_states=dict(N_STATES), max_iterations=4)
    for w in WORDS:
        ref = HiddenMarkovModelTrainable.from_data(w, labeled[w], num_of_states=N_STATES[w], max_iterations=4,
                                                   isMultiProcessingTraining=False, isTqdm=False)
        assert rel_close(got[w]._means, ref._means, rtol=1e-5, atol=1e-6), w
        assert rel_close(got[w]._covariances, ref._covariances, rtol=1e-4, atol=1e-6), w
        a, b = got[w]._log_transition_probs.to_dense(), ref._log_transition_probs.to_dense()
        assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)]), w
        assert rel_close(got[w]._means, golden[f"train_means_{w}"], rtol=1e-4, atol=1e-5), w
        assert got[w].num_of_states == N_STATES[w] and len(got[w]._multivariate_normals) == N_STATES[w]
    # identical frames: the best path jumps 0 -> 2 at once, state 1 is never visited -> HMMTrainMeanFail (both entry points)
    flat = np.tile(golden["train_feat_1_0"][:1], (12, 1))
    with pytest.raises(HiddenMarkovModelTrainable.HMMTrainMeanFail):
        HiddenMarkovModelTrainable.from_data_batch({"1": [flat]}, num_of_states=3, max_iterations=3)
    with pytest.raises(HiddenMarkovModelTrainable.HMMTrainMeanFail):
        HiddenMarkovModelTrainable.from_data("1", [flat], num_of_states=3, max_iterations=3, isTqdm=False)


def test_device_mstep_equals_host_mstep(eng, golden):
    """loe_mstep_dev against the host M-step (hidden_markov_model.py:320-350 restated in _update_from_statistics) on the
    statistics of a real E-step: float32 means / covariances / transition probabilities bit-identical, convergence flags
    equal, and the whitening data it writes into the 3xFP16 image scores like scipy's (same quadratic form)."""
    from loe_speech_recognition import HiddenMarkovModelTrainable
    labeled = {w: [golden[f"train_feat_{w}_{i}"] for i in range(64) if f"train_feat_{w}_{i}" in golden.files] for w in WORDS}
    dev, info = HiddenMarkovModelTrainable.from_data_batch(labeled, num_of_states=dict(N_STATES), max_iterations=4, return_info=True)
    host = HiddenMarkovModelTrainable.from_data_batch(labeled, num_of_states=dict(N_STATES), max_iterations=4, device_mstep=False)
    assert info["mstep"] == "device" and 1 <= info["iterations"] <= 4
    for w in WORDS:
        assert np.array_equal(dev[w]._means, host[w]._means), w
        assert np.array_equal(dev[w]._covariances, host[w]._covariances), w
        a, b = dev[w]._log_transition_probs.to_dense(), host[w]._log_transition_probs.to_dense()
        assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)]), w
        assert rel_close(dev[w]._means, golden[f"train_means_{w}"], rtol=1e-4, atol=1e-5), w
        assert rel_close(dev[w]._covariances, golden[f"train_covs_{w}"], rtol=1e-3, atol=1e-5), w
    # run to convergence: both stop on the same iteration for every word (the means-only test runs on the device)
    dev2, info2 = HiddenMarkovModelTrainable.from_data_batch(labeled, num_of_states=dict(N_STATES), max_iterations=40, return_info=True)
    host2 = HiddenMarkovModelTrainable.from_data_batch(labeled, num_of_states=dict(N_STATES), max_iterations=40, device_mstep=False)
    assert info2["iterations"] < 40
    for w in WORDS:
        assert np.array_equal(dev2[w]._means, host2[w]._means) and np.array_equal(dev2[w]._covariances, host2[w]._covariances), w
    with pytest.raises(HiddenMarkovModelTrainable.HMMTrainMeanFail):
        flat = np.tile(golden["train_feat_1_0"][:1], (12, 1))
        HiddenMarkovModelTrainable.from_data_batch({"1": [flat]}, num_of_states=3, max_iterations=3)


def test_device_mstep_kernel_against_numpy(eng):
    """The M-step kernels alone on synthetic statistics: exact float32 parameters, NaN rows for states that never leave,
    converged words untouched, an empty state reported, a non-finite covariance parked as suspect, and the image a word
    gets scores frames like the float64 kernel on the same (mean, covariance)."""
    import ctypes
    from loe_speech_recognition import _native, _trellis
    from loe_speech_recognition.hidden_markov_model import HiddenMarkovModelTrainable as T, MultivariateNormal
    torch = eng.torch
    rng = np.random.default_rng(12)
    D, sizes = 39, np.array([5, 3, 5, 5], np.int32)
    W, G = len(sizes), int(sizes.sum())
    first = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.int32)
    tile0 = np.arange(W, dtype=np.int32)
    stride = 1 + D + D * (D + 1) // 2
    iu = np.triu_indices(D)
    old = rng.normal(0, 2, size=(G, D)).astype(np.float32)
    stats = np.zeros((G, stride)); counts = np.zeros((G, G), np.int32)
    for g in range(G):
        n = int(rng.integers(60, 400))
        x = rng.normal(0, 1.5, size=(n, D)) @ np.diag(rng.uniform(0.3, 2.0, D)) + rng.normal(0, 0.1, D)
        stats[g, 0] = n; stats[g, 1:1 + D] = x.sum(0); stats[g, 1 + D:] = (x.T @ x)[iu]
    for i, (a, n) in enumerate(zip(first, sizes)):
        for s in range(n):
            counts[a + s, a + s] = rng.integers(5, 50)
            if s + 1 < n:
                counts[a + s, a + s + 1] = rng.integers(1, 20)
    stats[first[2]:first[2] + 5, 1:1 + D] *= 1e-9                 # word 2: means move by less than the allclose bar -> converged
    stats[first[3] + 1, 0] = 0.0                                   # word 3: an empty state -> mean fail
    means_d = eng._to_dev(old.copy()); cov_d = torch.full((G, D, D), 7.0, dtype=torch.float32, device=eng.device)
    tile_halves = eng.lib.loe_emission_h16_tile_bytes() // 2
    b_h16 = torch.zeros(W * tile_halves, dtype=torch.float16, device=eng.device); cst_pad = torch.zeros(W * 6, dtype=torch.float32, device=eng.device)
    band = torch.full((G, 3), 9.0, dtype=torch.float32, device=eng.device)
    active = torch.ones(W, dtype=torch.int32, device=eng.device); updated = torch.zeros_like(active); status = torch.zeros_like(active)
    applied = torch.zeros((G, G), dtype=torch.int32, device=eng.device)
    tabs = [eng._to_dev(np.repeat(np.arange(W, dtype=np.int32), sizes)), eng._to_dev(first), eng._to_dev(sizes), eng._to_dev(tile0)]
    sd, cd = eng._to_dev(stats), eng._to_dev(counts)
    _native.check(eng.lib.loe_mstep_dev(sd.data_ptr(), cd.data_ptr(), G, W, *[t.data_ptr() for t in tabs], means_d.data_ptr(), cov_d.data_ptr(),
                                        applied.data_ptr(), band.data_ptr(), b_h16.data_ptr(), cst_pad.data_ptr(), active.data_ptr(),
                                        updated.data_ptr(), status.data_ptr(), D, eng._stream()))
    st = status.cpu().numpy(); act = active.cpu().numpy()
    assert st.tolist() == [_native.LOE_MSTEP_UPDATED, _native.LOE_MSTEP_UPDATED, _native.LOE_MSTEP_CONVERGED, _native.LOE_MSTEP_MEAN_FAIL]
    assert act.tolist() == [1, 1, 0, -1]
    got_means, got_cov, got_band = means_d.cpu().numpy(), cov_d.cpu().numpy(), band.cpu().numpy()
    for i in (0, 1):
        a, n = int(first[i]), int(sizes[i])
        m = T(str(i)); m._means = old[a:a + n].copy(); m._covariances = None
        m._update_from_statistics(stats[a:a + n], counts[a:a + n, a:a + n].astype(np.int64), shift=old[a:a + n].astype(np.float64))
        assert np.array_equal(got_means[a:a + n], m._means) and np.array_equal(got_cov[a:a + n], m._covariances), i
        with np.errstate(divide="ignore", invalid="ignore"):
            ref_band = _trellis.build([np.log(m._transition_probs.to_dense())], [0], [0], "word").band
        assert np.array_equal(np.isnan(got_band[a:a + n]), np.isnan(ref_band))
        ok = ~np.isnan(ref_band)
        assert np.allclose(got_band[a:a + n][ok], ref_band[ok], rtol=2e-7, atol=0) or np.array_equal(got_band[a:a + n][ok], ref_band[ok])
        assert np.array_equal(applied.cpu().numpy()[a:a + n], counts[a:a + n])
        # the image scores like scipy on the same parameters
        normals = [MultivariateNormal.from_means_covariances(mm, cc) for mm, cc in zip(m._means, m._covariances)]
        gp = eng.pack_gaussians(normals)
        x = eng._to_dev((m._means[rng.integers(0, n, 300)] + rng.normal(0, 1.5, size=(300, D))).astype(np.float32))
        ref = eng.emission(x, gp, "fp64").cpu().numpy()
        out = torch.empty((300, n), dtype=torch.float32, device=eng.device)
        eng.emission_h16_into(x, b_h16, cst_pad, int(tile0[i]), n, out, 0)
        lps = [mn._core.cov_object._log_pdet for mn in normals]
        assert emission_close(out.cpu().numpy(), ref, lps), (i, np.abs(out.cpu().numpy() - ref).max())
    for i in (2, 3):                                               # converged / failed words keep everything
        a, n = int(first[i]), int(sizes[i])
        assert np.array_equal(got_means[a:a + n], old[a:a + n]) and np.all(got_cov[a:a + n] == 7.0) and np.all(got_band[a:a + n] == 9.0)
    # a non-finite covariance (one frame: N - 1 = 0) is parked for the host
    stats2 = stats.copy(); stats2[first[0] + 2, 0] = 1.0
    active.fill_(1); active[3] = 0
    _native.check(eng.lib.loe_mstep_dev(eng._to_dev(stats2).data_ptr(), cd.data_ptr(), G, W, *[t.data_ptr() for t in tabs], means_d.data_ptr(),
                                        cov_d.data_ptr(), applied.data_ptr(), band.data_ptr(), b_h16.data_ptr(), cst_pad.data_ptr(),
                                        active.data_ptr(), updated.data_ptr(), status.data_ptr(), D, eng._stream()))
    assert status.cpu().numpy()[0] & _native.LOE_MSTEP_SUSPECT and active.cpu().numpy()[0] == -2


@pytest.mark.gpu
def test_host_buffer_decoder_matches_flat_decode(eng, golden):
    """loe_decoder_decode_host (C pipeline, host buffers, no torch on the way) == decode_pcm_flat,
    for float32 and int16 PCM, one chunk and several, pinned and pageable; scores and paths too."""
    from loe_speech_recognition import MFCC
    from loe_speech_recognition._decoder import PinnedBuffer
    from loe_speech_recognition.hidden_markov_model import _penalty_args
    from loe_speech_recognition.synthetic import string_corpus
    inf = _loop_inference(golden)
    utts, _ = string_corpus(seed=31, n_utts=24, n_digits=5)
    utts.append(utts[0][:1700])                                   # 11 frames: close to the shortest allowed
    off = np.concatenate(([0], np.cumsum([len(u) for u in utts]))).astype(np.int64)
    flat = np.concatenate(utts).astype(np.float32)
    for penalty in (np.log(0.005), -100):
        inf._log_transition_probability_between_words = penalty
        want = inf.decode_pcm_flat(flat, off)
        assert inf.decode_pcm_host(flat, off) == want
        assert inf.decode_pcm_host(flat, off, n_chunks=5) == want
        assert inf.decode_pcm_host(flat.astype(np.int16), off, n_chunks=3) == inf.decode_pcm_flat(flat.astype(np.int16), off)
    pin = PinnedBuffer(flat.shape[0], np.float32)
    pin.array[:] = flat
    assert inf.decode_pcm_host(pin.array, off, n_chunks=4) == want
    sc, paths = inf.viterbi_batch(MFCC.batch(utts, 16000))
    pen, f64 = _penalty_args(inf._log_transition_probability_between_words)
    _, _, score, path = inf.native_decoder().decode(pin.array, off, pen, f64, -1, 32, 3, want_scores=True, want_path=True)
    np.testing.assert_array_equal(score, sc)
    np.testing.assert_array_equal(path, np.concatenate(paths))
    pin.close()
    assert inf.decode_pcm_host(flat[:0], off[:1]) == []
    with pytest.raises(ValueError):
        inf.decode_pcm_host(flat[:100], np.array([0, 100]))       # 1 frame < 9: the reference's savgol error
    # a second model gets its own decoder; replacing the penalty needs no rebuild
    assert inf.native_decoder() is inf.native_decoder()


@pytest.mark.gpu
def test_host_narrowing_of_float_pcm_is_lossless(eng, golden, monkeypatch):
    """loe_decoder_decode_host narrows float32 PCM that holds int16 values to int16 on the host (verified per
    sample, per chunk) when that beats the copy it saves; forced on here: same strings, scores and paths as the
    float32 path, also when one chunk holds an inexact sample and travels as float32."""
    from loe_speech_recognition._decoder import NativeDecoder
    from loe_speech_recognition.hidden_markov_model import _penalty_args
    from loe_speech_recognition.synthetic import string_corpus
    inf = _loop_inference(golden)
    utts, _ = string_corpus(seed=57, n_utts=18, n_digits=4)
    off = np.concatenate(([0], np.cumsum([len(u) for u in utts]))).astype(np.int64)
    flat = np.round(np.concatenate(utts)).astype(np.float32)
    pen, f64 = _penalty_args(inf._log_transition_probability_between_words)
    monkeypatch.setenv("LOE_B200_NARROW_THREADS", "0")
    inf.__dict__.pop("_native_decoder", None)
    w0, c0, s0, p0 = inf.native_decoder().decode(flat, off, pen, f64, -1, 32, 3, want_scores=True, want_path=True)
    assert inf.native_decoder().narrow_rate() == 0.0
    monkeypatch.setenv("LOE_B200_NARROW_THREADS", "4")
    monkeypatch.setenv("LOE_B200_NARROW_MIN_SAMPLES", "1000")
    inf.__dict__.pop("_native_decoder", None)
    dec = inf.native_decoder()
    dec.set_narrow("on")
    w1, c1, s1, p1 = dec.decode(flat, off, pen, f64, -1, 32, 3, want_scores=True, want_path=True)
    assert dec.narrow_rate() > 0.0
    st = dec.stats()
    assert st["narrow_on"] is True and st["narrow_threads"] == 4 and st["chunks"] == 3 and st["chunks_narrowed"] == 3
    assert st["pcm_bytes"] == flat.nbytes and st["wire_bytes"] == flat.nbytes // 2 + 2 * 8 * (len(off) - 1 + 3)
    assert st["copy_gbps"] > 0.0
    dec.set_narrow("off")
    w1b, c1b, s1b, p1b = dec.decode(flat, off, pen, f64, -1, 32, 3, want_scores=True, want_path=True)
    assert dec.stats()["chunks_narrowed"] == 0 and dec.stats()["wire_bytes"] > flat.nbytes and dec.narrow_rate() < 0.0
    def ids(w, c):                                           # ids past an utterance's count are unspecified
        return np.where(np.arange(32)[None, :] < c[:, None], w, 0)
    for a, b in ((ids(w1, c1), ids(w1b, c1b)), (c1, c1b), (s1, s1b), (p1, p1b)):
        np.testing.assert_array_equal(a, b)
    dec.set_narrow("on")
    for a, b in ((ids(w0, c0), ids(w1, c1)), (c0, c1), (s0, s1), (p0, p1)):
        np.testing.assert_array_equal(a, b)
    bent = flat.copy()
    bent[int(off[9]) + 1234] += 0.25                                # the middle chunk can no longer be narrowed
    inf.__dict__.pop("_native_decoder", None)
    monkeypatch.setenv("LOE_B200_NARROW_THREADS", "0")
    w2, c2, s2, p2 = inf.native_decoder().decode(bent, off, pen, f64, -1, 32, 3, want_scores=True, want_path=True)
    monkeypatch.setenv("LOE_B200_NARROW_THREADS", "4")
    monkeypatch.setenv("LOE_B200_NARROW", "on")
    inf.__dict__.pop("_native_decoder", None)
    w3, c3, s3, p3 = inf.native_decoder().decode(bent, off, pen, f64, -1, 32, 3, want_scores=True, want_path=True)
    assert inf.native_decoder().stats()["chunks_narrowed"] == 2
    for a, b in ((ids(w2, c2), ids(w3, c3)), (c2, c3), (s2, s3), (p2, p3)):
        np.testing.assert_array_equal(a, b)
    inf.__dict__.pop("_native_decoder", None)


def test_c_program_decodes_like_python(eng, golden, tmp_path):
    """examples/decode_host.c, compiled with gcc against include/loe_b200.h and libloe_b200.so, decodes a
    batch without Python in the process and prints the same word ids and scores."""
    import os
    import shutil
    import subprocess
    from loe_speech_recognition import _native
    from loe_speech_recognition._decoder import write_blob
    from loe_speech_recognition.hidden_markov_model import _penalty_args
    from loe_speech_recognition.synthetic import string_corpus
    if shutil.which("gcc") is None:
        pytest.skip("no gcc on this box")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "decode_host")
    lib_dir = os.path.dirname(_native.LIB_PATH)
    subprocess.run(["gcc", "-O2", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "decode_host.c"),
                    "-L", lib_dir, "-lloe_b200", f"-Wl,-rpath,{lib_dir}", "-o", exe], check=True)
    inf = _loop_inference(golden)
    utts, _ = string_corpus(seed=41, n_utts=12, n_digits=4)
    off = np.concatenate(([0], np.cumsum([len(u) for u in utts]))).astype(np.int64)
    flat = np.concatenate(utts).astype(np.int16)
    pen, f64 = _penalty_args(inf._log_transition_probability_between_words)
    dec = inf.native_decoder()
    words, count, score, _ = dec.decode(flat, off, pen, f64, -1, 32, 3, want_scores=True)
    blob = str(tmp_path / "batch.blob")
    write_blob(blob, dec, flat, off, pen, f64, -1, 32, 3)
    env = dict(os.environ)
    out = subprocess.run([exe, blob], check=True, capture_output=True, text=True, env=env, timeout=120).stdout.strip().splitlines()
    assert len(out) == len(utts)
    for i, line in enumerate(out):
        tok = line.split()
        assert int(tok[0]) == count[i]
        assert np.float32(tok[1]) == score[i]
        assert [int(t) for t in tok[2:]] == words[i, :count[i]].tolist()
