"""g1: diagonal-covariance Gaussian-mixture emission scoring with log-sum-exp over mixtures, and the
lexicon-expanded phone loop of BASELINE.json configs[4].  No live reference counterpart (deprecated/gaussian_mixture_model.py
is un-importable): parity is against the restated oracle (oracle/gmm.py) and is UNPINNED by construction."""
import numpy as np
import pytest

from oracle import gmm as OG
from oracle import hmm as O


def random_gmm(rng, S, M, D=39, spread=2.0, var_lo=0.05, var_hi=2.0):
    means = rng.normal(0, spread, size=(S, M, D))
    variances = rng.uniform(var_lo, var_hi, size=(S, M, D))
    w = rng.uniform(0.2, 1.0, size=(S, M))
    return w / w.sum(axis=1, keepdims=True), means, variances


def gmm_close(got, ref, cst, rtol=1e-4):
    """Same bar as the full-covariance kernels (tests/helpers.py::emission_close): 1e-4 of max(|score|, largest |c|)."""
    scale = np.maximum(np.abs(ref.astype(np.float64)), np.abs(cst).max())
    return bool(np.all(np.abs(got.astype(np.float64) - ref) <= rtol * scale))


# ------------------------------------------------------------------ CPU: the packed operand
@pytest.mark.parametrize("S,M", [(55, 4), (120, 16), (7, 3), (11, 1)])
def test_packed_image_reproduces_the_oracle_in_split_fp16_arithmetic(S, M):
    """NumPy walk through loe_emission_gmm_tc_dev: z = (x - shift) / t, A = [z^2, 1, z, 0] split into two binary16 parts,
    hi*hi + lo*hi + hi*lo against the packed image with float32 accumulation, log-sum-exp over each state's columns."""
    from loe_speech_recognition.gmm import GMM_K, GMM_TILE_N, gmm_tile_operand, pack_gmm_image
    rng = np.random.default_rng(S * 100 + M)
    w, mu, var = random_gmm(rng, S, M)
    pick = rng.integers(0, S, size=200), rng.integers(0, M, size=200)
    x = (mu[pick] + rng.normal(size=(200, 39)) * np.sqrt(var[pick]) * 1.5).astype(np.float32)
    ref = OG.gmm_emission_scores(x, w, mu, var)
    img, ss = pack_gmm_image(w, mu, var)
    MP = 1
    while MP < M:
        MP *= 2
    spt = GMM_TILE_N // MP
    n_tiles = (S + spt - 1) // spt
    assert img.shape == (n_tiles * 20 * GMM_TILE_N * 8,) and ss.shape == (n_tiles, 80)
    img = img.reshape(n_tiles, 2, GMM_K // 8, GMM_TILE_N, 8)
    got = np.zeros_like(ref)
    for t in range(n_tiles):
        B, shift, tk = gmm_tile_operand(w, mu, var, t)
        hi = img[t, 0].transpose(0, 2, 1).reshape(GMM_K, GMM_TILE_N).astype(np.float32)
        lo = img[t, 1].transpose(0, 2, 1).reshape(GMM_K, GMM_TILE_N).astype(np.float32)
        assert np.allclose(hi.astype(np.float64) + lo, B, rtol=2e-6, atol=1e-7)
        assert np.allclose(ss[t, :39], shift) and np.allclose(1.0 / ss[t, 40:79], tk)
        z = (x - ss[t, :39]) * ss[t, 40:79]
        assert np.abs(z).max() < 128
        A = np.zeros((len(x), GMM_K), dtype=np.float32)
        A[:, :39] = z * z; A[:, 39] = 1.0; A[:, 40:79] = z
        ah = A.astype(np.float16).astype(np.float32); al = (A - ah).astype(np.float16).astype(np.float32)
        y = (ah @ hi + al @ hi + ah @ lo).astype(np.float32)           # float32 accumulation like the TMEM accumulator
        for i in range(min(spt, S - t * spt)):
            comp = y[:, i * MP:i * MP + M].astype(np.float64)
            mx = comp.max(axis=1)
            got[:, t * spt + i] = mx + np.log(np.exp(comp - mx[:, None]).sum(axis=1))
    assert gmm_close(got, ref, OG.component_constants(w, var)), np.abs(got - ref).max()


def test_pack_refuses_models_outside_the_binary16_pair():
    from loe_speech_recognition.gmm import pack_gmm_image
    rng = np.random.default_rng(0)
    w, mu, var = random_gmm(rng, 5, 4)
    assert pack_gmm_image(w, mu, var) is not None
    var2 = var.copy(); var2[0, 0, 0] = 1e-7                            # 1 / var leaves the range
    assert pack_gmm_image(w, mu, var2) is None
    assert pack_gmm_image(w, mu[:, :, :13], var[:, :, :13]) is None   # built for 39 dimensions
    w0 = w.copy(); w0[1, 2] = 0.0; w0[1] /= w0[1].sum()
    img = pack_gmm_image(w0, mu, var)                                  # a zero-weight component is packed, never wins
    assert img is not None


def test_lexicon_trellis_shapes():
    phones = {"b": np.log(np.full((3, 3), 0.5, np.float32))}
    with np.errstate(divide="ignore"):
        phones["a"] = np.log(np.array([[.6, .4, 0], [0, .6, .4], [0, 0, .7]], np.float32))
    tr, col, sizes = OG.lexicon_trellis(phones, {"a": np.log(0.3), "b": np.log(0.5)}, {"a": 0, "b": 3}, {"X": ["a", "b"], "Y": ["b"]}, ["X", "Y"])
    assert sizes == [6, 3] and col.tolist() == [0, 1, 2, 3, 4, 5, 3, 4, 5]
    assert tr.band[3, 1] == np.float32(np.log(0.3)) and np.isinf(tr.band[3, 2])       # phone a's last state -> phone b's first
    assert tr.loop_starts.tolist() == [0, 6] and tr.loop_ends.tolist() == [5, 8]


# ------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def eng(built_lib):
    from loe_speech_recognition._engine import get_engine
    return get_engine()


@pytest.mark.gpu
@pytest.mark.parametrize("S,M", [(55, 4), (120, 16), (7, 3), (11, 1), (31, 8), (130, 2)])
def test_gmm_emission_matches_oracle(eng, S, M):
    rng = np.random.default_rng(S + M)
    w, mu, var = random_gmm(rng, S, M)
    gp = eng.pack_gmm(w, mu, var)
    assert gp.b_img is not None
    cst = OG.component_constants(w, var)
    for n_frames in (1, 127, 128, 129, 1000, 20000):
        pick = rng.integers(0, S, size=n_frames), rng.integers(0, M, size=n_frames)
        x = (mu[pick] + rng.normal(size=(n_frames, 39)) * np.sqrt(var[pick]) * 1.5).astype(np.float32)
        ref = OG.gmm_emission_scores(x[:2000], w, mu, var)
        xd = eng._to_dev(x)
        got64 = eng.emission_gmm(xd, gp, "fp64").cpu().numpy()
        assert np.allclose(got64[:2000], ref, rtol=1e-6, atol=1e-6), (n_frames, np.abs(got64[:2000] - ref).max())
        for precision in ("fp32", "tc", "auto"):
            got = eng.emission_gmm(xd, gp, precision).cpu().numpy()
            assert got.shape == (n_frames, S)
            assert gmm_close(got, got64.astype(np.float64), cst), (precision, n_frames, np.abs(got - got64).max())


@pytest.mark.gpu
def test_gmm_out_of_range_rows_and_guard_bands(eng):
    """Rows the binary16 operand cannot carry (|z| >= 128, inf, NaN) are evaluated in float32 inside the tensor-core kernel:
    same values as the SIMT kernel; neighbours untouched; nothing written outside the score matrix."""
    torch = eng.torch
    rng = np.random.default_rng(3)
    S, M = 40, 16
    w, mu, var = random_gmm(rng, S, M)
    gp = eng.pack_gmm(w, mu, var)
    n = 700
    x = (mu[rng.integers(0, S, n), rng.integers(0, M, n)] + rng.normal(size=(n, 39))).astype(np.float32)
    big = rng.choice(n, 40, replace=False)
    x[big] *= np.float32(10.0) ** rng.integers(2, 7, size=40)[:, None].astype(np.float32)
    x[5, 3] = np.nan
    x[6, 0] = np.inf
    xd = eng._to_dev(x)
    G = 2048
    buf = torch.full((G + n * S + G,), -777.0, dtype=torch.float32, device=eng.device)
    out = buf[G:G + n * S].view(n, S)
    eng.emission_gmm(xd, gp, "tc", out=out)
    torch.cuda.synchronize()
    assert bool((buf[:G] == -777.0).all()) and bool((buf[G + n * S:] == -777.0).all())
    got = out.cpu().numpy()
    ref = eng.emission_gmm(xd, gp, "fp32").cpu().numpy()
    assert np.all(np.isnan(got[5])) and np.all(np.isnan(ref[5]))
    rows = np.setdiff1d(np.arange(n), [5, 6])
    cst = OG.component_constants(w, var)
    assert np.array_equal(got[big], ref[big])                          # same float32 arithmetic
    assert gmm_close(got[rows], ref[rows].astype(np.float64), cst)
    assert np.array_equal(np.isfinite(got[6]), np.isfinite(ref[6]))


@pytest.mark.gpu
def test_phone_loop_decode_matches_oracle(eng):
    """configs[4] shape: 34 phone models x 3 states = 102 emission columns, 16 Gaussians each, digit loop expanded through a
    pronunciation lexicon (107 trellis positions sharing the phone states); strings and state paths against the oracle
    (oracle GMM scores -> reference loop Viterbi), mismatches adjudicated by the both-paths margin test."""
    from loe_speech_recognition.gmm import DiagGMM, PhoneLoopInference
    from oracle.adjudicate import compare_loop_decodes
    from phone_fixture import make_phone_task
    task = make_phone_task(seed=5, n_mix=16, n_utts=60)
    gmm = DiagGMM(task["weights"], task["means"], task["variances"])
    dec = PhoneLoopInference(gmm, task["phone_logA"], task["phone_log_exit"], task["phone_col"], task["lexicon"], task["order"], penalty=-50)
    strings, scores, paths = dec.decode_batch(task["feats"])
    tr, col, sizes = OG.lexicon_trellis(task["phone_logA"], task["phone_log_exit"], task["phone_col"], task["lexicon"], task["order"])
    assert tr.n_pos == len(col) and tr.n_pos > 64                      # four positions per lane in the Viterbi kernel
    ems = [OG.gmm_emission_scores(x, task["weights"], task["means"], task["variances"])[:, col] for x in task["feats"]]
    v = compare_loop_decodes(ems, tr, -50, sizes, task["order"], paths, strings)
    print("phone loop, 16 mixtures:", v.summary(), "accuracy vs truth", np.mean([a == b for a, b in zip(strings, task["truth"])]))
    assert not v.failed, v.failed[:3]
    assert v.identical_strings >= len(strings) - v.excused
    assert np.mean([a == b for a, b in zip(strings, task["truth"])]) >= 0.9
    # exact given the kernel's own scores
    sc = gmm.scores_batch(task["feats"])
    _, _, opaths = O.viterbi_batch([s[:, col] for s in sc], tr, penalty=-50)
    assert all(np.array_equal(a, b) for a, b in zip(opaths, paths))
