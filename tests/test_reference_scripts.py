"""The reference's own driver scripts, UNMODIFIED, against the drop-in package (SURVEY.md §2 row 18:
scripts/*.py are the acceptance harness).

Ten of the reference's twelve non-interactive drivers run this way (the rest need audio hardware, and project4_phone.py opens
its pool at module level without a __main__ guard, which only fork() can serve).  The first three:
project3_train.py (isolated training + save), project5_test_ndigits_with_sil.py (loop decode of 1 / 2 / 4 / 7-digit strings over
a ProcessPoolExecutor, accuracy lines, '|'-separated CSV tables) and project6_train.py (embedded training) are executed as
files in a scratch working directory that holds a synthetic ./ConvertedTIDigits tree and the seed models, with nothing but
PYTHONPATH pointing at this repository's package.  Expected outputs (tests/golden/golden_scripts.*) come from the same script
files run against the real reference package on the CPU (tests/golden/make_golden_scripts.py).

The script files are read from /root/reference/scripts, or from the staging copy oracle/_ref/scripts on the GPU box
(oracle/stage_reference_scripts.py; git-ignored).  Without either the tests skip.
"""
import json
import os
import sys

import numpy as np
import pytest

import ref_scripts_harness as H
from helpers import rel_close

pytestmark = pytest.mark.gpu
GOLD = os.path.join(H.ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def workspace(tmp_path_factory, built_lib, golden):
    if H.scripts_dir() is None:
        pytest.skip("reference scripts not available (neither /root/reference/scripts nor oracle/_ref/scripts)")
    ws = str(tmp_path_factory.mktemp("refscripts"))
    made = H.build_corpus(ws)
    record = json.load(open(os.path.join(GOLD, "golden_scripts.json")))
    assert made == record["corpus"]                       # same seeded corpus as the reference arm saw
    cwd = os.getcwd()
    os.chdir(ws)
    try:
        H.seed_all_models(golden)
    finally:
        os.chdir(cwd)
    return ws, record


def _models(folder):
    from loe_speech_recognition import HiddenMarkovModel
    out = {}
    for name in sorted(os.listdir(folder)):
        m = HiddenMarkovModel.from_folder(os.path.join(folder, name))
        out[name] = (np.stack([np.asarray(mn._core.mean) for mn in m._multivariate_normals]),
                     np.stack([np.asarray(mn._core.cov_object.covariance) for mn in m._multivariate_normals]),
                     m._log_transition_probs.to_dense())
    return out


def _same_models(got, gold, prefix, rtol_mean=1e-3, rtol_cov=1e-2):
    assert sorted(got) == sorted(k[len(prefix) + 7:] for k in gold.files if k.startswith(prefix + "_means_"))
    for name, (mean, cov, logA) in got.items():
        assert rel_close(mean, gold[f"{prefix}_means_{name}"], rtol=rtol_mean, atol=1e-4), name
        assert rel_close(cov, gold[f"{prefix}_covs_{name}"], rtol=rtol_cov, atol=1e-4), name
        ref = gold[f"{prefix}_logA_{name}"]
        assert np.array_equal(np.isnan(logA), np.isnan(ref)), name
        ok = np.isfinite(ref)
        assert np.array_equal(np.isfinite(logA), ok) and np.allclose(logA[ok], ref[ok], atol=2e-3), name


def test_project3_train_unmodified(workspace):
    ws, record = workspace
    r = H.run_script("project3_train.py", ws, [H.PKG])
    assert r.returncode == 0, r.stderr[-3000:]
    gold = np.load(os.path.join(GOLD, "golden_scripts.npz"))
    _same_models(_models(os.path.join(ws, ".cache", "big_model_no_silence")), gold, "p3")


def test_project5_test_ndigits_with_sil_unmodified(workspace):
    ws, record = workspace
    name = "project5_test_ndigits_with_sil.py"
    r = H.run_script(name, ws, [H.PKG], extra_env={"LOE_B200_WORKER_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-3000:]
    got = [l for l in r.stdout.splitlines() if l.startswith("In total")]
    print("\n".join(got))
    # every CSV the script writes: same rows as the reference wrote.  A differing prediction would have to be a near-tie of
    # the two decoders (tests/test_gpu_parity.py adjudicates those on the same models); none is expected on this corpus.
    n_rows = n_diff = 0
    for f, lines in record["csv"].items():
        mine = H.read_csv(os.path.join(ws, "plots", f))
        assert mine[0] == lines[0] and len(mine) == len(lines), f
        n_rows += len(lines) - 1
        n_diff += sum(a != b for a, b in zip(sorted(mine[1:]), sorted(lines[1:])))
    print(f"{n_rows} decoded utterances in {len(record['csv'])} CSV files, {n_diff} rows differ from the reference's")
    assert n_diff == 0
    assert got == record["stdout"][name]


def test_project6_train_unmodified(workspace):
    """Embedded training.  On this small synthetic corpus the REFERENCE dies in its second iteration with HMMTrainMeanFail (a
    state of one word collects no frame) and its `finally` saves the models of the first iteration: the drop-in must do the
    same -- same exception, same number of completed iterations, same saved models."""
    ws, record = workspace
    name = "project6_train.py"
    r = H.run_script(name, ws, [H.PKG])
    assert (r.returncode != 0) == (record["returncode"][name] != 0), r.stderr[-3000:]
    assert H.last_exception(r.stderr) == record["exception"][name], r.stderr[-3000:]
    assert H.iterations_done(r.stderr) == record["iterations"][name]
    gold = np.load(os.path.join(GOLD, "golden_scripts.npz"))
    _same_models(_models(os.path.join(ws, ".cache", "big_model_speech_only_continuous_2")), gold, "p6")


# ------------------------------------------------------------------------------------------------------------------------
# The other non-interactive drivers (ref_scripts_harness.EXTRA_SCRIPTS; golden_scripts_extra.* from
# `make_golden_scripts.py --extra`).  Both arms get tests' environment stubs on PYTHONPATH: a sitecustomize that seeds
# Python's global RNG (project4_2digits.py samples with an unseeded random.sample) and a permissive matplotlib.
# ------------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def extra(workspace):
    ws, _ = workspace
    record = json.load(open(os.path.join(GOLD, "golden_scripts_extra.json")))
    env = H.write_env_stubs(os.path.join(ws, "_env"))
    return ws, record, [H.PKG, env]


def _run_extra(name, extra, cwd=None):
    ws, record, pp = extra
    cwd = cwd or ws
    before = set(os.listdir(os.path.join(cwd, "plots")))
    start = H.log_size(cwd)
    r = H.run_script(name, cwd, pp)
    assert r.returncode == record["returncode"][name], r.stderr[-3000:]
    assert H.last_exception(r.stderr) == record["exception"][name], r.stderr[-3000:]
    csv = {f: H.read_csv(os.path.join(cwd, "plots", f)) for f in sorted(set(os.listdir(os.path.join(cwd, "plots"))) - before)
           if f.endswith(".csv")}
    return r, csv, start


def _same_tables(csv, gold):
    assert sorted(csv) == sorted(gold)
    n_rows = n_diff = 0
    for f, lines in gold.items():
        assert csv[f][0] == lines[0] and len(csv[f]) == len(lines), f
        n_rows += len(lines) - 1
        n_diff += sum(a != b for a, b in zip(sorted(csv[f][1:]), sorted(lines[1:])))
    print(f"{n_rows} decoded utterances in {len(gold)} CSV files, {n_diff} rows differ from the reference's")
    assert n_diff == 0


def test_project3_predict_simple_unmodified(extra):
    """Isolated-word classifier: ModelCollection.predict mapped over one ProcessPoolExecutor per label (22 pools)."""
    name = "project3_predict_simple.py"
    r, _, _ = _run_extra(name, extra)
    print("\n".join(H.stdout_record(r.stdout)))
    assert H.stdout_record(r.stdout) == extra[1]["stdout"][name]


def test_project4_2digits_unmodified(extra):
    """ModelCollection.predict on ten concatenated digit pairs: same pairs drawn (seeded RNG), same predictions logged."""
    name = "project4_2digits.py"
    r, _, start = _run_extra(name, extra)
    logged = H.logged_predictions(extra[0], start)
    assert len(logged) == 10 and logged == extra[1]["logged"][name], (logged, extra[1]["logged"][name])


@pytest.mark.parametrize("name", ["project5_test_1digit.py", "project5_test_ndigits_no_sil.py"])
def test_project5_test_without_silence_unmodified(extra, name):
    """Loop decode without the silence model: default float64 penalty np.log(0.005) (1digit) and the int penalty -250 (7 digits)."""
    r, csv, _ = _run_extra(name, extra)
    print("\n".join(H.stdout_record(r.stdout)))
    _same_tables(csv, extra[1]["csv"][name])
    assert H.stdout_record(r.stdout) == extra[1]["stdout"][name]


@pytest.mark.parametrize("name", ["project5_find_trans_ndigits_no_sil.py", "project5_find_trans_ndigits_with_sil.py"])
def test_project5_find_trans_unmodified(extra, name):
    """Penalty sweeps (20 / 100 values poked into _log_transition_probability_between_words, one process pool per value):
    the same accuracy at every penalty.  The 100-value sweep spends ~3 s per pool on CUDA start-up of the fresh workers
    (305 s on the B200 box; passed, log in profiles/r2d_reference_scripts.log), so it only runs with LOE_TEST_FULL_SWEEP=1."""
    if name.endswith("with_sil.py") and not os.environ.get("LOE_TEST_FULL_SWEEP"):
        pytest.skip("100 process pools (~5 min): set LOE_TEST_FULL_SWEEP=1")
    r, _, _ = _run_extra(name, extra)
    got, want = H.stdout_record(r.stdout), extra[1]["stdout"][name]
    assert len(got) == len(want) and len(got) in (40, 200)
    assert got == want, [(a, b) for a, b in zip(got, want) if a != b][:6]


def test_project5_train_no_empty_unmodified(extra):
    """Silence stripper -> MFCC -> segmental K-means for the 11 digits, then the 3-state silence model from the collected
    noise (SignalSeparation.get_all_noises): same outcome and same trained models as the reference."""
    name = "project5_train_no_empty.py"
    ws2 = H.second_workspace(extra[0])
    _run_extra(name, extra, cwd=ws2)
    gold = np.load(os.path.join(GOLD, "golden_scripts_extra.npz"))
    folder = os.path.join(ws2, ".cache", "big_model_speech_only")
    if not any(k.startswith("p5t_") for k in gold.files):
        assert not os.path.isdir(folder) or not os.listdir(folder)
        return
    _same_models(_models(folder), gold, "p5t")
