"""Template DTW: oracle vs the reference's outputs (golden, and the live class when
/root/reference exists), and the CUDA kernel vs both -- cost matrices, path codes, index and distance bit-exact."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from make_golden_dtw import CASES                 # noqa: E402
from oracle import dtw as OD                      # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_dtw.npz"))


def _templates(gold):
    return [gold[f"tmpl_feat_{i}"] for i in range(10)]


def test_oracle_matches_reference_golden(gold):
    feats = _templates(gold)
    for c, (w, pf, tb, pr) in enumerate(CASES):
        idx, dist, cost, path = OD.search(feats, gold[f"samp_feat_{c}"], pr, pf, tb)
        assert idx == gold[f"index_{c}"] and dist == gold[f"dist_{c}"]
        assert np.array_equal(cost, gold[f"cost_{c}"]) and np.array_equal(path, gold[f"path_{c}"])
    assert np.isinf(gold["dist_2"])                   # pruning factor 0.2 prunes every path


@pytest.mark.gpu
def test_kernel_matches_reference_bit_exact(built_lib, gold):
    from loe_speech_recognition import DynamicTimeWarping
    feats = _templates(gold)
    for c, (w, pf, tb, pr) in enumerate(CASES):
        d = DynamicTimeWarping.from_features(feats, gold[f"samp_feat_{c}"], trace_back=tb, pruning=pr, pruning_factor=pf)
        idx, dist = d.search()
        assert idx == gold[f"index_{c}"] and dist == gold[f"dist_{c}"] and isinstance(dist, np.float64)
        assert np.array_equal(d._cost_matrix, gold[f"cost_{c}"])
        assert np.array_equal(d._path_matrix, gold[f"path_{c}"])
    # batch entry point == one call per sample; 13-dimensional features take the generic kernel
    d = DynamicTimeWarping.from_features(feats, gold["samp_feat_0"], pruning_factor=4)
    samples = [gold[f"samp_feat_{c}"] for c in range(len(CASES))]
    bi, bd = d.search_batch(samples)
    for c, s in enumerate(samples):
        oi, od, _, _ = OD.search(feats, s, True, 4, False)
        assert bi[c] == oi and bd[c] == od
    f13 = [f[:, :13].copy() for f in feats]
    d13 = DynamicTimeWarping.from_features(f13, samples[1][:, :13].copy(), pruning_factor=7, trace_back=True)
    idx, dist = d13.search()
    oi, od, oc, op = OD.search(f13, samples[1][:, :13].copy(), True, 7, True)
    assert idx == oi and dist == od and np.array_equal(d13._cost_matrix, oc) and np.array_equal(d13._path_matrix, op)


@pytest.mark.gpu
def test_raw_signal_api(built_lib):
    from loe_speech_recognition import DynamicTimeWarping, MFCC
    from make_golden_dtw import dtw_signals
    templates, samples = dtw_signals()
    d = DynamicTimeWarping(templates, samples[0], pruning=True, pruning_factor=7)
    idx, dist = d.search()
    feats = MFCC.batch(templates + [samples[0]], 16000)
    oi, od, _, _ = OD.search(feats[:-1], feats[-1], True, 7, False)
    assert idx == oi and dist == od
    assert idx // 2 == 2                              # template pair of digit "3"


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/loe_speech_recognition"), reason="reference not present")
def test_oracle_matches_live_reference():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "tests", "golden", "make_golden_dtw.py")], text=True,
                                  cwd="/tmp", env=dict(os.environ, LOE_DTW_DRY="1"))
    assert "wrote golden_dtw.npz" in out
