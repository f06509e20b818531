"""Silence stripper: oracle vs the reference's outputs (golden, and the live class when
/root/reference exists), and the CUDA kernel vs the oracle -- bit-exact decisions and energies."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from make_golden_vad import vad_signals           # noqa: E402
from oracle import vad as OV                      # noqa: E402

KW = dict(sample_rate=16000, high=0.06, low=0.01)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_vad.npz"))


def test_oracle_matches_reference_golden(gold):
    sigs = vad_signals()
    st = OV.Stripper(**KW)
    for i, s in enumerate(sigs):
        r = OV.segment(s, **KW)
        assert np.array_equal(np.array([r["done"], r["start"], r["end"], len(r["energies"])]), gold[f"seg_{i}"])
        assert np.array_equal(r["noise_mask"], gold[f"noise_{i}"])
        assert np.array_equal(r["energies"], gold[f"energy_{i}"], equal_nan=True)
        out = st.remove_empty(s)
        assert (out is not None) == bool(gold[f"ok_{i}"])
        if out is not None:
            assert len(out) == gold[f"len_{i}"] and np.sum(out.astype(np.float64)) == gold[f"sum_{i}"]
    assert len(st.noises) == gold["n_noises"]
    assert [len(x) for x in st.noises] == gold["noise_lens"].tolist()
    assert np.array_equal([np.sum(x.astype(np.float64)) for x in st.noises], gold["noise_sums"])
    assert np.isnan(gold["energy_13"][-1])            # 3200 samples: the trailing partial frame is empty


@pytest.mark.gpu
def test_kernel_matches_oracle_bit_exact(built_lib):
    from loe_speech_recognition._engine import get_engine
    eng = get_engine()
    sigs = vad_signals()
    rng = np.random.default_rng(8)
    sigs += [rng.normal(0, 2000, size=n).round().astype(np.float32) for n in (1, 159, 160, 161, 1600, 20001)]
    for cast in (np.float32, np.int16):
        energy, noise, seg, mx, eoff = eng.silence([s.astype(cast) for s in sigs], 160, 0.06, 0.01, 2)
        for i, s in enumerate(sigs):
            r = OV.segment(s, **KW)
            assert np.array_equal(energy[eoff[i]:eoff[i + 1]], r["energies"], equal_nan=True), i
            assert seg[i].tolist() == [int(r["done"]), r["start"], r["end"], len(r["energies"])], i
            upto = r["end"] + 1 if r["done"] else len(r["energies"])
            assert np.array_equal(noise[eoff[i]:eoff[i + 1]][:upto], r["noise_mask"][:upto]), i
            assert mx[i] == np.float32(r["max_volume"])


@pytest.mark.gpu
def test_class_api_matches_reference_semantics(built_lib, gold):
    from loe_speech_recognition import SignalSeparation
    sigs = vad_signals()
    ss = SignalSeparation(sample_rate=16000, speech_high_threshold=0.06, speech_low_threshold=0.01)
    st = OV.Stripper(**KW)
    want = [st.remove_empty(s) for s in sigs]
    got = ss.remove_empty_batch(sigs)
    assert len(got) == sum(w is not None for w in want)
    for g, w in zip(got, [w for w in want if w is not None]):
        assert g.dtype == np.float32 and np.array_equal(g, w)
    assert len(ss.get_all_noises()) == len(st.noises) == gold["n_noises"]
    assert all(np.array_equal(a, b) for a, b in zip(ss.get_all_noises(), st.noises))
    one = SignalSeparation(sample_rate=16000, speech_high_threshold=0.06, speech_low_threshold=0.01)
    assert np.array_equal(one.remove_empty(sigs[0]), want[0])
    with pytest.raises(SignalSeparation.FailToProcess):
        one.remove_empty(sigs[13])
    assert one.frame_size == 160 and one.maximum_silence_frames == 2


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/loe_speech_recognition"), reason="reference not present")
def test_oracle_matches_live_reference():
    import subprocess
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "tests", "golden", "make_golden_vad.py")], text=True,
                                  cwd="/tmp", env=dict(os.environ, LOE_VAD_DRY="1"))
    assert "wrote golden_vad.npz" in out
