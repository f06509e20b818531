"""Golden vectors for the silence stripper: runs the UNMODIFIED reference SignalSeparation
(/root/reference, authoring container only) on seeded synthetic signals and stores its outputs.

    python tests/golden/make_golden_vad.py
"""
import importlib.util
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import vad as OV                      # noqa: E402
from oracle.ref_import import import_reference    # noqa: E402


def vad_signals():
    """Deterministic test signals (shared with tests/test_vad.py)."""
    path = os.path.join(ROOT, "cs-304-speech-recognition-code_b200", "loe_speech_recognition", "synthetic.py")
    spec = importlib.util.spec_from_file_location("loe_synth_for_vad", path)
    S = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = S
    spec.loader.exec_module(S)
    rng = np.random.default_rng(3)
    sigs = [S.synth_string(rng, [S.DIGITS[i % 11]]) for i in range(12)]
    sigs += [S.synth_isolated(rng, "5", 0.4),                                  # speech never ends -> fail, noise carried over
             rng.normal(0, 30, size=3200).round().astype(np.float32),            # noise only
             S.synth_string(rng, ["1", "2"])[:4000],                            # cut inside the word
             S.synth_string(rng, ["3"])[:4800 + 77],                            # trailing partial frame
             S.synth_string(rng, ["7", "O", "4"]),                              # stops after the first word
             np.concatenate([S.synth_string(rng, ["9"])[:3040], np.zeros(1600, np.float32)])]
    return sigs


def main():
    R = import_reference()
    sigs = vad_signals()
    kw = dict(sample_rate=16000, speech_high_threshold=0.06, speech_low_threshold=0.01)
    ref = R.SignalSeparation(**kw)
    mine = OV.Stripper(sample_rate=16000, high=0.06, low=0.01)
    out = {}
    for i, s in enumerate(sigs):
        try:
            a = ref.remove_empty(s)
        except R.SignalSeparation.FailToProcess:
            a = None
        b = mine.remove_empty(s)
        assert (a is None) == (b is None) and (a is None or np.array_equal(a, b)), i
        r = OV.segment(s, sample_rate=16000, high=0.06, low=0.01)
        out[f"ok_{i}"] = np.bool_(a is not None)
        out[f"len_{i}"] = np.int64(0 if a is None else len(a))
        out[f"sum_{i}"] = np.float64(0 if a is None else np.sum(a.astype(np.float64)))
        out[f"seg_{i}"] = np.array([r["done"], r["start"], r["end"], len(r["energies"])], dtype=np.int32)
        out[f"noise_{i}"] = r["noise_mask"]
        out[f"energy_{i}"] = r["energies"]
    noises = ref.get_all_noises()
    assert len(noises) == len(mine.noises) and all(np.array_equal(x, y) for x, y in zip(noises, mine.noises))
    out["n_noises"] = np.int64(len(noises))
    out["noise_lens"] = np.array([len(x) for x in noises])
    out["noise_sums"] = np.array([np.sum(x.astype(np.float64)) for x in noises])
    if not os.environ.get("LOE_VAD_DRY"):              # the live-reference test only re-checks the assertions
        np.savez_compressed(os.path.join(HERE, "golden_vad.npz"), **out)
    print("wrote golden_vad.npz:", sum(bool(out[f"ok_{i}"]) for i in range(len(sigs))), "of", len(sigs), "stripped;", len(noises), "noise clips")


if __name__ == "__main__":
    main()
