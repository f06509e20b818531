"""Golden vectors for template DTW: runs the UNMODIFIED reference DynamicTimeWarping (authoring container
only).  Its ``__post_init__`` computes MFCCs through librosa (absent), so the reference module's ``MFCC``
name is bound to the restated front end (oracle/mfcc.py) for this run -- the DTW code itself is untouched.

    python tests/golden/make_golden_dtw.py
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

from oracle import dtw as OD                      # noqa: E402
from oracle import mfcc as OM                     # noqa: E402
from oracle.ref_import import import_reference    # noqa: E402

CASES = [("3", 4, False, True), ("1", 7, True, True), ("5", 0.2, False, True), ("2", 4, True, True), ("4", 4, False, False)]


def dtw_signals():
    path = os.path.join(ROOT, "cs-304-speech-recognition-code_b200", "loe_speech_recognition", "synthetic.py")
    spec = importlib.util.spec_from_file_location("loe_synth_for_dtw", path)
    S = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = S
    spec.loader.exec_module(S)
    rng = np.random.default_rng(4)
    templates = [S.synth_isolated(rng, w, 0.25) for w in S.DIGITS[:5] for _ in range(2)]
    samples = [S.synth_isolated(rng, w, 0.3) for w, _, _, _ in CASES]
    return templates, samples


def main():
    import_reference()
    import loe_speech_recognition.dynamic_time_wrapping as RD

    class FrontEnd:
        def __init__(self, signal, sample_rate=16000):
            self.feature_vector = OM.mfcc_feature_vector(signal)
    RD.MFCC = FrontEnd
    templates, samples = dtw_signals()
    feats = [OM.mfcc_feature_vector(t).T for t in templates]
    out = {f"tmpl_feat_{i}": np.ascontiguousarray(f) for i, f in enumerate(feats)}
    for c, ((w, pf, tb, pr), sample) in enumerate(zip(CASES, samples)):
        d = RD.DynamicTimeWarping(templates, sample, pruning=pr, pruning_factor=pf, trace_back=tb)
        idx, dist = d.search()
        sf = OM.mfcc_feature_vector(sample).T
        oi, od, oc, op = OD.search(feats, sf, pr, pf, tb)
        assert idx == oi and dist == od and np.array_equal(oc, d._cost_matrix) and np.array_equal(op, d._path_matrix), c
        out[f"samp_feat_{c}"] = np.ascontiguousarray(sf)
        out[f"index_{c}"] = np.int64(idx)
        out[f"dist_{c}"] = np.float64(dist)
        out[f"cost_{c}"] = d._cost_matrix
        out[f"path_{c}"] = d._path_matrix.astype(np.int8)
    if not os.environ.get("LOE_DTW_DRY"):
        np.savez_compressed(os.path.join(HERE, "golden_dtw.npz"), **out)
    print("wrote golden_dtw.npz:", [int(out[f"index_{c}"]) for c in range(len(CASES))])


if __name__ == "__main__":
    main()
