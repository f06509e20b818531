"""Pin the MFCC oracle against the real librosa, on any box that has it.

    python tests/golden/make_golden_mfcc_librosa.py        (needs `import librosa`; writes tests/golden/golden_mfcc_librosa.npz)

The reference's front end is five librosa calls (src/loe_speech_recognition/mfcc.py:31-40); librosa is un-pinned in the
reference (pyproject.toml:17-21), un-vendored, and not installable in the authoring container, so oracle/mfcc.py is a
restatement and row a1 is "parity unpinned".  This generator closes that gap wherever librosa imports: it runs exactly the
reference's calls on seeded signals and stores inputs, outputs and the librosa / numpy / scipy versions.
tests/test_oracle_golden.py::test_mfcc_oracle_against_librosa_fixture compares the oracle with the stored vectors when the
file exists (and skips otherwise); the GPU parity tests then inherit the pin through the oracle.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_feature_vector(librosa, signal, sample_rate=16000, n_mfcc=13):
    """mfcc.py:31-43 of the reference, call for call."""
    mel = librosa.feature.melspectrogram(y=signal, sr=sample_rate, n_mels=40, n_fft=320, hop_length=160, fmin=133.33, fmax=6855.4976)
    log_mel = librosa.power_to_db(mel, ref=np.max)
    mfccs = librosa.feature.mfcc(S=log_mel, sr=sample_rate, n_mfcc=n_mfcc)
    d1 = librosa.feature.delta(mfccs)
    d2 = librosa.feature.delta(mfccs, order=2)
    mean = np.mean(mfccs, axis=0, keepdims=True)
    std = np.std(mfccs, axis=0, keepdims=True)
    return np.concatenate(((mfccs - mean) / (std + 1e-8), d1, d2), axis=0), mel, log_mel


def main():
    try:
        import librosa
    except Exception as e:                                  # noqa: BLE001
        print(f"librosa is not importable here ({e!r}): nothing written")
        return 1
    import scipy
    rng = np.random.default_rng(2024)
    out = {"versions": np.array([f"librosa {librosa.__version__}", f"numpy {np.__version__}", f"scipy {scipy.__version__}"])}
    t = np.arange(64000)
    for i, n in enumerate((1440, 1600, 16000, 23457, 64000)):
        f = rng.uniform(200, 4000, 3)
        sig = sum(3000 * np.sin(2 * np.pi * fi * t[:n] / 16000 + rng.uniform(0, 6)) for fi in f) + rng.normal(0, 30, n)
        sig = np.round(sig).astype(np.float32)              # float32 at int16 scale: what ti_digits.py:133 hands to MFCC
        feat, mel, log_mel = reference_feature_vector(librosa, sig)
        out[f"pcm{i}"] = sig
        out[f"feat{i}"] = np.asarray(feat)
        out[f"mel{i}"] = np.asarray(mel)
        out[f"logmel{i}"] = np.asarray(log_mel)
    path = os.path.join(HERE, "golden_mfcc_librosa.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "with", librosa.__version__)
    return 0


if __name__ == "__main__":
    sys.exit(main())
