"""Workspace for running the reference's own driver scripts UNMODIFIED (``scripts/*.py`` are the
acceptance harness).

The scripts (project3_train.py, project5_test_ndigits_with_sil.py, project6_train.py) read ``./ConvertedTIDigits``,
``./.cache/<model>`` and write ``./runtime.log``, ``./plots/*.csv`` and ``./.cache/<model>``, all relative to the working
directory, and import ``loe_speech_recognition`` from PYTHONPATH.  ``build_workspace`` lays down a synthetic corpus tree
and the seed models; ``run_script`` executes one script file with a chosen PYTHONPATH -- the drop-in package (GPU) or the
reference's ``src`` plus stub modules for its absent third-party imports (CPU, authoring container only).

Script files are taken from /root/reference/scripts when that exists, else from oracle/_ref/scripts (a git-ignored
staging copy that travels to the GPU box, made by oracle/stage_reference_scripts.py).  They are never modified.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cs-304-speech-recognition-code_b200")
SCRIPTS = ("project3_train.py", "project5_test_ndigits_with_sil.py", "project6_train.py")
# the rest of the reference's non-interactive drivers.  Not runnable under any replacement that owns a CUDA context:
# project4_phone.py opens its process pool at module level without a __main__ guard (only fork() can serve that, and a CUDA
# context does not survive fork); the *_interactive.py / record.py / mic_testing.py / play_all.py / project1.py drivers need
# audio hardware.
EXTRA_SCRIPTS = ("project3_predict_simple.py", "project4_2digits.py", "project5_test_1digit.py", "project5_test_ndigits_no_sil.py",
                 "project5_find_trans_ndigits_no_sil.py", "project5_find_trans_ndigits_with_sil.py", "project5_train_no_empty.py")
MODEL_ORDER = ("1", "2", "3", "4", "5", "6", "7", "8", "9", "O", "S", "Z")


def scripts_dir():
    for d in ("/root/reference/scripts", os.path.join(ROOT, "oracle", "_ref", "scripts")):
        if all(os.path.exists(os.path.join(d, s)) for s in SCRIPTS + EXTRA_SCRIPTS):
            return d
    return None


def write_seed_models(folder: str, golden, order=MODEL_ORDER) -> None:
    """Reference-format model folders (<label>/{log_trans_probs,multivariate_normals}.pickle) of the 12 word models the
    unmodified reference trained for tests/golden/golden_hmm.npz (``order``: which of them).  Written by whichever
    ``loe_speech_recognition`` is importable in the calling process (the pickles embed that module path, and both packages
    read each other's files)."""
    from loe_speech_recognition.hidden_markov_model import HiddenMarkovModel, HiddenMarkovModelTrainable
    from loe_speech_recognition.transition_probability import LogTransitionProbabilities
    for w in order:
        m = HiddenMarkovModel(w)
        m._multivariate_normals = HiddenMarkovModelTrainable.get_multivariate_normals(golden[f"train_means_{w}"], golden[f"train_covs_{w}"])
        ltp = LogTransitionProbabilities()
        dense = golden[f"train_logA_{w}"]
        ltp.num_of_states = int(dense.shape[0])
        for i in range(dense.shape[0]):
            for j in range(dense.shape[1]):
                ltp._core[(i, j)] = dense[i, j]
        m._log_transition_probs = ltp
        m.save(folder)


def build_corpus(root: str, seed: int = 7, n_train_iso: int = 6, n_test_iso: int = 2) -> dict:
    """``root/ConvertedTIDigits/Adults/TIDIGITS/{TRAIN,TEST}/<speaker>/<digits><production letter>.WAV`` -- int16 PCM from
    the seeded synthetic generator (file naming: ti_digits.py:125-129).  Returns {split: {label: n_files}}."""
    sys.path.insert(0, PKG)
    from scipy.io import wavfile
    # the generator module is plain NumPy and has no package-relative imports: load it by path so that this also works in
    # a process that has the REFERENCE package imported under the same name
    import importlib.util
    spec = importlib.util.spec_from_file_location("_loe_synth", os.path.join(PKG, "loe_speech_recognition", "synthetic.py"))
    synth = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synth)
    rng = np.random.default_rng(seed)
    plan = {"TRAIN": {"iso": n_train_iso, "strings": {2: 4, 3: 3, 4: 3, 5: 3, 6: 3, 7: 4}},
            "TEST": {"iso": n_test_iso, "strings": {2: 1, 4: 1, 7: 1}}}
    made = {}
    for split, p in plan.items():
        d = os.path.join(root, "ConvertedTIDigits", "Adults", "TIDIGITS", split, "SYN", "AA")
        os.makedirs(d, exist_ok=True)
        made[split] = {}
        for w in synth.DIGITS:
            for k in range(p["iso"]):
                pcm = synth.synth_string(rng, [w])                  # S + digit + S, like a TIDIGITS isolated-digit file
                wavfile.write(os.path.join(d, f"{w}{chr(65 + k)}.WAV"), 16000, pcm.astype(np.int16))
            made[split][w] = p["iso"]
        for n, count in p["strings"].items():
            for _ in range(count):
                while True:
                    ds = [synth.DIGITS[int(i)] for i in rng.integers(0, len(synth.DIGITS), size=n)]
                    lab = "".join(ds)
                    if lab not in made[split]:
                        break
                wavfile.write(os.path.join(d, f"{lab}A.WAV"), 16000, synth.synth_string(rng, ds).astype(np.int16))
                made[split][lab] = 1
    os.makedirs(os.path.join(root, "plots"), exist_ok=True)
    return made


def write_stubs(folder: str) -> str:
    """Stub modules for the reference's third-party imports that are absent in the authoring container.  ``librosa`` is
    NOT a no-op: it forwards the calls mfcc.py:31-40 makes to the restated oracle (oracle/mfcc.py), so the reference arm
    runs the reference's own HMM code on the oracle's features."""
    os.makedirs(os.path.join(folder, "librosa"), exist_ok=True)
    os.makedirs(os.path.join(folder, "matplotlib"), exist_ok=True)
    open(os.path.join(folder, "librosa", "__init__.py"), "w").write(f'''
import sys
sys.path.insert(0, {ROOT!r})
import numpy as np
import scipy.fft
from oracle import mfcc as _OM

def power_to_db(S, ref=1.0, **kw):
    assert ref is np.max
    return _OM.power_to_db(S)

class feature:
    @staticmethod
    def melspectrogram(y=None, sr=16000, n_mels=40, n_fft=320, hop_length=160, fmin=133.33, fmax=6855.4976):
        assert (n_mels, n_fft, hop_length) == (40, 320, 160)
        return np.einsum("ft,mf->mt", _OM.stft_power(y), _OM.mel_basis(sr, fmin=fmin, fmax=fmax), optimize=True)

    @staticmethod
    def mfcc(S=None, sr=16000, n_mfcc=13):
        return scipy.fft.dct(S, axis=-2, type=2, norm="ortho")[:n_mfcc, :]

    @staticmethod
    def delta(m, order=1):
        return _OM.delta(m, order)
''')
    for name in ("sounddevice.py", "uniplot.py", "soundfile.py", os.path.join("matplotlib", "__init__.py"), os.path.join("matplotlib", "pyplot.py")):
        open(os.path.join(folder, name), "w").write(_PERMISSIVE)
    return folder


_PERMISSIVE = '''
class _Any:
    def __call__(self, *a, **k): return _Any()
    def __getattr__(self, n):
        if n.startswith("__"): raise AttributeError(n)
        return _Any()
def __getattr__(n):
    if n.startswith("__"): raise AttributeError(n)
    return _Any()
'''


DIGIT_ORDER = tuple(w for w in MODEL_ORDER if w != "S")


def seed_all_models(golden) -> None:
    """Every model folder the scripts load, under ./.cache of the current directory: the 12-model sets of the silence-aware
    drivers and ``big_model`` = the 11 digit models (project3_predict_simple.py:42, project4_2digits.py:24,
    project5_test_1digit.py:66, project5_test_ndigits_no_sil.py:58)."""
    write_seed_models(".cache/big_model_speech_only_3", golden)
    write_seed_models(".cache/big_model_speech_only", golden)
    write_seed_models(".cache/big_model", golden, DIGIT_ORDER)


def write_env_stubs(folder: str) -> str:
    """What BOTH arms get on PYTHONPATH next to the package under test (test infrastructure, no arithmetic): a
    ``sitecustomize`` that seeds Python's global RNG -- project4_2digits.py draws its ten digit pairs with an unseeded
    ``random.sample`` -- and a permissive ``matplotlib`` for the plot calls (absent from this image)."""
    os.makedirs(os.path.join(folder, "matplotlib"), exist_ok=True)
    open(os.path.join(folder, "sitecustomize.py"), "w").write("import random\nrandom.seed(304)\n")
    for name in (os.path.join("matplotlib", "__init__.py"), os.path.join("matplotlib", "pyplot.py")):
        open(os.path.join(folder, name), "w").write(_PERMISSIVE)
    return folder


def second_workspace(ws: str, name: str = "ws_train") -> str:
    """A sibling working directory sharing the corpus (symlink) with its own ./.cache and ./plots: project5_train_no_empty.py
    WRITES .cache/big_model_speech_only, which project6_train.py reads as its seed."""
    ws2 = os.path.join(ws, name)
    os.makedirs(os.path.join(ws2, "plots"), exist_ok=True)
    link = os.path.join(ws2, "ConvertedTIDigits")
    if not os.path.exists(link):
        os.symlink(os.path.join(ws, "ConvertedTIDigits"), link)
    return ws2


def log_size(cwd: str) -> int:
    p = os.path.join(cwd, "runtime.log")
    return os.path.getsize(p) if os.path.exists(p) else 0


def logged_predictions(cwd: str, start: int):
    """(ground truth, prediction) pairs project4_2digits.py logs (scripts/project4_2digits.py:33), from byte ``start`` of
    ./runtime.log on."""
    import re
    with open(os.path.join(cwd, "runtime.log")) as f:
        f.seek(start)
        return [[gt, pred] for pred, gt in re.findall(r"Predict labels: (\S*), ground truth: (\S+)", f.read())]


def stdout_record(stdout: str):
    """The result lines the drivers print: accuracies and the penalty of a sweep step."""
    return [l for l in stdout.splitlines() if l.startswith(("In total", "Accuracy of", "For Log Transition"))]


def run_script(name: str, cwd: str, pythonpath, timeout: int = 3000, extra_env=None) -> subprocess.CompletedProcess:
    sd = scripts_dir()
    assert sd is not None, "reference scripts not found"
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(list(pythonpath))
    env.pop("LOE_REFERENCE_SRC", None)
    if extra_env:
        env.update(extra_env)
    return subprocess.run([sys.executable, os.path.join(sd, name)], cwd=cwd, env=env, capture_output=True, text=True, timeout=timeout)


def read_csv(path: str):
    return [line.rstrip("\n") for line in open(path)]


def copy_tree(src: str, dst: str) -> None:
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst)


def last_exception(stderr: str):
    """Name of the exception a script died with (last 'Xxx: message' / 'pkg.Xxx' line of the traceback), or None."""
    import re
    names = re.findall(r"^([A-Za-z_][\w.]*(?:Error|Exception|Fail|Converge|Interrupt))\b", stderr, flags=re.M)
    return names[-1].split(".")[-1] if names else None


def iterations_done(stderr: str):
    """Last 'n/200' the embedded trainer's progress bar printed (tqdm writes to stderr), or None."""
    import re
    hits = re.findall(r"Training Iteration:\s+\d+%\|[^|]*\|\s*(\d+)/(\d+)", stderr)
    return int(hits[-1][0]) if hits else None
