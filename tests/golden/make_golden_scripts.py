"""Generate tests/golden/golden_scripts.json + golden_scripts.npz: what the UNMODIFIED reference prints, writes and trains
when its own driver scripts run on the synthetic corpus of tests/ref_scripts_harness.py.

    python tests/golden/make_golden_scripts.py            (authoring container: needs /root/reference; ~2 min of CPU)
    python tests/golden/make_golden_scripts.py --extra    (the seven other non-interactive drivers -> golden_scripts_extra.*)

The three scripts are executed as files, unmodified, with PYTHONPATH = the reference's ``src`` + stub modules for its
absent third-party imports (``librosa`` forwards to the restated oracle front end, see the harness).  The GPU test
tests/test_reference_scripts.py runs the same files with PYTHONPATH = the drop-in package and compares.
"""
import json
import os
import pickle
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import ref_scripts_harness as H  # noqa: E402

REF_SRC = "/root/reference/src"


def model_arrays(folder, prefix, out):
    """means / covariances / dense log-transitions of every model folder under ``folder`` (read in a subprocess that has
    the reference package on its path: the pickles embed its classes)."""
    code = ("import os, sys, pickle, numpy as np\n"
            "from loe_speech_recognition import HiddenMarkovModel\n"
            "res = {}\n"
            f"for name in sorted(os.listdir({folder!r})):\n"
            f"    m = HiddenMarkovModel.from_folder(os.path.join({folder!r}, name))\n"
            "    S = len(m._multivariate_normals)\n"
            "    res[name] = (np.stack([np.asarray(mn._core.mean) for mn in m._multivariate_normals]),\n"
            "                 np.stack([np.asarray(mn._core.cov_object.covariance) for mn in m._multivariate_normals]),\n"
            "                 np.array([[m._log_transition_probs[(i, j)] for j in range(S)] for i in range(S)], dtype=np.float32))\n"
            "pickle.dump(res, sys.stdout.buffer)\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([REF_SRC, os.path.join(os.path.dirname(folder.rstrip('/')), "..", "_stubs")]))
    raw = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, check=True).stdout
    for name, (mean, cov, logA) in pickle.loads(raw).items():
        out[f"{prefix}_means_{name}"] = mean
        out[f"{prefix}_covs_{name}"] = cov
        out[f"{prefix}_logA_{name}"] = logA


def main():
    assert os.path.isdir(REF_SRC), "needs the reference checkout"
    with tempfile.TemporaryDirectory() as ws:
        made = H.build_corpus(ws)
        stubs = H.write_stubs(os.path.join(ws, "_stubs"))
        pp = [REF_SRC, stubs, os.path.join(ROOT, "tests"), ROOT]
        seed = ("import sys, numpy as np, ref_scripts_harness as H; "
                f"g = np.load({os.path.join(HERE, 'golden_hmm.npz')!r}); "
                "H.write_seed_models('.cache/big_model_speech_only_3', g); H.write_seed_models('.cache/big_model_speech_only', g)")
        subprocess.run([sys.executable, "-c", seed], cwd=ws, env=dict(os.environ, PYTHONPATH=os.pathsep.join(pp)), check=True)
        record = {"corpus": made, "stdout": {}, "csv": {}}
        for name in H.SCRIPTS:
            r = H.run_script(name, ws, pp, timeout=3000)
            # project6_train.py may legitimately die: on small synthetic corpora the reference's embedded trainer raises
            # HMMTrainMeanFail as soon as one state of one word collects no frame (hidden_markov_model.py:324-331); its
            # `finally` still saves the models.  The outcome is part of the golden record.
            assert r.returncode == 0 or name == "project6_train.py", (name, r.stderr[-2000:])
            record["stdout"][name] = [l for l in r.stdout.splitlines() if l.startswith("In total")]
            record.setdefault("returncode", {})[name] = r.returncode
            record.setdefault("exception", {})[name] = H.last_exception(r.stderr)
            record.setdefault("iterations", {})[name] = H.iterations_done(r.stderr)
            print(name, r.returncode, record["exception"][name], record["iterations"][name], record["stdout"][name])
        for f in sorted(os.listdir(os.path.join(ws, "plots"))):
            record["csv"][f] = H.read_csv(os.path.join(ws, "plots", f))
        arrays = {}
        model_arrays(os.path.join(ws, ".cache", "big_model_no_silence"), "p3", arrays)
        model_arrays(os.path.join(ws, ".cache", "big_model_speech_only_continuous_2"), "p6", arrays)
        json.dump(record, open(os.path.join(HERE, "golden_scripts.json"), "w"), indent=1, sort_keys=True)
        np.savez_compressed(os.path.join(HERE, "golden_scripts.npz"), **arrays)
        print("wrote golden_scripts.json / .npz:", len(record["csv"]), "csv files,", len(arrays), "arrays")


def main_extra():
    """The other non-interactive drivers (ref_scripts_harness.EXTRA_SCRIPTS) against the real reference: result lines,
    CSV tables, the digit pairs project4_2digits.py drew and what it predicted for them, and the models
    project5_train_no_empty.py trains (silence stripper -> MFCC -> segmental K-means, 11 digits + the silence model)."""
    assert os.path.isdir(REF_SRC), "needs the reference checkout"
    with tempfile.TemporaryDirectory() as ws:
        made = H.build_corpus(ws)
        stubs = H.write_stubs(os.path.join(ws, "_stubs"))
        env_stubs = H.write_env_stubs(os.path.join(ws, "_env"))
        pp = [REF_SRC, env_stubs, stubs, os.path.join(ROOT, "tests"), ROOT]
        seed = ("import sys, numpy as np, ref_scripts_harness as H; "
                f"g = np.load({os.path.join(HERE, 'golden_hmm.npz')!r}); H.seed_all_models(g)")
        subprocess.run([sys.executable, "-c", seed], cwd=ws, env=dict(os.environ, PYTHONPATH=os.pathsep.join(pp)), check=True)
        ws2 = H.second_workspace(ws)
        record = {"corpus": made, "stdout": {}, "csv": {}, "returncode": {}, "exception": {}, "logged": {}}
        for name in H.EXTRA_SCRIPTS:
            cwd = ws2 if name == "project5_train_no_empty.py" else ws
            before = set(os.listdir(os.path.join(cwd, "plots")))
            start = H.log_size(cwd)
            r = H.run_script(name, cwd, pp, timeout=3000)
            record["stdout"][name] = H.stdout_record(r.stdout)
            record["returncode"][name] = r.returncode
            record["exception"][name] = H.last_exception(r.stderr)
            record["logged"][name] = H.logged_predictions(cwd, start)
            record["csv"][name] = {f: H.read_csv(os.path.join(cwd, "plots", f))
                                   for f in sorted(set(os.listdir(os.path.join(cwd, "plots"))) - before) if f.endswith(".csv")}
            print(name, r.returncode, record["exception"][name], record["stdout"][name][:4], len(record["logged"][name]),
                  sorted(record["csv"][name]))
            if r.returncode != 0:
                print(r.stderr[-1500:])
        arrays = {}
        trained = os.path.join(ws2, ".cache", "big_model_speech_only")
        if os.path.isdir(trained):
            # model_arrays() finds the stub modules relative to <folder>/../../_stubs
            os.symlink(stubs, os.path.join(ws2, "_stubs"))
            model_arrays(trained, "p5t", arrays)
        json.dump(record, open(os.path.join(HERE, "golden_scripts_extra.json"), "w"), indent=1, sort_keys=True)
        np.savez_compressed(os.path.join(HERE, "golden_scripts_extra.npz"), **arrays)
        print("wrote golden_scripts_extra.json / .npz:", len(arrays), "arrays")


if __name__ == "__main__":
    main_extra() if "--extra" in sys.argv[1:] else main()
