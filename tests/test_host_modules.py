"""CPU: the host-side mirrors around the hot path: corpus walker, '|'-separated tables,
confusion counts.  Behaviour is pinned against the UNMODIFIED reference classes when /root/reference exists
(subprocess: the reference package has the same import name as the drop-in)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/src"
HAVE_REF = os.path.isdir(os.path.join(REF_SRC, "loe_speech_recognition"))


def _make_corpus(root):
    from scipy.io import wavfile
    rng = np.random.default_rng(0)
    files = {("Adults", "TRAIN", "MAN/AE"): ["1A.WAV", "1B.WAV", "25A.WAV", "Z9OA.WAV"],
             ("Adults", "TEST", "WOMAN/BC"): ["OA.WAV", "1A.WAV"],
             ("Children", "TRAIN", "BOY/XY"): ["1A.wav", "7B.wav", "notes.txt"],
             ("Children", "TEST", "GIRL/ZZ"): ["25B.WAV"]}
    for (grp, split, sub), names in files.items():
        d = os.path.join(root, grp, "TIDIGITS", split, sub)
        os.makedirs(d)
        for n in names:
            if n.endswith(".txt"):
                open(os.path.join(d, n), "w").write("x")
            else:
                wavfile.write(os.path.join(d, n), 16000, rng.integers(-20000, 20000, rng.integers(2000, 4000)).astype(np.int16))


def _summary(loader):
    return {k: sorted((int(len(v)), float(np.asarray(v, dtype=np.float64).sum())) for v in loader[k]) for k in sorted(loader.data)}


def test_tidigits_walker(tmp_path):
    from loe_speech_recognition import DataLoader, TIDigits
    root = str(tmp_path / "corpus")
    _make_corpus(root)
    ds = TIDigits(root)
    assert sorted(ds.train_dataset.data) == ["1", "25", "7", "Z9O"] and sorted(ds.test_dataset.data) == ["1", "25", "O"]
    assert len(ds.train_dataset["1"]) == 3 and len(ds.train_dataset) == 4
    clip = ds.train_dataset["7"][0]
    assert clip.dtype == np.float32 and clip.ndim == 1 and np.all(clip == np.round(clip))
    assert all(isinstance(p, str) for p in ds.train_dataset.data["1"])                    # lazy: paths until asked
    eager = TIDigits(root, isLazyLoading=False)
    assert all(isinstance(a, np.ndarray) for a in eager.train_dataset.data["1"])
    assert _summary(eager.train_dataset) == _summary(ds.train_dataset)
    assert sorted(TIDigits(root, include_children=False).train_dataset.data) == ["1", "25", "Z9O"]
    assert sorted(TIDigits(root, include_adult=False).test_dataset.data) == ["25"]
    with pytest.raises(Exception):
        TIDigits(root, include_adult=False, include_children=False)
    pairs = list(ds.test_dataset)
    assert len(pairs) == 3 and {lab for _, lab in pairs} == {"1", "25", "O"}
    two = ds.train_dataset.get_all_n_digits(2)
    assert list(two) == ["25"] and len(two["25"]) == 1
    comb = ds.train_dataset.get_combined("17")
    assert comb.shape[0] == ds.train_dataset["1"][0].shape[0] + ds.train_dataset["7"][0].shape[0]
    assert DataLoader.filename_parser("Z9OA.WAV") == "Z9O" and DataLoader.filename_parser("1b.x.wav") == "1"
    assert len(DataLoader.from_folder_path(str(tmp_path / "missing"))) == 0                # os.walk is silent
    with pytest.raises(NotImplementedError):
        DataLoader.lazy_loading(3)
    # the WAV samples can stay int16 all the way to the device
    try:
        DataLoader.sample_dtype = np.int16
        narrow = ds.train_dataset["7"][0]
        assert narrow.dtype == np.int16 and np.array_equal(narrow.astype(np.float32), clip)
    finally:
        DataLoader.sample_dtype = np.float32


def test_csv_round_trip_and_quirks(tmp_path):
    from loe_speech_recognition import CSVReader, CSVWriter
    w = CSVWriter(["truth", "pred", "n"])
    w.add_line(["123", "12O", 3])
    w.add_line(['say "hi"', None, 0])
    w.add_line(["", 2.5, -4])
    path = str(tmp_path / "t.csv")
    w.write(path)
    assert open(path, encoding="utf-8").read() == '"truth"|"pred"|"n"\n"123"|"12O"|3\n"say ""hi"""|None|0\n""|2.5|-4\n'
    assert str(w) == "Columns: truth, pred, n Size: 3" and len(w) == 3
    r = CSVReader(path)
    assert r.columns == ["truth", "pred", "n"] and len(r) == 3
    rows = list(r)
    assert rows[0] == {"truth": "123", "pred": "12O", "n": 3}                 # quoted digits stay strings
    assert rows[1] == {"truth": 'say "hi"', "pred": None, "n": 0}
    assert rows[2] == {"truth": "", "pred": "2.5", "n": "-4"}                 # only all-digit entries become ints
    assert CSVReader.line_parser('"a"|None|12|x') == ["a", None, 12, "x"]
    with pytest.raises(IndexError):
        CSVReader.line_parser("1||2")                                         # empty cell: the reference indexes entry[0]


def test_confusion_counts_and_plot_guards():
    from loe_speech_recognition import plot_line
    from loe_speech_recognition.visualizer import confusion_counts
    c = confusion_counts(["1", "2", "2", "O"], ["1", "2", "1", "O"], ["1", "2", "O"])
    assert c.tolist() == [[1, 1, 0], [0, 1, 0], [0, 0, 1]]                    # rows = truth, columns = prediction
    with pytest.raises(ValueError):
        confusion_counts(["9"], ["1"], ["1", "2"])
    with pytest.raises(ValueError):
        plot_line([1, 2], [1])


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not present (GPU box)")
def test_host_modules_match_reference(tmp_path):
    """Same corpus tree / same table through the reference's own classes (separate process) and through the
    mirrors: identical label maps, clip lengths and sums, file bytes and parsed rows."""
    root = str(tmp_path / "corpus")
    _make_corpus(root)
    csv_ref, csv_new = str(tmp_path / "ref.csv"), str(tmp_path / "new.csv")
    body = '''
import json, sys, numpy as np
{prelude}
from loe_speech_recognition.ti_digits import TIDigits
from loe_speech_recognition.csvnia import CSVReader, CSVWriter
ds = TIDigits({root!r})
def summary(loader):
    return {{k: sorted((int(len(v)), float(np.asarray(v, dtype=np.float64).sum())) for v in loader[k]) for k in sorted(loader.data)}}
w = CSVWriter(["truth", "pred", "n"])
for line in (["123", "12O", 3], ['say "hi"', None, 0], ["x", 2.5, -4]):
    w.add_line(line)
w.write({out!r})
rows = [dict(r) for r in CSVReader({out!r})]
print(json.dumps({{"train": summary(ds.train_dataset), "test": summary(ds.test_dataset), "rows": rows,
                  "combined": int(ds.train_dataset.get_combined("17").shape[0]),
                  "two": sorted(ds.train_dataset.get_all_n_digits(2)), "str": str(w)}}))
'''
    ref_prelude = ("import types\nfor m in ('librosa', 'sounddevice', 'uniplot', 'matplotlib', 'matplotlib.pyplot', 'soundfile'):\n"
                   "    mod = types.ModuleType(m); mod.__getattr__ = lambda name: object; sys.modules.setdefault(m, mod)\n"
                   f"sys.path.insert(0, {REF_SRC!r})")
    new_prelude = f"sys.path.insert(0, {os.path.join(ROOT, 'cs-304-speech-recognition-code_b200')!r})"
    outs = []
    for prelude, out in ((ref_prelude, csv_ref), (new_prelude, csv_new)):
        code = body.format(prelude=prelude, root=root, out=out)
        outs.append(json.loads(subprocess.check_output([sys.executable, "-c", code], text=True).strip().splitlines()[-1]))
    assert outs[0] == outs[1]
    assert open(csv_ref, "rb").read() == open(csv_new, "rb").read()
