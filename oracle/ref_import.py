"""Import the UNMODIFIED reference package from /root/reference (authoring container only).

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/make_golden.py`` (to generate the
committed fixtures) and by the ``test_oracle_vs_reference_*`` tests, which skip when
``/root/reference`` does not exist (e.g. on the GPU box).

The reference imports ``librosa``, ``sounddevice``, ``uniplot`` and ``matplotlib`` at
module level although the hot path (hidden_markov_model.py, signal.py, ...) never
calls them; those are absent here, so permissive stub modules are registered first.
``segmentation.py:85`` dereferences ``sd.InputStream`` at class-creation time, hence the
module-level ``__getattr__``.

The reference is imported under its own name ``loe_speech_recognition`` (its pickles
embed that module path), so a process that calls :func:`import_reference` must not
also import the drop-in package of the same name from this repo.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_SRC = "/root/reference/src"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "loe_speech_recognition"))


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _Anything:
    def __call__(self, *a, **k):
        raise RuntimeError("stubbed third-party dependency called on the oracle path")

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


def import_reference():
    """Return the real ``loe_speech_recognition`` package object."""
    if not reference_available():
        raise RuntimeError("/root/reference is not present on this machine")
    mod = sys.modules.get("loe_speech_recognition")
    if mod is not None:
        if not os.path.abspath(mod.__file__).startswith(REFERENCE_SRC):
            raise RuntimeError("the drop-in loe_speech_recognition is already imported in this process")
        return mod
    for name in ("librosa", "sounddevice", "uniplot", "matplotlib", "matplotlib.pyplot", "soundfile"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    sys.path.insert(0, REFERENCE_SRC)
    try:
        mod = importlib.import_module("loe_speech_recognition")
    finally:
        sys.path.remove(REFERENCE_SRC)
    assert os.path.abspath(mod.__file__).startswith(REFERENCE_SRC)
    return mod
