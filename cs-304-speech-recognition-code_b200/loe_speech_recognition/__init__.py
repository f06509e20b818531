"""loe_speech_recognition -- B200-native drop-in for the hot path of
loeeeee/CS-304-Speech-Recognition-Code: MFCC front end, Gaussian emission scoring, Viterbi
decoding (isolated words and the digit-loop grammar) and segmental K-means training.

Same class names and signatures as the reference package (src/loe_speech_recognition/
__init__.py:11-30) for everything on that path; the computation runs in hand-written sm_100a
CUDA kernels behind a C ABI (include/loe_b200.h).  There is no CPU fallback.

Host-side helpers around the path (corpus walker, '|'-separated result tables, result plots; SURVEY.md §8
f1 / f4) are plain Python mirrors of the reference's modules.  Out of scope: the live microphone
segmenter (``Segmentation``, needs sounddevice and audio hardware); accessing it raises NotImplementedError
unless LOE_REFERENCE_SRC points at a reference checkout.
"""
from .mfcc import MFCC, MFCCConfig
from .ti_digits import TIDigits, DataLoader, TI_DIGITS_LABELS, TI_DIGITS_LABEL_TYPE
from .hidden_markov_model import (Signal, HiddenMarkovModel, HiddenMarkovModelTrainable, HiddenMarkovModelInference,
                                  HiddenMarkovModelTrainContinuous)
from .model_collection import ModelCollection
from .signal_separation import SignalSeparation
from .dynamic_time_wrapping import DynamicTimeWarping
from .visualizer import plot_confusion_matrix_from_lists, plot_line
from .csvnia import CSVReader, CSVWriter

__all__ = [
    "MFCC",
    "MFCCConfig",
    "TIDigits",
    "DataLoader",
    "TI_DIGITS_LABELS",
    "TI_DIGITS_LABEL_TYPE",
    "plot_confusion_matrix_from_lists",
    "plot_line",
    "CSVReader",
    "CSVWriter",
    "HiddenMarkovModel",
    "HiddenMarkovModelTrainable",
    "HiddenMarkovModelInference",
    "HiddenMarkovModelTrainContinuous",
    "Signal",
    "ModelCollection",
    "SignalSeparation",
    "DynamicTimeWarping",
]

# name -> module of the reference that defines it (host-side I/O and tooling, not rebuilt here)
_OUT_OF_SCOPE = {"Segmentation": "segmentation"}


def __getattr__(name):
    """Names outside the accelerated path.  When LOE_REFERENCE_SRC points at the reference's
    ``src/loe_speech_recognition`` directory, the original host-side module is loaded from there (so
    the reference's scripts run unmodified with the hot path replaced); otherwise NotImplementedError."""
    if name in _OUT_OF_SCOPE:
        import importlib.util
        import os
        import sys
        src = os.environ.get("LOE_REFERENCE_SRC")
        mod_name = _OUT_OF_SCOPE[name]
        path = os.path.join(src, mod_name + ".py") if src else None
        if path and os.path.exists(path):
            full = f"{__name__}.{mod_name}"
            mod = sys.modules.get(full)
            if mod is None:
                spec = importlib.util.spec_from_file_location(full, path)
                mod = importlib.util.module_from_spec(spec)
                sys.modules[full] = mod
                spec.loader.exec_module(mod)
            return getattr(mod, name)
        raise NotImplementedError(f"loe_speech_recognition.{name} is host-side I/O / tooling outside the accelerated "
                                  "hot path and is not part of the B200 build (see DESIGN.md, 'Out of scope'); set "
                                  "LOE_REFERENCE_SRC=<reference>/src/loe_speech_recognition to load the original module")
    raise AttributeError(name)
