"""loe_speech_recognition -- B200-native drop-in for the hot path of
loeeeee/CS-304-Speech-Recognition-Code: MFCC front end, Gaussian emission scoring, Viterbi
decoding (isolated words and the digit-loop grammar) and segmental K-means training.

Same class names and signatures as the reference package (src/loe_speech_recognition/
__init__.py:11-30) for everything on that path; the computation runs in hand-written sm_100a
CUDA kernels behind a C ABI (include/loe_b200.h).  There is no CPU fallback.

Out of scope for this build: the TIDIGITS corpus walker, the
energy-based silence stripper, the live microphone segmenter, template DTW, CSV and plot
helpers.  Accessing those names raises NotImplementedError with this explanation.
"""
from .mfcc import MFCC
from .ti_digits import TI_DIGITS_LABELS, TI_DIGITS_LABEL_TYPE
from .hidden_markov_model import (Signal, HiddenMarkovModel, HiddenMarkovModelTrainable, HiddenMarkovModelInference,
                                  HiddenMarkovModelTrainContinuous)
from .model_collection import ModelCollection

__all__ = [
    "MFCC",
    "TI_DIGITS_LABELS",
    "TI_DIGITS_LABEL_TYPE",
    "HiddenMarkovModel",
    "HiddenMarkovModelTrainable",
    "HiddenMarkovModelInference",
    "HiddenMarkovModelTrainContinuous",
    "Signal",
    "ModelCollection",
]

_OUT_OF_SCOPE = {"Segmentation", "DynamicTimeWarping", "TIDigits", "DataLoader", "plot_confusion_matrix_from_lists",
                 "plot_line", "CSVReader", "CSVWriter", "SignalSeparation"}


def __getattr__(name):
    if name in _OUT_OF_SCOPE:
        raise NotImplementedError(f"loe_speech_recognition.{name} is host-side I/O / tooling outside the accelerated "
                                  "hot path and is not part of the B200 build (see DESIGN.md, 'Out of scope')")
    raise AttributeError(name)
