"""ctypes binding of libloe_b200.so (include/loe_b200.h).

This is the "reference-side binding" of the C ABI: the reference is Python, so the stub a
maintainer would add is a ctypes loader.  The library is built in-tree by
``cs-304-speech-recognition-code_b200/build.py`` (nvcc, sm_100a).  There is no CPU fallback:
if the library is missing, :func:`load` raises and every compute entry point of the package
fails loudly.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libloe_b200.so")

LOE_OK, LOE_ERR_CUDA, LOE_ERR_VALUE, LOE_ERR_OVERFLOW, LOE_ERR_DIM, LOE_ERR_UNSUPPORTED = range(6)
LOE_MAX_POS = 128
LOE_MEL_NA_MAX = 32
LOE_MEL_NB_MAX = 16
POS_INIT, POS_START, POS_END = 1, 2, 4
LOE_MSTEP_UPDATED, LOE_MSTEP_CONVERGED, LOE_MSTEP_MEAN_FAIL, LOE_MSTEP_SUSPECT = 1, 2, 4, 8

# every symbol include/loe_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "loe_abi_version": (c_int, []),
    "loe_last_error": (c_char_p, []),
    "loe_device_count": (c_int, []),
    "loe_mfcc_dev": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_int,
                             c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "loe_mfcc_phase_dev": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_int,
                                   c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "loe_mfcc_img_dev": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_int,
                                 c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "loe_mfcc_ex_dev": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "loe_emission_dev": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                 c_void_p, c_int, c_int, c_void_p]),
    "loe_emission_tc_tiles": (c_int, [c_int]),
    "loe_emission_tc_dev": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "loe_emission_h16_tile_bytes": (c_int, []),
    "loe_emission_h16_dev": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "loe_emission_h16_img_bytes": (c_int64, [c_int64]),
    "loe_emission_h16_img_dev": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "loe_emission_h16_multi_dev": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "loe_h16_image_dev": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "loe_emission_h16_multi_img_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "loe_emission_gmm_tile_bytes": (c_int, []),
    "loe_emission_gmm_tiles": (c_int, [c_int, c_int]),
    "loe_emission_gmm_dev": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "loe_emission_gmm_tc_dev": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                        c_void_p, c_int, c_void_p]),
    "loe_viterbi_bp_fits": (c_int, [c_int, c_int]),
    "loe_viterbi_dev": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                c_int, c_double, c_int,
                                c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "loe_labels_dev": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                               c_void_p, c_int, c_void_p, c_void_p]),
    "loe_silence_dev": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_double, c_double, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "loe_dtw_dev": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int,
                            c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "loe_align_dev": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "loe_kmeans_ws_doubles": (c_int64, [c_int64, c_int, c_int]),
    "loe_kmeans_dev": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "loe_mstep_dev": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "loe_decoder_create": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int,
                                   c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "loe_decoder_decode_host": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_double, c_int, c_int, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "loe_decoder_set_h16": (c_int, [c_void_p, c_void_p]),
    "loe_decoder_destroy": (None, [c_void_p]),
    "loe_pcm_narrow_host": (c_int, [c_void_p, c_void_p, c_int64]),
    "loe_decoder_narrow_rate": (c_double, [c_void_p]),
    "loe_decoder_set_narrow": (c_int, [c_void_p, c_int]),
    "loe_decoder_stats": (c_int, [c_void_p, c_void_p, c_int]),
    "loe_host_alloc": (c_int, [c_void_p, ctypes.c_size_t]),
    "loe_host_free": (c_int, [c_void_p]),
    "loe_host_fingerprint": (ctypes.c_uint64, [c_void_p, c_void_p, c_int]),
    "loe_labels_text_host": (c_int64, [c_void_p, c_void_p, c_int, c_int, c_char_p, c_int, ctypes.c_char, c_void_p]),
}



class MfccConfigStruct(ctypes.Structure):
    """loe_mfcc_config (include/loe_b200.h)."""
    _fields_ = [("n_fft", ctypes.c_int32), ("hop", ctypes.c_int32), ("n_mels", ctypes.c_int32), ("n_ceps", ctypes.c_int32),
                ("log_mode", ctypes.c_int32), ("norm_mode", ctypes.c_int32), ("preemph", ctypes.c_float), ("reserved", ctypes.c_float)]


LOG_MODES = {"db": 0, "ln": 1}
NORM_MODES = {"none": 0, "frame": 1, "cmn": 2, "cmvn": 3}

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the CUDA library (once per process).  Raises NativeLibraryMissing when it has not
    been built -- the package has no other compute path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: build it with `python cs-304-speech-recognition-code_b200/build.py` "
            "(or __graft_entry__.build()).  loe_speech_recognition (B200 build) has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.loe_abi_version() != 1:
        raise RuntimeError("libloe_b200.so ABI version mismatch")
    _lib = lib
    return lib


class _HMMTrainMeanFailPlaceholder(Exception):
    pass


def check(status: int) -> None:
    """Map a loe_status to the exception type the reference raises at the same API point."""
    if status == LOE_OK:
        return
    msg = (load().loe_last_error() or b"").decode("utf-8", "replace")
    if status == LOE_ERR_VALUE:
        raise ValueError(msg)
    if status == LOE_ERR_OVERFLOW:
        raise OverflowError(msg)
    if status == LOE_ERR_DIM:
        raise AssertionError(msg)
    if status == LOE_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(f"loe_b200: {msg}")
