"""CPU: the C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol the
public header declares (no compute calls here)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols(built_lib):
    from loe_speech_recognition import _native
    header = open(os.path.join(ROOT, "include", "loe_b200.h")).read()
    declared = set(re.findall(r"\b(loe_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    lib = _native.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.loe_abi_version() == 1
    out = subprocess.check_output(["nm", "-D", "--defined-only", built_lib], text=True)
    exported = set(re.findall(r" T (loe_[a-z0-9_]+)", out))
    assert declared <= exported


def test_sass_is_sm100a(built_lib):
    out = subprocess.check_output(["cuobjdump", "-lelf", built_lib], text=True)
    assert "sm_100a" in out


def test_host_side_argument_checks_need_no_gpu(built_lib):
    from loe_speech_recognition import _native
    import pytest
    lib = _native.load()
    # argument validation happens before any CUDA call
    st = lib.loe_mfcc_dev(0, 0, 0, 0, 1, 5, 5, 5, 0, 0, 11, 5, 0, 0, 0, 0)
    assert st == _native.LOE_ERR_VALUE
    with pytest.raises(ValueError):
        _native.check(st)
    st = lib.loe_viterbi_dev(0, 58, 0, 1, 10, 0, 0, 0, 0, 200, 0, 1, -100.0, 0, 0, 0, 12, 0, 0, 0, 0, 0, -1, 0, 0, 0, 0)
    assert st == _native.LOE_ERR_OVERFLOW
    with pytest.raises(OverflowError):
        _native.check(st)
    assert lib.loe_emission_dev(0, 10, 7, 0, 0, 0, 3, 0, 3, 0, 0) == _native.LOE_ERR_UNSUPPORTED
    assert lib.loe_viterbi_bp_fits(460, 58) == 1 and lib.loe_viterbi_bp_fits(100000, 128) == 0
    assert lib.loe_kmeans_ws_doubles(1000, 5, 39) == 5 * 820 + 1
