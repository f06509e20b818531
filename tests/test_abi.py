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
    assert lib.loe_mfcc_dev(0, 0, 0, 0, 1, 20, 20, 20, 0, 0, 11, 5, 4, 0, 0, 0) == _native.LOE_ERR_VALUE    # mel workspace not 16-byte aligned
    assert lib.loe_emission_dev(0, 10, 7, 0, 0, 0, 3, 0, 3, 0, 0) == _native.LOE_ERR_UNSUPPORTED
    assert lib.loe_viterbi_bp_fits(460, 58) == 1 and lib.loe_viterbi_bp_fits(100000, 128) == 0
    # sorted path: (1 work item per 2048 list entries + one per bucket) partials of 820 doubles + the integer tables
    assert lib.loe_kmeans_ws_doubles(1000, 5, 39) == 6 * 820 + (1 * 5 + 5 + 5 + 3 * 6 + 6 + 1000 + 1) // 2
    assert lib.loe_kmeans_ws_doubles(1000, 5, 13) == 5 * (1 + 13 + 91) + 1              # other dimensions: the scanning path


def test_pcm_narrowing_is_exact_or_refused(built_lib):
    """loe_pcm_narrow_host (host code, no GPU): float32 PCM that holds int16 values converts exactly at every
    length / alignment; a single fractional, out-of-range or non-finite sample makes the call report failure."""
    import numpy as np
    from loe_speech_recognition import _native
    lib = _native.load()
    rng = np.random.default_rng(5)
    for n in (0, 1, 7, 8, 15, 16, 17, 31, 33, 1000, 100003):
        a = rng.integers(-32768, 32768, n).astype(np.float32)
        for shift in (0, 1, 16):
            out = np.full(n + 32, 77, np.int16)
            assert lib.loe_pcm_narrow_host(a.ctypes.data, out.ctypes.data + 2 * shift, n) == 1
            assert np.array_equal(out[shift:shift + n], a.astype(np.int16))
            assert np.all(out[:shift] == 77) and np.all(out[shift + n:] == 77)
        if n:
            out = np.empty(n, np.int16)
            for bad in (0.5, -0.25, 32768.0, -32769.0, np.nan, np.inf, -np.inf, 1e20):
                b = a.copy()
                b[rng.integers(0, n)] = bad
                assert lib.loe_pcm_narrow_host(b.ctypes.data, out.ctypes.data, n) == 0, (n, bad)
    edge = np.array([-32768.0, 32767.0, 0.0, -0.0, 1.0, -1.0], np.float32)
    out = np.empty(edge.size, np.int16)
    assert lib.loe_pcm_narrow_host(edge.ctypes.data, out.ctypes.data, edge.size) == 1
    assert out.tolist() == [-32768, 32767, 0, 0, 1, -1]


def test_labels_text_host(built_lib):
    """loe_labels_text_host (host code): word-id tables -> newline-separated label text, exactly "".join(labels[k] ...) per
    utterance; counts outside the table (T == 1: -1, overflow: > max_words) leave an empty line for the caller's slow path."""
    import numpy as np
    from loe_speech_recognition import _native
    lib = _native.load()
    labels = "123456789OSZ"
    rng = np.random.default_rng(11)
    for n, mw in ((1, 1), (7, 4), (1000, 32)):
        words = rng.integers(0, len(labels), (n, mw)).astype(np.int8)
        count = rng.integers(-1, mw + 2, n).astype(np.int32)
        buf = np.full(n * (mw + 1) + 8, 0x7F, dtype=np.uint8)
        nb = lib.loe_labels_text_host(words.ctypes.data, count.ctypes.data, n, mw, labels.encode(), len(labels), b"\n", buf.ctypes.data)
        got = buf[:nb].tobytes().decode().split("\n")
        assert got[-1] == "" and len(got) == n + 1 and np.all(buf[nb:] == 0x7F)
        want = ["".join(labels[k] for k in words[i, :c]) if 0 <= c <= mw else "" for i, c in enumerate(count.tolist())]
        assert got[:n] == want
    assert lib.loe_labels_text_host(0, 0, 0, 32, labels.encode(), len(labels), b"\n", 0) == 0
