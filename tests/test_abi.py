"""The C-ABI library loads on a CPU-only box, exports every symbol that
include/ecgbyte.h declares, and fails loudly (no CPU fallback) when asked to compute."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ecgbyte.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ecgb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ecgbyte import _lib
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libecgbyte.so does not export %s" % n
    # and the Python binding declares a signature for each of them
    assert set(names) == set(_lib.EXPORTS)
    assert L.ecgb_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point reports ECGB_ENODEVICE."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from ecgbyte import _lib
    L = _lib.lib()
    h = C.c_void_p()
    assert L.ecgb_quantizer_create(0.0, 1.0, 0, 1e-3, 0, C.byref(h)) == _lib.ENODEVICE
    assert b"no CPU fallback" in L.ecgb_last_error()
    assert L.ecgb_trainer_create(0, 1024, 10, 0, C.byref(h)) == _lib.ENODEVICE
    import numpy as np
    seq = np.zeros(1, np.uint32)
    off = np.zeros(1, np.uint64)
    assert L.ecgb_vocab_create(seq.ctypes.data, off.ctypes.data, seq.ctypes.data, 0, 0, C.byref(h)) == _lib.ENODEVICE
    import rust_bpe
    with pytest.raises(Exception):
        rust_bpe.encode_text("abc", [])
    with pytest.raises(Exception):
        rust_bpe.byte_pair_encoding("abcabc", 2, 1)


def test_expand_merges_host_utility():
    import numpy as np
    from ecgbyte import _lib
    L = _lib.lib()
    pairs = np.array([[97, 98], [256, 99], [257, 256]], np.uint32)
    off = np.zeros(4, np.uint64)
    seq = np.zeros(16, np.uint32)
    assert L.ecgb_expand_merges(pairs.ctypes.data, 3, seq.ctypes.data, 16, off.ctypes.data) == 0
    assert off.tolist() == [0, 2, 5, 10]
    assert seq[:10].tolist() == [97, 98, 97, 98, 99, 97, 98, 99, 97, 98]
    assert L.ecgb_expand_merges(pairs.ctypes.data, 3, seq.ctypes.data, 4, off.ctypes.data) == _lib.ECAPACITY
    bad = np.array([[97, 300]], np.uint32)
    assert L.ecgb_expand_merges(bad.ctypes.data, 1, seq.ctypes.data, 16, off.ctypes.data) == _lib.EINVAL


def test_product_never_imports_oracle():
    """The package must not reference the oracle (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "ecg-byte_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(d, f)).read()
                assert "libecgb_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f
