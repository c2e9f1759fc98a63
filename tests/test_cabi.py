"""CPU: the C-ABI library loads and exports every symbol include/tehmm_b200.h
declares; without a GPU the product fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tehmm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tehmm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from tehmm_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libtehmm_b200.so does not export %s" % n
    # and the Python binding types exactly that set
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_error_string():
    from tehmm_b200 import _lib
    lib = _lib.load()
    assert lib.tehmm_abi_version() == 1
    assert isinstance(lib.tehmm_last_error(), bytes)


def test_no_silent_cpu_fallback():
    """On a box without a GPU every compute entry point must raise."""
    from tehmm_b200 import _lib
    lib = _lib.load()
    if lib.tehmm_device_count() > 0:
        pytest.skip("a GPU is visible")
    h = ctypes.c_void_p()
    rc = lib.tehmm_ctx_create(0, ctypes.byref(h))
    assert rc == _lib.TEHMM_ECUDA
    assert b"no CPU fallback" in lib.tehmm_last_error()
    import numpy as np
    from tehmm_b200 import _hmm
    with pytest.raises(_lib.TehmmError):
        _hmm._viterbi(1, 1, np.zeros(1), np.zeros((1, 1)), None, np.zeros((1, 1)))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tehmm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "tehmm_oracle" not in text and "ref_loader" not in text, f
