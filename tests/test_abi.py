"""The C-ABI library builds, loads on a GPU-less box and exports every symbol include/b200_ops.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from transformers_b200 import build

    return build.build()


def _declared():
    src = open(os.path.join(ROOT, "include", "b200_ops.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200_ops.h but not exported by libb200.so"


def test_python_binding_matches_header(lib_path):
    from transformers_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.load()
    assert lib.b200_abi_version() == 1


def test_fails_loudly_without_gpu(lib_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from transformers_b200 import _lib, ops

    assert _lib.load().b200_device_check() != 0
    with pytest.raises(_lib.B200Error):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    # shape validation happens before any launch: no GPU needed, error codes are errno-style
    rc = _lib.load().b200_gemm_bf16(None, None, None, 0, 8, 8, 8, 8, 8, 0, 0, 0, None)
    assert rc == -22 and "empty" in _lib.last_error()
