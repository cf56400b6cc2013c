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


def test_header_is_plain_c_and_links_against_the_library(lib_path, tmp_path):
    """include/b200_ops.h must be consumable by a C compiler (no C++ / torch / CUDA headers): a cgo / JNI / N-API binding
    would include it exactly like this.  The C program takes the address of every declared entry point (link check) and calls
    the GPU-independent ones."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    names = _declared()
    table = ", ".join(f"(fn_t){n}" for n in names)
    src = tmp_path / "abi_check.c"
    src.write_text('#include "b200_ops.h"\n#include <stdio.h>\ntypedef void (*fn_t)(void);\n'
                   f'static fn_t table[] = {{{table}}};\n'
                   'int main(void) {\n  unsigned i;\n'
                   '  for (i = 0; i < sizeof(table) / sizeof(table[0]); ++i) if (!table[i]) return 100;\n'
                   '  if (b200_abi_version() != 1) return 1;\n'
                   '  if (b200_gemm_bf16(0, 0, 0, 0, 8, 8, 8, 8, 8, 0, 0, 0, 0) != -22) return 2;\n'
                   '  printf("%s\\n", b200_last_error());\n  return 0;\n}\n')
    exe = tmp_path / "abi_check"
    libdir = os.path.dirname(lib_path)
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-lb200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "empty" in r.stdout, (r.returncode, r.stdout, r.stderr)
