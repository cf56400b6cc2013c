"""Repository contract checks that need no device: the product path never touches the oracle, no environment variable selects
a kernel, the reference install travels to the GPU box, and nothing run on the GPU box reads /root/reference."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py_files(d):
    for base, _, files in os.walk(os.path.join(ROOT, d)):
        for f in files:
            if f.endswith(".py"):
                yield os.path.join(base, f)


def test_product_path_never_imports_the_oracle():
    for path in _py_files("transformers_b200"):
        src = open(path).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{path} imports the oracle (test infrastructure only)"
    bench = open(os.path.join(ROOT, "bench.py")).read()
    # bench.py may execute the oracle only inside its CPU-baseline legs
    uses = [m.start() for m in re.finditer(r"from oracle import", bench)]
    assert len(uses) == 1 and bench.rfind("def cpu_reference_sample", 0, uses[0]) > bench.rfind("\ndef ", 0, bench.rfind("def cpu_reference_sample", 0, uses[0]))


def test_no_environment_variable_selects_a_kernel():
    csrc = os.path.join(ROOT, "transformers_b200", "csrc")
    for f in os.listdir(csrc):
        if f.endswith((".cu", ".cuh")):
            assert "getenv" not in open(os.path.join(csrc, f)).read(), f"{f} reads the environment"
    for path in _py_files("transformers_b200"):
        if os.path.basename(path) == "build.py":  # NVCC= picks the compiler at build time, not a kernel at run time
            continue
        assert "os.environ" not in open(path).read(), f"{path} reads the environment"


def test_reference_install_and_built_libraries_travel_with_the_snapshot():
    ignore = os.path.join(ROOT, ".gpurunignore")
    listed = open(ignore).read() if os.path.exists(ignore) else ""
    for must_travel in ("baseline/_ref", "oracle/_ref", "transformers_b200/lib"):
        assert must_travel not in listed
    gi = open(os.path.join(ROOT, ".gitignore")).read()
    assert "baseline/_ref/" in gi and "*.so" in gi  # ... but stay out of history


def test_gpu_side_code_does_not_depend_on_the_reference_checkout():
    """/root/reference does not exist on the GPU box: only the import helper may mention it (and falls back to baseline/_ref)."""
    allowed = {os.path.join(ROOT, "baseline", "ref_import.py"), os.path.join(ROOT, "tests", "golden", "make_golden.py"),
               os.path.abspath(__file__)}
    for d in ("transformers_b200", "tests"):  # (the oracle's docstrings cite reference paths; it runs no file access)
        for path in _py_files(d):
            if path in allowed:
                continue
            assert "/root/reference" not in open(path).read(), path
    for f in ("bench.py", "__graft_entry__.py"):
        assert "/root/reference" not in open(os.path.join(ROOT, f)).read(), f
