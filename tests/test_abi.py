"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/pb200.h declares,
keeps the reference's error behaviour for invalid domains, and refuses to run without a GPU (no CPU
fallback).  No compute call is made here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pb200.h")).read()
    return sorted(set(re.findall(r"PB200_API[^;(]*?\b(pb200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import plonk_prototype_b200 as pb
    if not os.path.exists(pb.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(pb.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert sorted(pb.EXPORTS) == names


def test_domain_size_errors_like_evaluation_domain_new():
    import plonk_prototype_b200 as pb
    d = pb.EvaluationDomain.__new__(pb.EvaluationDomain)  # no context needed for the size rule
    out = ctypes.c_uint32()
    lib = pb._native.lib()
    assert lib.pb200_domain_log_size(1, ctypes.byref(out)) == 0 and out.value == 0
    assert lib.pb200_domain_log_size(5, ctypes.byref(out)) == 0 and out.value == 3
    assert lib.pb200_domain_log_size(1 << 31, ctypes.byref(out)) == 0 and out.value == 31
    assert lib.pb200_domain_log_size((1 << 31) + 1, ctypes.byref(out)) != 0   # log2(size) = 32 → Err
    with pytest.raises(pb.InvalidEvalDomainSize):
        pb.EvaluationDomain(1 << 33)


def test_no_cpu_fallback_without_gpu():
    import torch
    import plonk_prototype_b200 as pb
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pb.Pb200Error):
        pb.Context(0)


def test_product_never_imports_oracle():
    """The product path must not import, include, link or load oracle/ (only tests, smoke() and bench's
    cpu_baseline may).  Comments that mention the directory are fine; code references are not."""
    pkg = os.path.join(ROOT, "plonk-prototype_b200")
    bad = re.compile(r"import\s+pyoracle|from\s+oracle|import\s+oracle|liboracle|#include\s+[\"<][^\n]*oracle|import\s+model\b")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", "Makefile")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), os.path.join(dirpath, f)


def test_cxx_host_layer_example_builds_and_refuses_to_run_without_gpu():
    """include/pb200.hpp + examples/mock_circuit.cpp (the reference's valid_balance circuit through the C++ host layer):
    compiles and links against libpb200.so; without a GPU it must fail loudly, not fall back."""
    import subprocess
    import torch
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "plonk-prototype_b200", "csrc"), "-s", "example"])
    exe = os.path.join(ROOT, "examples", "mock_circuit")
    assert os.path.exists(exe)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present (the gpu-marked test runs it)")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr
