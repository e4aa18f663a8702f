"""CPU-side checks of the drop-in boundary: libmsgpu.so builds, loads, and exports every symbol that
include/msgpu.h declares (no compute calls: there is no GPU on the build machine)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "msgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_symbols():
    syms = header_symbols()
    assert "msgpu_commit" in syms and "msgpu_dft_batch" in syms and len(syms) >= 30


def test_library_exports_every_header_symbol():
    import multi_stark_b200 as ms
    from multi_stark_b200 import _ffi
    L = ms.lib()
    for s in header_symbols():
        assert hasattr(L, s), "libmsgpu.so does not export %s" % s
        assert s in _ffi.SIGNATURES, "python binding misses %s" % s
    assert sorted(_ffi.SIGNATURES) == header_symbols()


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path must fail loudly, not fall back to CPU code."""
    import torch
    import multi_stark_b200 as ms
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ms.MsgpuError):
        ms.GpuContext(0)


def test_product_does_not_reference_oracle():
    """The oracle is test infrastructure: nothing under multi_stark_b200/ may include or load it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "multi_stark_b200")):
        if "build" in dirpath.split(os.sep)[-1:]:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                t = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"oracle/|liboracle|orc_", t):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_rust_bindings_cover_the_header():
    """integration/msgpu-sys/src/lib.rs is generated from include/msgpu.h (tools/gen_rust_bindings.py): it must be up to date
    and declare every function the header declares (there is no cargo here to compile it)."""
    import subprocess
    import sys
    assert subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_bindings.py"), "--check"]).returncode == 0, \
        "run python tools/gen_rust_bindings.py"
    src = open(os.path.join(ROOT, "integration", "msgpu-sys", "src", "lib.rs")).read()
    for s in header_symbols():
        assert re.search(r"pub fn %s\(" % s, src), "Rust binding misses %s" % s
    assert "pub pre_width: u32" in src and "pub main_width: u32" in src and "pub stage2_width: u32" in src
