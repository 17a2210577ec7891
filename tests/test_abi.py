"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol include/blr_cuda.h declares.
No compute call is made (there is no GPU here); the product path must fail loudly instead of falling back."""
import os
import re
import subprocess

import pytest

import blr_b200 as blr
from blr_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    blr.build()
    return blr.load()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "blr_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(blr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/blr_cuda.h but not exported by libblr_cuda.so"


def test_binding_covers_header(lib):
    assert sorted(L.SIGNATURES) == declared_symbols()


def test_version(lib):
    assert lib.blr_version() == 100


def test_sass_is_blackwell_native():
    """The Gram kernel must contain fp64 tensor-core MMAs and TMA bulk copies (cuobjdump needs no GPU)."""
    out = subprocess.run(["cuobjdump", "-sass", L.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out or "SM100" in out.upper()
    assert out.count("DMMA.8x8x4") > 100
    assert "UBLKCP" in out and "SYNCS.ARRIVE.TRANS64" in out


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(blr.BLRError):
        blr.Context()
    f = blr.BayesianLinearRegressor([0.0, 0.0], blr.Diagonal([1.0, 1.0]))
    import numpy as np

    with pytest.raises(blr.BLRError):
        blr.logpdf(f(blr.ColVecs(np.ones((2, 3))), 0.1), np.zeros(3))


def test_plain_c_client_compiles_links_and_fails_loudly(lib, tmp_path):
    """include/blr_cuda.h is a C header (not just C++): examples/minimal_client.c builds as C99 against the library.
    Without a GPU it must stop at blr_ctx_create with the no-fallback message (exit code 2)."""
    import torch

    exe = tmp_path / "minimal_client"
    libdir = os.path.dirname(L.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "minimal_client.c"), "-L", libdir, "-lblr_cuda", f"-Wl,-rpath,{libdir}",
                    "-o", str(exe)], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert res.returncode == 0, res.stderr
    else:
        assert res.returncode == 2 and "no CPU fallback" in res.stderr


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "bayesianlinearregressors.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in txt.lower().replace("# oracle", ""), f"{fn} mentions the oracle"
