"""The C-ABI library loads and exports every entry point include/quadray_b200.h
declares.  No compute calls here (no GPU in the build container); on a box
without a CUDA device qr_init must fail loudly -- there is no CPU fallback."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "quadray_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qr_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree(pkg):
    assert declared_symbols() == sorted(pkg.SYMBOLS)


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_library_is_sm100a_only():
    """cuobjdump lists exactly one cubin: sm_100a."""
    import subprocess
    lib = os.path.join(ROOT, "quadray-engine_b200", "lib", "libquadray_b200.so")
    tool = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not installed")
    out = subprocess.run([tool, "-lelf", lib], stdout=subprocess.PIPE, check=True).stdout.decode()
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, out


def test_init_without_gpu_fails_loudly(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.QuadRayError) as ei:
        pkg.Context([0])
    assert ei.value.code == pkg.QR_E_NODEV
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_reference_the_oracle():
    """Nothing under quadray-engine_b200/ or include/ may include, link or load
    anything from oracle/ or tests/hostsim."""
    bad = []
    for top in ("quadray-engine_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if not f.endswith((".cpp", ".cu", ".cuh", ".h", ".py", "Makefile")):
                    continue
                text = open(os.path.join(dp, f), errors="replace").read()
                code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
                code = re.sub(r"#.*", "", code) if f.endswith((".py", "Makefile")) else code
                if re.search(r"render0_oracle|libqr_oracle|qr_oracle_render|qr_hostsim", code):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
