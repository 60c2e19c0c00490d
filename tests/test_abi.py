"""The C-ABI library loads and exports every symbol include/autocycler_gpu.h declares (no compute calls:
this runs without a GPU), and refuses to run without a device instead of falling back to a CPU path."""
import os
import re
import subprocess

import pytest

from autocycler_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "autocycler_b200", "csrc")], check=True)
    return api.load_library()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "autocycler_gpu.h")).read()
    return sorted(set(re.findall(r"\b(ac_[a-z_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/autocycler_gpu.h but not exported"
    assert sorted(api.EXPORTS) == names


def test_built_for_sm_100a():
    out = subprocess.run(["cuobjdump", "--list-elf", os.path.join(ROOT, "autocycler_b200", "libautocycler_gpu.so")],
                         capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.AutocyclerGpuError) as e:
        api.KmerGraph(51, lib=lib)
    assert e.value.code == -2


def test_product_does_not_reference_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "autocycler_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in text and "liboracle" not in text and "autocycler_oracle" not in text, f
