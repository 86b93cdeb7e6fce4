"""CPU tests of the C-ABI boundary: the library loads, exports every symbol the header declares, and fails
loudly (no CPU fallback) when there is no GPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "scalaopt_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(so_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(native):
    lib = native.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    missing = [s for s in declared if not hasattr(lib, s)]
    assert missing == []
    assert sorted(native.EXPORTS) == declared
    assert lib.so_abi_version() == 2


def test_no_cpu_fallback_without_a_gpu(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(native.NativeError) as e:
        native.Context(1)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure; nothing under scalaopt_b200/ may reference it."""
    pkg = os.path.join(ROOT, "scalaopt_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", ".c")):
                src = open(os.path.join(dirpath, fn), errors="ignore").read()
                if re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M) or "liboracle" in src or "oracle.h" in src:
                    offenders.append(os.path.join(dirpath, fn))
    assert offenders == []
