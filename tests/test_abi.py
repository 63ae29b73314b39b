"""The C-ABI library builds, loads and exports every symbol include/xva_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "xva_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xva_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    from xva_trainer_b200 import capi

    names = _declared_symbols()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/xva_b200.h but not exported"
        assert n in capi.PROTOTYPES, f"{n} has no ctypes prototype in capi.py"
    assert set(capi.PROTOTYPES) == set(names)


def test_abi_version_and_struct_layout(lib):
    from xva_trainer_b200 import capi

    assert lib.xva_abi_version() == 1
    assert lib.xva_sizeof_gemm_args() == ctypes.sizeof(capi.GemmArgs)


def test_no_gpu_is_a_loud_error(lib):
    import torch

    if torch.cuda.is_available():
        return
    assert lib.xva_device_check(0) != 0
    assert lib.xva_last_error()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "xva-trainer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f"{f} mentions the oracle: the product path must not depend on it"
                assert "cabi_emu" not in src, f"{f} mentions the CPU stand-in of the C ABI (tests/cabi_emu.py): test infrastructure only"
