"""The drop-in boundary: libverilated_cuda.so loads, exports every symbol that
include/verilated_cuda_abi.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "verilated_cuda_abi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vlcu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from verilator_b200 import abi
    names = declared_symbols()
    assert len(names) >= 45
    for n in names:
        assert hasattr(abi.lib, n), f"libverilated_cuda.so lacks {n}"
    assert sorted(abi.SIGNATURES) == names, "abi.py and the header disagree"
    assert abi.lib.vlcu_abi_version() == 3


def test_designs_are_registered():
    from verilator_b200 import abi
    assert {"make_tracing_c", "make_tracing_c_cov"} <= set(abi.designs())


def test_no_cpu_fallback():
    from verilator_b200 import abi
    if abi.lib.vlcu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(abi.VlcuError, match="no CUDA device"):
        abi.Model("make_tracing_c", 4)


def test_unknown_design_is_an_error():
    from verilator_b200 import abi
    with pytest.raises(abi.VlcuError, match="unknown design"):
        abi.Model("no_such_design", 4)


def test_product_does_not_link_the_oracle():
    import subprocess
    out = subprocess.run(["ldd", os.path.join(ROOT, "verilator_b200", "libverilated_cuda.so")],
                         capture_output=True, text=True).stdout
    assert "vlref" not in out and "verilated_ref" not in out


def test_cxx_binding_compiles_and_fails_loudly_without_a_gpu():
    """include/verilated_cuda_model.h (the Vtop-shaped C++ binding) builds against the product
    library; without a device the testbench must stop with the no-CPU-path error."""
    import subprocess
    from verilator_b200 import abi
    exe = os.path.join(ROOT, "tests", "cxx", "sim_main_batch")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    if abi.lib.vlcu_device_count() > 0:
        pytest.skip("a GPU is present; the run itself is covered by the gpu tests")
    res = subprocess.run([exe, "4", "10", "/dev/null"], capture_output=True, text=True)
    assert res.returncode == 1 and "no CPU path" in res.stderr
