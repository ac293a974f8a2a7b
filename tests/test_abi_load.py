"""CPU: libmcpm.so loads and exports every symbol include/mcpm.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "mcpm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcpm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(mcpm):
    names = header_symbols()
    assert len(names) >= 25
    L = ctypes.CDLL(mcpm.lib_path())
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/mcpm.h but not exported by libmcpm.so"


def test_python_prototypes_cover_header(mcpm):
    from mcpm_b200 import abi
    assert sorted(abi.PROTOTYPES) == header_symbols()


def test_struct_layouts_match_oracle(mcpm):
    from conftest import O
    assert ctypes.sizeof(mcpm.SystemDesc) == ctypes.sizeof(O.System) == 8 * 4 + 16 * 8
    for (a, ta), (b, tb) in zip(mcpm.SystemDesc._fields_, O.System._fields_):
        assert getattr(mcpm.SystemDesc, a).offset == getattr(O.System, b).offset


def test_version_and_error_paths(mcpm):
    L = mcpm.lib()
    assert L.mcpm_version() == 200
    h = ctypes.c_void_p()
    assert L.mcpm_create(0, 0, 16, ctypes.byref(h)) != 0          # bad argument, no device needed
    assert b"positive" in L.mcpm_last_error()


def test_no_cpu_fallback_without_device(mcpm):
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    try:
        n = mcpm.device_count()
    except mcpm.McpmError:
        n = 0
    if n > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(mcpm.McpmError):
        mcpm.Engine(1, 64)


def test_product_does_not_link_oracle():
    import subprocess
    from mcpm_b200 import abi
    out = subprocess.run(["ldd", abi.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out and "mcref" not in out
    src = os.path.join(ROOT, "mc-porousmaterials_b200")
    for dirpath, _, files in os.walk(src):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py")):
                text = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in text and "libmcref" not in text and "from oracle" not in text, f
