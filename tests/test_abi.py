"""CPU: the C-ABI library loads and exports every symbol include/gplite_b200.h declares, the
ctypes table covers them all, and the product fails loudly (no CPU fallback) without a GPU."""
import ctypes
import os
import re

import pytest

import gplite_b200 as g
from gplite_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gplite_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpl_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in gplite_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == syms


def test_abi_version_and_error_string():
    lib = g.load_library()
    assert lib.gpl_abi_version() == 2
    assert isinstance(lib.gpl_last_error(), bytes)


def test_no_cpu_fallback():
    lib = g.load_library()
    if lib.gpl_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(g.GplError) as ei:
        g.Context(0)
    assert ei.value.status == _lib.GPL_ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gplite_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{f} references the oracle"


def test_header_is_plain_c_and_links(tmp_path):
    """The boundary is a C ABI: the header must compile as C (not only as C++), and a C translation
    unit that takes the address of every declared function must link against the library."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    syms = declared_symbols()
    src = tmp_path / "abi_check.c"
    src.write_text('#include "gplite_b200.h"\n#include <stddef.h>\n'
                   'typedef void (*fn)(void);\n'
                   'fn table[] = {' + ", ".join(f"(fn){s}" for s in syms) + '};\n'
                   'int main(void) { return table[0] == NULL; }\n')
    exe = tmp_path / "abi_check"
    inc = os.path.join(ROOT, "include")
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{inc}", str(src), "-o", str(exe),
           f"-L{libdir}", "-lgplite_b200", f"-Wl,-rpath,{libdir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
