"""CPU tests of the drop-in boundary: libnlist_b200.so loads, exports every symbol include/nlist_b200.h declares
(and the Python binding declares the same set), argument validation works, and compute entry points fail loudly —
never fall back — when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "nlist_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nlb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from md_neighbor_list_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 25
    L = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"libnlist_b200.so does not export {n}"
    assert sorted(_lib.SYMBOLS) == names, "python binding and header disagree"
    assert _lib.lib().nlb200_version() == 100


def test_header_is_plain_c():
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        with open(src, "w") as f:
            f.write('#include "nlist_b200.h"\nint main(void){return nlb200_version;}\n'.replace(
                "return nlb200_version;", "return NLB200_VERSION == 100 ? 0 : 1;"))
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                               "-c", src, "-o", os.path.join(d, "t.o")])


def test_create_validates_arguments():
    from md_neighbor_list_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    # fewer than 3 cells on an axis is rejected (the reference would emit duplicate pairs)
    assert L.nlb200_create(3.3, 9.0, 50.0, 50.0, _lib.F64, _lib.FULL_CSR, C.byref(h)) == _lib.ERR_INVALID
    assert L.nlb200_create(-1.0, 50.0, 50.0, 50.0, _lib.F64, _lib.FULL_CSR, C.byref(h)) == _lib.ERR_INVALID
    assert L.nlb200_create(3.3, 50.0, 50.0, 50.0, 7, _lib.FULL_CSR, C.byref(h)) == _lib.ERR_INVALID
    assert L.nlb200_create(3.3, 50.0, 50.0, 50.0, _lib.F64, _lib.FULL_CSR, C.byref(h)) == _lib.OK
    st = _lib.Stats()
    assert L.nlb200_get_stats(h, C.byref(st)) == _lib.OK
    assert list(st.mesh) == [15, 15, 15]  # neighlist_cpu.hpp:384-387 with L=50, SL=3.3
    assert L.nlb200_set_option(h, _lib.OPT_POSITION_STRIDE, 5) == _lib.ERR_INVALID
    assert L.nlb200_set_option(h, _lib.OPT_POSITION_STRIDE, 3) == _lib.OK
    assert L.nlb200_build(h, None, 0, None) == _lib.ERR_STATE  # build before initialize
    assert b"initialize" in L.nlb200_last_error(h)
    assert L.nlb200_destroy(h) == _lib.OK


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    from md_neighbor_list_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    assert L.nlb200_create(3.3, 50.0, 50.0, 50.0, _lib.F64, _lib.HALF_CSR, C.byref(h)) == _lib.OK
    assert L.nlb200_initialize(h, 1000, 0) == _lib.ERR_CUDA
    assert b"no CPU fallback" in L.nlb200_last_error(h)
    L.nlb200_destroy(h)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "md_neighbor_list_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    txt = f.read()
                assert "from oracle" not in txt and "import oracle" not in txt and "liboracle" not in txt, fn
