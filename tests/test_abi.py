"""CPU checks of the drop-in boundary: the library loads and exports exactly what
include/cerebro_b200.h declares; no compute calls."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_in_header():
    src = open(os.path.join(ROOT, "include", "cerebro_b200.h")).read()
    return sorted(set(re.findall(r"^CB_API [^;(]*?\b(cb_\w+)\(", src, flags=re.M)))


def test_header_symbols_exported(native_lib):
    names = _declared_in_header()
    assert len(names) >= 30
    for n in names:
        assert hasattr(native_lib, n), "symbol %s declared in the header but not exported" % n


def test_ctypes_table_matches_header(native_lib):
    from cerebro_b200 import _lib

    assert sorted(_lib.declared_symbols()) == _declared_in_header()


def test_version_and_error_string(native_lib):
    assert native_lib.cb_version() == 2
    assert isinstance(native_lib.cb_last_error(), bytes)


def test_create_fails_loudly_without_gpu(native_lib):
    """No CPU fallback: on a box without a GPU every create call must fail with a message."""
    import ctypes as C

    import torch

    if torch.cuda.is_available():
        return
    h = C.c_void_p()
    rc = native_lib.cb_index_create(C.byref(h), 4096, 100, 0, 0, 1)
    assert rc == -2
    assert b"no CPU fallback" in native_lib.cb_last_error()


def test_header_cites_reference():
    src = open(os.path.join(ROOT, "include", "cerebro_b200.h")).read()
    for cite in ("src/Cerebro.cpp:390", "src/DlsPnpWithRansac.cpp:132-245", "WholeImageDescriptorCompute.srv"):
        assert cite in src


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (cgo / FFI consumers) and as C++11, warning-free."""
    import subprocess

    src = tmp_path / "cabi.c"
    src.write_text('#include "cerebro_b200.h"\nint main(void){ cb_ransac_params p; cb_ir_block b; (void)p; (void)b; return cb_version() > 0 ? 0 : 1; }\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror"], ["g++", "-std=c++11", "-Wall", "-Werror", "-x", "c++"]):
        r = subprocess.run(cmd + ["-I" + inc, "-c", str(src), "-o", str(tmp_path / "cabi.o")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_package_root_exports_the_reference_named_mirrors():
    import cerebro_b200 as cb

    for name in cb.__all__:
        assert getattr(cb, name) is not None
    for name in ("HDF5ModelImageDescriptor", "IndexFlatIP", "StaticTheiaPoseCompute", "Cerebro", "LoopEdge"):
        assert name in cb.__all__
