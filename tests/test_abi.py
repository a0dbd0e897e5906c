"""CPU: the C-ABI shared library loads and exports every symbol include/sequila_cuda.h declares.
No compute call is made here (there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols(name="sequila_cuda.h"):
    text = open(os.path.join(ROOT, "include", name)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sq_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = header_symbols()
    for must in ("sq_ctx_create", "sq_index_build", "sq_probe_count", "sq_probe_emit_pairs",
                 "sq_gather_column", "sq_last_error", "sq_index_free", "sq_stream_free"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    import sequila_native_b200 as sn
    from sequila_native_b200 import _native
    assert os.path.exists(sn.LIB_PATH), "libsequila_cuda.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(sn.LIB_PATH)
    declared = header_symbols()
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    # the Python binding covers the same surface, so tests call exactly what a Rust -sys crate would
    assert sorted(_native.SIGNATURES) == declared
    # same for the exec-node layer (Arrow C Data Interface), include/sequila_exec.h
    declared_exec = [x for x in header_symbols("sequila_exec.h") if x.startswith("sq_exec_")]
    missing = [x for x in declared_exec if not hasattr(lib, x)]
    assert not missing, missing
    assert sorted(_native.EXEC_SIGNATURES) == declared_exec
    # and the scan side (delimited text -> device columns), include/sequila_scan.h
    declared_scan = [x for x in header_symbols("sequila_scan.h") if x.startswith("sq_scan_")]
    assert "sq_scan_text" in declared_scan and "sq_scan_key_ids_device" in declared_scan
    missing = [x for x in declared_scan if not hasattr(lib, x)]
    assert not missing, missing
    assert sorted(_native.SCAN_SIGNATURES) == declared_scan
    # and the native partition loop, include/sequila_driver.h
    declared_drv = [x for x in header_symbols("sequila_driver.h") if x.startswith("sq_driver_")]
    assert "sq_driver_run" in declared_drv and not [x for x in declared_drv if not hasattr(lib, x)]
    assert sorted(_native.DRIVER_SIGNATURES) == declared_drv


@pytest.mark.parametrize("header", ["sequila_cuda.h", "sequila_exec.h", "sequila_scan.h", "sequila_driver.h"])
def test_headers_are_plain_c(header):
    """the boundary is a C ABI: every header must compile as C11 on its own (what cgo / bindgen / a C host would see)"""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    p = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I",
                        os.path.join(ROOT, "include"), "-x", "c", "-"], input=f'#include "{header}"\n', text=True,
                       capture_output=True)
    assert p.returncode == 0, p.stderr


def test_abi_version_and_no_fallback_without_gpu():
    import sequila_native_b200 as sn
    from sequila_native_b200 import _native
    lib = _native.lib()
    assert lib.sq_abi_version() == 1
    if lib.sq_device_count() == 0:
        # product path must fail loudly, not fall back to a CPU implementation
        with pytest.raises(sn.SequilaCudaError) as e:
            sn.CudaContext(0)
        assert "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sequila_native_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "liborc" not in src and "coitrees_port" not in src, f
