"""CPU tests of the drop-in boundary: librsvdb.so loads, exports every symbol include/rsvdb.h declares, and refuses to
run without a GPU (no CPU fallback)."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    from rsvd_kamaneh_raganato_terrana_b200 import capi
    lib = capi.load()
    declared = capi.exported_symbols()
    assert len(declared) >= 32
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/rsvdb.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", str(capi.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (rsvdb_\w+)", out))
    assert exported == set(declared)
    assert lib.rsvdb_version().startswith(b"rsvdb")
    assert lib.rsvdb_pm_iterations(100) == 148 and lib.rsvdb_pm_iterations(20000) == 151   # reference src/PM.cpp:25-28


def test_header_is_plain_c():
    """include/rsvdb.h must compile as C (extern "C", plain pointers and sizes, no C++ or torch types)."""
    src = ROOT / "build" / "hdr_check.c"
    src.parent.mkdir(exist_ok=True)
    src.write_text('#include "rsvdb.h"\nint main(void){ return (int)sizeof(rsvdb_status) * 0; }\n')
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-pedantic", "-Werror", "-I", str(ROOT / "include"), "-fsyntax-only", str(src)], check=True)


def test_no_cpu_fallback():
    import torch
    from rsvd_kamaneh_raganato_terrana_b200 import capi, Engine
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible; the refusal path is exercised on the CPU box")
    lib = capi.load()
    h = ctypes.c_void_p()
    assert lib.rsvdb_create(ctypes.byref(h), 0) == capi.ERR_CUDA
    with pytest.raises(capi.RsvdbError):
        Engine(0)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing in the product package may import, load or link it."""
    pkg = ROOT / "rsvd_kamaneh_raganato_terrana_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + [q for q in (ROOT / "include").rglob("*") if q.is_file()]:
        text = p.read_text()
        for pat in (r"^\s*(from|import)\s+oracle", r"rsvd_oracle", r"liboracle", r"libref_rsvd", r"oracle/", r"oracle_c"):
            assert not re.search(pat, text, flags=re.M), f"{p} reaches into the oracle ({pat})"


def test_cpp_dropin_headers_compile_and_link():
    """The C++ drop-in headers (the reference's file names and symbols) compile with a plain host compiler and link against
    librsvdb.so -- no CUDA headers, no torch, no Eigen needed by a caller."""
    libdir = ROOT / "rsvd_kamaneh_raganato_terrana_b200"
    out = ROOT / "build"
    out.mkdir(exist_ok=True)
    for src in ("rsvd_dropin_test.cpp", "rsvd_test_main.cpp", "rsvd_v1_test.cpp", "pca_test.cpp", "pod_test.cpp", "image_test.cpp"):
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-o", str(out / (src[:-4] + "_cpu")),
                        str(ROOT / "tests" / "cpp" / src), "-L", str(libdir), "-lrsvdb", f"-Wl,-rpath,{libdir}"], check=True)
    # every reference header name on the path has a drop-in of the same name
    for h in ("rSVD.hpp", "SVD_class.hpp", "QR.hpp", "PM.hpp", "Jacobi_Class.hpp", "JacobiOperations.hpp", "matrixOperations.hpp", "PCA_class.hpp", "POD.hpp",
              "image_compression/rSVD.hpp", "image_compression/SVD.hpp", "image_compression/PowerMethod.hpp", "image_compression/QR.hpp", "image_compression/image_comp.hpp"):
        assert (ROOT / "include" / h).exists(), h


def test_reference_mains_compile_unmodified_over_dropin_headers():
    """The reference's own tests/rSVD_test.cpp, svd_test.cpp, QRTest.cpp, rSVD_test2.cpp, PCA/tests/*.cpp and POD.cpp build, unchanged,
    with -I include first (tests/cpp/Makefile).  Compile-and-link only on the CPU box; tests/test_gpu_parity.py runs them."""
    import subprocess
    from pathlib import Path
    import pytest
    root = Path(__file__).resolve().parent.parent
    if not Path("/root/reference/tests").is_dir():
        pytest.skip("/root/reference is not present on this box")
    r = subprocess.run(["make", "-B", "-C", str(root / "tests" / "cpp"), "refmains"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for name in ("rSVD_test", "svd_test", "QRTest", "rSVD_test2", "pca_test", "athletic_test", "pod_ref_class"):
        assert (root / "tests" / "cpp" / "_refbin" / name).exists(), name
