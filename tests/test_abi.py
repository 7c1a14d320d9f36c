"""The C-ABI library builds, loads and exports every symbol include/emdr2_b200.h declares
(no compute calls here: this file runs without a GPU)."""
import ctypes
import os
import re

import pytest

from emdr2_b200 import _lib, build as build_mod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for fn in sorted(os.listdir(inc)):
        if fn.endswith(".h"):
            text = open(os.path.join(inc, fn)).read()
            names.update(re.findall(r"EMDR2_API\s+[\w\s\*]+?\b(emdr2_\w+)\s*\(", text))
    return sorted(names)


def test_header_declares_the_documented_entry_points():
    names = declared_symbols()
    for must in ["emdr2_mips_create", "emdr2_mips_set_shard", "emdr2_mips_search",
                 "emdr2_mips_search_host", "emdr2_mips_merge", "emdr2_mips_destroy",
                 "emdr2_last_error", "emdr2_version"]:
        assert must in names


def test_library_builds_and_exports_every_declared_symbol():
    path = build_mod.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), "libemdr2_b200.so does not export %s" % name


def test_ctypes_binding_covers_the_header():
    assert set(_lib.exported_symbols()) == set(declared_symbols())


def test_version_string_names_the_target():
    lib = _lib.load()
    v = lib.emdr2_version().decode()
    assert "sm_100a" in v and v.startswith("emdr2_b200")


def test_sass_contains_tcgen05_and_tma():
    """The scan kernel is a real Blackwell kernel: UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld),
    UTMALDG (TMA load) appear in the SASS of the built library."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", build_mod.build()], capture_output=True, text=True).stdout
    for mnemonic in ["UTCHMMA", "LDTM", "UTMALDG"]:
        assert mnemonic in sass, mnemonic
    # the CTA-pair GEMM (cta_group::2 MMA, pair TMA loads, multicast commits) and the packed-fp32 GeLU
    for mnemonic in ["UTCHMMA.2CTA", "UTMALDG.2D.2CTA", "UTCBAR.2CTA.MULTICAST", "FFMA2", "FMUL2", "STTM"]:
        assert mnemonic in sass, mnemonic


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this test is about the GPU-less container")
    from emdr2_b200.mips import ShardSearcher
    with pytest.raises(RuntimeError):
        ShardSearcher(128, torch.float16, "cpu")
    with pytest.raises(_lib.Emdr2Error):
        ShardSearcher(128, torch.float16, "cuda:0")
    from emdr2_b200.index import B200BruteForceIndex
    with pytest.raises(RuntimeError):
        B200BruteForceIndex(128)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "emdr2_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), fn
                assert "liboracle" not in text, fn


def test_the_abi_is_usable_from_plain_c(tmp_path):
    """tests/c/abi_smoke.c compiled with gcc against include/emdr2_b200.h and linked to the library:
    version string, the host-side formatter on a tiny batch, the error convention."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    lib = build_mod.build()
    exe = str(tmp_path / "abi_smoke")
    src = os.path.join(ROOT, "tests", "c", "abi_smoke.c")
    libdir = os.path.dirname(lib)
    res = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                          "-L", libdir, "-lemdr2_b200", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0, (run.returncode, run.stderr)
    assert run.stdout.startswith("emdr2_b200")
