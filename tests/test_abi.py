"""CPU: the C-ABI library loads without a GPU and exports exactly what include/bpreveal_b200.h
declares; the ctypes binding mirrors the header one to one (no compute calls here)."""
import ctypes
import os
import re

import pytest

from bpreveal_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bpreveal_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(bpb_\w+)\s*\(([^;{]*?)\)\s*;", src):
        name, args = m.group(1), m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[name] = n
    return out


def test_library_is_built_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = _lib.load()
    assert lib.bpb_version() >= 1
    assert isinstance(lib.bpb_last_error(), bytes)


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    decl = header_functions()
    assert len(decl) >= 35
    missing = [name for name in decl if not hasattr(lib, name)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_binding_mirrors_header():
    decl = header_functions()
    assert set(_lib.SIGNATURES) == set(decl), (
        sorted(set(decl) - set(_lib.SIGNATURES)), sorted(set(_lib.SIGNATURES) - set(decl)))
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        assert len(argtypes) == decl[name], f"{name}: binding has {len(argtypes)} args, " \
                                            f"header {decl[name]}"


def test_struct_layouts_match_compiled_sizes():
    lib = _lib.load()
    for which, cls in enumerate((_lib.Conv3Args, _lib.Wgrad3Args, _lib.InputConvArgs,
                                 _lib.HeadsDgradArgs, _lib.PrepTable, _lib.PeerArgs)):
        assert lib.bpb_abi_struct_size(which) == ctypes.sizeof(cls), cls.__name__
    assert lib.bpb_abi_struct_size(99) == -1


def test_no_silent_fallback_without_gpu():
    """Without a CUDA device the entry points fail loudly with a message (no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the GPU-less build container")
    lib = _lib.load()
    a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = lib.bpb_device_info(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    assert rc != 0 and b"CUDA" in lib.bpb_last_error()
    assert lib.bpb_wgrad3_ws_bytes(4, 100, 64, 0) == -1
    with pytest.raises(_lib.BprevealB200Error):
        _lib.check(rc, "bpb_device_info")


def test_missing_library_raises(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libbpreveal_b200.so")
    with pytest.raises(_lib.BprevealB200Error):
        _lib.load()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under bpreveal_b200/ may reference it."""
    pkg = os.path.join(ROOT, "bpreveal_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), \
                    f"{f} imports the oracle"
                assert "import_module(\"oracle" not in text and "__import__(\"oracle" not in text


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C99 (no C++ constructs, no torch types),
    so that a cgo / JNI / ctypes / f2py-style binding can include it as is."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc on this box")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HEADER],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
