"""The C-ABI library builds, loads, and exports every symbol include/neuma_b200.h declares."""
import ctypes
import os
import re

from neuma_b200 import build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "neuma_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(neuma_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in neuma_b200.h but not exported"


def test_bindings_cover_header():
    assert sorted(_lib.SIGNATURES) == _declared()


def test_load_and_version():
    lib = _lib.load()
    assert lib.neuma_version() >= 100


def test_no_gpu_is_loud():
    import torch
    import pytest
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.NeumaError):
        _lib.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "neuma_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_every_module_imports_without_a_gpu():
    """A syntax error in a module that only the -m gpu tests import must not wait for the GPU box."""
    import importlib
    for m in ("neuma_b200.pipeline", "neuma_b200.slab", "neuma_b200.ops", "neuma_b200.material", "neuma_b200.glue",
              "neuma_b200.rasterizer", "neuma_b200.sim", "neuma_b200.svd", "neuma_b200.synthetic", "neuma_b200.build",
              "bench", "__graft_entry__"):
        importlib.import_module(m)
    import glob
    import os
    import py_compile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for f in glob.glob(os.path.join(root, "tools", "*.py")) + glob.glob(os.path.join(root, "tests", "*.py")):
        py_compile.compile(f, doraise=True)


def _header_params():
    """name -> list of parameter type strings, parsed from include/neuma_b200.h"""
    text = open(os.path.join(ROOT, "include", "neuma_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    out = {}
    for m in re.finditer(r"\b(neuma_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        params = [p.strip() for p in m.group(2).replace("\n", " ").split(",")]
        out[m.group(1)] = [] if params in ([""], ["void"]) else params
    return out


def test_ctypes_signatures_match_the_header():
    """Argument COUNT and KIND (pointer / integer / float) of every binding against the header: a mismatch
    would only show up as a crash on the GPU box."""
    decl = _header_params()
    for name, argtypes in _lib.SIGNATURES.items():
        assert name in decl, name
        params = decl[name]
        assert len(params) == len(argtypes), (name, len(params), len(argtypes))
        for p, a in zip(params, argtypes):
            is_ptr = "*" in p
            if is_ptr:
                assert a is ctypes.c_void_p or a is ctypes.c_char_p or hasattr(a, "contents") or \
                    issubclass(a, ctypes._Pointer), (name, p, a)
            elif re.search(r"\bfloat\b", p):
                assert a is ctypes.c_float, (name, p, a)
            elif re.search(r"\bsize_t\b", p):
                assert a is ctypes.c_size_t, (name, p, a)
            elif re.search(r"\blong long\b", p):
                assert a is ctypes.c_longlong, (name, p, a)
            else:
                assert a in (ctypes.c_int, ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint), (name, p, a)
