"""The C-ABI library loads and exports every symbol include/b200d.h declares; the product refuses to run without a
B200 (no silent fallback).  No compute calls here: this file runs on a CPU-only box."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "b200d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200d_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_bound_symbols():
    from whisper_nemo_b200 import _cabi

    declared = _declared()
    assert len(declared) >= 25
    assert sorted(_cabi.exported_symbols()) == [d for d in declared]


def test_ctypes_signatures_follow_the_header():
    """Same parameter count, and pointer / integer / float in the same positions, for every declaration of
    include/b200d.h and its ctypes binding (an argument added on one side only would shift every later one silently)."""
    from whisper_nemo_b200 import _cabi

    text = open(os.path.join(ROOT, "include", "b200d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = dict(re.findall(r"\b(b200d_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S))
    assert sorted(decls) == sorted(_cabi._SIGNATURES)

    def kind_c(param):
        param = param.strip()
        if "*" in param:
            return "ptr"
        if re.match(r"(const\s+)?float\b", param):
            return "float"
        return "int"

    def kind_py(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or isinstance(t, type(ctypes.POINTER(ctypes.c_int))):
            return "ptr"
        return "float" if t is ctypes.c_float else "int"

    for name, params in decls.items():
        params = params.strip()
        c_kinds = [] if params in ("", "void") else [kind_c(q) for q in params.split(",")]
        py_kinds = [kind_py(t) for t in _cabi._SIGNATURES[name][1]]
        assert c_kinds == py_kinds, (name, c_kinds, py_kinds)


def test_library_exports_every_declared_symbol():
    from whisper_nemo_b200 import _cabi, build

    build.build_library()  # no-op when the in-tree .so is current
    lib = _cabi.load()
    assert lib._b200d_missing == []
    raw = ctypes.CDLL(_cabi.LIB_PATH)
    for name in _declared():
        assert hasattr(raw, name), name
    assert b"sm_100a" in lib.b200d_version()


def test_workspace_queries_are_pure_host_functions():
    from whisper_nemo_b200 import _cabi

    lib = _cabi.load()
    assert lib.b200d_eigvals_workspace_bytes(30, 600) == 30 * 600 * 4 * 4 + 30 * 4
    assert lib.b200d_gram_workspace_bytes(1000, 32) == 4 * 32 * 32 * 4
    assert lib.b200d_kmeans_workspace_bytes(100, 4, 4, 30) > 0
    assert lib.b200d_eigvals_workspace_bytes(0, 10) == 0


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: the refusal path is only observable on a CPU-only box")
    from whisper_nemo_b200 import ClusteringDiarizer, config

    cfg = config.load_config("telephonic")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ClusteringDiarizer(cfg)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "whisper_nemo_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
