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


def test_titanet_pack_weights_is_a_host_function_and_matches_the_python_packing():
    """b200d_titanet_pack_weights (C++, host) against titanet.pack_weights (the per-operand Python packing the step-by-step
    reference orchestration uses): every folded / padded / fp16-converted operand bit for bit -- except the embedding bias,
    whose 6144-term double-precision dot product is summed in a different order (compared to 1e-6 relative)."""
    import numpy as np
    import torch

    from whisper_nemo_b200 import checkpoint
    from whisper_nemo_b200 import titanet as tn

    sd = checkpoint.seeded()
    desc, blob = tn.pack_weights_cabi(sd, device="cpu")
    pk = tn.pack_weights(sd, device="cpu")
    raw = blob.numpy()

    def region(off, like):
        n = like.numel() * like.element_size()
        return torch.from_numpy(raw[off : off + n].copy()).view(like.dtype).view(like.shape)

    assert desc.n_blocks == 5 and desc.feat_in == 80 and desc.feat_pad == 128 and desc.enc_out == 3072 and desc.attn == 128 and desc.emb == 192
    assert [desc.block[i].ksize for i in range(5)] == [3, 7, 11, 15, 1] and [desc.block[i].repeat for i in range(5)] == [1, 3, 3, 3, 1]
    for i, blk in enumerate(pk.blocks):
        d = desc.block[i]
        for r, sb in enumerate(blk.subs):
            if sb.dw is None:
                assert d.dw[r] == -1
            else:
                assert torch.equal(region(d.dw[r], sb.dw), sb.dw), (i, r, "dw")
            assert torch.equal(region(d.w[r], sb.w).view(torch.int16), sb.w.view(torch.int16)), (i, r, "w")
            assert torch.equal(region(d.bias[r], sb.bias), sb.bias), (i, r, "bias")
        assert torch.equal(region(d.se_w1, blk.se_w1).view(torch.int16), blk.se_w1.view(torch.int16))
        assert torch.equal(region(d.se_w2, blk.se_w2).view(torch.int16), blk.se_w2.view(torch.int16))
        if blk.res_w is not None:
            assert torch.equal(region(d.res_w, blk.res_w).view(torch.int16), blk.res_w.view(torch.int16))
            assert torch.equal(region(d.res_bias, blk.res_bias), blk.res_bias)
        else:
            assert d.residual == 0
    for name in ("tdnn_wx", "tdnn_wctx", "attn_w2", "emb_w"):
        a, b = region(getattr(desc, name), getattr(pk, name)), getattr(pk, name)
        assert torch.equal(a.view(torch.int16), b.view(torch.int16)), name
    for name in ("tdnn_b", "tdnn_scale", "tdnn_shift", "attn_b2", "fb_start", "fb_off", "window"):
        assert torch.equal(region(getattr(desc, name), getattr(pk, name)), getattr(pk, name)), name
    assert torch.equal(region(desc.fb_w, pk.fb_w), pk.fb_w) and desc.fb_nnz == pk.fb_w.numel()
    eb = region(desc.emb_b, pk.emb_b)
    assert torch.allclose(eb, pk.emb_b, rtol=1e-6, atol=1e-9) and (eb != pk.emb_b).float().mean().item() < 0.2
    # without the preprocessor buffers in the state_dict the library computes the slaney filterbank / hann window itself
    lib = tn._cabi.load()
    names = [k for k in sd if sd[k].dtype.is_floating_point]
    keep = [sd[k].float().contiguous() for k in names]
    c_names = (ctypes.c_char_p * len(names))(*[k.encode() for k in names])
    c_data = (ctypes.c_void_p * len(names))(*[t.data_ptr() for t in keep])
    c_numel = (ctypes.c_int64 * len(names))(*[t.numel() for t in keep])
    d2 = tn._cabi.TitaNetDesc()
    assert lib.b200d_titanet_pack_weights(len(names), c_names, c_data, c_numel, ctypes.byref(d2), None, 0) == 0
    blob2 = torch.empty(int(d2.packed_bytes), dtype=torch.uint8)
    assert lib.b200d_titanet_pack_weights(len(names), c_names, c_data, c_numel, ctypes.byref(d2), blob2.data_ptr(), blob2.numel()) == 0
    raw2 = blob2.numpy()
    own_fb = np.frombuffer(raw2[d2.fb_w : d2.fb_w + 4 * d2.fb_nnz].tobytes(), dtype=np.float32)
    own_win = np.frombuffer(raw2[d2.window : d2.window + 1600].tobytes(), dtype=np.float32)
    assert d2.fb_nnz == desc.fb_nnz
    assert np.abs(own_fb - pk.fb_w.numpy()).max() <= 2e-9 and np.abs(own_win - pk.window.numpy()).max() <= 1e-7
    assert lib.b200d_titanet_workspace_bytes(ctypes.byref(desc), 131072, 4096) > 131072 * 20000
    # errors: a truncated state_dict is refused with a message
    assert lib.b200d_titanet_pack_weights(3, c_names, c_data, c_numel, ctypes.byref(d2), None, 0) < 0
    assert b"b200d_titanet_pack_weights" in lib.b200d_last_error()


def test_eig_bottomk_queries():
    from whisper_nemo_b200 import _cabi

    lib = _cabi.load()
    assert [lib.b200d_eig_bottomk_block(k) for k in (1, 8, 24, 25, 50, 56, 57)] == [32, 32, 32, 64, 64, 64, 0]
    dense = lib.b200d_eig_bottomk_workspace_bytes(10000, 50, 684, None)
    sparse = lib.b200d_eig_bottomk_workspace_bytes(10000, 50, 11, None)
    assert dense > 4 * 10000 * 64 * 4 + 2 * 192 * 10000 * 2 and 4 * 10000 * 64 * 4 < sparse < dense  # CSR lists (22 per row) instead of two bf16 operand buffers
    assert lib.b200d_eig_bottomk_workspace_bytes(10000, 60, 0, None) == 0


def test_ctypes_structures_match_the_header_layout(tmp_path):
    """Every structure that crosses the C ABI by pointer: sizeof and the offset of every field as gcc lays them out from
    include/b200d.h equal the ctypes mirror in _cabi.py (a drifted field would silently shift everything behind it)."""
    import ctypes
    import subprocess

    from whisper_nemo_b200 import _cabi

    pairs = {"b200d_gemm_epilogue": _cabi.GemmEpilogue, "b200d_peer_group": _cabi.PeerGroup, "b200d_eig_options": _cabi.EigOptions,
             "b200d_eig_stats": _cabi.EigStats, "b200d_titanet_desc": _cabi.TitaNetDesc, "b200d_profile_span": _cabi.ProfileSpan}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200d.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} sizeof %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "layout"
    include = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.check_call(["gcc", "-I", include, str(src), "-o", str(exe)])
    got = {}
    for ln in subprocess.check_output([str(exe)], text=True).splitlines():
        cname, field, val = ln.split()
        got[(cname, field)] = int(val)
    for cname, cls in pairs.items():
        assert got[(cname, "sizeof")] == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
