"""Kernel-level parity on the B200: every C-ABI entry point of the clustering stage against torch fp64 /
the CPU oracle on seeded inputs.  Run with `pytest -m gpu`."""
import ctypes
import functools
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _unsplit(vt, b, n):
    """[3b (+ padding), ldvt] bf16 parts [hi | mid | lo] of V^T, part q of column j in row q * b + j -> float32 [n, b]"""
    return (vt[:b, :n].float() + vt[b : 2 * b, :n].float() + vt[2 * b : 3 * b, :n].float()).t()


def _lap_batch(n, batch, seed):
    g = torch.Generator().manual_seed(seed)
    mats = []
    for b in range(batch):
        a = (torch.rand(n, n, generator=g) < (0.02 + 0.03 * b / max(batch - 1, 1))).float()
        a = 0.5 * (a + a.t())
        a.fill_diagonal_(0)
        mats.append(torch.diag(a.sum(1)) - a)
    return torch.stack(mats)


@pytest.mark.parametrize("n,batch,n_low", [(2, 1, 1), (7, 3, 6), (100, 5, 9), (515, 30, 9), (600, 30, 9), (110, 10, 81), (1000, 12, 51)])
def test_eigvals_batched(dev, n, batch, n_low):
    from whisper_nemo_b200 import _cabi
    from whisper_nemo_b200._cabi import ptr

    n_low = min(n_low, n)
    lap = _lap_batch(n, batch, seed=n)
    ref = torch.linalg.eigvalsh(lap.double())
    a = lap.to(dev).contiguous()
    ws_bytes = _cabi.load().b200d_eigvals_workspace_bytes(batch, n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    ev = torch.empty(batch, n_low + 1, dtype=torch.float32, device=dev)
    _cabi.call("b200d_eigvals_batched", ptr(a), batch, n, n_low, ptr(ev), ptr(ws), ws_bytes, _cabi._stream())
    torch.cuda.synchronize()
    ev = ev.cpu().double()
    scale = ref.abs().max().item()
    err_low = (ev[:, :n_low] - ref[:, :n_low]).abs().max().item()
    err_max = (ev[:, n_low] - ref[:, -1]).abs().max().item()
    print(f"eigvals n={n} batch={batch}: low err {err_low:.3e} max err {err_max:.3e} (scale {scale:.1f})")
    assert err_low <= 2e-5 * scale + 1e-5 and err_max <= 2e-5 * scale + 1e-5


def _rank(mat, stride, dev):
    from whisper_nemo_b200 import _cabi
    from whisper_nemo_b200._cabi import ptr

    n = len(range(0, mat.shape[0], stride))
    rank = torch.empty(n, n, dtype=torch.int16, device=dev)
    rankT = torch.empty(n, n, dtype=torch.int16, device=dev)
    _cabi.call("b200d_row_rank", ptr(mat), mat.stride(0), stride, n, ptr(rank), ptr(rankT), _cabi._stream())
    return rank, rankT, n


@pytest.mark.parametrize("N,stride", [(300, 1), (1023, 1), (2399, 4), (3000, 5)])
def test_rank_laplacian_reach(dev, N, stride):
    from oracle import offline_clustering as oc
    from whisper_nemo_b200 import _cabi
    from whisper_nemo_b200._cabi import ptr

    g = torch.Generator().manual_seed(N)
    x = torch.randn(N, 16, generator=g)
    x[: N // 2] += 2.0
    mat = oc.getCosAffinityMatrix(x)
    mat[5, 7] = mat[5, 9]  # an exact tie inside a row
    sub = mat[::stride, ::stride].contiguous()
    md = mat.to(dev)
    rank, rankT, n = _rank(md, stride, dev)
    torch.cuda.synchronize()
    order = torch.argsort(sub, dim=1, descending=True, stable=True)
    want = torch.empty_like(order)
    want.scatter_(1, order, torch.arange(n).repeat(n, 1))
    assert torch.equal(rank.cpu().long(), want)
    assert torch.equal(rankT.cpu().long(), want.t())
    p_list = [1, 2, max(2, n // 40), max(3, n // 8)]
    lap = torch.empty(len(p_list), n, n, dtype=torch.float32, device=dev)
    arr = (ctypes.c_int32 * len(p_list))(*p_list)
    _cabi.call("b200d_laplacian_from_rank", ptr(rank), ptr(rankT), n, arr, len(p_list), ptr(lap), _cabi._stream())
    reach = torch.empty(len(p_list), dtype=torch.int32, device=dev)
    _cabi.call("b200d_graph_reach_rank", ptr(rank), ptr(rankT), n, arr, len(p_list), ptr(reach), _cabi._stream())
    torch.cuda.synchronize()
    for b, p in enumerate(p_list):
        aff = oc.getAffinityGraphMat(sub.clone(), p)
        comp = int(oc.getTheLargestComponent(aff, 0).sum())
        want_lap = oc.getLaplacian(aff.clone()).float()
        assert torch.equal(lap[b].cpu(), want_lap), f"laplacian mismatch at p={p}"
        assert int(reach[b]) == comp, f"reach mismatch at p={p}: {int(reach[b])} vs {comp}"


@pytest.mark.parametrize("N,p", [(200, 3), (2399, 40), (5000, 1300)])
def test_topp_binarize(dev, N, p):
    from oracle import offline_clustering as oc
    from whisper_nemo_b200 import clustering as cl

    g = torch.Generator().manual_seed(N + p)
    x = torch.randn(N, 24, generator=g)
    mat = oc.getCosAffinityMatrix(x)
    mat[3, 10] = mat[3, 20]
    a16, deg = cl.getAffinityGraphMat(mat.to(dev), p)
    torch.cuda.synchronize()
    aff = oc.getAffinityGraphMat(mat.clone(), p)
    want = aff.clone().float()
    want.fill_diagonal_(0)
    assert torch.equal(a16[:, :N].float().cpu(), want)
    assert a16[:, N:].abs().sum().item() == 0
    lap = oc.getLaplacian(aff.clone()).float()
    assert torch.equal(deg.cpu(), torch.diagonal(lap))


@pytest.mark.parametrize("n,b", [(500, 32), (5000, 64), (14399, 32)])
def test_gram_smalleig_rightmul_resid(dev, n, b):
    from whisper_nemo_b200 import _cabi
    from whisper_nemo_b200 import clustering as cl
    from whisper_nemo_b200._cabi import ptr

    g = torch.Generator().manual_seed(n + b)
    x = torch.randn(n, b, generator=g).to(dev)
    y = torch.randn(n, b, generator=g).to(dev)
    G = torch.empty(b, b, dtype=torch.float32, device=dev)
    wsb = _cabi.load().b200d_gram_workspace_bytes(n, b)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    _cabi.call("b200d_gram", ptr(x), ptr(y), n, b, b, ptr(G), ptr(ws), wsb, _cabi._stream())
    ref = (x.double().t() @ y.double())
    assert (G.double() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item() + 1e-4
    # CholQR: Q from X^T X makes X Q orthonormal
    _cabi.call("b200d_gram", ptr(x), ptr(x), n, b, b, ptr(G), ptr(ws), wsb, _cabi._stream())
    Q = torch.empty(b, b, dtype=torch.float32, device=dev)
    _cabi.call("b200d_small_eig", ptr(G), b, None, ptr(Q), 1, _cabi._stream())
    xo = torch.empty_like(x)
    ldvt = (n + 7) // 8 * 8
    vt = torch.zeros(cl.cheb_operand_rows(b), ldvt, dtype=torch.bfloat16, device=dev)
    _cabi.call("b200d_right_mul", ptr(x), n, b, b, ptr(Q), ptr(xo), ptr(vt), ldvt, _cabi._stream())
    torch.cuda.synchronize()
    eye = xo.double().t() @ xo.double()
    assert (eye - torch.eye(b, device=dev, dtype=torch.float64)).abs().max().item() < 5e-5
    assert (xo.double() - x.double() @ Q.double()).abs().max().item() < 1e-5
    rec = _unsplit(vt, b, n)
    assert (rec - xo).abs().max().item() <= 1e-7 * xo.abs().max().item() + 1e-9
    # symmetric eigen-decomposition
    S = torch.randn(b, b, generator=g)
    S = (S + S.t()).to(dev).contiguous()
    ev = torch.empty(b, dtype=torch.float32, device=dev)
    V = torch.empty(b, b, dtype=torch.float32, device=dev)
    _cabi.call("b200d_small_eig", ptr(S), b, ptr(ev), ptr(V), 0, _cabi._stream())
    torch.cuda.synchronize()
    ref_ev = torch.linalg.eigvalsh(S.double())
    assert (ev.double() - ref_ev).abs().max().item() < 1e-5
    assert (S.double() @ V.double() - V.double() * ev.double()[None, :]).abs().max().item() < 2e-5
    # residual norms
    theta = torch.rand(b, device=dev)
    out = torch.empty(b, dtype=torch.float32, device=dev)
    _cabi.call("b200d_resid_norms", ptr(y), ptr(x), ptr(theta), n, b, b, ptr(out), ptr(ws), wsb, _cabi._stream())
    ref_r = ((y.double() - x.double() * theta.double()[None, :]) ** 2).sum(0)
    assert ((out.double() - ref_r).abs() / ref_r).max().item() < 1e-4


@pytest.mark.parametrize("n,b,pair", [(300, 32, True), (2399, 32, True), (3000, 64, True), (5000, 64, True), (5000, 64, False),
                                      (10000, 64, False)])
def test_cheb_gemm_step(dev, n, b, pair):
    """One Chebyshev step on the tensor cores (N = 128 for 32 vectors, N = 192 for 64; the CTA-pair kernel from 4096 rows
    unless it is switched off, as inside multi-stream regions) against fp64."""
    import contextlib

    from whisper_nemo_b200 import _cabi
    from whisper_nemo_b200 import clustering as cl
    from whisper_nemo_b200._cabi import ptr

    g = torch.Generator().manual_seed(n)
    lda = (n + 7) // 8 * 8
    a = (torch.rand(n, n, generator=g) < 0.05).float()
    a = 0.5 * (a + a.t())
    a.fill_diagonal_(0)
    a16 = torch.zeros(n, lda, dtype=torch.bfloat16, device=dev)
    a16[:, :n] = a.to(dev).bfloat16()
    deg = a.sum(1).to(dev)
    x = torch.randn(n, b, generator=g).to(dev)
    xp = torch.randn(n, b, generator=g).to(dev)
    ldvt = lda
    nw = cl.cheb_operand_rows(b)
    vt_in = torch.zeros(nw, ldvt, dtype=torch.bfloat16, device=dev)
    vt_out = torch.zeros(nw, ldvt, dtype=torch.bfloat16, device=dev)
    _cabi.call("b200d_right_mul", ptr(x), n, b, b, None, None, ptr(vt_in), ldvt, _cabi._stream())
    out = torch.empty(n, b, dtype=torch.float32, device=dev)
    ca, cb, cc = 0.37, -1.2, 0.6
    with contextlib.nullcontext() if pair else _cabi.single_cta_gemms():
        cl._gemm_cheb(a16, lda, vt_in, ldvt, n, nw, out, deg, x, xp, ca, cb, cc, vt_out)
        torch.cuda.synchronize()
    ad = a.to(dev).double()
    want = ca * (deg.double()[:, None] * x.double() - ad @ x.double()) + cb * x.double() + cc * xp.double()
    err = (out.double() - want).abs().max().item()
    print(f"cheb step n={n} b={b}: max err {err:.3e} (max |want| {want.abs().max().item():.2f})")
    assert err <= 2e-5 * want.abs().max().item()
    rec = _unsplit(vt_out, b, n)
    assert (rec - out).abs().max().item() <= 1e-6 * out.abs().max().item()


@pytest.mark.parametrize("n,b", [(300, 32), (3000, 64), (10000, 64), (10000, 32), (1537, 32)])
def test_cheb_gemm_step_split_k(dev, n, b):
    """The split-K form of the Chebyshev step (what b200d_eig_bottomk uses): (row tile, 24-K-block segment) units + fix-up kernel.
    Against fp64, and ROW-INVARIANT: a launch on any subset of the rows gives those rows bit for bit (the property the row-sharded
    multi-GPU solver relies on), with and without a previous-step operand."""
    from whisper_nemo_b200 import _cabi
    from whisper_nemo_b200 import clustering as cl
    from whisper_nemo_b200._cabi import ptr

    g = torch.Generator().manual_seed(n + b)
    lda = (n + 7) // 8 * 8
    a = (torch.rand(n, n, generator=g) < 0.05).float()
    a = 0.5 * (a + a.t())
    a.fill_diagonal_(0)
    a16 = torch.zeros(n, lda, dtype=torch.bfloat16, device=dev)
    a16[:, :n] = a.to(dev).bfloat16()
    deg = a.sum(1).to(dev)
    x = torch.randn(n, b, generator=g).to(dev)
    xp = torch.randn(n, b, generator=g).to(dev)
    nw = cl.cheb_operand_rows(b)
    vt_in = torch.zeros(nw, lda, dtype=torch.bfloat16, device=dev)
    _cabi.call("b200d_right_mul", ptr(x), n, b, b, None, None, ptr(vt_in), lda, _cabi._stream())
    ws = torch.empty(int(_cabi.load().b200d_gemm_cheb_splitk_bytes(n, nw, n)), dtype=torch.uint8, device=dev)
    ca, cb, cc = 0.37, -1.2, 0.6
    ad = a.to(dev).double()
    for prev in (xp, None):
        out = torch.empty(n, b, dtype=torch.float32, device=dev)
        vt_out = torch.zeros(nw, lda, dtype=torch.bfloat16, device=dev)
        cl._gemm_cheb(a16, lda, vt_in, lda, n, nw, out, deg, x, prev, ca, cb, cc, vt_out, splitk=ws)
        want = ca * (deg.double()[:, None] * x.double() - ad @ x.double()) + cb * x.double() + (cc * xp.double() if prev is not None else 0.0)
        err = (out.double() - want).abs().max().item()
        assert err <= 2e-5 * want.abs().max().item()
        assert (_unsplit(vt_out, b, n) - out).abs().max().item() <= 1e-6 * out.abs().max().item()
        # rows [lo, hi) alone: operands offset exactly as b200d_eig_bottomk_sharded offsets them
        lo, hi = (n // 3) // 8 * 8, n - 5
        part = torch.full((n, b), float("nan"), dtype=torch.float32, device=dev)
        vt_part = torch.zeros(nw, lda, dtype=torch.bfloat16, device=dev)
        cl._gemm_cheb(a16[lo:hi], lda, vt_in, lda, hi - lo, nw, part[lo:hi], deg[lo:hi], x[lo:hi], None if prev is None else prev[lo:hi], ca, cb, cc,
                      vt_part[:, lo:], splitk=ws, k=n)
        torch.cuda.synchronize()
        assert torch.equal(part[lo:hi], out[lo:hi])
        assert torch.equal(vt_part[:, lo:hi], vt_out[:, lo:hi])
        # a short launch (takes the 128-row units where the full one takes the 256-row units): still the same bits
        hi2 = min(n, lo + 300)
        part2 = torch.full((n, b), float("nan"), dtype=torch.float32, device=dev)
        cl._gemm_cheb(a16[lo:hi2], lda, vt_in, lda, hi2 - lo, nw, part2[lo:hi2], deg[lo:hi2], x[lo:hi2], None if prev is None else prev[lo:hi2], ca, cb,
                      cc, None, splitk=ws, k=n)
        torch.cuda.synchronize()
        assert torch.equal(part2[lo:hi2], out[lo:hi2])
    print(f"split-K cheb step n={n} b={b}: max err {err:.3e}")


def _clustered_graph(n, k, p, seed):
    from oracle import offline_clustering as oc

    g = torch.Generator().manual_seed(seed)
    centers = torch.randn(k, 32, generator=g) * 3
    lab = torch.randint(0, k, (n,), generator=g)
    x = centers[lab] + torch.randn(n, 32, generator=g)
    return oc.getCosAffinityMatrix(x), lab


@pytest.mark.parametrize("n,p", [(64, 5), (1000, 11), (2399, 60), (10000, 11), (4001, 2500)])
def test_csr_from_dense_and_sparse_chebyshev_step(dev, n, p):
    """CSR lists of the binarised graph == its non-zeros (ascending columns, 0.5 / 1 flag); one sparse Chebyshev step
    == ca (D x - A x) + cb x + cc xprev in fp64."""
    from whisper_nemo_b200 import _cabi
    from whisper_nemo_b200 import clustering as cl
    from whisper_nemo_b200._cabi import ptr

    g = torch.Generator().manual_seed(n + p)
    x = torch.randn(n, 24, generator=g)
    x = x / x.norm(dim=1, keepdim=True)
    mat = (x @ x.t()).to(dev)
    a16, deg = cl.getAffinityGraphMat(mat, p)
    lda = a16.shape[1]
    capacity = min(2 * p, n) * n
    rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    colw = torch.full((capacity,), -1, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    _cabi.call("b200d_csr_from_dense", ptr(a16), n, lda, ptr(rowptr), ptr(colw), capacity, s)
    torch.cuda.synchronize()
    a = a16[:, :n].float().cpu()
    rp = rowptr.cpu().long()
    cw = colw.cpu().long() & 0xFFFFFFFF
    nz = a.nonzero()
    assert rp[0].item() == 0 and rp[-1].item() == nz.shape[0]
    assert torch.equal(rp[1:] - rp[:-1], (a != 0).sum(1))
    got_cols = cw[: nz.shape[0]] & 0x7FFFFFFF
    got_one = (cw[: nz.shape[0]] >> 31) == 1
    assert torch.equal(got_cols, nz[:, 1])          # nonzero() is row-major: ascending columns inside each row
    assert torch.equal(got_one, a[nz[:, 0], nz[:, 1]] == 1.0)
    for b in (32, 64):
        xx = torch.randn(n, b, generator=g)
        xp = torch.randn(n, b, generator=g)
        ca, cb, cc = 0.37, -1.25, 0.6
        for prev in (None, xp):
            out = torch.empty(n, b, device=dev)
            xd, pd = xx.to(dev), (prev.to(dev) if prev is not None else None)
            _cabi.call("b200d_spmm_cheb", ptr(rowptr), ptr(colw), n, b, ptr(deg), ptr(xd), ptr(pd), b, ca, cb, cc, ptr(out), b, s)
            torch.cuda.synchronize()
            d = deg.cpu().double()[:, None]
            ref = ca * (d * xx.double() - a.double() @ xx.double()) + cb * xx.double()
            if prev is not None:
                ref = ref + cc * prev.double()
            err = (out.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
            assert err < 2e-6, (n, p, b, err)


@functools.lru_cache(maxsize=2)
def _reference_bottom_eigvecs(n, k, p):
    from oracle import offline_clustering as oc

    mat, _ = _clustered_graph(n, k, p, seed=n + k)
    lap = oc.getLaplacian(oc.getAffinityGraphMat(mat, p)).double()
    lam, vec = torch.linalg.eigh(lap)
    return lam[: k + 1].clone(), vec[:, :k].clone()


@pytest.mark.parametrize("products", ["dense", "csr"])
@pytest.mark.parametrize("n,k,p", [(40, 3, 6), (64, 2, 8), (97, 8, 10), (128, 30, 12), (129, 30, 12), (150, 3, 12), (200, 25, 15),
                                   (2399, 4, 60), (6000, 8, 200), (4000, 50, 80), (6000, 50, 11)])
def test_spectral_embedding_subspace(dev, n, k, p, products, monkeypatch):
    """The k lowest eigenvectors span the same subspace as torch.linalg.eigh's (principal angles ~ 0), with the dense
    tcgen05 products and with the CSR row-gather products."""
    from oracle import offline_clustering as oc
    from whisper_nemo_b200 import clustering as cl

    if products == "csr" and (n <= cl.DENSE_EIG_MAX or 2 * p > 200):
        pytest.skip("Jacobi path / graph too dense for the CSR products")
    monkeypatch.setattr(cl, "SPARSE_MAX_ROW_NNZ", -1 if products == "dense" else 200)
    monkeypatch.setattr(cl, "SPARSE_MAX_DENSITY", -1.0)
    mat, _ = _clustered_graph(n, k, p, seed=n + k)
    graph = cl.getAffinityGraphMat(mat.to(dev), p)
    emb = cl.SpectralClustering(n_clusters=k).getSpectralEmbeddings(graph)
    torch.cuda.synchronize()
    st = cl.last_spectral_stats
    lam, ref = _reference_bottom_eigvecs(n, k, p)
    q, _ = torch.linalg.qr(emb.cpu().double())
    sv = torch.linalg.svdvals(ref.t() @ q)
    gap = (lam[k] - lam[k - 1]).item()
    print(f"spectral n={n} k={k}: method {st.method} outer {st.outer} gemms {st.gemms} resid {st.max_resid:.2e} "
          f"min cos(angle) {sv.min().item():.8f} gap {gap:.3e}")
    assert st.converged
    assert sv.min().item() > 1 - 1e-4
    if n > cl.DENSE_EIG_LIMIT:
        assert st.method.endswith("-csr") == (products == "csr")


@pytest.mark.parametrize("n,dim,k", [(50, 2, 2), (500, 4, 4), (2399, 3, 3), (14399, 8, 8), (10000, 50, 50)])
def test_kmeans_matches_oracle(dev, n, dim, k):
    from oracle import offline_clustering as oc
    from whisper_nemo_b200 import clustering as cl

    g = torch.Generator().manual_seed(n + dim)
    centers = torch.randn(k, dim, generator=g) * 4
    lab = torch.randint(0, k, (n,), generator=g)
    x = (centers[lab] + torch.randn(n, dim, generator=g)) / math.sqrt(n)
    state = torch.get_rng_state()
    want = oc.kmeans_torch(x.clone(), k)
    torch.set_rng_state(state)
    got = cl.kmeans_torch(x.to(dev), k).cpu()
    agree = (got == want).float().mean().item()
    print(f"kmeans n={n} dim={dim} k={k}: label agreement {agree:.6f}")
    assert agree == 1.0


def test_multiscale_affinity_matches_oracle(dev):
    from oracle import offline_clustering as oc
    from whisper_nemo_b200 import clustering as cl

    g = torch.Generator().manual_seed(3)
    dur = 400.0
    scales = [(1.5, 0.75), (1.0, 0.5), (0.5, 0.25)]
    embs, stamps = [], []
    for w, s in scales:
        n = int(math.ceil((dur - w) / s)) + 1
        t0 = torch.arange(n, dtype=torch.float64) * s
        ts = torch.stack([t0, torch.clamp(t0 + w, max=dur)], 1)
        spk = ((t0 + w / 2) // 20).long() % 3
        e = torch.randn(3, 192, generator=g)[spk] * 2 + torch.randn(n, 192, generator=g)
        embs.append(e)
        stamps.append(ts.float())
    weights = torch.tensor([[1.0, 1.0, 1.0]])
    want = oc.getMultiScaleCosAffinityMatrix(weights, embs, stamps)
    got = cl.getMultiScaleCosAffinityMatrix(weights, [e.to(dev) for e in embs], stamps)
    torch.cuda.synchronize()
    want_map = oc.get_argmin_mat(stamps)
    got_map = cl.get_argmin_mat(stamps)
    for a, b in zip(want_map, got_map):
        assert np.array_equal(a.numpy(), b)
    err = (got.cpu() - want).abs().max().item()
    print(f"fused affinity [0, {weights.sum().item():.0f}] range: max abs err {err:.3e}")
    assert err <= 1e-4  # BASELINE.json: fused affinity within 1e-4 absolute (range [0, sum w] = [0, 3])


@pytest.mark.parametrize("M,N,K", [(300, 256, 128), (1000, 1024, 1024), (20011, 1024, 1024), (19000, 3072, 128), (513, 384, 3072)])
def test_gemm_epilogues_match_torch(dev, M, N, K):
    """tcgen05 GEMM (1-CTA and CTA-pair kernels; the pair kernel takes the shapes with >= 74 tiles of 256 x 256) against
    torch fp32 on the same fp16 operands: every fused epilogue, ragged M."""
    from whisper_nemo_b200 import _cabi
    from whisper_nemo_b200 import titanet as tn

    g = torch.Generator().manual_seed(M + N)
    T = 151
    A = (torch.randn(M, K, generator=g) * 0.5).half().to(dev)
    W = (torch.randn(N, K, generator=g) * (1.0 / K ** 0.5)).half().to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    ref = A.float() @ W.float().t()
    n_seg = (M + T - 1) // T
    out = torch.empty(M, N, dtype=torch.float16, device=dev)
    tol = 2e-3 * max(1.0, ref.abs().max().item())  # fp16 output rounding
    tn.gemm(A, W, out, _cabi.EPI_BIAS, bias=bias)
    assert (out.float() - (ref + bias)).abs().max().item() <= tol
    tn.gemm(A, W, out, _cabi.EPI_BIAS_RELU, bias=bias)
    assert (out.float() - torch.relu(ref + bias)).abs().max().item() <= tol
    aux = torch.randn(M, N, generator=g).half().to(dev)
    gate = torch.rand(n_seg, N, generator=g).to(dev)
    tn.gemm(A, W, out, _cabi.EPI_SE_RES, bias=bias, rowvec=gate, aux16=aux, rows_per_seg=T)
    want = torch.relu(aux.float() * gate.repeat_interleave(T, 0)[:M] + ref + bias)
    assert (out.float() - want).abs().max().item() <= tol
    out32 = torch.empty(M, N, dtype=torch.float32, device=dev)
    tn.gemm(A, W, out32, _cabi.EPI_BIAS_F32, bias=bias)
    assert (out32 - (ref + bias)).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
    tn.gemm(A, W, out32, _cabi.EPI_SIGMOID_F32)
    assert (out32 - torch.sigmoid(ref)).abs().max().item() <= 1e-5
    if N % 128 == 0 and N <= 256:
        scale, shift = (torch.rand(N, generator=g) + 0.5).to(dev), (torch.randn(N, generator=g) * 0.1).to(dev)
        rb = torch.randn(n_seg, N, generator=g).to(dev)
        tn.gemm(A, W, out, _cabi.EPI_TDNN, scale=scale, shift=shift, rowvec=rb, rows_per_seg=T)
        want = torch.tanh(scale * torch.relu(ref + rb.repeat_interleave(T, 0)[:M]) + shift)
        assert (out.float() - want).abs().max().item() <= 2e-3


@pytest.mark.parametrize("n,k,p", [(2399, 4, 60), (4000, 50, 80), (6000, 50, 11)])
def test_eig_bottomk_composite_equals_the_stepwise_iteration(dev, n, k, p):
    """b200d_eig_bottomk (the whole Chebyshev-filtered subspace iteration inside the library) reproduces the same iteration
    driven call by call from Python bit for bit: same kernels, same host decisions (polynomial degree, stopping test)."""
    from whisper_nemo_b200 import clustering as cl

    mat, _ = _clustered_graph(n, k, p, seed=n + k)
    a16, deg = cl.getAffinityGraphMat(mat.to(dev), p)
    got = cl.bottom_eigvecs(a16, deg, k, p=p)
    st = cl.SpectralStats()
    st.__dict__.update(cl.last_spectral_stats.__dict__)
    want = cl.bottom_eigvecs(a16, deg, k, p=p, stepwise=True)
    st2 = cl.last_spectral_stats
    torch.cuda.synchronize()
    print(f"eig_bottomk n={n} k={k} p={p}: {st.method} outer {st.outer} gemms {st.gemms} resid {st.max_resid:.2e} | stepwise outer {st2.outer} gemms {st2.gemms}")
    assert st.converged and (st.method, st.outer, st.gemms) == (st2.method, st2.outer, st2.gemms)
    assert torch.equal(got, want)


@pytest.mark.parametrize("fixed_len,n,max_frames", [(24000, 12, 8192), (8000, 300, 4096), (48000, 9, 1024), (20000, 7, 131072)])
def test_titanet_forward_composite_equals_the_stepwise_orchestration(dev, weights, fixed_len, n, max_frames):
    """b200d_titanet_forward (featurizer + encoder + decoder launched from C++ on the library's packed blob, windows processed
    in groups that fit the workspace) against the step-by-step Python orchestration over the fine-grained entry points and
    the Python-packed operands: identical embeddings."""
    from tools import workload as synth
    from whisper_nemo_b200 import titanet as tn

    wav, _ = synth.synth_recording(60.0, 3, seed=7)
    wav_d = torch.from_numpy(wav).to(dev)
    net = tn.TitaNetB200(weights, dev, max_frames=max_frames)
    starts = torch.tensor([500 + 2960 * i for i in range(n)], dtype=torch.int32, device=dev)
    lens = torch.full((n,), fixed_len, dtype=torch.int32, device=dev)
    lens[-1] = fixed_len // 3
    got = net.embed_segments(wav_d, starts, lens, fixed_len)
    taps = {}
    want = net.embed_segments(wav_d, starts, lens, fixed_len, taps=taps)
    torch.cuda.synchronize()
    assert "encoder" in taps
    assert torch.equal(got, want)


def test_plain_c_host_drives_the_clustering_half(dev, tmp_path):
    """tools/cabi_host_example.c: a host that is not Python (plain C, include/b200d.h + the CUDA runtime, its own cudaMalloc'ed
    buffers) runs affinity -> fusion -> top-p graph -> b200d_eig_bottomk -> k-means through the C ABI and recovers planted
    clusters exactly (the boundary is the C ABI, not the Python wrapper)."""
    import os
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no gcc on this host")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    lib_dir = os.path.join(root, "whisper_nemo_b200")
    exe = tmp_path / "cabi_host"
    subprocess.check_call(["gcc", "-O2", "-I", os.path.join(root, "include"), "-I", os.path.join(cuda, "include"),
                           os.path.join(root, "tools", "cabi_host_example.c"), "-o", str(exe), "-L", lib_dir, "-lb200d",
                           "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm", f"-Wl,-rpath,{lib_dir}", f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}"])
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    print(run.stdout.strip())
    assert run.returncode == 0, run.stdout + run.stderr
    assert "labels match the planted clusters on 3000 of 3000 points" in run.stdout
