"""Row-sharded clustering of one long recording (whisper_nemo_b200/rowshard.py) against the single-GPU path, on ONE GPU:
rowshard.LocalComm runs `world` ranks as host threads with a stream and a peer buffer each, so the peer stores of the GEMM
epilogue and the device-side barriers are the real ones (tools/multi_gpu_check.py repeats the comparison across processes and
GPUs).  Everything must be bit-for-bit the single-GPU arithmetic."""
import threading

import numpy as np
import pytest
import torch

from tests.util import synthetic_multiscale_embeddings

pytestmark = pytest.mark.gpu


def run_ranks(world, fn):
    """fn(comm) on `world` threads (a LocalComm and a CUDA stream each); returns the per-rank results."""
    from whisper_nemo_b200 import rowshard

    comms = rowshard.LocalComm.make(world)
    streams = [torch.cuda.Stream() for _ in range(world)]
    results, errors = [None] * world, []
    torch.cuda.synchronize()

    def work(r):
        torch.cuda.set_device(0)
        try:
            with torch.cuda.stream(streams[r]), torch.no_grad():
                results[r] = fn(comms[r])
                streams[r].synchronize()
        except BaseException as exc:  # noqa: BLE001 -- re-raised below; release the ranks waiting on this one
            errors.append(exc)
            comms[r].shared.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    real = [e for e in errors if not isinstance(e, threading.BrokenBarrierError)]
    if real or errors:
        raise (real or errors)[0]
    return results


def _multiscale(dev, duration, scales, k, seed):
    from whisper_nemo_b200 import clustering as cl

    emb, ts, counts, _ = synthetic_multiscale_embeddings(duration, scales, k, seed)
    embs, stamps = cl.split_input_data(emb.to(dev), ts, counts)
    return embs, stamps, cl.get_argmin_mat(stamps)


@pytest.mark.parametrize("world", [2, 3])
def test_affinity_and_graph_rows_equal_full_matrix(dev, world):
    """Fused affinity rows, threshold-based symmetric binarisation and degrees == the full-matrix kernels, bit for bit
    (1 200 s on three scales: N = 4 799; ties included through duplicated embeddings)."""
    from whisper_nemo_b200 import clustering as cl
    from whisper_nemo_b200 import rowshard

    embs, stamps, mapping = _multiscale(dev, 1200.0, [(1.5, 0.75), (1.0, 0.5), (0.5, 0.25)], 4, seed=3)
    embs = [e.clone() for e in embs]
    embs[-1][100:110] = embs[-1][100]  # exact ties inside rows: the "lower column first" rule decides
    weights = torch.tensor([1.0, 1.0, 1.0])
    full = cl.getMultiScaleCosAffinityMatrix(weights, embs, stamps, mapping)
    assert torch.equal(full, full.t())  # the symmetry sym_combine_rows relies on
    n = full.shape[0]
    graphs = {p: cl.getAffinityGraphMat(full, p) for p in (7, 150)}
    shards = rowshard.row_shards(n, world)

    def rank_fn(comm):
        lo, hi = shards[comm.rank]
        rows = rowshard.fused_affinity_rows(comm, weights.tolist(), embs, mapping, lo, hi)
        out = {"fused": rows.clone()}
        for p in graphs:
            out[p] = rowshard.graph_rows(comm, rows, p, lo, hi, shards)
        return out

    res = run_ranks(world, rank_fn)
    for r, (lo, hi) in enumerate(shards):
        assert torch.equal(res[r]["fused"], full[lo:hi])
        for p, (a16, deg) in graphs.items():
            a_rows, deg_all = res[r][p]
            assert torch.equal(a_rows, a16[lo:hi])
            assert torch.equal(deg_all, deg)


@pytest.mark.parametrize("world,k", [(2, 4), (3, 30)])
def test_sharded_eigensolver_equals_single_gpu(dev, world, k, monkeypatch):
    """b200d_eig_bottomk_sharded (peer stores from the GEMM epilogue + flag barriers) == b200d_eig_bottomk with dense products:
    the same Ritz block on every rank, bit for bit (32- and 64-vector blocks)."""
    from whisper_nemo_b200 import clustering as cl
    from whisper_nemo_b200 import rowshard

    emb, _, _, _ = synthetic_multiscale_embeddings(1500.0, [(0.5, 0.25)], max(k, 2), seed=5)
    mat = cl.getCosAffinityMatrix(emb.to(dev))
    n = mat.shape[0]
    a16, deg = cl.getAffinityGraphMat(mat, 40)
    monkeypatch.setattr(cl, "SPARSE_MAX_ROW_NNZ", 0)
    monkeypatch.setattr(cl, "SPARSE_MAX_DENSITY", 0.0)
    want = cl.bottom_eigvecs(a16, deg, k)
    want_stats = (cl.last_spectral_stats.outer, cl.last_spectral_stats.gemms)
    shards = rowshard.row_shards(n, world)

    def rank_fn(comm):
        lo, hi = shards[comm.rank]
        x = rowshard.bottom_eigvecs_sharded(comm, a16[lo:hi].contiguous(), deg, k, lo, hi)
        return x, (cl.last_spectral_stats.outer, cl.last_spectral_stats.gemms)

    res = run_ranks(world, rank_fn)
    for x, stats in res:
        assert stats == want_stats
        assert torch.equal(x, want)


def test_eigenvalues_of_a_subset_with_the_full_batch_layout(dev):
    """b200d_eigvals_batched_layout: every third Laplacian of a 30-matrix sweep diagonalised alone gives bit for bit the eigenvalues
    the full batch gives (the property that lets the ranks of a row-sharded recording share the NME sweep)."""
    from whisper_nemo_b200 import _cabi
    from whisper_nemo_b200._cabi import ptr

    g = torch.Generator().manual_seed(4)
    n, batch, n_low = 515, 30, 9
    a = torch.rand(batch, n, n, generator=g)
    a = (a + a.transpose(1, 2)).to(dev).contiguous()
    lib = _cabi.load()

    def run(mats, layout):
        m = mats.clone()
        ws = torch.empty(int(lib.b200d_eigvals_workspace_bytes(m.shape[0], n)), dtype=torch.uint8, device=dev)
        out = torch.empty(m.shape[0], n_low + 1, dtype=torch.float32, device=dev)
        _cabi.call("b200d_eigvals_batched_layout", ptr(m), m.shape[0], layout, n, n_low, ptr(out), ptr(ws), ws.numel(), _cabi._stream())
        torch.cuda.synchronize()
        return out

    full = run(a, batch)
    for r in range(3):
        assert torch.equal(run(a[r::3].contiguous(), batch), full[r::3])
    ref = torch.linalg.eigvalsh(a[:2].double().cpu())
    assert (full[:2, :n_low].double().cpu() - ref[:, :n_low]).abs().max().item() <= 1e-3 * ref.abs().max().item()


@pytest.mark.parametrize("world", [2, 4])
def test_row_sharded_clustering_gives_single_gpu_labels(dev, world):
    """forward_infer with everything quadratic row-sharded: same speaker count, p-hat and labels as SpeakerClustering.forward_infer
    on one GPU (N = 7 199 on five scales)."""
    from whisper_nemo_b200 import clustering as cl
    from whisper_nemo_b200 import rowshard

    scales = [(1.5, 0.75), (1.25, 0.625), (1.0, 0.5), (0.75, 0.375), (0.5, 0.25)]
    emb, ts, counts, truth = synthetic_multiscale_embeddings(1800.0, scales, 5, seed=9)
    weights = torch.ones(1, 5)
    single = cl.SpeakerClustering()
    want = single.forward_infer(emb.to(dev), ts, counts, weights, max_rp_threshold=0.25, max_num_speakers=8, sparse_search_volume=30).cpu()

    def rank_fn(comm):
        sc = cl.SpeakerClustering()
        sc.row_comm = comm
        lab = sc.forward_infer(emb.to(dev), ts, counts, weights, max_rp_threshold=0.25, max_num_speakers=8, sparse_search_volume=30).cpu()
        return lab, dict(sc.debug)

    res = run_ranks(world, rank_fn)
    for lab, debug in res:
        assert debug["row_sharded"] == world
        assert debug["n_clusters"] == single.debug["n_clusters"] == 5 and debug["p_hat"] == single.debug["p_hat"]
        assert torch.equal(lab, want)


def test_peer_barrier_times_out_instead_of_hanging(dev):
    """A rank that never arrives ends the barrier with an error after timeout_ms, not with a hung GPU."""
    import ctypes

    from whisper_nemo_b200 import _cabi, rowshard

    bases = [rowshard._alloc(1 << 20)[0] for _ in range(2)]
    grp = _cabi.PeerGroup()
    grp.rank, grp.world, grp.bytes, grp.epoch, grp.timeout_ms = 0, 2, 1 << 20, 0, 200
    for r in range(2):
        grp.base[r] = bases[r]
    lib = _cabi.load()
    _cabi.check(lib.b200d_peer_barrier(ctypes.byref(grp), _cabi._stream()), "b200d_peer_barrier")
    assert lib.b200d_peer_status(ctypes.byref(grp), _cabi._stream()) != 0
    assert b"timed out" in lib.b200d_last_error()
    torch.cuda.synchronize()
    for b in bases:
        _cabi.check(lib.b200d_peer_free(ctypes.c_void_p(b)), "b200d_peer_free")
