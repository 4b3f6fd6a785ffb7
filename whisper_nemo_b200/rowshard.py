"""One long recording on the GPUs of one NVSwitch box: the N x N affinity, its p-neighbour graph and the products of the
iterative eigensolver ROW-SHARDED over the ranks ("Affinity / fusion / binarize -- by rows", "Final
spectral embedding -- row-sharded A V products"; BASELINE.json north_star: "for long recordings the affinity is row-sharded
after an NCCL all-gather of embeddings over NVLink").  The reference has no multi-GPU code; upstream's functions this
distributes are offline_clustering.getMultiScaleCosAffinityMatrix / NMESC / getAffinityGraphMat / SpectralClustering.

Every rank holds all embeddings (all-gathered by the diarizer) and rows [lo, hi) of everything quadratic:

  per-scale cosine rows  b200d_cos_affinity_rows      + all-reduce of the per-scale (min, max)            [2 S floats]
  fused affinity rows    b200d_fuse_scales_rows
  NME sweep              the strided subsample's rows gathered from their owners (<= 1024^2 floats); the p values of the sweep
                         dealt to the ranks (b200d_eigvals_batched_layout: the single-GPU reduction order) + all-gather of
                         the eigenvalues [~10 floats per p]
  top-p binarisation     b200d_topp_select_rows       + all-gather of each row's threshold / tie cut-off   [8 B per row]
                         b200d_sym_combine_rows (transposed term from the symmetric entry, no all-to-all)
                         + all-gather of the degrees                                                       [4 B per row]
  spectral embedding     b200d_eig_bottomk_sharded: every Chebyshev product runs on the local rows and its epilogue stores
                         the result straight into every peer's buffer over NVLink; device-side flag barrier per product
  k-means                replicated (N x k floats)

All of it is bit-for-bit the single-GPU arithmetic (see include/b200d.h), so the labels equal the single-GPU labels -- with one
caveat: the sharded solver always takes dense tcgen05 products, while one GPU switches to fp32 CSR gathers for very sparse graphs
(at most 32 neighbours or 1/64 of a row; clustering._use_csr_products), where the two agree to fp32 rounding instead of bit for bit.
The small collectives go through a `comm` object: DistComm (torch.distributed, NCCL on the GPUs) or LocalComm (several
"ranks" as host threads with a stream each on ONE GPU -- how the single-GPU test suite exercises this path).
"""
import ctypes
import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi
from ._cabi import ptr

ROW_ALIGN = 128        # shard boundaries on GEMM row tiles
MIN_ROWS_TO_SHARD = 4096


def row_shards(n: int, world: int, align: int = ROW_ALIGN) -> List[Tuple[int, int]]:
    """[lo, hi) of every rank: equal multiples of `align` rows, the last shard takes the remainder; the alignment is relaxed
    (128 -> 32 -> 8 -> 1) when rounding up would leave a rank without rows."""
    while True:
        per = -(-(-(-n // world)) // align) * align
        if per * (world - 1) < n or align == 1:
            return [(min(n, r * per), min(n, (r + 1) * per)) for r in range(world)]
        align = max(1, align // 4)


# ----------------------------------------------------------------------------- peer buffers
class PeerBuffers:
    """The peer buffers of a group of ranks (include/b200d.h b200d_peer_group) as seen from one rank."""

    def __init__(self, group: _cabi.PeerGroup, owned: int, opened: Sequence[int]):
        self.group, self._owned, self._opened = group, owned, list(opened)

    @property
    def nbytes(self) -> int:
        return int(self.group.bytes)

    def close(self):
        lib = _cabi.load()
        torch.cuda.synchronize()
        for p in self._opened:
            _cabi.check(lib.b200d_peer_close(ctypes.c_void_p(p)), "b200d_peer_close")
        if self._owned:
            _cabi.check(lib.b200d_peer_free(ctypes.c_void_p(self._owned)), "b200d_peer_free")
        self._opened, self._owned = [], 0


def _alloc(nbytes: int):
    handle = (ctypes.c_ubyte * _cabi.PEER_HANDLE_BYTES)()
    p = ctypes.c_void_p()
    _cabi.check(_cabi.load().b200d_peer_alloc(nbytes, ctypes.byref(p), handle), "b200d_peer_alloc")
    return int(p.value), bytes(handle)


class DistComm:
    """torch.distributed ranks, one process per GPU (NCCL).  Collectives here are the small ones (thresholds, degrees, NME rows)."""

    def __init__(self):
        import torch.distributed as dist

        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self._peers: Optional[PeerBuffers] = None

    def peers(self, nbytes: int) -> PeerBuffers:
        """Peer buffers of at least nbytes on every rank (collective: every rank asks for the same size)."""
        if self._peers is not None and self._peers.nbytes >= nbytes:
            return self._peers
        if self._peers is not None:
            self.dist.barrier()
            self._peers.close()
        nbytes = max(int(nbytes * 1.25), 1 << 20)
        own, handle = _alloc(nbytes)
        handles: List[Optional[bytes]] = [None] * self.world
        self.dist.all_gather_object(handles, handle)
        grp = _cabi.PeerGroup()
        grp.rank, grp.world, grp.bytes, grp.epoch, grp.timeout_ms = self.rank, self.world, nbytes, 0, 0
        opened = []
        for r in range(self.world):
            if r == self.rank:
                grp.base[r] = own
                continue
            q = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * _cabi.PEER_HANDLE_BYTES).from_buffer_copy(handles[r])
            _cabi.check(_cabi.load().b200d_peer_open(buf, ctypes.byref(q)), "b200d_peer_open")
            grp.base[r] = q.value
            opened.append(int(q.value))
        self.dist.barrier()
        self._peers = PeerBuffers(grp, own, opened)
        return self._peers

    def all_gather_rows(self, local: torch.Tensor, counts: Sequence[int]) -> torch.Tensor:
        """Concatenate per-rank row blocks (rank r contributes counts[r] rows) in rank order on every rank."""
        pad = max(max(counts), 1)
        buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        buf[: local.shape[0]] = local
        out = torch.empty((self.world * pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        self.dist.all_gather_into_tensor(out, buf)
        return torch.cat([out[r * pad : r * pad + counts[r]] for r in range(self.world)], dim=0)

    def all_reduce_minmax(self, mm: torch.Tensor) -> torch.Tensor:
        """mm float32 [S, 2] = per-scale (min, max) over this rank's rows -> over all rows."""
        t = torch.stack([-mm[:, 0], mm[:, 1]], dim=1).contiguous()
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return torch.stack([-t[:, 0], t[:, 1]], dim=1).contiguous()


class LocalComm:
    """`world` ranks as host threads of ONE process, each on its own stream of the same GPU: the row-sharded path with real
    peer stores and device-side barriers, runnable (and tested) on a single GPU.  Build with LocalComm.make(world)."""

    class _Shared:
        def __init__(self, world):
            self.world = world
            self.barrier = threading.Barrier(world)
            self.slots = [None] * world
            self.bases: List[int] = []
            self.nbytes = 0

    def __init__(self, shared, rank):
        self.shared, self.rank, self.world = shared, rank, shared.world
        self._group: Optional[_cabi.PeerGroup] = None

    @staticmethod
    def make(world: int) -> List["LocalComm"]:
        sh = LocalComm._Shared(world)
        return [LocalComm(sh, r) for r in range(world)]

    def peers(self, nbytes: int) -> PeerBuffers:
        sh = self.shared
        if self._group is not None and sh.nbytes >= nbytes:
            return PeerBuffers(self._group, 0, [])
        sh.barrier.wait()
        if self.rank == 0:
            torch.cuda.synchronize()
            for b in sh.bases:
                _cabi.check(_cabi.load().b200d_peer_free(ctypes.c_void_p(b)), "b200d_peer_free")
            sh.nbytes = max(int(nbytes * 1.25), 1 << 20)
            sh.bases = [_alloc(sh.nbytes)[0] for _ in range(self.world)]
        sh.barrier.wait()
        grp = _cabi.PeerGroup()
        grp.rank, grp.world, grp.bytes, grp.epoch, grp.timeout_ms = self.rank, self.world, sh.nbytes, 0, 0
        for r in range(self.world):
            grp.base[r] = sh.bases[r]
        self._group = grp
        return PeerBuffers(grp, 0, [])

    def _exchange(self, value, combine):
        """Every thread publishes `value` (its stream drained first), combines all of them on its own stream and drains that
        too before anyone may release or overwrite what it published."""
        sh = self.shared
        torch.cuda.current_stream().synchronize()
        sh.slots[self.rank] = value
        sh.barrier.wait()
        out = combine(list(sh.slots))
        torch.cuda.current_stream().synchronize()
        sh.barrier.wait()
        return out

    def all_gather_rows(self, local: torch.Tensor, counts: Sequence[int]) -> torch.Tensor:
        return self._exchange(local, lambda parts: torch.cat([t[:c] for t, c in zip(parts, counts)], dim=0))

    def all_reduce_minmax(self, mm: torch.Tensor) -> torch.Tensor:
        def combine(parts):
            st = torch.stack(parts)
            return torch.stack([st[:, :, 0].min(dim=0).values, st[:, :, 1].max(dim=0).values], dim=1).contiguous()

        return self._exchange(mm, combine)


# ----------------------------------------------------------------------------- the sharded stages
def _stream():
    return _cabi._stream()


def fused_affinity_rows(comm, weights, embeddings_in_scales, mapping, lo: int, hi: int) -> torch.Tensor:
    """Rows [lo, hi) of getMultiScaleCosAffinityMatrix (float32 [hi - lo, N_base]); every rank passes the same full inputs."""
    from .clustering import COS_EPS

    dev = embeddings_in_scales[0].device
    n_base = int(mapping[-1].shape[0])
    S = len(embeddings_in_scales)
    cos_list, mm_list, map_list, row0 = [], [], [], []
    for emb, m in zip(embeddings_in_scales, mapping):
        ms = np.sort(m).astype(np.int32)
        emb = emb.float().contiguous()
        ns, d = emb.shape
        r0, r1 = int(ms[lo]), int(ms[hi - 1]) + 1
        mm = torch.empty(2, dtype=torch.float32, device=dev)
        if ns == 1:  # cos_affinity's single-point case: the 1 x 1 matrix [[1]], (min, max) = (0, 1)
            cos, r0 = torch.ones(1, 1, device=dev), 0
            mm.copy_(torch.tensor([0.0, 1.0]))
        else:
            xn = torch.empty_like(emb)
            _cabi.call("b200d_l2_normalize", ptr(emb), ptr(xn), ns, d, COS_EPS, _stream())
            cos = torch.empty(r1 - r0, ns, dtype=torch.float32, device=dev)
            _cabi.call("b200d_cos_affinity_rows", ptr(xn), ns, d, r0, r1, ptr(cos), ptr(mm), _stream())
        cos_list.append(cos)
        mm_list.append(mm)
        row0.append(r0)
        map_list.append(torch.from_numpy(ms).to(dev))
    mm_all = comm.all_reduce_minmax(torch.stack(mm_list))
    # a scale with one point keeps cos_affinity's (0, 1)
    fused = torch.empty(hi - lo, n_base, dtype=torch.float32, device=dev)
    cos_p = (ctypes.c_void_p * S)(*[c.data_ptr() for c in cos_list])
    map_p = (ctypes.c_void_p * S)(*[m.data_ptr() for m in map_list])
    mm_rows = [mm_all[s].contiguous() for s in range(S)]
    mm_p = (ctypes.c_void_p * S)(*[m.data_ptr() for m in mm_rows])
    ns_a = (ctypes.c_int32 * S)(*[int(e.shape[0]) for e in embeddings_in_scales])
    r0_a = (ctypes.c_int32 * S)(*row0)
    w = (ctypes.c_float * S)(*[float(x) for x in weights])
    _cabi.call("b200d_fuse_scales_rows", S, cos_p, ns_a, r0_a, map_p, mm_p, w, ptr(fused), n_base, lo, hi, _stream())
    return fused


def graph_rows(comm, fused_rows: torch.Tensor, p_value: int, lo: int, hi: int, shards):
    """Rows [lo, hi) of getAffinityGraphMat (bf16 [hi - lo, lda]) and the full degree vector (float32 [N])."""
    m, n = fused_rows.shape
    dev = fused_rows.device
    p_value = min(int(p_value), n)
    if p_value <= 0:
        raise ValueError("p_value must be positive on the device path")
    lda = (n + 7) // 8 * 8
    sel = torch.empty(m, n, dtype=torch.uint8, device=dev)
    tc = torch.empty(2, m, dtype=torch.int32, device=dev)  # [threshold code ; tie cut-off column] of the local rows
    _cabi.call("b200d_topp_select_rows", ptr(fused_rows), m, n, p_value, ptr(sel), ptr(tc[0]), ptr(tc[1]), _stream())
    counts = [h - l for l, h in shards]
    tc_all = comm.all_gather_rows(tc.t().contiguous(), counts).t().contiguous()  # [2, N]
    a_rows = torch.empty(m, lda, dtype=torch.bfloat16, device=dev)
    deg_rows = torch.empty(m, dtype=torch.float32, device=dev)
    _cabi.call("b200d_sym_combine_rows", ptr(fused_rows), ptr(sel), ptr(tc_all[0]), ptr(tc_all[1]), lo, m, n, ptr(a_rows), lda, ptr(deg_rows), _stream())
    deg = comm.all_gather_rows(deg_rows.unsqueeze(1), counts).squeeze(1).contiguous()
    return a_rows, deg


def bottom_eigvecs_sharded(comm, a_rows: torch.Tensor, deg: torch.Tensor, k: int, lo: int, hi: int, tol: float = 2e-6, max_outer: int = 40,
                           seed: int = 0) -> torch.Tensor:
    """clustering.bottom_eigvecs on the row-sharded graph (b200d_eig_bottomk_sharded); same start block on every rank."""
    from . import clustering as cl

    n = int(deg.shape[0])
    lda = int(a_rows.shape[1])
    dev = deg.device
    lib = _cabi.load()
    b = int(lib.b200d_eig_bottomk_block(int(k)))
    if b == 0 or n < 2 * b:
        raise NotImplementedError(f"row-sharded spectral embedding of {k} clusters on {n} points")
    peers = comm.peers(int(lib.b200d_eig_bottomk_sharded_peer_bytes(n, int(k))))
    opt = _cabi.EigOptions(tol=float(tol), max_outer=int(max_outer), gemm_flags=_cabi.gemm_flags(), sparse_max_row_nnz=0, sparse_max_density=0.0)
    stats = _cabi.EigStats()
    X = torch.empty(n, b, dtype=torch.float32, device=dev)
    gen = torch.Generator(device="cpu").manual_seed(1000 + seed)
    X.copy_(torch.randn(n, b, generator=gen))
    _cabi.call("b200d_eig_bottomk_sharded", ptr(a_rows) if hi > lo else None, lda, ptr(deg), n, lo, hi, int(k), ptr(X), b, ctypes.byref(opt),
               ctypes.byref(stats), ctypes.byref(peers.group), _stream())
    st = cl.SpectralStats()
    st.outer, st.gemms, st.n, st.max_resid, st.converged = stats.outer, stats.gemms, n, float(stats.max_resid), bool(stats.converged)
    st.method = f"chfsi{b}-rows{comm.world}"
    cl.last_spectral_stats.__dict__.update(st.__dict__)
    if len(cl.spectral_log) < 256:
        cl.spectral_log.append({"n": n, "k": k, "p": None, "method": st.method, "block": b, "outer": st.outer, "gemms": st.gemms, "resid": st.max_resid,
                                "converged": st.converged, "history": [float(f"{stats.history[i]:.2e}") for i in range(min(stats.outer, _cabi.EIG_HISTORY))]})
    return X[:, :k].contiguous()


def forward_infer_rows(sc, comm, embeddings_in_scales: List[torch.Tensor], timestamps_in_scales: List[torch.Tensor], multiscale_weights,
                       oracle_num_speakers: int, max_rp_threshold: float, max_num_speakers: int, sparse_search_volume: int, fixed_thres: float,
                       scale_mapping=None) -> torch.Tensor:
    """SpeakerClustering.forward_infer + forward_unit_infer for one long recording with everything quadratic row-sharded over
    comm's ranks (`sc`: the SpeakerClustering whose knobs / debug dict this fills).  Returns int64 labels [N] on every rank."""
    from . import clustering as cl

    weights = torch.as_tensor(multiscale_weights).reshape(-1).tolist()
    mapping = scale_mapping if scale_mapping is not None else cl.get_argmin_mat(timestamps_in_scales)
    n = int(timestamps_in_scales[-1].shape[0])
    shards = row_shards(n, comm.world)
    lo, hi = shards[comm.rank]
    if any(h <= l for l, h in shards):
        raise ValueError(f"row sharding {n} windows over {comm.world} ranks leaves a rank without rows")
    fused = fused_affinity_rows(comm, weights, embeddings_in_scales, mapping, lo, hi)
    sc.fused_affinity = None
    # NME sweep on the strided subsample: its rows come from their owners, the sweep itself (<= 1024 x 1024) runs replicated
    ratio = max(1, int(n / sc.nme_mat_size))
    sub_idx = np.arange(0, n, ratio)
    counts = [int(((sub_idx >= l) & (sub_idx < h)).sum()) for l, h in shards]
    mine = torch.from_numpy(sub_idx[(sub_idx >= lo) & (sub_idx < hi)] - lo).to(fused.device)
    sub_rows = fused.index_select(0, mine)[:, ::ratio].contiguous()
    sub = comm.all_gather_rows(sub_rows, counts).contiguous()
    nmesc = cl.NMESC(sub, max_num_speakers=max_num_speakers if oracle_num_speakers <= 0 else oracle_num_speakers, max_rp_threshold=max_rp_threshold,
                     sparse_search=sc.sparse_search, sparse_search_volume=sparse_search_volume, fixed_thres=fixed_thres,
                     nme_mat_size=sc.nme_mat_size, maj_vote_spk_count=sc.maj_vote_spk_count, presampled_ratio=ratio, comm=comm)
    est_num_of_spk, p_hat_value = nmesc.forward()
    n_clusters = int(oracle_num_speakers) if oracle_num_speakers > 0 else int(est_num_of_spk)
    sc.debug = {"est_num_of_spk": int(est_num_of_spk), "p_hat": int(p_hat_value), "n_clusters": n_clusters,
                "g_p": getattr(nmesc, "eig_ratio_list", None), "p_list": list(nmesc.p_value_list), "row_sharded": comm.world}
    a_rows, deg = graph_rows(comm, fused, p_hat_value, lo, hi, shards)
    del fused
    emb = bottom_eigvecs_sharded(comm, a_rows, deg, n_clusters, lo, hi)
    return cl.kmeans_torch(emb, n_clusters, random_state=0)
