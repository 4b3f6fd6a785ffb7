"""Multi-scale affinity + NME-SC spectral clustering on B200 (host orchestration over libb200d.so).

Mirrors the operator interface of upstream NeMo's
`nemo/collections/asr/parts/utils/offline_clustering.py` -- same function / class names,
argument meaning and error behaviour -- for the code the reference reaches through
`ClusteringDiarizer.diarize()` -> `perform_clustering` (diarize.py:200-201,
nemo_process.py:31-32; knobs: nemo_msdd_configs/diar_infer_*.yaml:47-56).  Every tensor
that upstream keeps on `device` stays in HBM here; the arithmetic runs in the hand-written
kernels of csrc/{affinity,graph,eig,spectral,kmeans,gemm_tcgen05}.cu:

  getMultiScaleCosAffinityMatrix  l2_normalize + cos_affinity (per scale) + fuse_scales (one pass)
  NMESC.forward                   row_rank (once) -> laplacian_from_rank (all p) -> eigvals_batched
  getAffinityGraphMat (final)     topp_binarize (radix select, bf16 {0, .5, 1} graph + fp16-rounded degree)
  getSpectralEmbeddings           Chebyshev-filtered subspace iteration on the tcgen05 GEMM
  kmeans_torch                    kmeans (host supplies the torch.manual_seed(0) draws)

Random numbers (k-means++ draws, anchor embeddings of the enhanced speaker count) come from
the CPU torch generator exactly as upstream draws them; everything else is device work.
There is no CPU fallback: without libb200d.so / an sm_100 device every entry point raises.
"""
import ctypes
import os
import math
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _cabi
from ._cabi import GemmEpilogue, ptr

COS_EPS = 3.5e-4  # offline_clustering.cos_similarity
MIN_SAMPLES_FOR_NMESC = 6
NME_MAT_SIZE = 512
ENHANCED_COUNT_THRES = 80
DENSE_EIG_MAX = 96   # spectral embedding of graphs up to this size: one Jacobi solve instead of subspace iteration
DENSE_EIG_LIMIT = 128  # ... and up to this size when the 64-vector block of the iteration would not fit (n < 128, k > 24)


def _s():
    return _cabi._stream()


def _i32_array(vals):
    return (ctypes.c_int32 * len(vals))(*[int(v) for v in vals])


# ----------------------------------------------------------------------------- affinity
def cos_affinity(emb: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """cos_similarity(emb, emb) with the diagonal forced to 1, plus its global (min, max):
    the two ingredients of getCosAffinityMatrix.  emb float32 [n, d] on device."""
    n, d = emb.shape
    if n == 1:
        dev = emb.device
        return torch.ones(1, 1, device=dev), torch.tensor([0.0, 1.0], device=dev)
    emb = emb.float().contiguous()
    xn = torch.empty_like(emb)
    _cabi.call("b200d_l2_normalize", ptr(emb), ptr(xn), n, d, COS_EPS, _s())
    cos = torch.empty(n, n, dtype=torch.float32, device=emb.device)
    mm = torch.empty(2, dtype=torch.float32, device=emb.device)
    _cabi.call("b200d_cos_affinity", ptr(xn), n, d, ptr(cos), ptr(mm), _s())
    return cos, mm


def getCosAffinityMatrix(emb: torch.Tensor) -> torch.Tensor:
    """ScalerMinMax(cos_similarity(emb, emb)) as one fused-scale pass with an identity mapping."""
    n = emb.shape[0]
    if n == 1:
        return torch.tensor([[1]], device=emb.device)
    cos, mm = cos_affinity(emb)
    ident = torch.arange(n, dtype=torch.int32, device=emb.device)
    return _fuse([cos], [ident], [mm], [1.0], n)


def get_argmin_mat(timestamps_in_scales: List[torch.Tensor]) -> List[np.ndarray]:
    """For every base-scale segment the index of the nearest segment (by centre time, fp32
    arithmetic, first minimum) of each scale.  Upstream materialises an N_base x N_s |diff|
    matrix per scale; segment centres are sorted in time, so the same argmin is a binary
    search plus an exact fp32 comparison of the neighbouring candidates (host, O(N log N))."""
    anchors = [np.asarray(t.detach().cpu().float().numpy()).astype(np.float32) for t in timestamps_in_scales]
    anchors = [((a[:, 0] + a[:, 1]) / np.float32(2.0)).astype(np.float32) for a in anchors]
    base = anchors[-1]
    out = []
    for cur in anchors:
        ns = cur.shape[0]
        if ns > 1 and np.any(np.diff(cur) < 0):
            diff = np.abs(cur[None, :] - base[:, None])  # unsorted centres: upstream's dense form
            out.append(np.argmin(diff, axis=1).astype(np.int32))
            continue
        pos = np.searchsorted(cur, base, side="left")
        best = np.zeros(base.shape[0], dtype=np.int64)
        best_d = np.full(base.shape[0], np.inf, dtype=np.float32)
        for off in (-2, -1, 0, 1, 2):  # ascending index order + strict '<' keeps the first minimum
            j = np.clip(pos + off, 0, ns - 1)
            dj = np.abs(cur[j] - base).astype(np.float32)
            better = (dj < best_d) | ((dj == best_d) & (j < best))
            best = np.where(better, j, best)
            best_d = np.where(better, dj, best_d)
        # equal centres: argmin returns the first of the run
        best = np.searchsorted(cur, cur[best], side="left")
        out.append(best.astype(np.int32))
    return out


def _fuse(cos_list, map_list, mm_list, weights, n_base) -> torch.Tensor:
    S = len(cos_list)
    dev = cos_list[0].device
    fused = torch.empty(n_base, n_base, dtype=torch.float32, device=dev)
    cos_p = (ctypes.c_void_p * S)(*[c.data_ptr() for c in cos_list])
    map_p = (ctypes.c_void_p * S)(*[m.data_ptr() for m in map_list])
    mm_p = (ctypes.c_void_p * S)(*[m.data_ptr() for m in mm_list])
    ns = _i32_array([c.shape[0] for c in cos_list])
    w = (ctypes.c_float * S)(*[float(x) for x in weights])
    _cabi.call("b200d_fuse_scales", S, cos_p, ns, map_p, mm_p, w, ptr(fused), n_base, _s())
    return fused


def getMultiScaleCosAffinityMatrix(multiscale_weights, embeddings_in_scales, timestamps_in_scales, scale_mapping=None) -> torch.Tensor:
    """Fused N_base x N_base affinity, the unnormalised weighted sum of the per-scale min-max
    scaled cosine matrices (range [0, sum w]).  The per-scale N x N expansions of upstream's
    repeat_interleave are never materialised."""
    weights = torch.as_tensor(multiscale_weights).reshape(-1).tolist()
    # `scale_mapping`: get_argmin_mat(timestamps) computed by the caller while the GPU was still embedding
    mapping = scale_mapping if scale_mapping is not None else get_argmin_mat(timestamps_in_scales)
    dev = embeddings_in_scales[0].device
    n_base = int(timestamps_in_scales[-1].shape[0])
    cos_list, mm_list, map_list = [], [], []
    for emb, m in zip(embeddings_in_scales, mapping):
        cos, mm = cos_affinity(emb)
        cos_list.append(cos)
        mm_list.append(mm)
        # getRepeatedList + repeat_interleave expand by counts: the sorted form of the mapping
        map_list.append(torch.from_numpy(np.sort(m)).to(dev))
    return _fuse(cos_list, map_list, mm_list, weights, n_base)


# ----------------------------------------------------------------------------- graph
class _Graph(tuple):
    """(A bf16 [n, lda], deg float32 [n]) plus `.p`, the neighbour count it was built from (at most 2 p non-zeros per row)."""


def getAffinityGraphMat(affinity_mat_raw: torch.Tensor, p_value: int):
    """0.5 * (B + B^T) of the p-neighbour binarisation.  Returns (A bf16 [n, lda], deg float32 [n])."""
    n = affinity_mat_raw.shape[0]
    dev = affinity_mat_raw.device
    p_value = int(p_value)
    if p_value <= 0:
        raise ValueError("p_value must be positive on the device path")
    p_value = min(p_value, n)
    lda = (n + 7) // 8 * 8
    a16 = torch.empty(n, lda, dtype=torch.bfloat16, device=dev)
    deg = torch.empty(n, dtype=torch.float32, device=dev)
    sel = torch.empty(n, n, dtype=torch.uint8, device=dev)
    mat = affinity_mat_raw.contiguous()
    _cabi.call("b200d_topp_binarize", ptr(mat), n, p_value, ptr(a16), lda, ptr(deg), ptr(sel), _s())
    graph = _Graph((a16, deg))
    graph.p = p_value
    return graph


# ----------------------------------------------------------------------------- NME analysis
def nme_ratios(evals: torch.Tensor, p_value_list: List[int], n: int, max_num_speakers: int, eps: float = 1e-10):
    """getEigRatio for every p of the sweep at once (host, float32 as upstream's per-p loop): evals [np, n_low + 1] holds
    the n_low lowest eigenvalues of each Laplacian and, last, its largest.  Returns (speaker counts, g_p) as float32
    [np]: count = argmax of the first max_num_speakers eigengaps + 1, g_p = (p / n) / (max gap / (lambda_max + eps) + eps)."""
    lambdas, lam_max = evals[:, :-1], evals[:, -1]
    gaps = (lambdas[:, 1:] - lambdas[:, :-1])[:, :max_num_speakers]
    num_of_spk = torch.argmax(gaps, dim=1) + 1
    max_key = torch.argsort(gaps, dim=1, descending=True)[:, :1]
    max_eig_gap = gaps.gather(1, max_key)[:, 0] / (lam_max.double() + eps).float()
    # upstream's `(p / n) / (tensor + eps)` is Tensor.__rtruediv__: reciprocal(tensor + eps) * float32(p / n)
    p_over_n = (torch.tensor([float(p) for p in p_value_list], dtype=torch.float64) / n).float()
    g_p = torch.reciprocal(max_eig_gap + eps) * p_over_n
    return num_of_spk.float(), g_p


class NMESC:
    """Normalized-maximum-eigengap analysis (upstream class of the same name): p-neighbour sweep on
    a strided subsample.  The subsample is ranked once; every p of the sweep is then an
    element-wise function of the rank matrices, and all Laplacians are diagonalised together."""

    def __init__(self, mat: torch.Tensor, max_num_speakers: int = 10, max_rp_threshold: float = 0.15, sparse_search: bool = True,
                 sparse_search_volume: int = 30, nme_mat_size: int = 512, use_subsampling_for_nme: bool = True,
                 fixed_thres: float = -1.0, maj_vote_spk_count: bool = False, presampled_ratio: Optional[int] = None, comm=None):
        """presampled_ratio: `mat` already IS the strided subsample mat_full[::r, ::r] with r = presampled_ratio (the
        row-sharded path gathers only those rows from the ranks); p-hat is still reported on the full matrix's scale.
        comm (rowshard.DistComm / LocalComm): the p values of the sweep are dealt to the communicator's ranks -- every rank
        builds and diagonalises the Laplacians of its own p values with the single-GPU reduction layout
        (b200d_eigvals_batched_layout) and the eigenvalues are all-gathered: bit for bit the replicated sweep."""
        self.presampled_ratio = presampled_ratio
        self.comm = comm
        self.max_num_speakers = int(max_num_speakers)
        self.max_rp_threshold = max_rp_threshold
        self.use_subsampling_for_nme = use_subsampling_for_nme
        self.nme_mat_size = nme_mat_size
        self.sparse_search = sparse_search
        self.sparse_search_volume = sparse_search_volume
        self.min_p_value = 2
        self.fixed_thres = fixed_thres
        self.eps = 1e-10
        self.max_N = 0
        self.mat = mat
        self.maj_vote_spk_count = maj_vote_spk_count
        self.p_value_list: List[int] = [2]

    def getPvalueList(self, n: int) -> List[int]:
        if self.fixed_thres is not None and self.fixed_thres > 0.0:
            self.max_N = max(int(math.floor(np.float32(n * self.fixed_thres))), self.min_p_value)
            return [self.max_N]
        self.max_N = max(int(math.floor(np.float32(n * self.max_rp_threshold))), self.min_p_value)
        if self.sparse_search:
            search_volume = min(self.max_N, int(self.sparse_search_volume))
            steps = min(self.max_N, max(search_volume, 2))
            return [int(v) for v in torch.linspace(start=1, end=self.max_N, steps=int(steps)).type(torch.int).tolist()]
        return list(range(1, self.max_N + 1))

    def forward(self) -> Tuple[int, int]:
        mat = self.mat
        N = mat.shape[0]
        dev = mat.device
        if self.presampled_ratio is not None:
            ratio, n, stride = int(self.presampled_ratio), N, 1
        else:
            ratio = max(1, int(N / self.nme_mat_size)) if self.use_subsampling_for_nme else 1
            n, stride = len(range(0, N, ratio)), ratio
        if n > 1024:
            raise ValueError(f"NME sweep matrix of size {n} exceeds the 1024 limit of the rank kernel (nme_mat_size too large)")
        self.p_value_list = self.getPvalueList(n)
        p_list = [min(p, n) for p in self.p_value_list]
        np_ = len(p_list)
        rank = torch.empty(n, n, dtype=torch.int16, device=dev)
        rankT = torch.empty(n, n, dtype=torch.int16, device=dev)
        _cabi.call("b200d_row_rank", ptr(mat), mat.stride(0), stride, n, ptr(rank), ptr(rankT), _s())
        m = min(self.max_num_speakers, n - 1)  # gaps[:max_num_speakers] needs lambda_0 .. lambda_m
        n_low = m + 1
        evals_all = []
        world = self.comm.world if self.comm is not None else 1
        for b0 in range(0, np_, 64):
            pl_all = p_list[b0 : b0 + 64]
            mine = list(range(len(pl_all)))[self.comm.rank :: world] if world > 1 else list(range(len(pl_all)))
            pl = [pl_all[i] for i in mine]
            evals = torch.empty(len(pl), n_low + 1, dtype=torch.float32, device=dev)
            if pl:
                lap = torch.empty(len(pl), n, n, dtype=torch.float32, device=dev)
                _cabi.call("b200d_laplacian_from_rank", ptr(rank), ptr(rankT), n, _i32_array(pl), len(pl), ptr(lap), _s())
                ws_bytes = _cabi.load().b200d_eigvals_workspace_bytes(len(pl), n)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                _cabi.call("b200d_eigvals_batched_layout", ptr(lap), len(pl), len(pl_all), n, n_low, ptr(evals), ptr(ws), ws_bytes, _s())
            if world > 1:  # rank r holds the p values r, r + world, ...: gather in rank order, then back to sweep order
                from .sharding import round_robin_counts_and_order

                counts, order = round_robin_counts_and_order(len(pl_all), world)
                gathered = self.comm.all_gather_rows(evals, counts)
                evals = torch.empty_like(gathered)
                evals[torch.tensor(order, device=dev)] = gathered
            evals_all.append(evals)
        evals = torch.cat(evals_all).cpu()  # [np, n_low + 1]  (the sweep's only device->host read)
        est_num_of_spk_list, eig_ratio_list = nme_ratios(evals, self.p_value_list, n, self.max_num_speakers, self.eps)
        index_nn = int(torch.argmin(eig_ratio_list))
        rp_p_value = self.p_value_list[index_nn]
        reach = self._reach(rank, rankT, n, [rp_p_value])
        if reach[0] != n:  # getMinimumConnection (connectivity is tested on the previous p's graph)
            cand = [1] + list(self.p_value_list)
            reach_all = self._reach(rank, rankT, n, cand[:-1])
            for fully_connected, p_neighbors in zip([r == n for r in reach_all], self.p_value_list):
                rp_p_value = p_neighbors
                if fully_connected or p_neighbors > self.max_N:
                    break
        p_hat_value = int(ratio * rp_p_value)
        if self.maj_vote_spk_count:
            est_num_of_spk = int(torch.mode(est_num_of_spk_list.clone())[0].item())
        else:
            est_num_of_spk = int(est_num_of_spk_list[index_nn].item())
        self.eig_ratio_list = eig_ratio_list
        self.est_num_of_spk_list = est_num_of_spk_list
        self.subsample_ratio, self.sweep_size = ratio, n
        return est_num_of_spk, p_hat_value

    @staticmethod
    def _reach(rank, rankT, n, p_list) -> List[int]:
        out: List[int] = []
        for b0 in range(0, len(p_list), 64):
            pl = [min(int(p), n) for p in p_list[b0 : b0 + 64]]
            reach = torch.empty(len(pl), dtype=torch.int32, device=rank.device)
            _cabi.call("b200d_graph_reach_rank", ptr(rank), ptr(rankT), n, _i32_array(pl), len(pl), ptr(reach), _s())
            out.extend(reach.cpu().tolist())
        return out


# ----------------------------------------------------------------------------- spectral embedding
class SpectralStats:
    """Counters of the last getSpectralEmbeddings call (reported by bench.py / tests)."""

    def __init__(self):
        self.outer = 0
        self.gemms = 0
        self.max_resid = float("nan")
        self.converged = True
        self.n = 0
        self.method = ""


last_spectral_stats = SpectralStats()
spectral_log: List[dict] = []  # one entry per bottom_eigvecs call (bounded; bench.py reports and clears it)


def cheb_operand_rows(b: int) -> int:
    """Rows of the W operand of the Chebyshev GEMM for a block of b vectors: [hi | mid | lo] parts of b rows each
    (b = 64: N = 192; b = 32: N = 128 with 32 zero rows, the narrowest tile of the kernel)."""
    return 192 if b == 64 else 128


def _gemm_cheb(A16, lda, vt_in, ldvt, n, nw, out, deg, x, xprev, ca, cb, cc, vt_out, splitk=None, k=None):
    """One Chebyshev / Laplacian step as a b200d_gemm_f16 call; `splitk`: uint8 workspace of b200d_gemm_cheb_splitk_bytes(n, nw, n)
    bytes -> the split-K form (what b200d_eig_bottomk always uses).  n: rows of this launch; k: columns of A (default n)."""
    epi = GemmEpilogue()
    epi.mode = _cabi.EPI_CHEB
    epi.deg = deg.data_ptr()
    epi.x32 = x.data_ptr()
    epi.xprev32 = xprev.data_ptr() if xprev is not None else None
    epi.ca, epi.cb, epi.cc = float(ca), float(cb), float(cc)
    epi.ldx = x.stride(0)
    epi.vt = vt_out.data_ptr() if vt_out is not None else None
    epi.ldvt = ldvt
    epi.flags = _cabi.gemm_flags()
    if splitk is not None:
        epi.splitk_ws, epi.splitk_ws_bytes = splitk.data_ptr(), splitk.numel()
    _cabi.call("b200d_gemm_f16", ptr(A16), lda, ptr(vt_in), ldvt, n, nw, n if k is None else k, ptr(out), out.stride(0), ctypes.byref(epi), _s())


def _spmm_cheb(csr, n, b, out, deg, x, xprev, ca, cb, cc):
    rowptr, colw = csr
    _cabi.call("b200d_spmm_cheb", ptr(rowptr), ptr(colw), n, b, ptr(deg), ptr(x), ptr(xprev), x.stride(0), float(ca), float(cb),
               float(cc), ptr(out), out.stride(0), _s())


def _dense_bottom_eigvecs(lap: torch.Tensor, k: int) -> torch.Tensor:
    """All eigenvectors of a small (<= 64) Laplacian by the fp64 Jacobi kernel; returns the k lowest."""
    n = lap.shape[0]
    dev = lap.device
    b = n + (n & 1)
    g = torch.zeros(b, b, dtype=torch.float32, device=dev)
    g[:n, :n] = lap
    if b > n:
        g[n, n] = 4.0 * float(lap.abs().sum().item()) + 1.0  # decoupled padding node, sorted last
    evals = torch.empty(b, dtype=torch.float32, device=dev)
    evecs = torch.empty(b, b, dtype=torch.float32, device=dev)
    _cabi.call("b200d_small_eig", ptr(g), b, ptr(evals), ptr(evecs), 0, _s())
    st = last_spectral_stats
    st.outer, st.gemms, st.max_resid, st.converged, st.n, st.method = 0, 0, 0.0, True, n, "jacobi"
    return evecs[:n, :k].contiguous()


# Graphs built from few neighbours take fp32 row-gather products over their CSR lists instead of the dense tcgen05
# GEMM.  Both are bandwidth bound (n * 2p * b * 4 bytes of gathers against n^2 * 2 bytes of operand + the re-read V
# tiles); measured on the 1-hour meeting's chunks (n = 10 000, b = 64, two chunks in flight) the two cross at ~3 %
# non-zeros per row (p = 171: 0.154 ms per product against 0.125 ms dense); the switch sits a factor two below that.
SPARSE_MAX_ROW_NNZ = int(os.environ.get("B200D_SPARSE_MAX_ROW_NNZ", "32"))  # always below this many non-zeros per row
SPARSE_MAX_DENSITY = float(os.environ.get("B200D_SPARSE_MAX_DENSITY", str(1.0 / 64.0)))  # ... or this fraction of a row


def _use_csr_products(n: int, p: Optional[int]) -> bool:
    if p is None:
        return False
    nnz = 2 * int(p)
    return nnz <= SPARSE_MAX_ROW_NNZ or nnz <= SPARSE_MAX_DENSITY * n


def bottom_eigvecs(a16: torch.Tensor, deg: torch.Tensor, k: int, tol: float = 2e-6, max_outer: int = 40, seed: int = 0,
                   p: Optional[int] = None, stepwise: bool = False) -> torch.Tensor:
    """The k lowest eigenvectors of L = diag(deg) - A for the binarised graph A (bf16 [n, lda]): `b200d_eig_bottomk`, the
    Chebyshev-filtered subspace iteration as one library call (upstream: a dense eigh(N x N) with k columns kept; k-means
    only sees the k-dimensional invariant subspace, which is what converges here).  The random start block comes from the
    CPU torch generator (seed 1000 + seed).  When the graph was built from few neighbours (`p` given, see _use_csr_products)
    the products run in fp32 over its CSR lists instead of the tcgen05 GEMM.  `stepwise=True` runs the same iteration call by
    call from Python (`bottom_eigvecs_stepwise`, the cross-check of the composite)."""
    if stepwise:
        return bottom_eigvecs_stepwise(a16, deg, k, tol, max_outer, seed, p)
    n, lda = a16.shape
    dev = a16.device
    lib = _cabi.load()
    b = int(lib.b200d_eig_bottomk_block(int(k)))
    if b == 0:
        raise NotImplementedError(f"spectral embedding for {k} clusters: the subspace block is limited to 64 vectors")
    if n < 2 * b:
        raise NotImplementedError(f"spectral embedding of {k} clusters on {n} points: needs n >= {2 * b} (or n <= {DENSE_EIG_MAX})")
    opt = _cabi.EigOptions(tol=float(tol), max_outer=int(max_outer), gemm_flags=_cabi.gemm_flags(), sparse_max_row_nnz=SPARSE_MAX_ROW_NNZ,
                           sparse_max_density=SPARSE_MAX_DENSITY)
    stats = _cabi.EigStats()
    pp = int(p) if p is not None else 0
    ws_bytes = int(lib.b200d_eig_bottomk_workspace_bytes(n, int(k), pp, ctypes.byref(opt)))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    X = torch.empty(n, b, dtype=torch.float32, device=dev)
    gen = torch.Generator(device="cpu").manual_seed(1000 + seed)
    X.copy_(torch.randn(n, b, generator=gen))
    _cabi.call("b200d_eig_bottomk", ptr(a16), lda, ptr(deg), n, int(k), pp, ptr(X), b, ctypes.byref(opt), ctypes.byref(stats), ptr(ws), ws_bytes, _s())
    st = SpectralStats()
    st.outer, st.gemms, st.n, st.max_resid, st.converged = stats.outer, stats.gemms, n, float(stats.max_resid), bool(stats.converged)
    st.method = f"chfsi{b}" + ("-csr" if stats.sparse else "")
    last_spectral_stats.__dict__.update(st.__dict__)
    if len(spectral_log) < 256:
        spectral_log.append({"n": n, "k": k, "p": p, "method": st.method, "block": b, "outer": st.outer, "gemms": st.gemms, "resid": st.max_resid,
                             "converged": st.converged, "history": [float(f"{stats.history[i]:.2e}") for i in range(min(stats.outer, _cabi.EIG_HISTORY))]})
    return X[:, :k].contiguous()


def bottom_eigvecs_stepwise(a16: torch.Tensor, deg: torch.Tensor, k: int, tol: float = 2e-6, max_outer: int = 40, seed: int = 0,
                            p: Optional[int] = None) -> torch.Tensor:
    """The iteration of b200d_eig_bottomk driven call by call from Python over the fine-grained entry points (gemm / spmm /
    gram / small_eig / right_mul / resid_norms).  Kept as the readable statement of the algorithm and as the test that the
    composite reproduces it bit for bit."""
    n, lda = a16.shape
    dev = a16.device
    if k + 8 <= 32:
        b = 32
    elif k + 8 <= 64:
        b = 64
    else:
        raise NotImplementedError(f"spectral embedding for {k} clusters: the subspace block is limited to 64 vectors")
    if n < 2 * b:
        raise NotImplementedError(f"spectral embedding of {k} clusters on {n} points: needs n >= {2 * b} (or n <= {DENSE_EIG_MAX})")
    nw = cheb_operand_rows(b)
    ldvt = (n + 7) // 8 * 8
    st = SpectralStats()  # per call: long-form chunks run this from several threads
    sparse = _use_csr_products(n, p)
    st.outer, st.gemms, st.n, st.method, st.converged = 0, 0, n, f"chfsi{b}" + ("-csr" if sparse else ""), False
    f32 = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
    X, W = f32(n, b), f32(n, b)
    Y = [f32(n, b) for _ in range(3)]
    if sparse:
        capacity = min(2 * int(p), n) * n
        csr = (torch.empty(n + 1, dtype=torch.int32, device=dev), torch.empty(capacity, dtype=torch.int32, device=dev))
        _cabi.call("b200d_csr_from_dense", ptr(a16), n, lda, ptr(csr[0]), ptr(csr[1]), capacity, _s())
        vt = [None, None]

        def step(vin, out, x, xprev, ca, cb, cc, vout):
            _spmm_cheb(csr, n, b, out, deg, x, xprev, ca, cb, cc)
    else:
        vt = [torch.zeros(nw, ldvt, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        splitk = torch.empty(int(_cabi.load().b200d_gemm_cheb_splitk_bytes(n, nw, n)), dtype=torch.uint8, device=dev)

        def step(vin, out, x, xprev, ca, cb, cc, vout):
            _gemm_cheb(a16, lda, vin, ldvt, n, nw, out, deg, x, xprev, ca, cb, cc, vout, splitk)
    G, Q, theta, resid = f32(b, b), f32(b, b), f32(b), f32(b)
    gws_bytes = _cabi.load().b200d_gram_workspace_bytes(n, b)
    gws = torch.empty(gws_bytes, dtype=torch.uint8, device=dev)
    gen = torch.Generator(device="cpu").manual_seed(1000 + seed)
    X.copy_(torch.randn(n, b, generator=gen))
    up = 2.0 * float(deg.max().item()) * 1.01 + 1e-3  # Gershgorin bound on lambda_max(L)

    def gram(a, c, out):
        _cabi.call("b200d_gram", ptr(a), ptr(c), n, b, b, ptr(out), ptr(gws), gws_bytes, _s())

    def cholqr(V, want_vt=None):
        gram(V, V, G)
        _cabi.call("b200d_small_eig", ptr(G), b, None, ptr(Q), 1, _s())
        _cabi.call("b200d_right_mul", ptr(V), n, b, b, ptr(Q), ptr(V), ptr(want_vt) if want_vt is not None else None, ldvt, _s())

    cholqr(X)
    cholqr(X, vt[0])
    history: List[float] = []
    for outer in range(max_outer):
        st.outer = outer + 1
        # Rayleigh-Ritz on span(X): W = L X, H = X^T W, X <- X Q, W <- W Q
        step(vt[0], W, X, None, 1.0, 0.0, 0.0, None)
        st.gemms += 1
        gram(X, W, G)
        _cabi.call("b200d_small_eig", ptr(G), b, ptr(theta), ptr(Q), 0, _s())
        _cabi.call("b200d_right_mul", ptr(X), n, b, b, ptr(Q), ptr(X), ptr(vt[0]), ldvt, _s())
        _cabi.call("b200d_right_mul", ptr(W), n, b, b, ptr(Q), ptr(W), None, ldvt, _s())
        _cabi.call("b200d_resid_norms", ptr(W), ptr(X), ptr(theta), n, b, b, ptr(resid), ptr(gws), gws_bytes, _s())
        th = theta.cpu().double().numpy()
        rs = np.sqrt(np.maximum(resid.cpu().double().numpy(), 0.0))
        st.max_resid = float(rs[:k].max() / up)
        history.append(st.max_resid)
        if st.max_resid <= tol:
            st.converged = True
            break
        if len(history) >= 8 and st.max_resid < 5e-5 and st.max_resid > 0.5 * history[-4]:
            st.converged = True  # stagnated at the fp32 floor of the products
            break
        # filter interval [a, up]: damp everything above the block's largest Ritz value
        a = float(min(max(th[b - 1], 1e-3 * up), 0.95 * up))
        e = (up - a) / 2.0
        c = (up + a) / 2.0
        t_k = (c - max(float(th[k - 1]), 0.0)) / e
        t_0 = c / e
        m = int(round(math.acosh(1e3) / max(math.acosh(max(t_k, 1.0 + 1e-9)), 1e-6)))
        m_cap = int(math.acosh(1e7) / math.acosh(t_0))
        m = max(3, min(m, max(m_cap, 3), 48))
        # scaled three-term recurrence (gain 1 at lambda = 0):  Y1 = (s1/e)(L - c) X,
        # Y_{i+1} = (2 s_{i+1}/e)(L - c) Y_i - s_i s_{i+1} Y_{i-1}
        sigma1 = e / (0.0 - c)
        sigma = sigma1
        tau = 2.0 / sigma1
        step(vt[0], Y[0], X, None, sigma1 / e, -c * sigma1 / e, 0.0, vt[1])
        st.gemms += 1
        prev, cur = X, Y[0]
        vin = 1
        for i in range(2, m + 1):
            nxt = Y[(i - 1) % 3]
            sigma_new = 1.0 / (tau - sigma)
            ca = 2.0 * sigma_new / e
            step(vt[vin], nxt, cur, prev, ca, -c * ca, -sigma * sigma_new, vt[1 - vin])
            st.gemms += 1
            sigma = sigma_new
            prev, cur = cur, nxt
            vin = 1 - vin
        X.copy_(cur)
        cholqr(X)
        cholqr(X, vt[0])
    last_spectral_stats.__dict__.update(st.__dict__)
    if len(spectral_log) < 256:
        spectral_log.append({"n": n, "k": k, "p": p, "method": st.method, "block": b, "outer": st.outer, "gemms": st.gemms, "resid": st.max_resid,
                             "converged": st.converged, "history": [float(f"{h:.2e}") for h in history]})
    return X[:, :k].contiguous()


class SpectralClustering:
    def __init__(self, n_clusters: int = 8, random_state: int = 0, n_random_trials: int = 1):
        self.n_clusters = int(n_clusters)
        self.random_state = random_state
        self.n_random_trials = max(n_random_trials, 1)

    def forward(self, graph) -> torch.Tensor:
        """`graph` is (A bf16 [n, lda], deg) from getAffinityGraphMat, or a raw float32 affinity [n, n]
        (the <= min_samples_for_nmesc branch of upstream, which clusters the un-binarised matrix)."""
        emb = self.getSpectralEmbeddings(graph)
        labels_set = [kmeans_torch(emb, self.n_clusters, random_state=seed)
                      for seed in range(self.random_state, self.random_state + self.n_random_trials)]
        if len(labels_set) == 1:
            return labels_set[0]
        stacked = torch.stack([l.cpu() for l in labels_set])
        label_index = torch.mode(torch.mode(stacked, 0)[1])[0]
        return labels_set[int(label_index)]

    def getSpectralEmbeddings(self, graph) -> torch.Tensor:
        if isinstance(graph, tuple):
            a16, deg = graph
            n = a16.shape[0]
            if n <= DENSE_EIG_MAX or (n <= DENSE_EIG_LIMIT and self.n_clusters + 8 > 32):
                lap = torch.diag(deg) - a16[:, :n].float()
                return _dense_bottom_eigvecs(lap, self.n_clusters)
            return bottom_eigvecs(a16, deg, self.n_clusters, p=getattr(graph, "p", None))
        mat = graph.float().clone()
        n = mat.shape[0]
        if n > DENSE_EIG_LIMIT:
            raise NotImplementedError("raw-affinity spectral embedding is only reached for <= min_samples_for_nmesc points")
        mat.fill_diagonal_(0)
        lap = torch.diag(mat.abs().sum(dim=1)) - mat
        return _dense_bottom_eigvecs(lap, self.n_clusters)


# ----------------------------------------------------------------------------- k-means
def kmeans_torch(X: torch.Tensor, num_clusters: int, threshold: float = 1e-4, iter_limit: int = 15, random_state: int = 0,
                 n_local_trials: int = 30) -> torch.Tensor:
    """kmeans_plusplus_torch + kmeans_torch on device; the CPU torch generator provides the draws
    upstream makes after torch.manual_seed(random_state).  Returns int64 labels on device."""
    X = X.float().contiguous()
    n, dim = X.shape
    dev = X.device
    num_clusters = int(num_clusters)
    if num_clusters == 1 or n == 1:
        return torch.zeros(n, dtype=torch.int64, device=dev)
    gen = torch.Generator(device="cpu").manual_seed(int(random_state))
    first = int(torch.randint(0, n, (1,), generator=gen).item())
    # one draw per center / per fallback index upstream; the CPU generator fills a batched draw in the same order
    # (tests/test_cpu_host.py pins that), and 250 tiny torch calls were ~2 ms on each chunk's critical path
    rands = torch.rand(num_clusters - 1, n_local_trials, generator=gen)
    n_fb = 4 * num_clusters
    fallback = torch.randint(n, (n_fb,), generator=gen).to(torch.int32)
    rands_d, fb_d = rands.to(dev), fallback.to(dev)
    labels = torch.empty(n, dtype=torch.int32, device=dev)
    ws_bytes = _cabi.load().b200d_kmeans_workspace_bytes(n, dim, num_clusters, n_local_trials)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _cabi.call("b200d_kmeans", ptr(X), n, dim, num_clusters, first, ptr(rands_d), n_local_trials, ptr(fb_d), n_fb, iter_limit,
               float(threshold), ptr(labels), ptr(ws), ws_bytes, _s())
    return labels.long()


# ----------------------------------------------------------------------------- enhanced count
def addAnchorEmb(emb: torch.Tensor, anchor_sample_n: int, anchor_spk_n: int, sigma: float, generator=None) -> torch.Tensor:
    """Host-side (CPU torch RNG, as upstream): synthetic anchor speakers for very short recordings.  `generator`: a CPU
    torch.Generator seeded like upstream's torch.manual_seed(seed) -- same stream of draws without touching (or racing on)
    the process-global RNG when several recordings are clustered from separate threads."""
    emb_dim = emb.shape[1]
    std_org = torch.std(emb, dim=0)
    sigma = torch.tensor(sigma)
    new_emb_list = []
    for _ in range(anchor_spk_n):
        emb_m = torch.tile(torch.randn(1, emb_dim, generator=generator), (anchor_sample_n, 1))
        emb_noise = torch.randn(anchor_sample_n, emb_dim, generator=generator).T
        emb_noise = torch.matmul(torch.diag(std_org), emb_noise / torch.max(torch.abs(emb_noise), dim=0)[0].unsqueeze(0)).T
        new_emb_list.append(emb_m + sigma * emb_noise)
    new_emb_list.append(emb)
    return torch.vstack(new_emb_list)


def getEnhancedSpeakerCount(emb: torch.Tensor, random_test_count: int = 5, anchor_spk_n: int = 3, anchor_sample_n: int = 10,
                            sigma: float = 50) -> int:
    """Speaker count for <= enhanced_count_thres segments: 5 seeded trials with anchor embeddings.
    Anchor generation is RNG-bound host work on a <= 80 x 192 matrix; each trial's affinity and NME
    sweep run on device."""
    dev = emb.device
    emb_cpu = emb.detach().float().cpu()
    est: List[int] = []
    for seed in range(random_test_count):
        gen = torch.Generator(device="cpu").manual_seed(seed)  # == torch.manual_seed(seed) followed by global draws
        emb_aug = addAnchorEmb(emb_cpu, anchor_sample_n, anchor_spk_n, sigma, generator=gen).to(dev)
        mat = getCosAffinityMatrix(emb_aug)
        nmesc = NMESC(mat, max_num_speakers=emb.shape[0], max_rp_threshold=0.15, sparse_search=True, sparse_search_volume=10,
                      fixed_thres=-1.0, nme_mat_size=300)
        est_num_of_spk, _ = nmesc.forward()
        est.append(int(est_num_of_spk))
    return max(int(torch.mode(torch.tensor(est))[0].item()) - anchor_spk_n, 1)


def split_input_data(embeddings_in_scales, timestamps_in_scales, multiscale_segment_counts):
    split_index = [int(x) for x in multiscale_segment_counts.tolist()]
    return list(torch.split(embeddings_in_scales, split_index, dim=0)), list(torch.split(timestamps_in_scales, split_index, dim=0))


# ----------------------------------------------------------------------------- top level
class SpeakerClustering:
    """offline_clustering.SpeakerClustering: forward_infer (multi-scale) / forward_unit_infer (one matrix)."""

    def __init__(self, min_samples_for_nmesc: int = MIN_SAMPLES_FOR_NMESC, nme_mat_size: int = NME_MAT_SIZE, sparse_search: bool = True,
                 maj_vote_spk_count: bool = False):
        self.min_samples_for_nmesc = min_samples_for_nmesc
        self.nme_mat_size = nme_mat_size
        self.sparse_search = sparse_search
        self.maj_vote_spk_count = maj_vote_spk_count
        self.embeddings_in_scales: List[torch.Tensor] = []
        self.timestamps_in_scales: List[torch.Tensor] = []
        self.debug = {}
        self.keep_affinity = False  # parity tests set this to read the fused N x N matrix back (13 GB for a 4-hour recording)
        self.fused_affinity: Optional[torch.Tensor] = None
        # rowshard.DistComm / LocalComm: one long recording with the affinity, its graph and the eigensolver's products
        # row-sharded over the ranks of the communicator (recordings of >= rowshard.MIN_ROWS_TO_SHARD base windows)
        self.row_comm = None

    def forward_unit_infer(self, mat: torch.Tensor, oracle_num_speakers: int = -1, max_num_speakers: int = 8,
                           max_rp_threshold: float = 0.15, sparse_search_volume: int = 30, est_num_of_spk_enhanced: int = -1,
                           fixed_thres: float = -1.0, kmeans_random_trials: int = 1) -> torch.Tensor:
        nmesc = NMESC(mat, max_num_speakers=max_num_speakers, max_rp_threshold=max_rp_threshold, sparse_search=self.sparse_search,
                      sparse_search_volume=sparse_search_volume, fixed_thres=fixed_thres, nme_mat_size=self.nme_mat_size,
                      maj_vote_spk_count=self.maj_vote_spk_count)
        if mat.shape[0] > self.min_samples_for_nmesc:
            est_num_of_spk, p_hat_value = nmesc.forward()
            graph = getAffinityGraphMat(mat, p_hat_value)
        else:
            nmesc.fixed_thres = max_rp_threshold
            est_num_of_spk, p_hat_value = nmesc.forward()
            graph = mat
        if oracle_num_speakers > 0:
            n_clusters = int(oracle_num_speakers)
        elif est_num_of_spk_enhanced > 0:
            n_clusters = int(est_num_of_spk_enhanced)
        else:
            n_clusters = int(est_num_of_spk)
        self.debug = {"est_num_of_spk": int(est_num_of_spk), "p_hat": int(p_hat_value), "n_clusters": n_clusters,
                      "g_p": getattr(nmesc, "eig_ratio_list", None), "p_list": list(nmesc.p_value_list)}
        spectral_model = SpectralClustering(n_clusters=n_clusters, n_random_trials=kmeans_random_trials)
        return spectral_model.forward(graph)

    def forward_infer(self, embeddings_in_scales: torch.Tensor, timestamps_in_scales: torch.Tensor,
                      multiscale_segment_counts: torch.Tensor, multiscale_weights: torch.Tensor, oracle_num_speakers: int = -1,
                      max_rp_threshold: float = 0.15, max_num_speakers: int = 8, enhanced_count_thres: int = ENHANCED_COUNT_THRES,
                      sparse_search_volume: int = 30, fixed_thres: float = -1.0, scale_mapping=None) -> torch.Tensor:
        self.embeddings_in_scales, self.timestamps_in_scales = split_input_data(embeddings_in_scales, timestamps_in_scales,
                                                                                multiscale_segment_counts)
        emb = self.embeddings_in_scales[-1]
        if emb.shape[0] == 1:
            return torch.zeros((1,), dtype=torch.int64, device=emb.device)
        elif emb.shape[0] <= max(enhanced_count_thres, self.min_samples_for_nmesc) and oracle_num_speakers < 0:
            est_num_of_spk_enhanced = getEnhancedSpeakerCount(emb=emb)
        else:
            est_num_of_spk_enhanced = -1
        if oracle_num_speakers > 0:
            max_num_speakers = oracle_num_speakers
        if self.row_comm is not None and self.row_comm.world > 1 and est_num_of_spk_enhanced < 0 and not self.keep_affinity:
            from . import rowshard

            if emb.shape[0] >= rowshard.MIN_ROWS_TO_SHARD:
                return rowshard.forward_infer_rows(self, self.row_comm, self.embeddings_in_scales, self.timestamps_in_scales, multiscale_weights,
                                                   oracle_num_speakers, max_rp_threshold, max_num_speakers, sparse_search_volume, fixed_thres,
                                                   scale_mapping)
        mat = getMultiScaleCosAffinityMatrix(multiscale_weights, self.embeddings_in_scales, self.timestamps_in_scales, scale_mapping)
        self.fused_affinity = mat if self.keep_affinity else None
        return self.forward_unit_infer(mat=mat, oracle_num_speakers=oracle_num_speakers, max_rp_threshold=max_rp_threshold,
                                       max_num_speakers=max_num_speakers, sparse_search_volume=sparse_search_volume,
                                       est_num_of_spk_enhanced=est_num_of_spk_enhanced, fixed_thres=fixed_thres)
