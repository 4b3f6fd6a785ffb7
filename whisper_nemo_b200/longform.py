"""Long-form clustering on B200: recordings with more than `embeddings_per_chunk` base-scale windows.

Mirrors upstream NeMo's `nemo/collections/asr/parts/utils/longform_clustering.py`
(`LongFormSpeakerClustering`) and the reducer helpers it borrows from
`online_clustering.py` (`get_merge_quantity`, `calculate_removable_counts`,
`get_closest_embeddings`, `merge_vectors`, `run_reducer`) -- the behaviour the reference
gets from the shipped knobs `chunk_cluster_count: 50`, `embeddings_per_chunk: 10000`
(nemo_msdd_configs/diar_infer_*.yaml:55-56) on any recording longer than ~42 min.

Per chunk the O(N^2)/O(N^3) work stays on device: scale interpolation (interp_scales), the
10 000 x 10 000 cosine affinity, NME sweep, binarisation, the 50-vector spectral embedding and
k-means (clustering.py), and the within-cluster affinity mass that ranks merge candidates
(masked_rowsum, one pass for all 50 clusters).  What remains on the host is integer
bookkeeping on <= 50 cluster sizes and index lists.
"""
import ctypes
import os
from typing import List, Tuple

import numpy as np
import torch

from . import _cabi
from ._cabi import ptr
from .clustering import SpeakerClustering, _s, get_argmin_mat, cos_affinity, _fuse, split_input_data


_CHUNK_STREAMS: List[torch.cuda.Stream] = []


def get_scale_interpolated_embs(multiscale_weights, embeddings_in_scales, timestamps_in_scales, scale_mapping=None):
    """One embedding per base-scale window: the weighted sum over scales of the embedding of the
    nearest window of each scale.  Returns (float32 [n_base, d] on device, mapping list)."""
    mapping = scale_mapping if scale_mapping is not None else get_argmin_mat(timestamps_in_scales)
    dev = embeddings_in_scales[0].device
    n_base = int(timestamps_in_scales[-1].shape[0])
    d = int(embeddings_in_scales[0].shape[1])
    S = len(embeddings_in_scales)
    embs = [e.float().contiguous() for e in embeddings_in_scales]
    maps = [torch.from_numpy(np.sort(m)).to(dev) for m in mapping]
    weights = torch.as_tensor(multiscale_weights).reshape(-1).tolist()
    out = torch.empty(n_base, d, dtype=torch.float32, device=dev)
    emb_p = (ctypes.c_void_p * S)(*[e.data_ptr() for e in embs])
    map_p = (ctypes.c_void_p * S)(*[m.data_ptr() for m in maps])
    w = (ctypes.c_float * S)(*[float(x) for x in weights])
    _cabi.call("b200d_interp_scales", S, emb_p, map_p, w, ptr(out), n_base, d, _s())
    return out, mapping


def calculate_removable_counts(removable_counts_mat: torch.Tensor, remain_count: int, num_clus: int) -> torch.Tensor:
    """Spread `remain_count` kept vectors over the clusters as evenly as their sizes allow
    (water-filling from the largest cluster down); returns how many vectors each cluster gives up.
    The two orderings come from torch.sort (upstream's tie order); the arithmetic on <= 50 integers is numpy --
    as torch CPU ops it cost milliseconds per chunk on the critical path of the long-form stage."""
    asc = torch.sort(removable_counts_mat)[0].numpy()
    order = torch.sort(removable_counts_mat, descending=True)[1].numpy()
    counts = removable_counts_mat.numpy().copy()
    padded = np.concatenate([[0], asc, [0]])
    steps = (padded[1:] - padded[:-1])[:num_clus]
    level_cost = np.cumsum(np.arange(num_clus, 0, -1) * steps).tolist()
    level = 0
    for level, cost in enumerate(level_cost):
        if remain_count < cost:
            break
    left = remain_count
    for j in range(level):
        counts[order[: num_clus - j]] -= steps[j]
        left -= int(steps[j]) * (num_clus - j)
    share, extra = divmod(left, num_clus - level)
    counts[order[: num_clus - level]] -= share
    counts[order[:extra]] -= 1
    return torch.from_numpy(counts).int()


def get_merge_quantity(num_to_be_removed: int, pre_clus_labels: torch.Tensor, min_count_per_cluster: int) -> torch.Tensor:
    if num_to_be_removed > pre_clus_labels.shape[0] - 1:
        raise ValueError(f"num_to_be_removed: {num_to_be_removed} should be less than pre_clus_labels length - 1")
    remain_count = pre_clus_labels.shape[0] - num_to_be_removed
    spk_freq_count = torch.bincount(pre_clus_labels)
    num_clus = len(torch.unique(pre_clus_labels))
    if remain_count < min_count_per_cluster * num_clus:
        raise ValueError("The remaining embedding vectors should be more than the minimum quantity")
    floor = torch.minimum(torch.full_like(spk_freq_count, min_count_per_cluster), spk_freq_count)
    remain_count -= int(floor.sum())
    removable = calculate_removable_counts(spk_freq_count - floor, remain_count, num_clus)
    if int(removable.sum()) != num_to_be_removed:
        raise ValueError("Sum of `removable_counts_mat` is not equal to `num_to_be_removed` variable.")
    if not torch.all(removable >= 0) or not torch.all(spk_freq_count - floor >= removable):
        raise ValueError("removable_counts_mat out of range")
    return removable


def plan_cluster_merges(y_np: np.ndarray, vols: List[int], mass_np, n: int, offset_index: int):
    """Index bookkeeping of run_reducer for all clusters of one chunk (host, no device work).
    y_np: int64 labels [n]; vols[c]: how many vectors cluster c gives up; mass_np[i]: within-cluster affinity mass of
    window i (column sum of its cluster's affinity block), needed when any vols[c] > 0.
    Per cluster, in label order: the merge_quantity + 1 windows of largest mass are replaced by their mean, the others
    are kept in ascending order.  Returns (mapping_list [(kept + offset, merged + offset)], sizes, sel_idx (members of
    each mean), seg_off (their segment offsets), order (gather order over [chunk rows ; means]), n_avg)."""
    # members of every cluster in ascending window order (== torch.where(labels == c)[0]) from one stable sort
    by_label = np.argsort(y_np, kind="stable")
    bounds = np.searchsorted(y_np[by_label], np.arange(len(vols) + 1))
    mapping_list, sizes = [], []
    sel_idx, seg_off, order = [], [0], []
    n_avg = 0
    none = torch.arange(0)
    for spk_idx, merge_quantity in enumerate(vols):
        target = by_label[bounds[spk_idx] : bounds[spk_idx + 1]]
        if merge_quantity > 0:
            if merge_quantity > target.shape[0] - 1:
                raise ValueError("merge_quantity is larger than the half of targeted speaker's labels")
            # the ranking is upstream's torch call (its tie order); the index bookkeeping around it is numpy
            rank = torch.argsort(torch.from_numpy(mass_np[target]), descending=True).numpy()
            selected, rest_sorted = rank[: merge_quantity + 1], np.sort(rank[merge_quantity + 1 :])
            sel_idx.append(target[selected])
            seg_off.append(seg_off[-1] + selected.size)
            order.append(target[rest_sorted])
            order.append(np.array([n + n_avg]))  # row of the merged vector in [emb_part ; means]
            n_avg += 1
            mapping_list.append((torch.from_numpy(target[rest_sorted] + offset_index), torch.from_numpy(target[selected] + offset_index)))
            sizes.append(int(rest_sorted.size) + 1)
            if target.shape[0] - merge_quantity != sizes[-1]:
                raise ValueError("Reducer output is not matched to the target quantity")
        else:
            order.append(target)
            mapping_list.append((torch.from_numpy(target + offset_index), none))
            sizes.append(int(target.size))
    return mapping_list, sizes, sel_idx, seg_off, order, n_avg


class LongFormSpeakerClustering:
    def __init__(self, shard_chunks: bool = False, chunk_streams: int = None):
        """shard_chunks: under torch.distributed, deal the (independent) chunks of the long-form path to the ranks and
        gather their reduced embeddings; everything else runs replicated and gives the same labels on every rank.
        chunk_streams: chunks clustered concurrently on this many streams (default: B200D_CHUNK_STREAMS, 2)."""
        self.shard_chunks = bool(shard_chunks)
        self.chunk_streams = chunk_streams
        self.speaker_clustering = SpeakerClustering()
        self.embeddings_in_scales: List[torch.Tensor] = []
        self.timestamps_in_scales: List[torch.Tensor] = []
        self.chunk_labels = {}  # long-form path: chunk index -> (first window, int64 over-clustering labels on the host)

    @staticmethod
    def get_div_ceil_count(numer: int, denomin: int) -> int:
        return int(torch.ceil(torch.tensor(numer / denomin)).item())

    def check_input(self, embeddings_per_chunk, chunk_cluster_count, max_num_speakers) -> None:
        if chunk_cluster_count is None or embeddings_per_chunk is None:
            raise ValueError(f"chunk_cluster_count ({chunk_cluster_count}) and embeddings_per_chunk ({embeddings_per_chunk}) should be set.")
        if chunk_cluster_count >= embeddings_per_chunk:
            raise ValueError("chunk_cluster_count should be smaller than embeddings_per_chunk.")
        if max_num_speakers <= 1:
            raise ValueError("max_num_speakers should be greater than 1.")
        if chunk_cluster_count <= max_num_speakers:
            raise ValueError("chunk_cluster_count should be greater than max_num_speakers.")

    def forward_infer(self, embeddings_in_scales, timestamps_in_scales, multiscale_segment_counts, multiscale_weights,
                      oracle_num_speakers: int = -1, max_rp_threshold: float = 0.15, max_num_speakers: int = 8,
                      sparse_search_volume: int = 30, fixed_thres: float = -1.0, chunk_cluster_count=50,
                      embeddings_per_chunk=10000, scale_mapping=None) -> torch.Tensor:
        """scale_mapping: optional precomputed clustering.get_argmin_mat(timestamps) (pure index arithmetic on the window
        times; the diarizer computes it on the host while the GPU is still embedding)."""
        if embeddings_per_chunk is not None and int(torch.max(multiscale_segment_counts)) > embeddings_per_chunk:
            return self.long_forward_infer(embeddings_in_scales, timestamps_in_scales, multiscale_segment_counts, multiscale_weights,
                                           oracle_num_speakers, max_rp_threshold, max_num_speakers, sparse_search_volume, fixed_thres,
                                           int(chunk_cluster_count), int(embeddings_per_chunk), scale_mapping)
        labels = self.speaker_clustering.forward_infer(
            embeddings_in_scales=embeddings_in_scales, timestamps_in_scales=timestamps_in_scales,
            multiscale_segment_counts=multiscale_segment_counts, multiscale_weights=multiscale_weights,
            oracle_num_speakers=oracle_num_speakers, max_rp_threshold=max_rp_threshold, max_num_speakers=max_num_speakers,
            sparse_search_volume=sparse_search_volume, fixed_thres=fixed_thres, scale_mapping=scale_mapping)
        self.timestamps_in_scales = self.speaker_clustering.timestamps_in_scales
        return labels

    def _reduce_chunk(self, emb_part: torch.Tensor, y_host: torch.Tensor, mass_np, class_target_vol: torch.Tensor, offset_index: int):
        """run_reducer for every cluster of one chunk.  Returns ([merged_embs per cluster], [index mappings]).
        The index bookkeeping (<= 50 clusters) is host work on the labels and the within-cluster affinity mass; the
        embedding traffic is two device operations for the whole chunk: merged means, one gather."""
        n, d = emb_part.shape
        dev = emb_part.device
        y_np = y_host.numpy()
        vols = [int(v) for v in class_target_vol.tolist()]
        mapping_list, sizes, sel_idx, seg_off, order, n_avg = plan_cluster_merges(y_np, vols, mass_np if sum(vols) > 0 else None, n, offset_index)
        src = emb_part
        if n_avg > 0:
            idx_d = torch.from_numpy(np.concatenate(sel_idx).astype(np.int32)).to(dev)
            off_d = torch.tensor(seg_off, dtype=torch.int32).to(dev)
            means = torch.empty(n_avg, d, dtype=torch.float32, device=dev)
            x = emb_part.contiguous()
            _cabi.call("b200d_gather_segment_mean", ptr(x), d, ptr(idx_d), ptr(off_d), n_avg, ptr(means), _s())
            src = torch.cat([emb_part, means], dim=0)
        merged_all = src.index_select(0, torch.from_numpy(np.concatenate(order)).to(dev)) if order else src[:0]
        merged_list = list(torch.split(merged_all, sizes, dim=0))
        return merged_list, mapping_list

    def long_forward_infer(self, embeddings_in_scales, timestamps_in_scales, multiscale_segment_counts, multiscale_weights,
                           oracle_num_speakers, max_rp_threshold, max_num_speakers, sparse_search_volume, fixed_thres,
                           chunk_cluster_count, embeddings_per_chunk, scale_mapping=None) -> torch.Tensor:
        self.check_input(embeddings_per_chunk, chunk_cluster_count, max_num_speakers)
        self.embeddings_in_scales, self.timestamps_in_scales = split_input_data(embeddings_in_scales, timestamps_in_scales,
                                                                                multiscale_segment_counts)
        emb, _ = get_scale_interpolated_embs(multiscale_weights, self.embeddings_in_scales, self.timestamps_in_scales, scale_mapping)
        n_total = emb.shape[0]
        total_emb: List[torch.Tensor] = []
        window_range_list: List[Tuple[int, int]] = []
        absolute_merge_mapping = []
        window_offset = 0
        n_chunks = self.get_div_ceil_count(n_total, embeddings_per_chunk)
        rank, world = (0, 1)
        if self.shard_chunks:
            from . import sharding

            rank, world = sharding.rank_world()

        def chunk_range(win_index: int):
            if embeddings_per_chunk * (win_index + 1) > n_total:  # last chunk is aligned to the end (overlaps the previous one)
                offset_index = n_total - embeddings_per_chunk
            else:
                offset_index = embeddings_per_chunk * win_index
            return offset_index, emb[offset_index : offset_index + embeddings_per_chunk]

        def cluster_chunk(win_index: int, clusterer: SpeakerClustering):
            """The expensive, per-chunk part (the unit dealt to ranks / streams): affinity, NME sweep, binarisation, spectral
            embedding, k-means(50) and the within-cluster affinity mass.  Returns (labels int32 [m], mass float32 [m]) on device."""
            _, emb_part = chunk_range(win_index)
            m = emb_part.shape[0]
            if m == 1:
                return torch.zeros((1,), dtype=torch.int32, device=emb.device), torch.zeros((1,), dtype=torch.float32, device=emb.device)
            cos, mm = cos_affinity(emb_part)
            ident = torch.arange(m, dtype=torch.int32, device=emb.device)
            mat = _fuse([cos], [ident], [mm], [1.0], m)
            del cos
            overcluster_count = min(chunk_cluster_count, mat.shape[0])
            Y_part = clusterer.forward_unit_infer(
                mat=mat, oracle_num_speakers=overcluster_count, max_rp_threshold=max_rp_threshold,
                max_num_speakers=chunk_cluster_count, sparse_search_volume=sparse_search_volume)
            y32 = Y_part.to(torch.int32).contiguous()
            mass = torch.empty(m, dtype=torch.float32, device=emb.device)
            _cabi.call("b200d_masked_rowsum", ptr(mat), m, ptr(y32), ptr(mass), _s())
            return y32, mass

        mine = [w for w in range(n_chunks) if w % world == rank]  # sharding.chunks_of_rank
        want = self.chunk_streams if self.chunk_streams is not None else int(os.environ.get("B200D_CHUNK_STREAMS", "2"))
        n_streams = min(len(mine), max(1, want))
        per_chunk = {}
        if n_streams <= 1:
            for w in mine:
                per_chunk[w] = cluster_chunk(w, self.speaker_clustering)
        else:
            # chunks are independent and each one's kernels (80-CTA GEMMs, one-CTA Jacobi, host round trips) leave most of
            # the GPU idle: run them on separate streams from separate host threads (ctypes releases the GIL)
            from concurrent.futures import ThreadPoolExecutor

            main = torch.cuda.current_stream()
            if len(_CHUNK_STREAMS) < n_streams:
                _CHUNK_STREAMS.extend(torch.cuda.Stream() for _ in range(n_streams - len(_CHUNK_STREAMS)))
            streams = _CHUNK_STREAMS[:n_streams]
            for st in streams:
                st.wait_stream(main)
            dev_index = emb.device.index

            free = list(streams)

            def work(w):
                torch.cuda.set_device(dev_index)
                st = free.pop()  # each worker thread takes a stream of its own (list.pop / append are atomic under the GIL)
                try:
                    with torch.cuda.stream(st), torch.no_grad(), _cabi.single_cta_gemms():
                        y32, mass = cluster_chunk(w, SpeakerClustering())
                        y32.record_stream(main)
                        mass.record_stream(main)
                        return y32, mass
                finally:
                    free.append(st)

            main.synchronize()  # nothing of the single-stream phase (CTA-pair GEMMs) may still be running
            with _cabi.short_gil_switch(), ThreadPoolExecutor(max_workers=n_streams) as pool:
                futures = {w: pool.submit(work, w) for w in mine}
                for w, fut in futures.items():
                    per_chunk[w] = fut.result()
                for st in streams:
                    st.synchronize()  # the multi-stream region is drained before the pair kernel is allowed again
        # every rank needs every chunk's over-clustering: 2 x 4 bytes per window (labels + affinity mass), gathered as one
        # device tensor per rank (NCCL over NVLink; gloo in the CPU tests) -- the merge bookkeeping below is then replicated
        m_chunk = min(embeddings_per_chunk, n_total)
        if world > 1:
            from . import sharding

            per_chunk = sharding.all_gather_chunk_results(per_chunk, n_chunks, m_chunk, [chunk_range(w)[1].shape[0] for w in range(n_chunks)])
        for win_index in range(n_chunks):
            offset_index, emb_part = chunk_range(win_index)
            y32, mass = per_chunk[win_index]
            y_host = y32.cpu().long()
            self.chunk_labels[win_index] = (offset_index, y_host)
            num_to_be_merged = int(min(embeddings_per_chunk, emb_part.shape[0]) - chunk_cluster_count)
            min_count_per_cluster = self.get_div_ceil_count(chunk_cluster_count, len(torch.unique(y_host)))
            class_target_vol = get_merge_quantity(num_to_be_merged, y_host, min_count_per_cluster)
            merged_list, mapping_list = self._reduce_chunk(emb_part, y_host, mass.cpu().numpy(), class_target_vol, offset_index)
            for merged, mapping in zip(merged_list, mapping_list):
                total_emb.append(merged)
                absolute_merge_mapping.append(mapping)
                window_range_list.append((window_offset, window_offset + merged.shape[0]))
                window_offset += merged.shape[0]
        reduced_embs = torch.cat(total_emb)
        cos, mm = cos_affinity(reduced_embs)
        ident = torch.arange(reduced_embs.shape[0], dtype=torch.int32, device=emb.device)
        reduced_mat = _fuse([cos], [ident], [mm], [1.0], reduced_embs.shape[0])
        Y_aggr = self.speaker_clustering.forward_unit_infer(
            mat=reduced_mat, oracle_num_speakers=oracle_num_speakers, max_rp_threshold=max_rp_threshold,
            max_num_speakers=max_num_speakers, sparse_search_volume=sparse_search_volume, fixed_thres=fixed_thres)
        if reduced_embs.shape[0] != Y_aggr.shape[0]:
            raise ValueError("The number of embeddings and labels should be same")
        # unpack: every original window takes the label of the reduced vector it was kept as / merged into
        y_aggr = Y_aggr.cpu()
        Y_unpack = torch.zeros((n_total,), dtype=torch.int64)
        for (lo, hi), (kept, merged_idx) in zip(window_range_list, absolute_merge_mapping):
            part = y_aggr[lo:hi]
            if len(merged_idx) > 0:
                Y_unpack[merged_idx] = part[-1].clone()
                if len(kept) > 0:
                    Y_unpack[kept] = part[:-1].clone()
            else:
                Y_unpack[kept] = part.clone()
        return Y_unpack.to(emb.device)
