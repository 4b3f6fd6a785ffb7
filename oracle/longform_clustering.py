"""Oracle: long-form clustering.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates upstream `nemo/collections/asr/parts/utils/longform_clustering.py`
(`LongFormSpeakerClustering`) and the reducer helpers it takes from
`nemo/collections/asr/parts/utils/online_clustering.py` (`get_merge_quantity`,
`calculate_removable_counts`, `get_closest_embeddings`, `merge_vectors`,
`run_reducer`) and `offline_clustering.get_scale_interpolated_embs`.
Active when max(multiscale_segment_counts) > embeddings_per_chunk
(diar_infer_*.yaml:55-56: chunk_cluster_count 50, embeddings_per_chunk 10000);
SURVEY.md section 8 row a18.  LOWEST-CONFIDENCE part of the restatement: the
merge order and label unpacking are recalled from memory.
"""
from typing import List, Tuple

import torch

from .offline_clustering import (
    SpeakerClustering,
    get_argmin_mat,
    getCosAffinityMatrix,
    getRepeatedList,
    split_input_data,
)


def get_scale_interpolated_embs(multiscale_weights, embeddings_in_scales, timestamps_in_scales):
    rep_mat_list = []
    session_scale_mapping_list = get_argmin_mat(timestamps_in_scales)
    for scale_idx in range(len(timestamps_in_scales)):
        mapping_argmat = session_scale_mapping_list[scale_idx]
        emb_t = embeddings_in_scales[scale_idx]
        repeat_list = getRepeatedList(mapping_argmat, torch.tensor(emb_t.shape[0]))
        rep_mat_list.append(torch.repeat_interleave(emb_t, repeats=repeat_list, dim=0))
    stacked_scale_embs = torch.stack(rep_mat_list)
    context_emb = torch.matmul(stacked_scale_embs.permute(2, 1, 0), multiscale_weights.t()).squeeze().t()
    if len(context_emb.shape) < 2:
        context_emb = context_emb.unsqueeze(0)
    return context_emb, session_scale_mapping_list


def calculate_removable_counts(removable_counts_mat: torch.Tensor, remain_count: int, num_clus: int) -> torch.Tensor:
    zero = torch.tensor([0])
    zero_padded_counts = torch.cat([zero, removable_counts_mat.sort()[0], zero], dim=0)
    removable_count_args = removable_counts_mat.sort(descending=True)[1]
    diff_counts = (zero_padded_counts[1:] - zero_padded_counts[:-1])[:num_clus]
    gradual_counts = torch.arange(num_clus, 0, -1) * diff_counts
    cumsum_counts = torch.cumsum(gradual_counts, dim=0)
    remain_count_rem = remain_count
    ind = 0
    for ind, num in enumerate(cumsum_counts):
        if remain_count < num:
            break
    if ind > 0:
        for knd in range(ind):
            removable_counts_mat[removable_count_args[: num_clus - knd]] -= diff_counts[knd]
            remain_count_rem -= int(diff_counts[knd].item()) * (num_clus - knd)
    assert remain_count >= 0
    num_labels = remain_count_rem // (num_clus - ind)
    rem_labels = remain_count_rem % (num_clus - ind)
    removable_counts_mat[removable_count_args[: (num_clus - ind)]] -= num_labels
    removable_counts_mat[removable_count_args[:rem_labels]] -= 1
    return removable_counts_mat.int()


def get_merge_quantity(num_to_be_removed: int, pre_clus_labels: torch.Tensor, min_count_per_cluster: int) -> torch.Tensor:
    if num_to_be_removed > pre_clus_labels.shape[0] - 1:
        raise ValueError(f"num_to_be_removed: {num_to_be_removed} should be less than pre_clus_labels length - 1")
    remain_count = pre_clus_labels.shape[0] - num_to_be_removed
    spk_freq_count = torch.bincount(pre_clus_labels)
    num_clus = len(torch.unique(pre_clus_labels))
    if remain_count < min_count_per_cluster * num_clus:
        raise ValueError("The remaining embedding vectors should be more than the minimum quantity")
    min_seg_count = torch.tensor([min_count_per_cluster] * len(spk_freq_count))
    min_seg_count_mat = torch.stack((min_seg_count, spk_freq_count)).min(0)[0]
    remain_count -= int(torch.sum(min_seg_count_mat))
    removable_counts_mat = spk_freq_count - min_seg_count_mat
    removable_counts_mat = calculate_removable_counts(removable_counts_mat, remain_count, num_clus)
    if int(removable_counts_mat.sum()) != num_to_be_removed:
        raise ValueError("Sum of `removable_counts_mat` is not equal to `num_to_be_removed` variable.")
    if not torch.all(removable_counts_mat >= 0) or not torch.all(spk_freq_count - min_seg_count_mat >= removable_counts_mat):
        raise ValueError("removable_counts_mat out of range")
    return removable_counts_mat


def get_closest_embeddings(affinity_mat: torch.Tensor, n_closest: int = 2) -> Tuple[torch.Tensor, torch.Tensor]:
    comb_limit = int(affinity_mat.shape[0] - 1)
    if n_closest > comb_limit:
        raise ValueError(f"Got n_closest of {n_closest}: {n_closest} is bigger than comb_limit {comb_limit}")
    sum_cmat = affinity_mat.sum(0)
    order = torch.argsort(sum_cmat, descending=True)
    return order[: (n_closest + 1)], order[(n_closest + 1) :]


def merge_vectors(selected_inds: torch.Tensor, emb_ndx: torch.Tensor, pre_cluster_labels: torch.Tensor):
    if emb_ndx.shape[0] != pre_cluster_labels.shape[0]:
        raise ValueError("pre_cluster_labels and emb_ndx have mismatch in dimension")
    avg_emb = torch.mean(emb_ndx[selected_inds, :], dim=0)
    merged_clus_labels = pre_cluster_labels[selected_inds]
    selected_inds_list: List[int] = selected_inds.tolist()
    sel = set(selected_inds_list)
    bypass_inds = torch.tensor([k for k in range(emb_ndx.shape[0]) if k not in sel], dtype=torch.long)
    if bypass_inds.shape[0] == 0:
        merged_vecs = avg_emb.unsqueeze(0)
        merged_clus_labels = merged_clus_labels[:1]
    else:
        merged_vecs = torch.vstack((emb_ndx[bypass_inds], avg_emb))
        merged_clus_labels = torch.hstack((pre_cluster_labels[bypass_inds], merged_clus_labels[0]))
    return merged_vecs, merged_clus_labels


def run_reducer(pre_embs, target_spk_idx: int, merge_quantity: int, pre_clus_labels, total_affinity_mat=None):
    """online_clustering.run_reducer.  Upstream recomputes getCosAffinityMatrix(pre_embs) on every
    call; the matrix is identical across calls on one chunk, so the caller may pass it in."""
    if pre_embs.shape[0] != pre_clus_labels.shape[0]:
        raise ValueError("Dimension mismatch between `pre_embs` and `pre_clus_labels`.")
    target_emb_index = torch.where(pre_clus_labels == target_spk_idx)[0]
    org_size = target_emb_index.shape[0]
    if merge_quantity > 0:
        if merge_quantity > (target_emb_index.shape[0] - 1):
            raise ValueError("merge_quantity is larger than the half of targeted speaker's labels")
        if total_affinity_mat is None:
            total_affinity_mat = getCosAffinityMatrix(pre_embs)
        affinity_mat = total_affinity_mat[:, target_emb_index][target_emb_index, :]
        selected_inds, rest_inds = get_closest_embeddings(affinity_mat, merge_quantity)
        spk_cluster_labels, selected_embs = pre_clus_labels[target_emb_index], pre_embs[target_emb_index]
        index_mapping = (target_emb_index[rest_inds.sort()[0]], target_emb_index[selected_inds])
        merged_embs, merged_clus_labels = merge_vectors(selected_inds, selected_embs, spk_cluster_labels)
        if (org_size - merge_quantity) != merged_embs.shape[0]:
            raise ValueError("Reducer output is not matched to the target quantity")
    else:
        merged_embs = pre_embs[target_emb_index]
        merged_clus_labels = pre_clus_labels[target_emb_index]
        index_mapping = (target_emb_index, torch.arange(0))
    return merged_embs, merged_clus_labels, index_mapping


class LongFormSpeakerClustering:
    def __init__(self):
        self.speaker_clustering = SpeakerClustering()
        self.embeddings_in_scales: List[torch.Tensor] = []
        self.timestamps_in_scales: List[torch.Tensor] = []
        self.chunk_labels = {}  # long-form path: chunk index -> (first window, over-clustering labels); read by the parity tests

    @staticmethod
    def get_div_ceil_count(numer: int, denomin: int) -> int:
        return int(torch.ceil(torch.tensor(numer / denomin)).item())

    def check_input(self, embeddings_per_chunk, chunk_cluster_count, max_num_speakers) -> None:
        if chunk_cluster_count is None or embeddings_per_chunk is None:
            raise ValueError("chunk_cluster_count and embeddings_per_chunk should be set.")
        if chunk_cluster_count >= embeddings_per_chunk:
            raise ValueError("chunk_cluster_count should be smaller than embeddings_per_chunk.")
        if max_num_speakers <= 1:
            raise ValueError("max_num_speakers should be greater than 1.")
        if chunk_cluster_count <= max_num_speakers:
            raise ValueError("chunk_cluster_count should be greater than max_num_speakers.")

    def unpack_labels(self, Y_aggr, window_range_list, absolute_merge_mapping, org_len) -> torch.LongTensor:
        Y_unpack = torch.zeros((org_len,)).long()
        for win_rng, abs_mapping in zip(window_range_list, absolute_merge_mapping):
            inferred_merged_embs = Y_aggr[win_rng[0] : win_rng[1]]
            if len(abs_mapping[1]) > 0:
                Y_unpack[abs_mapping[1]] = inferred_merged_embs[-1].clone()  # merged vector's label
                if len(abs_mapping[0]) > 0:
                    Y_unpack[abs_mapping[0]] = inferred_merged_embs[:-1].clone()  # bypassed vectors
            else:
                Y_unpack[abs_mapping[0]] = inferred_merged_embs.clone()
        return Y_unpack

    def split_embs_to_windows(self, index: int, emb: torch.Tensor, embeddings_per_chunk: int):
        if embeddings_per_chunk * (index + 1) > emb.shape[0]:
            emb_part = emb[-1 * embeddings_per_chunk :]
            offset_index = emb.shape[0] - embeddings_per_chunk
        else:
            emb_part = emb[embeddings_per_chunk * index : embeddings_per_chunk * (index + 1)]
            offset_index = embeddings_per_chunk * index
        return emb_part, offset_index

    def forward_infer(
        self,
        embeddings_in_scales,
        timestamps_in_scales,
        multiscale_segment_counts,
        multiscale_weights,
        oracle_num_speakers: int = -1,
        max_rp_threshold: float = 0.15,
        max_num_speakers: int = 8,
        sparse_search_volume: int = 30,
        fixed_thres: float = -1.0,
        chunk_cluster_count=50,
        embeddings_per_chunk=10000,
    ) -> torch.LongTensor:
        if embeddings_per_chunk is not None and torch.max(multiscale_segment_counts) > embeddings_per_chunk:
            return self.long_forward_infer(
                embeddings_in_scales, timestamps_in_scales, multiscale_segment_counts, multiscale_weights,
                oracle_num_speakers, max_rp_threshold, max_num_speakers, sparse_search_volume, fixed_thres,
                int(chunk_cluster_count), int(embeddings_per_chunk),
            )
        cluster_labels = self.speaker_clustering.forward_infer(
            embeddings_in_scales=embeddings_in_scales,
            timestamps_in_scales=timestamps_in_scales,
            multiscale_segment_counts=multiscale_segment_counts,
            multiscale_weights=multiscale_weights,
            oracle_num_speakers=oracle_num_speakers,
            max_rp_threshold=max_rp_threshold,
            max_num_speakers=max_num_speakers,
            sparse_search_volume=sparse_search_volume,
            fixed_thres=fixed_thres,
        )
        self.timestamps_in_scales = self.speaker_clustering.timestamps_in_scales
        return cluster_labels

    def long_forward_infer(
        self, embeddings_in_scales, timestamps_in_scales, multiscale_segment_counts, multiscale_weights,
        oracle_num_speakers, max_rp_threshold, max_num_speakers, sparse_search_volume, fixed_thres,
        chunk_cluster_count, embeddings_per_chunk,
    ) -> torch.LongTensor:
        self.check_input(embeddings_per_chunk, chunk_cluster_count, max_num_speakers)
        self.embeddings_in_scales, self.timestamps_in_scales = split_input_data(embeddings_in_scales, timestamps_in_scales, multiscale_segment_counts)
        emb, _ = get_scale_interpolated_embs(multiscale_weights, self.embeddings_in_scales, self.timestamps_in_scales)
        offset_index, window_offset = 0, 0
        total_emb: List[torch.Tensor] = []
        window_range_list: List[List[int]] = []
        absolute_merge_mapping: List[List[torch.Tensor]] = []
        total_window_count = self.get_div_ceil_count(numer=emb.shape[0], denomin=embeddings_per_chunk)
        for win_index in range(total_window_count):
            emb_part, offset_index = self.split_embs_to_windows(index=win_index, emb=emb, embeddings_per_chunk=embeddings_per_chunk)
            if emb_part.shape[0] == 1:
                Y_part = torch.zeros((1,), dtype=torch.int64)
                mat = None
            else:
                mat = getCosAffinityMatrix(emb_part)
                overcluster_count = min(chunk_cluster_count, mat.shape[0])
                Y_part = self.speaker_clustering.forward_unit_infer(
                    mat=mat.clone(),
                    oracle_num_speakers=overcluster_count,
                    max_rp_threshold=max_rp_threshold,
                    max_num_speakers=chunk_cluster_count,
                    sparse_search_volume=sparse_search_volume,
                )
            self.chunk_labels[win_index] = (int(offset_index), Y_part.clone())
            num_to_be_merged = int(min(embeddings_per_chunk, emb_part.shape[0]) - chunk_cluster_count)
            min_count_per_cluster = self.get_div_ceil_count(numer=chunk_cluster_count, denomin=len(torch.unique(Y_part)))
            class_target_vol = get_merge_quantity(num_to_be_removed=num_to_be_merged, pre_clus_labels=Y_part, min_count_per_cluster=min_count_per_cluster)
            for spk_idx, merge_quantity in enumerate(list(class_target_vol)):
                merged_embs, _, index_mapping = run_reducer(
                    pre_embs=emb_part, target_spk_idx=spk_idx, merge_quantity=int(merge_quantity.item()), pre_clus_labels=Y_part, total_affinity_mat=mat
                )
                total_emb.append(merged_embs)
                absolute_merge_mapping.append([x + offset_index for x in index_mapping])
                window_range_list.append([window_offset, window_offset + merged_embs.shape[0]])
                window_offset += merged_embs.shape[0]
        reduced_embs = torch.cat(total_emb)
        reduced_mat = getCosAffinityMatrix(reduced_embs)
        Y_aggr = self.speaker_clustering.forward_unit_infer(
            mat=reduced_mat,
            oracle_num_speakers=oracle_num_speakers,
            max_rp_threshold=max_rp_threshold,
            max_num_speakers=max_num_speakers,
            sparse_search_volume=sparse_search_volume,
            fixed_thres=fixed_thres,
        )
        if reduced_embs.shape[0] != Y_aggr.shape[0]:
            raise ValueError("The number of embeddings and labels should be same")
        Y_unpack = self.unpack_labels(Y_aggr, window_range_list, absolute_merge_mapping, org_len=emb.shape[0])
        if Y_unpack.shape[0] != emb.shape[0]:
            raise ValueError("The number of raw input embeddings and labels should be same")
        return Y_unpack
