"""Oracle: multi-scale affinity + NME-SC spectral clustering.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates upstream `nemo/collections/asr/parts/utils/offline_clustering.py`
(SURVEY.md section 8 rows a9-a17): cos_similarity / ScalerMinMax /
getCosAffinityMatrix, get_argmin_mat / getRepeatedList /
getMultiScaleCosAffinityMatrix, getKneighborsConnections / getAffinityGraphMat,
getLaplacian / eigDecompose / getLamdaGaplist / estimateNumofSpeakers, the
connectivity helpers, NMESC, SpectralClustering, kmeans_plusplus_torch /
kmeans_torch, addAnchorEmb / getEnhancedSpeakerCount, split_input_data and
SpeakerClustering.{forward_infer, forward_unit_infer}.
The reference's YAML knobs feeding it: nemo_msdd_configs/diar_infer_*.yaml:47-56.
"""
from typing import List, Tuple

import torch

from . import switches


# ----------------------------------------------------------------------------- affinity
def cos_similarity(emb_a: torch.Tensor, emb_b: torch.Tensor, eps=None) -> torch.Tensor:
    if eps is None:
        eps = torch.tensor(switches.COS_EPS)
    if emb_a.shape[0] == 1 or emb_b.shape[0] == 1:
        raise ValueError("Number of feature vectors should be greater than 1")
    a_norm = emb_a / (torch.norm(emb_a, dim=1).unsqueeze(1) + eps)
    b_norm = emb_b / (torch.norm(emb_b, dim=1).unsqueeze(1) + eps)
    res = torch.mm(a_norm, b_norm.transpose(0, 1))
    res.fill_diagonal_(1)
    return res


def ScalerMinMax(X: torch.Tensor) -> torch.Tensor:
    v_min, v_max = X.min(), X.max()
    return (X - v_min) / (v_max - v_min)


def getCosAffinityMatrix(emb: torch.Tensor) -> torch.Tensor:
    if emb.shape[0] == 1:
        return torch.tensor([[1]]).to(emb.device)
    emb = emb.float()
    sim_d = cos_similarity(emb, emb)
    return ScalerMinMax(sim_d)


def get_argmin_mat(timestamps_in_scales: List[torch.Tensor]) -> List[torch.Tensor]:
    scale_list = list(range(len(timestamps_in_scales)))
    segment_anchor_list = [torch.mean(timestamps_in_scales[s], dim=1) for s in scale_list]
    base_scale_anchor = segment_anchor_list[max(scale_list)]
    session_scale_mapping_list = []
    for scale_idx in scale_list:
        curr_scale_anchor = segment_anchor_list[scale_idx]
        curr_mat = torch.tile(curr_scale_anchor, (base_scale_anchor.shape[0], 1))
        base_mat = torch.tile(base_scale_anchor, (curr_scale_anchor.shape[0], 1)).t()
        argmin_mat = torch.argmin(torch.abs(curr_mat - base_mat), dim=1)
        session_scale_mapping_list.append(argmin_mat)
    return session_scale_mapping_list


def getRepeatedList(mapping_argmat: torch.Tensor, score_mat_size: torch.Tensor) -> torch.Tensor:
    repeat_list = torch.zeros(int(score_mat_size), dtype=torch.int32)
    idxs, counts = torch.unique(mapping_argmat, return_counts=True)
    repeat_list[idxs] = counts.int()
    return repeat_list


def getMultiScaleCosAffinityMatrix(multiscale_weights, embeddings_in_scales, timestamps_in_scales) -> torch.Tensor:
    """Per scale: min-max scaled cosine matrix expanded to base resolution by
    repeat_interleave on both axes; fused as the *unnormalised* weighted sum (range [0, sum w])."""
    multiscale_weights = torch.squeeze(multiscale_weights, dim=0)
    session_scale_mapping_list = get_argmin_mat(timestamps_in_scales)
    n_base = len(timestamps_in_scales[-1])
    fused_sim_d = torch.zeros(n_base, n_base)
    for embeddings, weight, map_argmin in zip(embeddings_in_scales, multiscale_weights, session_scale_mapping_list):
        cosine_affinity_matrix = getCosAffinityMatrix(embeddings)
        repeat_list = getRepeatedList(map_argmin, torch.tensor(cosine_affinity_matrix.shape[0]))
        repeated_tensor_0 = torch.repeat_interleave(cosine_affinity_matrix, repeats=repeat_list, dim=0)
        repeated_tensor_1 = torch.repeat_interleave(repeated_tensor_0, repeats=repeat_list, dim=1)
        fused_sim_d += weight * repeated_tensor_1
    return fused_sim_d


# ----------------------------------------------------------------------------- graph
def getKneighborsConnections(affinity_mat: torch.Tensor, p_value: int) -> torch.Tensor:
    """Top-p per row via argsort(descending); scattered column-wise AND row-wise (mask_method 'binary')."""
    dim = affinity_mat.shape
    binarized = torch.zeros_like(affinity_mat)
    if switches.BINARIZE_HALF:
        binarized = binarized.half()
    p_value = int(p_value)
    sorted_matrix = torch.argsort(affinity_mat, dim=1, descending=True, stable=switches.ARGSORT_STABLE)[:, :p_value]
    binarized[sorted_matrix.T, torch.arange(affinity_mat.shape[0])] = 1
    indices_row = sorted_matrix[:, :p_value].flatten()
    indices_col = torch.arange(dim[1]).repeat(p_value, 1).T.flatten()
    binarized[indices_row, indices_col] = 1
    return binarized


def getAffinityGraphMat(affinity_mat_raw: torch.Tensor, p_value: int) -> torch.Tensor:
    X = affinity_mat_raw if p_value <= 0 else getKneighborsConnections(affinity_mat_raw, p_value)
    return 0.5 * (X + X.T)


def getLaplacian(X: torch.Tensor) -> torch.Tensor:
    """Unnormalised graph Laplacian L = D - A with the diagonal of A zeroed first."""
    X.fill_diagonal_(0)
    D = torch.sum(torch.abs(X), dim=1)
    D = torch.diag_embed(D)
    return D - X


def eigDecompose(laplacian: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    laplacian = laplacian.float()
    lambdas, diffusion_map = torch.linalg.eigh(laplacian)
    return lambdas, diffusion_map


def getLamdaGaplist(lambdas: torch.Tensor) -> torch.Tensor:
    if torch.is_complex(lambdas):
        lambdas = torch.real(lambdas)
    return lambdas[1:] - lambdas[:-1]


def estimateNumofSpeakers(affinity_mat: torch.Tensor, max_num_speakers: int):
    laplacian = getLaplacian(affinity_mat)
    lambdas, _ = eigDecompose(laplacian)
    lambdas = torch.sort(lambdas)[0]
    lambda_gap = getLamdaGaplist(lambdas)
    num_of_spk = torch.argmax(lambda_gap[: min(max_num_speakers, lambda_gap.shape[0])]) + 1
    return num_of_spk, lambdas, lambda_gap


def getTheLargestComponent(affinity_mat: torch.Tensor, seg_index: int) -> torch.Tensor:
    """Nodes reachable from `seg_index`.  Upstream is a python loop over frontier nodes OR-ing
    adjacency rows; this is the same breadth-first closure with the inner loop vectorised."""
    num_of_segments = affinity_mat.shape[0]
    adj = affinity_mat != 0
    connected_nodes = torch.zeros(num_of_segments, dtype=torch.bool)
    nodes_to_explore = torch.zeros(num_of_segments, dtype=torch.bool)
    nodes_to_explore[seg_index] = True
    for _ in range(num_of_segments):
        last_num_component = connected_nodes.sum()
        connected_nodes = torch.logical_or(connected_nodes, nodes_to_explore)
        if last_num_component >= connected_nodes.sum():
            break
        neighbors = adj[nodes_to_explore].any(dim=0)
        nodes_to_explore = torch.logical_or(nodes_to_explore, neighbors)
    return connected_nodes


def isGraphFullyConnected(affinity_mat: torch.Tensor) -> bool:
    return bool(getTheLargestComponent(affinity_mat, 0).sum() == affinity_mat.shape[0])


def getMinimumConnection(mat: torch.Tensor, max_N: torch.Tensor, n_list: torch.Tensor):
    """Raise p until the graph is connected (upstream quirk kept: connectivity is tested on
    the matrix of the *previous* p before the current one is built)."""
    p_neighbors = 1
    affinity_mat = getAffinityGraphMat(mat, p_neighbors)
    for p_neighbors in n_list:
        fully_connected = isGraphFullyConnected(affinity_mat)
        affinity_mat = getAffinityGraphMat(mat, int(p_neighbors))
        if fully_connected or p_neighbors > max_N:
            break
    return affinity_mat, p_neighbors


# ----------------------------------------------------------------------------- k-means
def getEuclideanDistance(specEmbA: torch.Tensor, specEmbB: torch.Tensor) -> torch.Tensor:
    A, B = specEmbA.unsqueeze(dim=1), specEmbB.unsqueeze(dim=0)
    dis = (A - B) ** 2.0
    return dis.sum(dim=-1).squeeze()


def kmeans_plusplus_torch(X: torch.Tensor, n_clusters: int, random_state: int, n_local_trials: int = 30):
    torch.manual_seed(random_state)
    n_samples, n_features = X.shape
    centers = torch.zeros(n_clusters, n_features, dtype=X.dtype)
    center_id = torch.randint(0, n_samples, (1,)).long()
    indices = torch.full([n_clusters], -1, dtype=torch.int)
    centers[0] = X[center_id].squeeze(0)
    indices[0] = center_id.squeeze(0)
    closest_dist_diff = centers[0, None].repeat(1, X.shape[0]).view(X.shape[0], -1) - X
    closest_dist_sq = closest_dist_diff.pow(2).sum(dim=1).unsqueeze(dim=0)
    current_pot = closest_dist_sq.sum()
    for c in range(1, n_clusters):
        rand_vals = torch.rand(n_local_trials) * current_pot.item()
        if len(closest_dist_sq.shape) > 1:
            torch_cumsum = torch.cumsum(closest_dist_sq, dim=1)[0]
        else:
            torch_cumsum = torch.cumsum(closest_dist_sq, dim=0)
        candidate_ids = torch.searchsorted(torch_cumsum, rand_vals)
        # upstream indexes X[candidate_ids] unguarded; an id == n_samples (rand beyond the
        # rounded cumsum) would raise there -- clamp, as sklearn's k-means++ does.
        candidate_ids = torch.clamp(candidate_ids, max=n_samples - 1)
        N_ci = candidate_ids.shape[0]
        distance_diff = X[candidate_ids].repeat(1, X.shape[0]).view(X.shape[0] * N_ci, -1) - X.repeat(N_ci, 1)
        distance = distance_diff.pow(2).sum(dim=1).view(N_ci, -1)
        distance_to_candidates = torch.minimum(closest_dist_sq, distance)
        candidates_pot = distance_to_candidates.sum(dim=1)
        best_candidate = torch.argmin(candidates_pot)
        current_pot = candidates_pot[best_candidate]
        closest_dist_sq = distance_to_candidates[best_candidate]
        best_candidate = candidate_ids[best_candidate]
        centers[c] = X[best_candidate]
        indices[c] = best_candidate
    return centers, indices


def kmeans_torch(X: torch.Tensor, num_clusters: int, threshold: float = 1e-4, iter_limit: int = 15, random_state: int = 0):
    X = X.float()
    input_size = X.shape[0]
    centers, _ = kmeans_plusplus_torch(X, n_clusters=num_clusters, random_state=random_state)
    selected_cluster_indices = torch.zeros(input_size).long()
    for _ in range(iter_limit):
        euc_dist = getEuclideanDistance(X, centers)
        if len(euc_dist.shape) <= 1:
            break
        selected_cluster_indices = torch.argmin(euc_dist, dim=1)
        center_inits = centers.clone()
        for index in range(num_clusters):
            selected_cluster = torch.nonzero(selected_cluster_indices == index).squeeze(1)
            chosen_indices = torch.index_select(X, 0, selected_cluster)
            if chosen_indices.shape[0] == 0:
                chosen_indices = X[torch.randint(len(X), (1,))]
            centers[index] = chosen_indices.mean(dim=0)
        center_delta_pow = torch.pow((centers - center_inits), 2)
        center_shift_pow = torch.pow(torch.sum(torch.sqrt(torch.sum(center_delta_pow, dim=1))), 2)
        if center_shift_pow < threshold:
            break
    return selected_cluster_indices


# ----------------------------------------------------------------------------- spectral clustering
class SpectralClustering:
    def __init__(self, n_clusters: int = 8, random_state: int = 0, n_random_trials: int = 1):
        self.n_clusters = n_clusters
        self.random_state = random_state
        self.n_random_trials = max(n_random_trials, 1)

    def forward(self, X) -> torch.Tensor:
        if X.shape[0] != X.shape[1]:
            raise ValueError("The affinity matrix is not a square matrix.")
        return self.clusterSpectralEmbeddings(X)

    def clusterSpectralEmbeddings(self, affinity) -> torch.Tensor:
        spectral_emb = self.getSpectralEmbeddings(affinity)
        labels_set = []
        for seed in range(self.random_state, self.random_state + self.n_random_trials):
            labels_set.append(kmeans_torch(X=spectral_emb, num_clusters=self.n_clusters, random_state=seed))
        stacked_labels = torch.stack(labels_set)
        label_index = torch.mode(torch.mode(stacked_labels, 0)[1])[0]
        return stacked_labels[label_index]

    def getSpectralEmbeddings(self, affinity_mat: torch.Tensor) -> torch.Tensor:
        laplacian = getLaplacian(affinity_mat)
        if switches.SPECTRAL_EIGH_FP64:  # test probe (off = upstream): how far do the LABELS depend on eigh's fp32 rounding?
            diffusion_map_ = torch.linalg.eigh(laplacian.double())[1].float()
        else:
            _, diffusion_map_ = eigDecompose(laplacian)
        diffusion_map = diffusion_map_[:, : self.n_clusters]
        inv_idx = torch.arange(diffusion_map.size(1) - 1, -1, -1).long()
        embedding = diffusion_map.T[inv_idx, :]
        return embedding[: self.n_clusters].T


# ----------------------------------------------------------------------------- NME analysis
class NMESC:
    """Normalized-maximum-eigengap analysis: p-neighbour sweep on a strided subsample."""

    def __init__(
        self,
        mat: torch.Tensor,
        max_num_speakers: int = 10,
        max_rp_threshold: float = 0.15,
        sparse_search: bool = True,
        sparse_search_volume: int = 30,
        nme_mat_size: int = 512,
        use_subsampling_for_nme: bool = True,
        fixed_thres: float = -1.0,
        maj_vote_spk_count: bool = False,
    ):
        self.max_num_speakers = max_num_speakers
        self.max_rp_threshold = max_rp_threshold
        self.use_subsampling_for_nme = use_subsampling_for_nme
        self.nme_mat_size = nme_mat_size
        self.sparse_search = sparse_search
        self.sparse_search_volume = sparse_search_volume
        self.min_p_value = torch.tensor(2)
        self.fixed_thres = fixed_thres
        self.eps = 1e-10
        self.max_N = torch.tensor(0)
        self.mat = mat
        self.p_value_list = self.min_p_value.unsqueeze(0)
        self.maj_vote_spk_count = maj_vote_spk_count

    def forward(self) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.use_subsampling_for_nme:
            subsample_ratio = self.subsampleAffinityMat(self.nme_mat_size)
        else:
            subsample_ratio = torch.tensor(1)
        self.p_value_list = self.getPvalueList()
        p_volume = self.p_value_list.shape[0]
        eig_ratio_list = torch.zeros(p_volume)
        est_num_of_spk_list = torch.zeros(p_volume)
        for p_idx, p_value in enumerate(self.p_value_list):
            est_spk_n, g_p = self.getEigRatio(p_value)
            est_num_of_spk_list[p_idx], eig_ratio_list[p_idx] = est_spk_n, g_p
        index_nn = torch.argmin(eig_ratio_list)
        rp_p_value = self.p_value_list[index_nn]
        affinity_mat = getAffinityGraphMat(self.mat, int(rp_p_value))
        if not isGraphFullyConnected(affinity_mat):
            affinity_mat, rp_p_value = getMinimumConnection(self.mat, self.max_N, self.p_value_list)
        p_hat_value = (subsample_ratio * rp_p_value).type(torch.int)
        if self.maj_vote_spk_count:
            est_num_of_spk = torch.mode(est_num_of_spk_list.clone())[0]
        else:
            est_num_of_spk = est_num_of_spk_list[index_nn]
        self.eig_ratio_list = eig_ratio_list  # kept for margin reporting in tests
        self.est_num_of_spk_list = est_num_of_spk_list
        return est_num_of_spk, p_hat_value

    def subsampleAffinityMat(self, nme_mat_size: int) -> torch.Tensor:
        subsample_ratio = torch.max(torch.tensor(1), torch.tensor(self.mat.shape[0] / nme_mat_size)).type(torch.int)
        self.mat = self.mat[:: subsample_ratio.item(), :: subsample_ratio.item()]
        return subsample_ratio

    def getEigRatio(self, p_neighbors) -> torch.Tensor:
        affinity_mat = getAffinityGraphMat(self.mat, int(p_neighbors))
        est_num_of_spk, lambdas, lambda_gap_list = estimateNumofSpeakers(affinity_mat, self.max_num_speakers)
        arg_sorted_idx = torch.argsort(lambda_gap_list[: self.max_num_speakers], descending=True)
        max_key = arg_sorted_idx[0]
        max_eig_gap = lambda_gap_list[max_key] / (torch.max(lambdas).item() + self.eps)
        g_p = (p_neighbors / self.mat.shape[0]) / (max_eig_gap + self.eps)
        return torch.stack([est_num_of_spk, g_p])

    def getPvalueList(self) -> torch.Tensor:
        if self.fixed_thres is not None and self.fixed_thres > 0.0:
            self.max_N = torch.max(torch.floor(torch.tensor(self.mat.shape[0] * self.fixed_thres)).type(torch.int), self.min_p_value)
            p_value_list = self.max_N.unsqueeze(0).int()
        else:
            self.max_N = torch.max(torch.floor(torch.tensor(self.mat.shape[0] * self.max_rp_threshold)).type(torch.int), self.min_p_value)
            if self.sparse_search:
                search_volume = torch.min(self.max_N, torch.tensor(self.sparse_search_volume).type(torch.int))
                N = torch.max(search_volume, torch.tensor(2))
                steps = min(self.max_N, N)
                p_value_list = torch.linspace(start=1, end=self.max_N, steps=int(steps)).type(torch.int)
            else:
                p_value_list = torch.arange(1, self.max_N + 1)
        return p_value_list


# ----------------------------------------------------------------------------- enhanced count
def addAnchorEmb(emb: torch.Tensor, anchor_sample_n: int, anchor_spk_n: int, sigma: float) -> torch.Tensor:
    emb_dim = emb.shape[1]
    std_org = torch.std(emb, dim=0)
    sigma = torch.tensor(sigma)
    new_emb_list = []
    for _ in range(anchor_spk_n):
        emb_m = torch.tile(torch.randn(1, emb_dim), (anchor_sample_n, 1))
        emb_noise = torch.randn(anchor_sample_n, emb_dim).T
        emb_noise = torch.matmul(torch.diag(std_org), emb_noise / torch.max(torch.abs(emb_noise), dim=0)[0].unsqueeze(0)).T
        emb_gen = emb_m + sigma * emb_noise
        new_emb_list.append(emb_gen)
    new_emb_list.append(emb)
    return torch.vstack(new_emb_list)


def getEnhancedSpeakerCount(emb: torch.Tensor, random_test_count: int = 5, anchor_spk_n: int = 3, anchor_sample_n: int = 10, sigma: float = 50) -> torch.Tensor:
    est_num_of_spk_list: List[int] = []
    for seed in range(random_test_count):
        torch.manual_seed(seed)
        emb_aug = addAnchorEmb(emb, anchor_sample_n, anchor_spk_n, sigma)
        mat = getCosAffinityMatrix(emb_aug)
        nmesc = NMESC(mat, max_num_speakers=emb.shape[0], max_rp_threshold=0.15, sparse_search=True, sparse_search_volume=10, fixed_thres=-1.0, nme_mat_size=300)
        est_num_of_spk, _ = nmesc.forward()
        est_num_of_spk_list.append(int(est_num_of_spk.item()))
    comp_est_num_of_spk = torch.tensor(max(torch.mode(torch.tensor(est_num_of_spk_list))[0].item() - anchor_spk_n, 1))
    return comp_est_num_of_spk


def split_input_data(embeddings_in_scales, timestamps_in_scales, multiscale_segment_counts):
    split_index = multiscale_segment_counts.tolist()
    return list(torch.split(embeddings_in_scales, split_index, dim=0)), list(torch.split(timestamps_in_scales, split_index, dim=0))


# ----------------------------------------------------------------------------- top level
class SpeakerClustering:
    def __init__(self, min_samples_for_nmesc=None, nme_mat_size=None, sparse_search=True, maj_vote_spk_count=False):
        self.min_samples_for_nmesc = switches.MIN_SAMPLES_FOR_NMESC if min_samples_for_nmesc is None else min_samples_for_nmesc
        self.nme_mat_size = switches.NME_MAT_SIZE if nme_mat_size is None else nme_mat_size
        self.sparse_search = sparse_search
        self.maj_vote_spk_count = maj_vote_spk_count
        self.embeddings_in_scales: List[torch.Tensor] = []
        self.timestamps_in_scales: List[torch.Tensor] = []
        self.debug = {}

    def forward_unit_infer(
        self,
        mat: torch.Tensor,
        oracle_num_speakers: int = -1,
        max_num_speakers: int = 8,
        max_rp_threshold: float = 0.15,
        sparse_search_volume: int = 30,
        est_num_of_spk_enhanced: torch.Tensor = torch.tensor(-1),
        fixed_thres: float = -1.0,
        kmeans_random_trials: int = 1,
    ) -> torch.LongTensor:
        nmesc = NMESC(
            mat,
            max_num_speakers=max_num_speakers,
            max_rp_threshold=max_rp_threshold,
            sparse_search=self.sparse_search,
            sparse_search_volume=sparse_search_volume,
            fixed_thres=fixed_thres,
            nme_mat_size=self.nme_mat_size,
            maj_vote_spk_count=self.maj_vote_spk_count,
        )
        if mat.shape[0] > self.min_samples_for_nmesc:
            est_num_of_spk, p_hat_value = nmesc.forward()
            affinity_mat = getAffinityGraphMat(mat, int(p_hat_value))
        else:
            nmesc.fixed_thres = max_rp_threshold
            est_num_of_spk, p_hat_value = nmesc.forward()
            affinity_mat = mat
        if oracle_num_speakers > 0:
            n_clusters = int(oracle_num_speakers)
        elif est_num_of_spk_enhanced > 0:
            n_clusters = int(est_num_of_spk_enhanced.item())
        else:
            n_clusters = int(est_num_of_spk.item())
        self.debug = {
            "est_num_of_spk": int(est_num_of_spk.item()),
            "p_hat": int(p_hat_value),
            "n_clusters": n_clusters,
            "g_p": getattr(nmesc, "eig_ratio_list", None),
            "p_list": nmesc.p_value_list,
        }
        spectral_model = SpectralClustering(n_clusters=n_clusters, n_random_trials=kmeans_random_trials)
        return spectral_model.forward(affinity_mat)

    def forward_infer(
        self,
        embeddings_in_scales: torch.Tensor,
        timestamps_in_scales: torch.Tensor,
        multiscale_segment_counts: torch.LongTensor,
        multiscale_weights: torch.Tensor,
        oracle_num_speakers: int = -1,
        max_rp_threshold: float = 0.15,
        max_num_speakers: int = 8,
        enhanced_count_thres: int = None,
        sparse_search_volume: int = 30,
        fixed_thres: float = -1.0,
    ) -> torch.LongTensor:
        if enhanced_count_thres is None:
            enhanced_count_thres = switches.ENHANCED_COUNT_THRES
        self.embeddings_in_scales, self.timestamps_in_scales = split_input_data(embeddings_in_scales, timestamps_in_scales, multiscale_segment_counts)
        emb = self.embeddings_in_scales[-1]
        if emb.shape[0] == 1:
            return torch.zeros((1,), dtype=torch.int64)
        elif emb.shape[0] <= max(enhanced_count_thres, self.min_samples_for_nmesc) and oracle_num_speakers < 0:
            est_num_of_spk_enhanced = getEnhancedSpeakerCount(emb=emb)
        else:
            est_num_of_spk_enhanced = torch.tensor(-1)
        if oracle_num_speakers > 0:
            max_num_speakers = oracle_num_speakers
        mat = getMultiScaleCosAffinityMatrix(multiscale_weights, self.embeddings_in_scales, self.timestamps_in_scales)
        self.fused_affinity = mat
        return self.forward_unit_infer(
            mat=mat,
            oracle_num_speakers=oracle_num_speakers,
            max_rp_threshold=max_rp_threshold,
            max_num_speakers=max_num_speakers,
            sparse_search_volume=sparse_search_volume,
            est_num_of_spk_enhanced=est_num_of_spk_enhanced,
            fixed_thres=fixed_thres,
        )
