"""Named switches for upstream-NeMo details recalled from memory.

Each constant selects between behaviours that different NeMo 1.x/2.x releases
(or an imperfect recollection) could have.  The product path implements the
default value of every switch; tests flip them only on the oracle side to show
which ones change results.
"""

# features.FilterbankFeatures.forward: 2.x masks the pre-emphasised signal beyond
# seq_len ("timemask") -- irrelevant under fixed_seq collate (all lengths equal).
PREEMPH_TIMEMASK = True

# features.FilterbankFeatures: torch.stft(center=True) with its default reflect padding, and
# get_seq_len = floor((len + 2 * (n_fft // 2) - n_fft) / hop) + 1.  Hugging Face's port of a 2025 NeMo main
# (transformers 5.5 ParakeetFeatureExtractor, in this image) has pad_mode="constant" and no "+ 1", so later 2.x
# releases may differ in the first / last two frames of every window and in the frame count the per-feature
# statistics run over.  tests/test_cpu_oracle.py flips both and reproduces that port's features; the CUDA featurizer
# implements the defaults below.
STFT_PAD_MODE = "reflect"  # or "constant"
SEQ_LEN_PLUS_ONE = True

# label_models.EncDecSpeakerLabelModel.__setup_dataloader_from_config uses
# dataset.fixed_seq_collate_fn: short segments in a batch are *tiled* (repeated)
# up to the batch max length, and every length becomes that max.
# "pad" would be the generic _speech_collate_fn (zero pad, true lengths).
COLLATE = "fixed_seq"  # or "pad"

# speaker_utils.get_subsegments: "classic" python loop (<= 1.2x) or the
# torch.arange/round(decimals=2) variant that appeared in 2.x.
SUBSEGMENT_RULE = "classic"  # or "v2"
MIN_SUBSEGMENT_DURATION = 0.05

# offline_clustering.getKneighborsConnections builds the binarized matrix in
# half precision; the degree vector of getLaplacian is therefore rounded to fp16.
BINARIZE_HALF = True

# offline_clustering.getLaplacian: unnormalised L = D - A.
LAPLACIAN = "unnormalized"

# offline_clustering.cos_similarity eps added to the norm.
COS_EPS = 3.5e-4

# SpeakerClustering / NMESC defaults (not forwarded from YAML by perform_clustering)
MIN_SAMPLES_FOR_NMESC = 6
NME_MAT_SIZE = 512
ENHANCED_COUNT_THRES = 80

# offline_clustering.getKneighborsConnections calls torch.argsort(descending=True) WITHOUT stable=True.
# On CPU that sort is not stable (tie order is an implementation detail of the torch build: it differs
# between CPU and CUDA and between releases), so upstream's choice among exactly equal affinities at the
# p-th position is unspecified.  The restatement pins it to the stable order (lower column first), which
# is also what the CUDA path implements; set False to get torch's native (unspecified) tie order.
ARGSORT_STABLE = True

# TEST PROBE, not an upstream behaviour: run the eigh of SpectralClustering.getSpectralEmbeddings in float64 (result cast
# back to float32).  Upstream's is fp32; where the k lowest eigenvalues have no gap to the next ones (the long-form
# path over-clusters 10 000 windows into 50), fp32 LAPACK rounding decides which way near-degenerate eigenvectors mix, and
# a handful of k-means assignments with it.  Tests flip this to measure how many labels of the ORACLE ITSELF hang on that
# rounding (the yardstick for label differences between the CPU and the B200 path at that size).
SPECTRAL_EIGH_FP64 = False
