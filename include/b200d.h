/*
 * b200d.h -- C ABI of the B200-native diarization hot path (libb200d.so).
 *
 * Drop-in boundary for the path the reference reaches at
 *   /root/reference/diarize.py:200-201 and /root/reference/nemo_process.py:31-32
 *   (`NeuralDiarizer(cfg=create_config(dir)).to(device).diarize()` ->
 *    NeMo `ClusteringDiarizer(cfg).diarize()`),
 * configured by /root/reference/nemo_msdd_configs/diar_infer_*.yaml:39-56 and
 * /root/reference/helpers.py:252-303.  The reference itself has no FFI for this
 * path: the arithmetic lives in nemo-toolkit (PyTorch ops -> cuFFT/cuDNN/cuBLAS/
 * cuSOLVER).  Each entry point below therefore cites the upstream NeMo function
 * whose device work it replaces; INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *    its name ends in _host; the caller (PyTorch) owns all memory;
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*); no
 *    allocation, no retained pointers, no process-wide mutable state (kernel
 *    choice is a per-call flag; one-time kernel attributes are per device).
 *    Exceptions, each stated at its declaration: b200d_eig_bottomk(_sharded)
 *    synchronise `stream` once per outer iteration (host decisions on 2 b
 *    floats); b200d_peer_alloc / b200d_peer_free own the one buffer the
 *    library allocates (the peer buffer of the multi-GPU solver);
 *  - return 0 on success, a negative B200D_E* code on failure;
 *    b200d_last_error() returns a thread-local message;
 *  - sm_100a only; there is no CPU path.
 */
#ifndef B200D_H_
#define B200D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200D_OK 0
#define B200D_EINVAL (-1)   /* bad shape / alignment / null pointer          */
#define B200D_EARCH (-2)    /* device is not sm_100                           */
#define B200D_ELAUNCH (-3)  /* CUDA launch or driver error (see last_error)   */
#define B200D_EWORKSPACE (-4) /* workspace too small                          */

const char* b200d_version(void);
const char* b200d_last_error(void);
/* 0 if the current device is sm_100 and the driver entry points were resolved. */
int b200d_check_device(void);

#define B200D_FEAT_ZERO_PAD 1    /* constant-zero instead of reflect STFT padding */
#define B200D_FEAT_NO_PLUS_ONE 2 /* frame count floor(len / hop) instead of floor(len / hop) + 1 */
/* ------------------------------------------------------------------------------------------
 * Featurizer: replaces AudioToSpeechLabelDataset slicing + fixed_seq collate (tiling) and
 * FilterbankFeatures.forward (nemo/collections/asr/parts/preprocessing/features.py):
 * pre-emphasis 0.97 -> STFT(n_fft 512, hann 400, hop 160, center/reflect) -> |.|^2 ->
 * 80 slaney mels -> log(x + 2^-24) -> per-feature mean / unbiased std over the segment.
 *
 *  wav          float32 [n_wav]      whole recording(s), resident in HBM
 *  seg_start    int32   [n_seg]      first sample of each segment in `wav`
 *  seg_len      int32   [n_seg]      true sample count of each segment (>= 1)
 *  fixed_len    samples every segment is tiled up to (batch max, fixed_seq collate);
 *               frames T = fixed_len / 160 + 1   (fixed_len / 160 with B200D_FEAT_NO_PLUS_ONE)
 *  fb_start     int32   [80]         first FFT bin of each mel filter's support
 *  fb_off       int32   [81]         offsets of each filter's weights in fb_w (fb_off[80] == fb_nnz)
 *  fb_w         float32 [fb_nnz]     packed non-zero filterbank weights (librosa slaney), fb_nnz <= 1024
 *  window       float32 [400]        hann(400, periodic=False)
 *  variant      0 = FilterbankFeatures as in NeMo 1.x / early 2.x: torch.stft's reflect padding, frame count with "+ 1";
 *               B200D_FEAT_ZERO_PAD | B200D_FEAT_NO_PLUS_ONE = the later behaviour (the choice of transformers'
 *               ParakeetFeatureExtractor port): constant-zero padding of the pre-emphasised window, no "+ 1"
 *  out_f16      __half  [n_seg*T][ldo]  channels-last, channels 80..ldo-1 zeroed (ldo >= 80, ldo % 8 == 0)
 *  out_f32      float32 [n_seg][T][80] or NULL (parity taps)
 *  (these are the arguments of b200d_featurize_windows below; there is no separate per-window entry point)
 * ------------------------------------------------------------------------------------------ */
/* Every 10-ms frame is computed ONCE per recording instead of once per window (windows of all scales
 * overlap: upstream recomputes each STFT frame ~10 times).  A frame in the interior of a window depends only on where its 400
 * samples lie, not on the window, so:
 *   b200d_mel_stream         log(mel + 2^-24) of the frames of `n_streams` STREAMS: stream s holds the frames centred on samples
 *                            stream_start[s] + g * 160, g = 0 .. stream_off[s+1] - stream_off[s] - 1, as rows stream_off[s] + g of
 *                            logmel float32 [total_rows][80].  stream_off[] are multiples of 32 (pad each stream's frame count;
 *                            padding rows are computed from whatever samples lie there and never used).  Samples outside
 *                            [0, n_wav) read as 0.  stream_start int64 [n_streams], stream_off int32 [n_streams + 1] (device).
 *   b200d_featurize_windows  the featurizer above, with seg_row0 int32 [n_seg]: the row in `logmel` of the stream frame centred on the
 *                            segment's first sample (the host plans streams so that every full-length window starts on a frame of
 *                            one), or < 0.  Segments with seg_row0 >= 0 and seg_len == fixed_len copy their interior frames
 *                            (t*160 - 201 >= 0 and t*160 + 200 <= fixed_len) from the stream and compute only the two frames at
 *                            each end (reflect / zero padding, first-sample pre-emphasis); all others compute every frame.
 *                            n_on_stream: the caller orders the segments so that the first n_on_stream are the ones on a stream
 *                            (their edge frames: one CTA per window; the rest: every frame, spread over the chip); a segment placed
 *                            in the wrong part is still computed correctly, only slower.
 *                            logmel == NULL and seg_row0 == NULL together (n_on_stream ignored): every segment takes the generic path.
 *                            scratch float32 [n_seg][T][80]: un-normalised log-mel of the frames computed per window (two
 *                            kernels: frames spread over the chip one warp each, then one CTA per window gathers + normalises). */
int b200d_mel_stream(const float* wav, int64_t n_wav, const int64_t* stream_start, const int32_t* stream_off, int32_t n_streams,
                     int32_t total_rows, const int32_t* fb_start, const int32_t* fb_off, const float* fb_w, int32_t fb_nnz,
                     const float* window, float* logmel, void* stream);
int b200d_featurize_windows(const float* wav, int64_t n_wav, const float* logmel, const int32_t* seg_start, const int32_t* seg_len,
                            const int32_t* seg_row0, int32_t n_on_stream, int32_t n_seg, int32_t fixed_len, const int32_t* fb_start,
                            const int32_t* fb_off, const float* fb_w, int32_t fb_nnz, const float* window, int32_t variant, float* scratch,
                            void* out_f16, int32_t ldo, float* out_f32, void* stream);

/* ------------------------------------------------------------------------------------------
 * TitaNet-L building blocks (nemo/collections/asr/parts/submodules/jasper.py JasperBlock,
 * MaskedConv1d, SqueezeExcite; modules/conv_asr.py ConvASREncoder, SpeakerDecoder;
 * parts/submodules/tdnn_attention.py AttentivePoolLayer).  Activations are fp16,
 * channels-last [n_seg*T][C]; every segment of a call has the same frame count T.
 * ------------------------------------------------------------------------------------------ */

/* Depthwise 1-D conv over time (groups == C, 'same' zero padding inside each segment).
 *  x, y __half [n_seg*T][C];  w float32 [ksize][C] (tap-major);  C % 8 == 0; ksize odd <= 15 */
int b200d_depthwise_conv(const void* x, void* y, const float* w, int32_t n_seg, int32_t T, int32_t C,
                         int32_t ksize, void* stream);

/* Pointwise conv == GEMM on tcgen05 tensor cores, TMA-fed, TMEM accumulators:
 *    acc[m][n] = sum_k A[m][k] * W[n][k]      A __half [M][lda], W __half [N][ldw] (both K-major)
 * followed by one fused epilogue (fp32 math):
 *   B200D_EPI_BIAS       out16 = acc + bias[n]
 *   B200D_EPI_BIAS_RELU  out16 = relu(acc + bias[n])
 *   B200D_EPI_SE_RES     out16 = relu(aux16[m][n] * rowvec[m / rows_per_seg][n] + acc + bias[n])
 *                        (SE excitation * main branch + BN-folded residual conv, JasperBlock tail)
 *   B200D_EPI_TDNN       out16 = tanh(scale[n] * relu(acc + rowvec[m / rows_per_seg][n]) + shift[n])
 *   B200D_EPI_BIAS_F32   out32 = acc + bias[n]
 *   B200D_EPI_SIGMOID_F32 out32 = 1 / (1 + exp(-acc))          (SqueezeExcite gate)
 *   B200D_EPI_CHEB       Chebyshev / Laplacian step of the spectral solver (bf16 operands; see the struct below)
 * N % 128 == 0 (N % 256 == 0 uses 128x256 tiles), lda/ldw/ldo % 8 == 0; K tails are zero-filled by TMA. */
enum {
  B200D_EPI_BIAS = 0,
  B200D_EPI_BIAS_RELU = 1,
  B200D_EPI_SE_RES = 2,
  B200D_EPI_TDNN = 3,
  B200D_EPI_BIAS_F32 = 4,
  B200D_EPI_CHEB = 5,
  B200D_EPI_SIGMOID_F32 = 6
};

typedef struct b200d_gemm_epilogue {
  int32_t mode;
  int32_t rows_per_seg;   /* T for SE_RES / TDNN */
  const float* bias;      /* [N]  (BIAS*, SE_RES) */
  const float* scale;     /* [N]  (TDNN) */
  const float* shift;     /* [N]  (TDNN) */
  const float* rowvec;    /* [M / rows_per_seg][N] (SE_RES: sigmoid gate; TDNN: per-segment bias) */
  const void* aux16;      /* __half [M][ldo]  (SE_RES main branch) */
  /* CHEB: out32 = ca * (deg[m] * x32[m][c] - (A V)[m][c]) + cb * x32[m][c] + cc * xprev32[m][c] for the b block
   *       columns c; W (and, when vt != NULL, the output vt for the next step) is the 3-way bf16 split of V^T,
   *       part q of column c in row q * b + c ([hi | mid | lo]): b = 64 -> N = 192; b = 32 -> N = 128, the last
   *       32 rows zero.  vt [N][ldvt].                                                                          */
  const float* deg;       /* [M] */
  const float* x32;       /* [M][ldx] */
  const float* xprev32;   /* [M][ldx] or NULL */
  float ca, cb, cc;
  int32_t ldx;
  void* vt;               /* __nv_bfloat16 [N][ldvt] */
  int32_t ldvt;
  int32_t flags;          /* B200D_GEMM_* bits, 0 = default kernel choice */
  /* CHEB on a row shard (b200d_eig_bottomk_sharded): n_peers > 0 -> the launch's rows are one rank's share of a product whose
   * operands live at the same offsets of every rank's peer buffer (b200d_peer_group); vt -- and out with B200D_GEMM_PEER_OUT32 --
   * is stored at (address + peer_delta[r]) for r < n_peers (byte distance from this rank's buffer to rank r's mapping of its
   * own; one entry is 0 = the local copy).  n_peers == 0: plain local stores.                                                 */
  int32_t n_peers;
  int32_t pad_;
  int64_t peer_delta[8];
  float* colsum;          /* B200D_EPI_BIAS on the CTA-pair kernel only, else must be NULL: float32 [ceil(M / 32)][2][N] partial
                             column sums of the fp16-rounded output over each group of 32 rows -- [g][0] the rows of the window
                             (rows_per_seg >= 32 rows each) that row 32 g belongs to, [g][1] the rows of the next window; reduced
                             to per-window means by b200d_se_mean_from_colsum (SqueezeExcite's pool without re-reading the output) */
  /* CHEB: split-K form of the product.  splitk_ws != NULL (b200d_gemm_cheb_splitk_bytes(M, N, K) bytes, 16-byte aligned): the
   * launch is cut into (128-row tile, segment of 24 x 64 columns of K) units that fill the chip whatever M is; their raw fp32
   * accumulators go to the workspace and a second kernel sums them in ascending K order and applies the epilogue.  The
   * summation order of an output row then depends on K only -- the same bits for any M, i.e. on any number of GPUs.       */
  void* splitk_ws;
  uint64_t splitk_ws_bytes;
} b200d_gemm_epilogue;

/* Kernel choice is a PER-CALL property (no process-wide state): large GEMMs run on a cluster-launched CTA-pair kernel
 * (tcgen05 cta_group::2).  On B200 / driver 580.159 that kernel deadlocked the device when it shared the GPU with a
 * register-heavy kernel of ANOTHER stream (tools/concurrency_check.py), so a caller that keeps several streams busy at
 * once passes B200D_GEMM_NO_PAIR on the calls made inside that region and every such GEMM takes the 1-CTA kernel. */
#define B200D_GEMM_NO_PAIR 1
#define B200D_GEMM_PEER_OUT32 2 /* CHEB with n_peers > 0: also store the fp32 rows into every peer's buffer (default: bf16 split only) */

int b200d_gemm_f16(const void* A, int32_t lda, const void* W, int32_t ldw, int32_t M, int32_t N, int32_t K,
                   void* out, int32_t ldo, const b200d_gemm_epilogue* epi, void* stream);
size_t b200d_gemm_cheb_splitk_bytes(int32_t M, int32_t N, int32_t K);
/* SqueezeExcite = b200d_time_stats(with_std=0) -> b200d_gemm_f16(fc.0, BIAS_RELU with zero bias)
 * -> b200d_gemm_f16(fc.2, SIGMOID_F32) -> gate float32 [n_seg][C].                                 */

/* mean16 __half [n_seg][C] = per-window time mean from the colsum partials a B200D_EPI_BIAS GEMM left behind
 * (== b200d_time_stats(with_std = 0) of that GEMM's output up to fp32 summation order).  colsum float32 [ceil(n_seg*T/32)][2][C]. */
int b200d_se_mean_from_colsum(const float* colsum, int32_t n_seg, int32_t T, int32_t C, void* mean16, void* stream);

/* y = relu(x * gate[seg]) elementwise (block without residual: B0, B4).  x,y __half [n_seg*T][C] */
int b200d_se_apply_relu(const void* x, const float* gate, void* y, int32_t n_seg, int32_t T, int32_t C, void* stream);

/* y = relu(x * gate[seg]) and, in the same pass, stats16 __half [n_seg][2*C] = [mean | std] over time of y
 * (block 4 of the encoder feeding AttentivePoolLayer: SE scaling + get_statistics_with_mask fused).      */
int b200d_se_apply_relu_stats(const void* x, const float* gate, void* y, int32_t n_seg, int32_t T, int32_t C, void* stats16,
                              void* stream);

/* Per-segment statistics over time.  with_std == 0: mean, __half [n_seg][C] (SqueezeExcite pool).
 * with_std != 0: [mean | std] __half [n_seg][2*C] (get_statistics_with_mask with uniform weights;
 * input of the hoisted TDNN context term of AttentivePoolLayer).                                  */
int b200d_time_stats(const void* x, int32_t n_seg, int32_t T, int32_t C, int32_t with_std, void* out16, void* stream);

/* Attentive statistics pooling: alpha = softmax_t(e), mu = sum alpha x, sg = sqrt(clamp(sum alpha (x-mu)^2, 1e-10)).
 *  x, e __half [n_seg*T][C];  out16 __half [n_seg][2*C] = [mu | sg]                              */
int b200d_attn_pool(const void* x, const void* e, int32_t n_seg, int32_t T, int32_t C, void* out16, void* stream);

/* ------------------------------------------------------------------------------------------
 * TitaNet-L as two composite calls (label_models.EncDecSpeakerLabelModel.forward: preprocessor -> ConvASREncoder ->
 * SpeakerDecoder; upstream's `_extract_embeddings` dataloader loop calls it once per batch of 64 windows).
 *
 * b200d_titanet_pack_weights (HOST, one-time): upstream's state_dict as parallel arrays -- names[i] (NeMo's keys:
 *   `encoder.encoder.<b>.mconv.<j>.conv.weight`, `...mconv.<j>.{weight,bias,running_mean,running_var}`, `...mconv.<j>.fc.{0,2}.weight`,
 *   `encoder.encoder.<b>.res.0.{0.conv.weight,1.*}`, `decoder._pooling.attention_layer.{0.conv_layer.*,0.bn.*,2.*}`,
 *   `decoder.emb_layers.0.{0.*,1.*}`; optional `preprocessor.featurizer.{fb,window}`, computed when absent), data_host[i] float32
 *   host arrays in upstream's layouts, numel[i] their element counts (all shapes follow from them).  Writes the folded /
 *   padded / fp16 operands into packed_host (position independent: copy it to the device as is, 256-byte aligned) and the
 *   byte offset of every operand into *desc.  packed_host == NULL: only fills *desc (desc->packed_bytes is the size needed).
 * b200d_titanet_forward: embeddings of n_seg windows that all share fixed_len (see the featurizer section for wav / seg_* / variant,
 *   b200d_featurize_windows for logmel / seg_row0, both may be NULL).  Windows are processed in groups that fit `ws`
 *   (b200d_titanet_workspace_bytes(desc, max_frames, max_segs) holds groups of max_segs windows / max_frames frames; any
 *   size that holds one window works).  emb_out float32 [n_seg][ld_emb], first desc->emb columns written.  flags: B200D_GEMM_*.
 * b200d_titanet_mel_stream: b200d_mel_stream with the filterbank / window tables of the packed blob.
 * ------------------------------------------------------------------------------------------ */
#define B200D_TITANET_MAX_BLOCKS 8
typedef struct b200d_titanet_desc {
  int32_t n_blocks, feat_in, feat_pad, enc_out, attn, emb, emb_pad, fb_nnz;
  struct {
    int32_t cin, cin_pad, cout, repeat, ksize, residual, se_hidden, pad_;
    int64_t dw[3];    /* float32 [ksize][cin_pad] tap-major, -1 when ksize == 1 (folded into w) */
    int64_t w[3];     /* __half [cout][cin_pad], BatchNorm scale folded in */
    int64_t bias[3];  /* float32 [cout], BatchNorm shift */
    int64_t se_w1, se_w2, res_w, res_bias;
  } block[B200D_TITANET_MAX_BLOCKS];
  int64_t tdnn_wx, tdnn_wctx, tdnn_b, tdnn_scale, tdnn_shift, attn_w2, attn_b2, emb_w, emb_b, zeros;
  int64_t fb_start, fb_off, fb_w, window;
  int64_t packed_bytes;
} b200d_titanet_desc;

int b200d_titanet_pack_weights(int32_t n_tensors, const char* const* names, const float* const* data_host, const int64_t* numel,
                               b200d_titanet_desc* desc, void* packed_host, size_t packed_bytes);
size_t b200d_titanet_workspace_bytes(const b200d_titanet_desc* desc, int32_t max_frames, int32_t max_segs);
/* Windows of `fixed_len` samples that b200d_titanet_forward processes per launch group with a workspace of ws_bytes (a caller that
 * splits one batch of windows over several GPUs at multiples of this count reproduces the single-GPU launches exactly). */
int32_t b200d_titanet_group_windows(const b200d_titanet_desc* desc, size_t ws_bytes, int32_t fixed_len, int32_t variant);
int b200d_titanet_forward(const b200d_titanet_desc* desc, const void* packed_dev, const float* wav, int64_t n_wav, const float* logmel,
                          const int32_t* seg_start, const int32_t* seg_len, const int32_t* seg_row0, int32_t n_on_stream, int32_t n_seg,
                          int32_t fixed_len, int32_t variant, int32_t flags, float* emb_out, int32_t ld_emb, void* ws, size_t ws_bytes,
                          void* stream);
int b200d_titanet_mel_stream(const b200d_titanet_desc* desc, const void* packed_dev, const float* wav, int64_t n_wav,
                             const int64_t* stream_start, const int32_t* stream_off, int32_t n_streams, int32_t total_rows, float* logmel,
                             void* stream);

/* Per-kernel device time INSIDE the composite entry points (b200d_titanet_forward, b200d_eig_bottomk): between start and stop
 * every kernel they launch is bracketed by a cudaEvent pair on the caller's stream.  A measurement aid (bench.py's roofline
 * leg); adds event overhead, never active in a timed region.  stop synchronises the device, fills up to `capacity` spans and
 * sets *n_spans to the number recorded.  work = 2 M N K for GEMM launches, else 0.                                          */
typedef struct b200d_profile_span {
  char name[48];
  float ms;
  double work;
} b200d_profile_span;
/* Kernels launched so far from inside composite entry points (process-wide counter; callers count their own fine-grained calls). */
int64_t b200d_launch_count(void);
int b200d_profile_start(void);
int b200d_profile_stop(b200d_profile_span* spans, int32_t capacity, int32_t* n_spans);

/* ------------------------------------------------------------------------------------------
 * Affinity (nemo/collections/asr/parts/utils/offline_clustering.py):
 * cos_similarity + getCosAffinityMatrix + ScalerMinMax per scale, get_argmin_mat/getRepeatedList
 * mapping, getMultiScaleCosAffinityMatrix fusion.
 * ------------------------------------------------------------------------------------------ */

/* xn[i] = x[i] / (||x[i]|| + eps)                            x, xn float32 [n][d]           */
int b200d_l2_normalize(const float* x, float* xn, int32_t n, int32_t d, float eps, void* stream);

/* cos[i][j] = <xn_i, xn_j>, diagonal forced to 1; minmax[0] = min, minmax[1] = max over the whole
 * matrix (minmax must be initialised by the call itself).  cos float32 [n][n].                  */
int b200d_cos_affinity(const float* xn, int32_t n, int32_t d, float* cos, float* minmax, void* stream);

/* fused[i][j] = sum_s w_s * (cos_s[map_s[i]][map_s[j]] - min_s) / (max_s - min_s), all scales in one pass
 * over the N x N output (upstream materialises two repeat_interleave copies per scale).
 *  n_scales <= 8; the *_host arrays are HOST arrays of n_scales entries holding DEVICE pointers / sizes:
 *  cos_host[s] float32 [ns][ns], map_host[s] int32 [n_base] (non-decreasing coarse index of every base
 *  segment, get_argmin_mat + getRepeatedList), minmax_host[s] float32 [2] as written by b200d_cos_affinity. */
int b200d_fuse_scales(int32_t n_scales, const float* const* cos_host, const int32_t* ns_host, const int32_t* const* map_host,
                      const float* const* minmax_host, const float* weights_host, float* fused, int32_t n_base, void* stream);

/* Long-form path (longform_clustering.LongFormSpeakerClustering, offline_clustering.get_scale_interpolated_embs,
 * online_clustering.run_reducer / get_closest_embeddings):
 *  interp_scales  out[i] = sum_s w_s * emb_s[map_s[i]]      out float32 [n_base][d]; *_host arrays as in fuse_scales
 *  masked_rowsum  out[j] = sum_i [labels[i] == labels[j]] * mat[j][i]   (within-cluster affinity mass, all clusters at once) */
int b200d_interp_scales(int32_t n_scales, const float* const* emb_host, const int32_t* const* map_host, const float* weights_host,
                        float* out, int32_t n_base, int32_t d, void* stream);
int b200d_masked_rowsum(const float* mat, int32_t n, const int32_t* labels, float* out, void* stream);
/*  gather_segment_mean  out[s] = mean over i in [seg_off[s], seg_off[s+1]) of x[idx[i]] (online_clustering.merge_vectors for
 *                       every cluster of a chunk in one launch; fixed summation order).  x float32 [*][d], out [n_seg][d]  */
int b200d_gather_segment_mean(const float* x, int32_t d, const int32_t* idx, const int32_t* seg_off, int32_t n_seg, float* out,
                              void* stream);

/* ------------------------------------------------------------------------------------------
 * NME-SC graph + spectrum (getKneighborsConnections, getAffinityGraphMat, getLaplacian,
 * estimateNumofSpeakers, isGraphFullyConnected, SpectralClustering.getSpectralEmbeddings).
 * ------------------------------------------------------------------------------------------ */

/* rank[i][j] = position of column j in row i of mat sorted descending (ties: lower column first, the
 * order of torch.argsort(descending=True) on CPU), on the strided view mat[i*stride][j*stride], i,j < n
 * (n <= 1024: NMESC.subsampleAffinityMat keeps the sweep matrix below 2*nme_mat_size).
 * rank, rankT uint16 [n][n]; rankT is the transpose of rank.                                      */
int b200d_row_rank(const float* mat, int64_t ld, int32_t stride, int32_t n, void* rank_u16, void* rankT_u16, void* stream);

/* For each p in p_list (host array, np <= 64 entries): Laplacian L = D - A of A_p = 0.5*(B_p + B_p^T),
 * B_p[i][j] = rank[j][i] < p  (column-wise scatter of upstream), diagonal of A zeroed, degree
 * rounded through fp16 as upstream's half-precision graph does.  lap float32 [np][n][n].         */
int b200d_laplacian_from_rank(const void* rank_u16, const void* rankT_u16, int32_t n, const int32_t* p_list_host, int32_t np,
                              float* lap, void* stream);

/* reach[b] = number of nodes reachable from node 0 in the p_list[b]-neighbour graph given by the rank
 * matrices (isGraphFullyConnected / getTheLargestComponent; the whole list of getMinimumConnection in
 * one launch): frontier expansion, one CTA per p.  p_list_host: HOST array, np <= 64; reach int32 [np]. */
int b200d_graph_reach_rank(const void* rank_u16, const void* rankT_u16, int32_t n, const int32_t* p_list_host, int32_t np,
                           int32_t* reach, void* stream);

/* Batched symmetric eigenvalues (values only): chip-wide cooperative Householder tridiagonalisation
 * followed by Sturm multisection in fp64.
 *  a float32 [batch][n][n] (destroyed), batch <= 148, n <= 2048; evals float32 [batch][n_low + 1] = the
 *  n_low smallest eigenvalues ascending followed by the largest.  ws: b200d_eigvals_workspace_bytes(). */
size_t b200d_eigvals_workspace_bytes(int32_t batch, int32_t n);
int b200d_eigvals_batched(float* a, int32_t batch, int32_t n, int32_t n_low, float* evals, void* ws, size_t ws_bytes,
                          void* stream);
/* The same for a SUBSET of a larger batch: layout_batch >= batch is the size of the batch whose per-matrix work split (and with it
 * the fp32 reduction order) is used, so the subset's eigenvalues are bit for bit those the full batch gives -- the p values of the
 * NME sweep dealt to the ranks of a row-sharded recording ("NME sweep").                                   */
int b200d_eigvals_batched_layout(float* a, int32_t batch, int32_t layout_batch, int32_t n, int32_t n_low, float* evals, void* ws,
                                 size_t ws_bytes, void* stream);

/* Top-p binarisation of the full matrix (getAffinityGraphMat on the N x N fused affinity):
 * a[i][j] = 0.5*([j in top_p(row i)] + [i in top_p(row j)]), diagonal zeroed, as __nv_bfloat16 [n][lda]
 * (exact: entries are 0, 0.5, 1; columns n..lda-1 zeroed); deg float32 [n] = row sums, fp16-rounded as
 * upstream.  Radix select per row (no sort); sel uint8 [n][n] scratch.                             */
int b200d_topp_binarize(const float* mat, int32_t n, int32_t p, void* a_bf16, int32_t lda, float* deg, void* sel_u8,
                        void* stream);

/* One step of the Chebyshev-filtered subspace iteration for the bottom eigenvectors of L = D - A is
 * b200d_gemm_f16(A, vt, ..., B200D_EPI_CHEB).  The small dense helpers around it:                  */
/* g[b][b] = X^T Y (fp32 block partials, fixed-order fp64 combine); x,y float32 [n][ld]; b in {32, 64} */
size_t b200d_gram_workspace_bytes(int32_t n, int32_t b);
int b200d_gram(const float* x, const float* y, int32_t n, int32_t b, int32_t ld, float* g, void* ws, size_t ws_bytes, void* stream);
/* Symmetric eigendecomposition of g[b][b] (b even, <= 128) by parallel cyclic Jacobi in fp64:
 * evals ascending float32 [b], evecs float32 [b][b] (columns).  mode 1: scaled Cholesky instead --
 * returns Q = S R^-1 in evecs with (Y Q)^T (Y Q) = I for g = Y^T Y (CholQR).                       */
int b200d_small_eig(const float* g, int32_t b, float* evals, float* evecs, int32_t mode, void* stream);
/* y[n][b] = x[n][b] * q[b][b] (q NULL: y = x; y may alias x or be NULL); when vt_bf16 != NULL also the
 * 3-way bf16 split of y^T into vt_bf16 [3*b][ldvt], part q of column j in row q*b + j ([hi | mid | lo]):
 * the W operand of the next CHEB GEMM (N = 192 for b = 64; N = 128 for b = 32, rows 96..127 zero).       */
int b200d_right_mul(const float* x, int32_t n, int32_t b, int32_t ld, const float* q, float* y, void* vt_bf16, int32_t ldvt,
                    void* stream);
/* Sparse form of the same step, for graphs built from few neighbours (at most 2 p non-zeros per row):
 * b200d_csr_from_dense lists the non-zeros of a_bf16 [n][lda] (b200d_topp_binarize's output) row by row in
 * ascending column order: rowptr int32 [n + 1], colw uint32 [capacity] = column | (1u << 31 when the entry is 1,
 * else it is 0.5); capacity must be >= the number of non-zeros (2 n p always is).
 * b200d_spmm_cheb: out = ca * (deg .* x - A x) + cb * x + cc * xprev  (xprev may be NULL), x / xprev float32
 * [n][ldx], out float32 [n][ldo], b in {32, 64} columns: the arithmetic of the B200D_EPI_CHEB epilogue with the
 * product taken in fp32 over the CSR lists (upstream: the L @ V products inside torch.linalg.eigh,
 * offline_clustering.py getSpectralEmbeddings).                                                            */
int b200d_csr_from_dense(const void* a_bf16, int32_t n, int32_t lda, int32_t* rowptr, uint32_t* colw, int64_t capacity,
                         void* stream);
int b200d_spmm_cheb(const int32_t* rowptr, const uint32_t* colw, int32_t n, int32_t b, const float* deg, const float* x,
                    const float* xprev, int32_t ldx, float ca, float cb, float cc, float* out, int32_t ldo, void* stream);
/* out[j] = sum_r (w[r][j] - theta[j] * x[r][j])^2 : squared residual norms of the Ritz pairs (fixed summation order).
 * ws: at least ceil(n / 256) * 64 floats (b200d_gram_workspace_bytes(n, b) is always enough).              */
int b200d_resid_norms(const float* w, const float* x, const float* theta, int32_t n, int32_t b, int32_t ld, float* out,
                      void* ws, size_t ws_bytes, void* stream);

/* The whole solver as ONE call (SpectralClustering.getSpectralEmbeddings: upstream's eigh(N x N) + k columns kept): the k lowest
 * eigenvectors of L = diag(deg) - A for A = b200d_topp_binarize's output, by Chebyshev-filtered subspace iteration over the
 * kernels above (B200D_EPI_CHEB GEMMs, or b200d_spmm_cheb when the graph was built from p neighbours with 2 p <=
 * sparse_max_row_nnz or 2 p <= sparse_max_density * n; p <= 0: always dense).
 *  x     float32 [n][b], b = b200d_eig_bottomk_block(k) (32 for k <= 24, 64 for k <= 56): IN a random start block (the caller's
 *        RNG, e.g. torch.randn under a fixed seed), OUT orthonormal Ritz vectors, ascending; columns 0..k-1 span the answer.
 *  n >= 2 b.  Synchronises `stream` once per outer iteration (the polynomial degree and the stopping test are host decisions
 *  on 2 b floats read back).  opt == NULL: defaults; stats may be NULL.                                                     */
#define B200D_EIG_HISTORY 40
typedef struct b200d_eig_options {
  double tol;                /* stop when max_j<k ||L x_j - theta_j x_j|| <= tol * (2 max deg); default 2e-6 */
  int32_t max_outer;         /* default 40 */
  int32_t gemm_flags;        /* B200D_GEMM_* for the products */
  int32_t sparse_max_row_nnz;   /* default 32 */
  int32_t pad_;
  double sparse_max_density;    /* default 1/64 */
} b200d_eig_options;
typedef struct b200d_eig_stats {
  int32_t block, outer, gemms, converged, sparse;
  float max_resid;
  float history[B200D_EIG_HISTORY];  /* relative residual after each outer iteration */
} b200d_eig_stats;
int32_t b200d_eig_bottomk_block(int32_t k);
size_t b200d_eig_bottomk_workspace_bytes(int32_t n, int32_t k, int32_t p, const b200d_eig_options* opt);
int b200d_eig_bottomk(const void* a_bf16, int32_t lda, const float* deg, int32_t n, int32_t k, int32_t p, float* x, int32_t ldx,
                      const b200d_eig_options* opt, b200d_eig_stats* stats, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * One long recording on the GPUs of one NVSwitch box (the reference has no multi-GPU code): the N x N
 * affinity and its graph are ROW-SHARDED -- rank r holds rows [lo_r, hi_r) -- and the iterative eigensolver's products run on
 * the row shards, each rank storing its rows of the result straight into every peer's buffer over NVLink from the GEMM
 * epilogue (no NCCL call per product), with a device-side flag barrier between products.
 *
 * Peer memory: every rank allocates ONE buffer of the same size (b200d_peer_alloc: cudaMalloc + CUDA IPC handle, the one
 * allocation this library makes; released by b200d_peer_free), exchanges the 64-byte handles through the host (e.g.
 * torch.distributed.all_gather_object) and maps the others (b200d_peer_open / b200d_peer_close).  The first
 * B200D_PEER_HEADER_BYTES of each buffer are the barrier flags, the rest is laid out by the entry point that uses it.
 * b200d_peer_group: rank / world (<= 8), base[r] = rank r's buffer as mapped into THIS process (base[rank] = the local
 * allocation), bytes = size of each buffer, epoch = barrier counter advanced by the library (every rank must make the same
 * sequence of calls on the group), timeout_ms = how long a barrier waits for a missing rank before the call fails (0: 10 s).
 * ------------------------------------------------------------------------------------------ */
#define B200D_MAX_PEERS 8
#define B200D_PEER_HANDLE_BYTES 64
#define B200D_PEER_HEADER_BYTES 4096
#define B200D_PEER_STATUS_WORD 16 /* uint32 index in the header: set to 1 by a barrier that timed out */
typedef struct b200d_peer_group {
  int32_t rank, world;
  void* base[B200D_MAX_PEERS];
  uint64_t bytes;
  uint32_t epoch;
  uint32_t timeout_ms;
} b200d_peer_group;
int b200d_peer_alloc(size_t bytes, void** ptr, void* handle_host /* 64 bytes out, may be NULL */);
int b200d_peer_open(const void* handle_host, void** ptr);
int b200d_peer_close(void* ptr);
int b200d_peer_free(void* ptr);
/* Device-side barrier of the group on `stream` (one tiny kernel: system-scope release of this rank's epoch into every peer's
 * header, acquire-spin on its own); orders the peer stores of everything enqueued before it on every rank's stream. */
int b200d_peer_barrier(b200d_peer_group* grp, void* stream);
/* Synchronises `stream` and fails if a barrier of the group timed out. */
int b200d_peer_status(const b200d_peer_group* grp, void* stream);

/* Row-sharded forms of the affinity / graph steps: same arithmetic, bit for bit, as the full-matrix entry points above.
 *  cos_affinity_rows   rows [row_lo, row_hi) of b200d_cos_affinity: cos_rows float32 [row_hi - row_lo][n]; minmax[2] = min / max over
 *                      THOSE rows (the caller reduces over ranks: min of mins, max of maxes).
 *  fuse_scales_rows    rows [row_lo, row_hi) of b200d_fuse_scales; cos_host[s] holds the rows [cos_row0_host[s], ...) of scale s's
 *                      matrix (every row map_s[i], row_lo <= i < row_hi, must be there); minmax_host[s]: the GLOBAL min / max.
 *  topp_select_rows    the radix select of b200d_topp_binarize on m rows: sel uint8 [m][n] plus, per row, the threshold (as the
 *                      order-preserving uint32 code of the float) and the last column selected among the entries equal to it.
 *  sym_combine_rows    a_rows[i][j] = 0.5 ([j in top_p(row i)] + [i in top_p(row j)]) for the local rows i = row_lo + r, using the
 *                      symmetry of the affinity (mat[i][j] == mat[j][i] bit for bit: both sides of the cosine are the same fused
 *                      multiply-add chain) and the thresholds of ALL rows (thr_all / cut_all [n], gathered over the ranks);
 *                      deg_rows float32 [m], fp16-rounded.                                                                       */
int b200d_cos_affinity_rows(const float* xn, int32_t n, int32_t d, int32_t row_lo, int32_t row_hi, float* cos_rows, float* minmax,
                            void* stream);
int b200d_fuse_scales_rows(int32_t n_scales, const float* const* cos_host, const int32_t* ns_host, const int32_t* cos_row0_host,
                           const int32_t* const* map_host, const float* const* minmax_host, const float* weights_host,
                           float* fused_rows, int32_t n_base, int32_t row_lo, int32_t row_hi, void* stream);
int b200d_topp_select_rows(const float* mat_rows, int32_t m, int32_t n, int32_t p, void* sel_u8, uint32_t* thr, int32_t* cut,
                           void* stream);
int b200d_sym_combine_rows(const float* mat_rows, const void* sel_u8, const uint32_t* thr_all, const int32_t* cut_all, int32_t row_lo,
                           int32_t m, int32_t n, void* a_rows_bf16, int32_t lda, float* deg_rows, void* stream);

/* b200d_eig_bottomk on a row-sharded graph: a_rows_bf16 [row_hi - row_lo][lda] = this rank's rows of A, deg float32 [n] (all of
 * it), x float32 [n][b] IN the same start block on every rank, OUT the same Ritz vectors on every rank -- bit for bit those of
 * b200d_eig_bottomk on the whole graph (each output row's accumulation order does not depend on which rank computes it, and the
 * small dense steps run replicated).  Always dense products.  grp->bytes >= b200d_eig_bottomk_sharded_peer_bytes(n, k).
 * Synchronises `stream` once per outer iteration, like b200d_eig_bottomk.                                                   */
size_t b200d_eig_bottomk_sharded_peer_bytes(int32_t n, int32_t k);
int b200d_eig_bottomk_sharded(const void* a_rows_bf16, int32_t lda, const float* deg, int32_t n, int32_t row_lo, int32_t row_hi,
                              int32_t k, float* x, int32_t ldx, const b200d_eig_options* opt, b200d_eig_stats* stats,
                              b200d_peer_group* grp, void* stream);

/* k-means of kmeans_torch / kmeans_plusplus_torch with the RNG draws supplied by the host
 * (torch.manual_seed(0) stream: first-centre index, rand(30) per further centre, fallback randints).
 *  x float32 [n][dim] (dim, n_clusters <= 128; n_trials <= 32); labels int32 [n].                  */
int b200d_kmeans(const float* x, int32_t n, int32_t dim, int32_t n_clusters, int32_t first_center,
                 const float* rand_vals /*[n_clusters-1][n_trials]*/, int32_t n_trials,
                 const int32_t* fallback_idx /*[n_fallback]*/, int32_t n_fallback, int32_t iter_limit,
                 float threshold, int32_t* labels, void* ws, size_t ws_bytes, void* stream);
size_t b200d_kmeans_workspace_bytes(int32_t n, int32_t dim, int32_t n_clusters, int32_t n_trials);

#ifdef __cplusplus
}
#endif
#endif /* B200D_H_ */
