// Log-mel featurizer: pre-emphasis -> STFT -> |.|^2 -> mel -> log -> per-feature normalise, every 10-ms frame computed ONCE.
//
// Upstream featurizes each window of each scale on its own, so a recording's STFT frames are recomputed ~10 times: windows
// start every 0.25 s (25 hops) and overlap by half, at 5-6 scales.  A frame in the INTERIOR of a window does not depend on
// the window at all -- it sees neither the window's reflect padding nor its "first sample has no predecessor" pre-emphasis
// rule -- only on where its 400 samples lie.  So:
//
//   mel_stream_kernel        log-mel frames on the 160-sample grid of a "stream" (a contiguous stretch of speech at one grid
//                            phase; the host plans streams so that every full-length window starts on a frame of one).  One
//                            CTA = 32 consecutive frames, one warp per frame; the 5.4 k samples the CTA's frames share are
//                            staged in shared memory with coalesced loads (each sample read from HBM once per stream).
//   featurize_windows_kernel one CTA per window: gathers the window's T frames from its stream (interior frames), computes
//                            only the <= 2 + 2 EDGE frames whose support crosses the window's ends (reflect / zero padding,
//                            pre-emphasis start), then the exact two-pass per-feature mean / unbiased std over the window and
//                            the fp16 channels-last rows for the first TitaNet conv.  Windows that are not on a stream (tiled
//                            up to the batch length by fixed_seq collate, or off-grid) take the generic path: all T frames
//                            computed by the CTA, as the round-1 kernel did for every window.
//
// The 512-point real FFT of a frame is a 256-point complex FFT held in registers (8-point DFT per lane) + warp shuffles
// (32-point DIF across lanes) and one shared-memory hop for the real-FFT recombination; interior and edge frames run the
// same code on the same sample values, so a window's features do not depend on which path produced a frame.
//
// Replaces AudioToSpeechLabelDataset.__getitem__ + fixed_seq collate (audio_to_label.py) and
// FilterbankFeatures.forward (parts/preprocessing/features.py) of upstream NeMo.
#include "common.cuh"

namespace b200d {

constexpr int kNFFT = 512;
constexpr int kHop = 160;
constexpr int kWin = 400;
constexpr int kWinOff = (kNFFT - kWin) / 2;  // 56: torch.stft centres the window inside n_fft
constexpr int kMels = 80;
constexpr int kBins = kNFFT / 2 + 1;  // 257
constexpr int kMaxFbNnz = 1024;
constexpr int kWarpScratch = 512 + kBins + 3;  // floats: 256 complex spectrum + 257 power bins (+ pad)
constexpr int kStreamWarps = 32;               // frames per CTA of the stream kernel
constexpr int kStreamSpan = (kStreamWarps - 1) * kHop + kWin + 1;  // samples the 32 frames share (+ 1 for the pre-emphasis)
constexpr int kWinWarps = 8;                   // warps per CTA of the window kernel

struct FbTables {
  const int* fb_start;  // [80] first bin of each filter
  const int* fb_off;    // [81] offsets into fb_w
  const float* fb_w;    // packed non-zero weights
  const float* window;  // [400]
};

struct FbShared {
  float2 tw[256];  // e^{-2 pi i k / 512}
  float win[kWin];
  float fbw[kMaxFbNnz];
  int fbs[kMels], fbo[kMels + 1];
};

__device__ __forceinline__ void load_tables(FbShared& s, const FbTables& t) {
  for (int k = threadIdx.x; k < 256; k += blockDim.x) {
    float sn, cs;
    sincospif(static_cast<float>(k) / 256.f, &sn, &cs);
    s.tw[k] = make_float2(cs, -sn);
  }
  for (int k = threadIdx.x; k < kWin; k += blockDim.x) s.win[k] = t.window[k];
  for (int k = threadIdx.x; k < kMels; k += blockDim.x) s.fbs[k] = t.fb_start[k];
  for (int k = threadIdx.x; k <= kMels; k += blockDim.x) s.fbo[k] = t.fb_off[k];
  const int nnz = t.fb_off[kMels];
  for (int k = threadIdx.x; k < nnz; k += blockDim.x) s.fbw[k] = t.fb_w[k];
}

__device__ __forceinline__ float preemph(float x0, float xm1) { return x0 - 0.97f * xm1; }

// Lane-constant twiddles of the in-register / cross-lane FFT.
struct LaneTwiddles {
  float2 tw2[8];  // W256^(lane * k1)
  float2 tw3[5];  // stage twiddles of the cross-lane FFT: W32^((lane & (h-1)) * (16/h)), h = 16, 8, 4, 2, 1
  int k2;         // bitrev5(lane)
  __device__ __forceinline__ void init(int lane) {
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) {
      float sn, cs;
      sincospif(static_cast<float>(lane * k1) / 128.f, &sn, &cs);
      tw2[k1] = make_float2(cs, -sn);
    }
#pragma unroll
    for (int st = 0; st < 5; ++st) {
      const int h = 16 >> st;
      float sn, cs;
      sincospif(static_cast<float>((lane & (h - 1)) * (16 / h)) / 16.f, &sn, &cs);
      tw3[st] = make_float2(cs, -sn);
    }
    k2 = static_cast<int>(__brev(static_cast<unsigned>(lane)) >> 27);
  }
};

// One frame by one warp: `sample(i)` is the pre-emphasised (padded) signal at FFT position i in [56, 456) of this frame;
// writes log(mel + 2^-24) for the 80 filters to out[0..79] (shared or global memory).
//   256-point complex FFT of z[n] = v[2n] + i v[2n+1] without shared-memory stages: n = lane + 32 j.
//     step 1  8-point DFT over j in registers                 A[k1]  = sum_j z[j] W8^(j k1)
//     step 2  twiddle                                          B[k1]  = A[k1] W256^(lane k1)
//     step 3  32-point DIF FFT across the lanes by shuffles    X[k1 + 8 k2], k2 = bitrev5(lane)
template <typename SampleFn>
__device__ __forceinline__ void frame_logmel(const FbShared& s, const LaneTwiddles& lt, float* scratch, int lane, SampleFn sample, float* out) {
  float2* cbuf = reinterpret_cast<float2*>(scratch);
  float* pbuf = scratch + 512;
  float2 z[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int i0 = 2 * (lane + 32 * j), i1 = i0 + 1;
    float v0 = 0.f, v1 = 0.f;
    if (i0 >= kWinOff && i0 < kWinOff + kWin) v0 = s.win[i0 - kWinOff] * sample(i0);
    if (i1 >= kWinOff && i1 < kWinOff + kWin) v1 = s.win[i1 - kWinOff] * sample(i1);
    z[j] = make_float2(v0, v1);
  }
  // ---- step 1: 8-point DIF DFT in registers (outputs in bit-reversed order, undone by the index map below)
  auto bf = [](float2& a, float2& b) { const float2 sm = make_float2(a.x + b.x, a.y + b.y), d = make_float2(a.x - b.x, a.y - b.y); a = sm; b = d; };
  auto mul = [](float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); };
  const float r2 = 0.70710678118654752f;
  bf(z[0], z[4]); bf(z[1], z[5]); bf(z[2], z[6]); bf(z[3], z[7]);
  z[5] = mul(z[5], make_float2(r2, -r2));      // W8^1
  z[6] = make_float2(z[6].y, -z[6].x);         // W8^2 = -i
  z[7] = mul(z[7], make_float2(-r2, -r2));     // W8^3
  bf(z[0], z[2]); bf(z[1], z[3]); bf(z[4], z[6]); bf(z[5], z[7]);
  z[3] = make_float2(z[3].y, -z[3].x);
  z[7] = make_float2(z[7].y, -z[7].x);
  bf(z[0], z[1]); bf(z[2], z[3]); bf(z[4], z[5]); bf(z[6], z[7]);
  // z[q] now holds A[bitrev3(q)]
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int k1 = ((q & 1) << 2) | (q & 2) | ((q >> 2) & 1);
    float2 b = mul(z[q], lt.tw2[k1]);
    // ---- step 3: 32-point DIF FFT across lanes
#pragma unroll
    for (int st = 0; st < 5; ++st) {
      const int h = 16 >> st;
      const float ox = __shfl_xor_sync(0xffffffffu, b.x, h), oy = __shfl_xor_sync(0xffffffffu, b.y, h);
      if (lane & h) b = mul(make_float2(ox - b.x, oy - b.y), lt.tw3[st]);
      else b = make_float2(b.x + ox, b.y + oy);
    }
    cbuf[k1 + 8 * lt.k2] = b;
  }
  __syncwarp();
  // real-FFT recombination: X[k] = E[k] + e^{-2 pi i k/512} O[k], power spectrum |X|^2
#pragma unroll
  for (int q = 0; q < 9; ++q) {
    const int k = lane + 32 * q;
    if (k <= 256) {
      const float2 zk = cbuf[k & 255];
      const float2 zc = cbuf[(256 - k) & 255];
      const float er = 0.5f * (zk.x + zc.x), ei = 0.5f * (zk.y - zc.y);
      const float orr = 0.5f * (zk.y + zc.y), oi = -0.5f * (zk.x - zc.x);
      const float2 w = (k < 256) ? s.tw[k] : make_float2(-1.f, 0.f);
      const float xr = er + w.x * orr - w.y * oi;
      const float xi = ei + w.x * oi + w.y * orr;
      pbuf[k] = xr * xr + xi * xi;
    }
  }
  __syncwarp();
  for (int m = lane; m < kMels; m += 32) {
    const int k0 = s.fbs[m], o0 = s.fbo[m], n = s.fbo[m + 1] - o0;
    float acc = 0.f;
    for (int i = 0; i < n; ++i) acc = fmaf(s.fbw[o0 + i], pbuf[k0 + i], acc);
    out[m] = logf(acc + 5.9604644775390625e-08f);  // log(x + 2^-24)
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------ stream frames
struct StreamParams {
  const float* wav;
  long long n_wav;
  const long long* stream_start;  // [n_streams] sample at the centre of frame 0 of each stream
  const int* stream_off;          // [n_streams + 1] first row of each stream in `logmel` (multiples of 32)
  int n_streams;
  FbTables tab;
  float* logmel;                  // [stream_off[n_streams]][80]
};

__global__ void __launch_bounds__(kStreamWarps * 32) mel_stream_kernel(const StreamParams p) {
  extern __shared__ float dyn[];  // samples [kStreamSpan (+pad)] | per-warp FFT scratch
  __shared__ FbShared tabs;
  load_tables(tabs, p.tab);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kStreamWarps;
  // the stream this CTA's 32 rows belong to: last s with stream_off[s] <= row0 (uniform across the CTA)
  int lo = 0, hi = p.n_streams - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(p.stream_off + mid) <= row0) lo = mid; else hi = mid - 1;
  }
  const int g0 = row0 - __ldg(p.stream_off + lo);
  // smp[i] = wav[first + i], first = centre of frame g0 - 200 - 1
  const long long first = __ldg(p.stream_start + lo) + static_cast<long long>(g0) * kHop - kWin / 2 - 1;
  float* smp = dyn;
  for (int i = threadIdx.x; i < kStreamSpan; i += blockDim.x) {
    const long long idx = first + i;
    smp[i] = (idx >= 0 && idx < p.n_wav) ? __ldg(p.wav + idx) : 0.f;
  }
  __syncthreads();
  LaneTwiddles lt;
  lt.init(lane);
  float* scratch = dyn + ((kStreamSpan + 3) & ~3) + warp * kWarpScratch;
  const float* fs = smp + warp * kHop + 1 - kWinOff;  // FFT position i of this warp's frame is sample fs[i]
  frame_logmel(tabs, lt, scratch, lane, [&](int i) { return preemph(fs[i], fs[i - 1]); },
               p.logmel + static_cast<size_t>(row0 + warp) * kMels);
}

// ------------------------------------------------------------------------------------ windows
struct FeatParams {
  const float* wav;
  long long n_wav;
  const int* seg_start;
  const int* seg_len;
  const int* seg_row0;   // row in `logmel` of the stream frame centred on each segment's first sample, or < 0; may be NULL
  const float* logmel;   // stream frames (mel_stream_kernel), may be NULL
  int n_seg;
  int fixed_len;
  int T;
  int zero_pad;  // STFT padding of the (pre-emphasised) window: 0 = reflect (torch.stft's default), 1 = zeros
  FbTables tab;
  float* scratch;        // [n_seg][T][80] un-normalised log-mel of the frames computed per window
  __half* out16;
  int ldo;
  float* out32;
};

// Interior frames of a window of F samples / T frames: support [t*160 - 200 - 1, t*160 + 200) inside [0, F) -- those are
// identical to the stream's frames.  Empty range (lo > hi) when the window has no stream.
__device__ __forceinline__ void interior_range(const FeatParams& p, int seg, int& t_lo, int& t_hi, int& row0) {
  const int F = p.fixed_len;
  row0 = (p.seg_row0 && p.logmel && p.seg_len[seg] >= F) ? p.seg_row0[seg] : -1;
  t_lo = p.T;
  t_hi = -1;
  if (row0 >= 0) {
    t_lo = (kWin / 2 + 1 + kHop - 1) / kHop;  // smallest t with t*160 - 201 >= 0  (= 2)
    t_hi = (F - kWin / 2) / kHop;             // largest t with t*160 + 200 <= F
    if (t_hi > p.T - 1) t_hi = p.T - 1;
  }
}

// Kernel 1: every frame that cannot be taken from a stream, one warp per frame, un-normalised log-mel to `scratch`.
//   EDGE  == true : segments [0, n_on_stream) -- windows on a stream: one CTA per window computes its 2 + 2 edge frames;
//   EDGE  == false: segments [n_on_stream, n_seg) -- windows without one (tiled up by fixed_seq collate / off-grid): all T
//                   frames, spread over the whole chip (grid: ceil(T / 8) x windows).
template <bool EDGE>
__global__ void __launch_bounds__(kWinWarps * 32) window_frames_kernel(const FeatParams p, int seg_base) {
  extern __shared__ float dyn[];  // per-warp spectrum scratch: kWinWarps x (256 float2 + 260 float)
  __shared__ FbShared tabs;
  const int seg = seg_base + blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int t_lo, t_hi, row0;
  interior_range(p, seg, t_lo, t_hi, row0);
  load_tables(tabs, p.tab);
  __syncthreads();
  const int n_int = t_hi >= t_lo ? t_hi - t_lo + 1 : 0;
  if (EDGE && n_int == p.T) return;
  // this warp's first frame; EDGE: e-th non-interior frame, e = warp, warp + 8, ...
  int e = EDGE ? warp : blockIdx.x * kWinWarps + warp;
  auto frame_of = [&](int q) { return (n_int == 0 || q < t_lo) ? q : q + n_int; };
  int t = frame_of(e);
  if (t >= p.T) return;

  const int F = p.fixed_len;
  const int len = p.seg_len[seg];
  const float* src = p.wav + p.seg_start[seg];
  const int rep_end = (F / len) * len;  // tiled region [0, rep_end); tail copies the last F % len samples
  const int tail_base = len - (F - rep_end);
  auto x_at = [&](int i) -> float {  // fixed_seq collate: segment tiled up to F samples
    const int s = (i < rep_end) ? (i % len) : (tail_base + (i - rep_end));
    return __ldg(src + s);
  };
  const bool zero_pad = p.zero_pad != 0;
  const bool tiled = len < F;
  auto sample = [&](int j) -> float {  // pre-emphasised signal with torch.stft's reflect (or constant-zero) padding
    if (zero_pad && (j < 0 || j >= F)) return 0.f;
    if (j < 0) j = -j;
    if (j >= F) j = 2 * (F - 1) - j;
    if (j < 0) j = 0;  // only for degenerate F < 257; never on this path
    if (!tiled) {
      const float x0 = __ldg(src + j);
      return j == 0 ? x0 : preemph(x0, __ldg(src + j - 1));
    }
    const float x0 = x_at(j);
    return j == 0 ? x0 : preemph(x0, x_at(j - 1));
  };
  LaneTwiddles lt;
  lt.init(lane);
  for (;;) {
    const int base = t * kHop - kNFFT / 2;
    frame_logmel(tabs, lt, dyn + warp * kWarpScratch, lane, [&](int i) { return sample(base + i); },
                 p.scratch + (static_cast<size_t>(seg) * p.T + t) * kMels);
    if (!EDGE) break;
    e += kWinWarps;
    t = frame_of(e);
    if (t >= p.T) break;
  }
}

// Kernel 2: one CTA per window gathers its T frames (interior ones from the stream, the others from `scratch`), takes the
// exact two-pass per-feature mean / unbiased std over the window and writes the normalised features.  Pure data movement.
__global__ void __launch_bounds__(kWinWarps * 32) window_normalise_kernel(const FeatParams p) {
  extern __shared__ float logmel[];  // [T][80]
  __shared__ float s_mean[kMels], s_inv[kMels];
  const int seg = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int t_lo, t_hi, row0;
  interior_range(p, seg, t_lo, t_hi, row0);
  {
    const float4* own = reinterpret_cast<const float4*>(p.scratch + static_cast<size_t>(seg) * p.T * kMels);
    const float4* shared = row0 >= 0 ? reinterpret_cast<const float4*>(p.logmel + static_cast<size_t>(row0) * kMels) : own;
    float4* dst = reinterpret_cast<float4*>(logmel);
    const int lo4 = t_lo * (kMels / 4), hi4 = (t_hi + 1) * (kMels / 4);
    for (int i = tid; i < p.T * (kMels / 4); i += blockDim.x) dst[i] = __ldg(((i >= lo4 && i < hi4) ? shared : own) + i);
  }
  __syncthreads();

  // per-feature statistics over the T frames of this segment
  for (int m = warp; m < kMels; m += kWinWarps) {
    float s = 0.f;
    for (int t = lane; t < p.T; t += 32) s += logmel[t * kMels + m];
    const float mean = warp_sum(s) / static_cast<float>(p.T);
    float ss = 0.f;
    for (int t = lane; t < p.T; t += 32) {
      const float d = logmel[t * kMels + m] - mean;
      ss = fmaf(d, d, ss);
    }
    ss = warp_sum(ss);
    float sd = (p.T > 1) ? sqrtf(ss / static_cast<float>(p.T - 1)) : 0.f;
    if (lane == 0) {
      s_mean[m] = mean;
      s_inv[m] = 1.f / (sd + 1e-5f);
    }
  }
  __syncthreads();

  const size_t orow0 = static_cast<size_t>(seg) * p.T;
  const int cpr = p.ldo / 2;  // half2 columns per row
  for (int idx = tid; idx < p.T * cpr; idx += blockDim.x) {
    const int t = idx / cpr, c = (idx - t * cpr) * 2;
    float a = 0.f, b = 0.f;
    if (c < kMels) a = (logmel[t * kMels + c] - s_mean[c]) * s_inv[c];
    if (c + 1 < kMels) b = (logmel[t * kMels + c + 1] - s_mean[c + 1]) * s_inv[c + 1];
    reinterpret_cast<__half2*>(p.out16 + (orow0 + t) * p.ldo)[idx - t * cpr] = __floats2half2_rn(a, b);
  }
  if (p.out32) {
    float* o = p.out32 + orow0 * kMels;
    for (int idx = tid; idx < p.T * kMels; idx += blockDim.x) {
      const int m = idx % kMels;
      o[idx] = (logmel[idx] - s_mean[m]) * s_inv[m];
    }
  }
}

static int check_tables(const int32_t* fb_start, const int32_t* fb_off, const float* fb_w, int32_t fb_nnz, const float* window) {
  B200D_CHECK_ARG(fb_start && fb_off && fb_w && window);
  B200D_CHECK_ARG(fb_nnz > 0 && fb_nnz <= kMaxFbNnz);
  return B200D_OK;
}

}  // namespace b200d

using namespace b200d;

extern "C" int b200d_mel_stream(const float* wav, int64_t n_wav, const int64_t* stream_start, const int32_t* stream_off, int32_t n_streams,
                                int32_t total_rows, const int32_t* fb_start, const int32_t* fb_off, const float* fb_w, int32_t fb_nnz,
                                const float* window, float* logmel, void* stream) {
  B200D_CHECK_ARG(wav && stream_start && stream_off && logmel);
  B200D_CHECK_ARG(n_streams > 0 && total_rows > 0 && total_rows % kStreamWarps == 0);
  if (int rc = check_tables(fb_start, fb_off, fb_w, fb_nnz, window)) return rc;
  StreamParams p;
  p.wav = wav; p.n_wav = n_wav; p.stream_start = reinterpret_cast<const long long*>(stream_start); p.stream_off = stream_off;
  p.n_streams = n_streams; p.tab = {fb_start, fb_off, fb_w, window}; p.logmel = logmel;
  constexpr size_t smem = (((kStreamSpan + 3) & ~3) + static_cast<size_t>(kStreamWarps) * kWarpScratch) * sizeof(float);
  B200D_CHECK_CUDA(cudaFuncSetAttribute(mel_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  mel_stream_kernel<<<total_rows / kStreamWarps, kStreamWarps * 32, smem, as_stream(stream)>>>(p);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_featurize_windows(const float* wav, int64_t n_wav, const float* logmel, const int32_t* seg_start, const int32_t* seg_len,
                                       const int32_t* seg_row0, int32_t n_on_stream, int32_t n_seg, int32_t fixed_len,
                                       const int32_t* fb_start, const int32_t* fb_off, const float* fb_w, int32_t fb_nnz, const float* window,
                                       int32_t variant, float* scratch, void* out_f16, int32_t ldo, float* out_f32, void* stream) {
  B200D_CHECK_ARG(wav && seg_start && seg_len && scratch && out_f16);
  B200D_CHECK_ARG((logmel == nullptr) == (seg_row0 == nullptr));
  B200D_CHECK_ARG(n_seg > 0 && n_on_stream >= 0 && n_on_stream <= n_seg && fixed_len >= kNFFT / 2 + 1);
  B200D_CHECK_ARG(ldo >= kMels && ldo % 8 == 0);
  B200D_CHECK_ARG(variant >= 0 && variant <= 3);
  if (int rc = check_tables(fb_start, fb_off, fb_w, fb_nnz, window)) return rc;
  FeatParams p;
  p.wav = wav; p.n_wav = n_wav; p.seg_start = seg_start; p.seg_len = seg_len; p.seg_row0 = seg_row0; p.logmel = logmel; p.n_seg = n_seg;
  p.fixed_len = fixed_len; p.T = fixed_len / kHop + ((variant & B200D_FEAT_NO_PLUS_ONE) ? 0 : 1);
  p.zero_pad = (variant & B200D_FEAT_ZERO_PAD) ? 1 : 0;
  B200D_CHECK_ARG(p.T >= 2);
  p.tab = {fb_start, fb_off, fb_w, window};
  p.scratch = scratch;
  p.out16 = reinterpret_cast<__half*>(out_f16); p.ldo = ldo; p.out32 = out_f32;
  const size_t smem_n = ((static_cast<size_t>(p.T) * kMels + 3) & ~static_cast<size_t>(3)) * sizeof(float);
  B200D_CHECK_ARG(smem_n <= 200 * 1024);  // T <= 640 frames (6.4 s windows)
  constexpr size_t smem_f = static_cast<size_t>(kWinWarps) * kWarpScratch * sizeof(float);
  cudaStream_t st = as_stream(stream);
  const int n_fast = (logmel != nullptr) ? n_on_stream : 0;
  for (int s0 = 0; s0 < n_fast; s0 += 65535)  // grid.y limit
    window_frames_kernel<true><<<dim3(1, n_fast - s0 < 65535 ? n_fast - s0 : 65535), kWinWarps * 32, smem_f, st>>>(p, s0);
  for (int s0 = n_fast; s0 < n_seg; s0 += 65535)
    window_frames_kernel<false><<<dim3((p.T + kWinWarps - 1) / kWinWarps, n_seg - s0 < 65535 ? n_seg - s0 : 65535), kWinWarps * 32, smem_f, st>>>(p, s0);
  B200D_CHECK_LAUNCH();
  B200D_CHECK_CUDA(cudaFuncSetAttribute(window_normalise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(200 * 1024)));
  window_normalise_kernel<<<n_seg, kWinWarps * 32, smem_n, st>>>(p);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}
