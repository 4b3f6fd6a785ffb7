// Fused segment gather -> pre-emphasis -> STFT -> |.|^2 -> mel -> log -> per-feature normalise.
//
// One CTA per segment; each warp owns whole frames (512-point real FFT as a 256-point complex
// radix-2 FFT in the warp's private shared-memory buffer).  The segment's un-normalised log-mel
// [T][80] stays in shared memory, so the per-feature mean / unbiased std over the segment
// (normalize_batch 'per_feature') is an exact two-pass reduction and the features are written to
// HBM exactly once, as fp16 channels-last rows for the first TitaNet conv.  The waveform is read
// straight from HBM/L2 with coalesced loads: consecutive lanes read consecutive samples, and the
// 60 % overlap between neighbouring frames / segments / scales is served by L1/L2.
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
constexpr int kFeatWarps = 32;  // the kernel is latency-bound: more resident warps per CTA, their FFT scratch in dynamic smem
constexpr int kMaxFbNnz = 1024;
constexpr int kWarpScratch = 512 + kBins + 3;  // floats: 256 complex spectrum + 257 power bins (+ pad)

struct FeatParams {
  const float* wav;
  long long n_wav;
  const int* seg_start;
  const int* seg_len;
  int n_seg;
  int fixed_len;
  int T;
  int zero_pad;  // STFT padding of the (pre-emphasised) window: 0 = reflect (torch.stft's default), 1 = zeros
  const int* fb_start;  // [80] first bin of each filter
  const int* fb_off;    // [81] offsets into fb_w
  const float* fb_w;    // packed non-zero weights
  const float* window;  // [400]
  __half* out16;
  int ldo;
  float* out32;
};

__global__ void __launch_bounds__(kFeatWarps * 32) featurize_kernel(const FeatParams p) {
  extern __shared__ float logmel[];  // [T][80] | per-warp spectrum scratch: kFeatWarps x (256 float2 + 260 float)
  __shared__ float2 s_tw[256];       // e^{-2 pi i k / 512}
  __shared__ float s_win[kWin];
  __shared__ float s_fbw[kMaxFbNnz];
  __shared__ int s_fbs[kMels], s_fbo[kMels + 1];
  __shared__ float s_mean[kMels], s_inv[kMels];

  const int seg = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int k = tid; k < 256; k += blockDim.x) {
    float s, c;
    sincospif(static_cast<float>(k) / 256.f, &s, &c);
    s_tw[k] = make_float2(c, -s);
  }
  for (int k = tid; k < kWin; k += blockDim.x) s_win[k] = p.window[k];
  for (int k = tid; k < kMels; k += blockDim.x) s_fbs[k] = p.fb_start[k];
  for (int k = tid; k <= kMels; k += blockDim.x) s_fbo[k] = p.fb_off[k];
  const int nnz = p.fb_off[kMels];
  for (int k = tid; k < nnz; k += blockDim.x) s_fbw[k] = p.fb_w[k];
  __syncthreads();

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
  auto y_at = [&](int j) -> float {  // pre-emphasised signal with torch.stft's reflect (or constant-zero) padding
    if (zero_pad && (j < 0 || j >= F)) return 0.f;
    if (j < 0) j = -j;
    if (j >= F) j = 2 * (F - 1) - j;
    if (j < 0) j = 0;  // only for degenerate F < 257; never on this path
    const float x0 = x_at(j);
    return j == 0 ? x0 : x0 - 0.97f * x_at(j - 1);
  };

  float* scratch = logmel + ((p.T * kMels + 3) & ~3) + warp * kWarpScratch;
  float2* cbuf = reinterpret_cast<float2*>(scratch);
  float* pbuf = scratch + 512;
  // 256-point complex FFT of z[n] = v[2n] + i v[2n+1] without shared-memory stages: n = lane + 32 j.
  //   step 1  8-point DFT over j in registers                 A[k1]  = sum_j z[j] W8^(j k1)
  //   step 2  twiddle                                          B[k1]  = A[k1] W256^(lane k1)
  //   step 3  32-point DIF FFT across the lanes by shuffles    X[k1 + 8 k2], k2 = bitrev5(lane)
  // lane-constant twiddles live in registers for the whole kernel.
  float2 tw2[8];   // W256^(lane * k1)
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) {
    float sn, cs;
    sincospif(static_cast<float>(lane * k1) / 128.f, &sn, &cs);
    tw2[k1] = make_float2(cs, -sn);
  }
  float2 tw3[5];   // stage twiddles of the cross-lane FFT: W32^((lane & (h-1)) * (16/h)), h = 16, 8, 4, 2, 1
#pragma unroll
  for (int st = 0; st < 5; ++st) {
    const int h = 16 >> st;
    float sn, cs;
    sincospif(static_cast<float>((lane & (h - 1)) * (16 / h)) / 16.f, &sn, &cs);
    tw3[st] = make_float2(cs, -sn);
  }
  const int k2 = static_cast<int>(__brev(static_cast<unsigned>(lane)) >> 27);
  const bool tiled = len < F;
  auto sample = [&](int j) -> float {  // pre-emphasised, reflect-padded signal at frame-relative position
    if (!tiled) {
      if (zero_pad && (j < 0 || j >= F)) return 0.f;
      if (j < 0) j = -j;
      if (j >= F) j = 2 * (F - 1) - j;
      const float x0 = __ldg(src + j);
      return j == 0 ? x0 : x0 - 0.97f * __ldg(src + j - 1);
    }
    return y_at(j);
  };
  for (int t = warp; t < p.T; t += kFeatWarps) {
    const int base = t * kHop - kNFFT / 2;
    float2 z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i0 = 2 * (lane + 32 * j), i1 = i0 + 1;
      float v0 = 0.f, v1 = 0.f;
      if (i0 >= kWinOff && i0 < kWinOff + kWin) v0 = s_win[i0 - kWinOff] * sample(base + i0);
      if (i1 >= kWinOff && i1 < kWinOff + kWin) v1 = s_win[i1 - kWinOff] * sample(base + i1);
      z[j] = make_float2(v0, v1);
    }
    // ---- step 1: 8-point DIF DFT in registers (outputs in bit-reversed order, undone by the index map below)
    auto bf = [](float2& a, float2& b) { const float2 s = make_float2(a.x + b.x, a.y + b.y), d = make_float2(a.x - b.x, a.y - b.y); a = s; b = d; };
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
      float2 b = mul(z[q], tw2[k1]);
      // ---- step 3: 32-point DIF FFT across lanes
#pragma unroll
      for (int st = 0; st < 5; ++st) {
        const int h = 16 >> st;
        const float ox = __shfl_xor_sync(0xffffffffu, b.x, h), oy = __shfl_xor_sync(0xffffffffu, b.y, h);
        if (lane & h) b = mul(make_float2(ox - b.x, oy - b.y), tw3[st]);
        else b = make_float2(b.x + ox, b.y + oy);
      }
      cbuf[k1 + 8 * k2] = b;
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
        const float2 w = (k < 256) ? s_tw[k] : make_float2(-1.f, 0.f);
        const float xr = er + w.x * orr - w.y * oi;
        const float xi = ei + w.x * oi + w.y * orr;
        pbuf[k] = xr * xr + xi * xi;
      }
    }
    __syncwarp();
    for (int m = lane; m < kMels; m += 32) {
      const int k0 = s_fbs[m], o0 = s_fbo[m], n = s_fbo[m + 1] - o0;
      float acc = 0.f;
      for (int i = 0; i < n; ++i) acc = fmaf(s_fbw[o0 + i], pbuf[k0 + i], acc);
      logmel[t * kMels + m] = logf(acc + 5.9604644775390625e-08f);  // log(x + 2^-24)
    }
    __syncwarp();
  }
  __syncthreads();

  // per-feature statistics over the T frames of this segment
  for (int m = warp; m < kMels; m += kFeatWarps) {
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

  const size_t row0 = static_cast<size_t>(seg) * p.T;
  const int cpr = p.ldo / 2;  // half2 columns per row
  for (int idx = tid; idx < p.T * cpr; idx += blockDim.x) {
    const int t = idx / cpr, c = (idx - t * cpr) * 2;
    float a = 0.f, b = 0.f;
    if (c < kMels) a = (logmel[t * kMels + c] - s_mean[c]) * s_inv[c];
    if (c + 1 < kMels) b = (logmel[t * kMels + c + 1] - s_mean[c + 1]) * s_inv[c + 1];
    reinterpret_cast<__half2*>(p.out16 + (row0 + t) * p.ldo)[idx - t * cpr] = __floats2half2_rn(a, b);
  }
  if (p.out32) {
    float* o = p.out32 + row0 * kMels;
    for (int idx = tid; idx < p.T * kMels; idx += blockDim.x) {
      const int m = idx % kMels;
      o[idx] = (logmel[idx] - s_mean[m]) * s_inv[m];
    }
  }
}

}  // namespace b200d

using namespace b200d;

extern "C" int b200d_featurize(const float* wav, int64_t n_wav, const int32_t* seg_start, const int32_t* seg_len, int32_t n_seg,
                               int32_t fixed_len, const int32_t* fb_start, const int32_t* fb_off, const float* fb_w, int32_t fb_nnz,
                               const float* window, int32_t variant, void* out_f16, int32_t ldo, float* out_f32, void* stream) {
  B200D_CHECK_ARG(wav && seg_start && seg_len && fb_start && fb_off && fb_w && window && out_f16);
  B200D_CHECK_ARG(n_seg > 0 && fixed_len >= kNFFT / 2 + 1);
  B200D_CHECK_ARG(ldo >= kMels && ldo % 8 == 0);
  B200D_CHECK_ARG(fb_nnz > 0 && fb_nnz <= kMaxFbNnz);
  B200D_CHECK_ARG(variant >= 0 && variant <= 3);
  FeatParams p;
  p.wav = wav; p.n_wav = n_wav; p.seg_start = seg_start; p.seg_len = seg_len; p.n_seg = n_seg;
  p.fixed_len = fixed_len; p.T = fixed_len / kHop + ((variant & B200D_FEAT_NO_PLUS_ONE) ? 0 : 1);
  p.zero_pad = (variant & B200D_FEAT_ZERO_PAD) ? 1 : 0;
  B200D_CHECK_ARG(p.T >= 2);
  p.fb_start = fb_start; p.fb_off = fb_off; p.fb_w = fb_w; p.window = window;
  p.out16 = reinterpret_cast<__half*>(out_f16); p.ldo = ldo; p.out32 = out_f32;
  const size_t smem = (((static_cast<size_t>(p.T) * kMels + 3) & ~static_cast<size_t>(3)) + static_cast<size_t>(kFeatWarps) * kWarpScratch) * sizeof(float);
  B200D_CHECK_ARG(smem <= 215 * 1024);  // T <= 362 frames (3.6 s windows)
  static size_t configured = 0;
  if (smem > configured) {
    B200D_CHECK_CUDA(cudaFuncSetAttribute(featurize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(215 * 1024)));
    configured = 215 * 1024;
  }
  featurize_kernel<<<n_seg, kFeatWarps * 32, smem, as_stream(stream)>>>(p);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}
