// CUDA-core kernels of the TitaNet-L forward (everything that is not a dense contraction):
// masked depthwise conv over time, squeeze-excite pooling / gating, time statistics and attentive
// statistics pooling.  Activations are fp16 channels-last [n_seg*T][C]; all segments of a call
// share the frame count T (fixed_seq collate), so "masking beyond the sequence length" of upstream
// MaskedConv1d reduces to zero padding at the segment edges.
// Upstream: nemo/collections/asr/parts/submodules/jasper.py (MaskedConv1d, SqueezeExcite),
//           nemo/collections/asr/parts/submodules/tdnn_attention.py (AttentivePoolLayer).
#include "common.cuh"

#include <cstdlib>

namespace b200d {

__device__ __forceinline__ void load4h(const __half* p, float (&v)[4]) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void store4h(__half* p, const float (&v)[4]) {
  __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// ------------------------------------------------------------------------------------ depthwise
// Thread = 2 channels x a strip of kDwTT consecutive output frames of one window.  All kDwTT + k - 1 input rows of the
// strip are requested up front (independent 4-byte loads, a warp touches 128 contiguous bytes per row), so the DRAM /
// L2 latency is paid once per strip; every input row is loaded once per strip (read amplification (TT + k - 1) / TT,
// the halo from L1/L2) and the output is written once.
// Arithmetic: the fp32 version of this kernel was ISSUE-bound (ncu: 127 M warp instructions per launch, 57 % issue
// utilisation at 23 % of DRAM peak; a third of the instructions were address / predicate arithmetic), so
//  * the taps are applied with packed half2 FMAs straight on the loaded fp16 pairs, in partial sums of at most 5 taps
//    (relative rms error 5e-4 against 2e-4 for fp32 accumulation + fp16 store, whose rounding dominates both; upstream
//    runs this conv under fp16 autocast on CUDA),
//  * the channel count is a template parameter, so row addresses are immediates off one base pointer, and
//  * strips that do not touch a window edge take a predicate-free path.
constexpr int kDwTT = 32;

template <int KS, int CT, bool EDGE>
__device__ __forceinline__ void depthwise_strip(const __half* __restrict__ xs, __half* __restrict__ ys, const __half2 (&wt)[KS], int t0,
                                                int T, int C_rt) {
  constexpr int PAD = KS / 2;
  constexpr int ROWS = kDwTT + KS - 1;
  constexpr int GROUP = 5;
  const size_t C = CT > 0 ? static_cast<size_t>(CT) : static_cast<size_t>(C_rt);
  const __half* x0 = xs + (static_cast<long long>(t0) - PAD) * static_cast<long long>(C);
  __half2 raw[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    uint32_t u;
    if constexpr (EDGE) {
      const int t = t0 - PAD + r;
      u = (t >= 0 && t < T) ? __ldg(reinterpret_cast<const uint32_t*>(x0 + r * C)) : 0u;
    } else {
      u = __ldg(reinterpret_cast<const uint32_t*>(x0 + r * C));
    }
    raw[r] = *reinterpret_cast<__half2*>(&u);
  }
  __half* y0 = ys + static_cast<size_t>(t0) * C;
#pragma unroll
  for (int tt = 0; tt < kDwTT; ++tt) {
    if (!EDGE || t0 + tt < T) {
      __half2 acc;
#pragma unroll
      for (int j0 = 0; j0 < KS; j0 += GROUP) {
        __half2 part = __hmul2(wt[j0], raw[tt + j0]);
#pragma unroll
        for (int j = j0 + 1; j < j0 + GROUP && j < KS; ++j) part = __hfma2(wt[j], raw[tt + j], part);
        acc = (j0 == 0) ? part : __hadd2(acc, part);
      }
      *reinterpret_cast<__half2*>(y0 + tt * C) = acc;
    }
  }
}

template <int KS, int CT>
__global__ void __launch_bounds__(256, 2) depthwise_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                           const float* __restrict__ w, int n_seg, int T, int C, int strips) {
  constexpr int PAD = KS / 2;
  const int cp = blockIdx.y * blockDim.x + threadIdx.x;  // channel pair
  if (cp * 2 >= C) return;
  const int strip = blockIdx.x;
  const int seg = strip / strips;
  const int t0 = (strip - seg * strips) * kDwTT;
  const int c = cp * 2;
  const __half* xs = x + (static_cast<size_t>(seg) * T) * C + c;
  __half* ys = y + (static_cast<size_t>(seg) * T) * C + c;
  __half2 wt[KS];
#pragma unroll
  for (int j = 0; j < KS; ++j) {
    const float2 f = __ldg(reinterpret_cast<const float2*>(w + static_cast<size_t>(j) * C + c));
    wt[j] = __floats2half2_rn(f.x, f.y);
  }
  if (t0 >= PAD && t0 + kDwTT + PAD <= T) depthwise_strip<KS, CT, false>(xs, ys, wt, t0, T, C);
  else depthwise_strip<KS, CT, true>(xs, ys, wt, t0, T, C);
}

// ------------------------------------------------------------------------------------ time statistics
// Block = (window, 256-channel chunk): 64 channel groups of 4 x kTsSlices time slices; each thread reduces the rows
// t == slice (mod kTsSlices), the slices are combined through shared memory in a fixed order.
// WITH_STD == false: mean only (SqueezeExcite pool).
// WITH_STD == true: [mean | sqrt(clamp(mean((x-mean)^2), 1e-10))] (AttentivePoolLayer context), exact two-pass form.
constexpr int kTsSlices = 4;

template <bool WITH_STD>
__global__ void __launch_bounds__(64 * kTsSlices) time_stats_kernel(const __half* __restrict__ x, int T, int C, __half* __restrict__ out16) {
  __shared__ float s_part[kTsSlices][64][4];
  __shared__ float s_mean[64][4];
  const int seg = blockIdx.x;
  const int g = threadIdx.x & 63, slice = threadIdx.x >> 6;
  const int c = (blockIdx.y * 64 + g) * 4;
  const bool active = c < C;
  const __half* xs = x + static_cast<size_t>(seg) * T * C + c;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  if (active) {
    int t = slice;
    for (; t + 3 * kTsSlices < T; t += 4 * kTsSlices) {  // four independent row requests in flight per thread
      float v0[4], v1[4], v2[4], v3[4];
      load4h(xs + static_cast<size_t>(t) * C, v0);
      load4h(xs + static_cast<size_t>(t + kTsSlices) * C, v1);
      load4h(xs + static_cast<size_t>(t + 2 * kTsSlices) * C, v2);
      load4h(xs + static_cast<size_t>(t + 3 * kTsSlices) * C, v3);
#pragma unroll
      for (int q = 0; q < 4; ++q) s[q] += (v0[q] + v1[q]) + (v2[q] + v3[q]);
    }
    for (; t < T; t += kTsSlices) {
      float v[4];
      load4h(xs + static_cast<size_t>(t) * C, v);
#pragma unroll
      for (int q = 0; q < 4; ++q) s[q] += v[q];
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) s_part[slice][g][q] = s[q];
  __syncthreads();
  const float inv = 1.f / static_cast<float>(T);
  const int ld = WITH_STD ? 2 * C : C;
  if (slice == 0) {
    float mean[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float tot = 0.f;
#pragma unroll
      for (int sl = 0; sl < kTsSlices; ++sl) tot += s_part[sl][g][q];
      mean[q] = tot * inv;
      s_mean[g][q] = mean[q];
    }
    if (active) store4h(out16 + static_cast<size_t>(seg) * ld + c, mean);
  }
  if constexpr (WITH_STD) {
    __syncthreads();
    float mean[4], ss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; ++q) mean[q] = s_mean[g][q];
    if (active) {
      for (int t = slice; t < T; t += kTsSlices) {
        float v[4];
        load4h(xs + static_cast<size_t>(t) * C, v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float d = v[q] - mean[q];
          ss[q] = fmaf(d, d, ss[q]);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) s_part[slice][g][q] = ss[q];
    __syncthreads();
    if (slice == 0 && active) {
      float sd[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float tot = 0.f;
#pragma unroll
        for (int sl = 0; sl < kTsSlices; ++sl) tot += s_part[sl][g][q];
        sd[q] = sqrtf(fmaxf(tot * inv, 1e-10f));
      }
      store4h(out16 + static_cast<size_t>(seg) * ld + C + c, sd);
    }
  }
}

// ------------------------------------------------------------------------------------ SE mean from GEMM partials
// The producing GEMM's epilogue left, per group of 32 rows, the column sums of the rows before and from the window
// boundary inside the group (b200d_gemm_epilogue.colsum).  Window s owns part[g][0] of the groups that start inside it and
// part[g][1] of the group that started in window s - 1 and runs into it; summed in group order (deterministic).
__global__ void se_mean_from_colsum_kernel(const float* __restrict__ part, int n_seg, int T, int C, __half* __restrict__ mean16) {
  const int seg = blockIdx.x;
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (c >= C) return;
  const long long r0 = static_cast<long long>(seg) * T, r1 = r0 + T - 1;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long g = r0 >> 5; g <= (r1 >> 5); ++g) {
    const int which = ((g << 5) / T == seg) ? 0 : 1;
    const float4 v = __ldg(reinterpret_cast<const float4*>(part + (static_cast<size_t>(g) * 2 + which) * C + c));
    s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
  }
  const float inv = 1.f / static_cast<float>(T);
  const float m[4] = {s[0] * inv, s[1] * inv, s[2] * inv, s[3] * inv};
  store4h(mean16 + static_cast<size_t>(seg) * C + c, m);
}

// ------------------------------------------------------------------------------------ SE apply
__global__ void se_apply_relu_kernel(const __half* __restrict__ x, const float* __restrict__ gate, __half* __restrict__ y,
                                     size_t total4, int T, int C) {
  const int c4 = C / 4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t row = i / c4;
    const int c = static_cast<int>(i - row * c4) * 4;
    const size_t seg = row / T;
    float v[4];
    load4h(x + row * C + c, v);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gate + seg * C + c));
    v[0] = fmaxf(v[0] * g.x, 0.f); v[1] = fmaxf(v[1] * g.y, 0.f);
    v[2] = fmaxf(v[2] * g.z, 0.f); v[3] = fmaxf(v[3] * g.w, 0.f);
    store4h(y + row * C + c, v);
  }
}

// ------------------------------------------------------------------------------------ SE apply + statistics
// Last encoder block: y = relu(x * gate[window]) AND, in the same pass over the 3072-channel activation, the per-window
// [mean | std] of y that AttentivePoolLayer's TDNN context needs (moments of the fp16-rounded y, accumulated in fp32;
// T <= 362 values per channel, so the one-pass variance is exact to ~1e-6).  Saves two full reads of y.
__global__ void __launch_bounds__(64 * kTsSlices) se_apply_stats_kernel(const __half* __restrict__ x, const float* __restrict__ gate,
                                                                        __half* __restrict__ y, int T, int C, __half* __restrict__ stats16) {
  __shared__ float s_1[kTsSlices][64][4], s_2[kTsSlices][64][4];
  const int seg = blockIdx.x;
  const int g = threadIdx.x & 63, slice = threadIdx.x >> 6;
  const int c = (blockIdx.y * 64 + g) * 4;
  const bool active = c < C;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if (active) {
    const __half* xs = x + static_cast<size_t>(seg) * T * C + c;
    __half* ys = y + static_cast<size_t>(seg) * T * C + c;
    const float4 gt = __ldg(reinterpret_cast<const float4*>(gate + static_cast<size_t>(seg) * C + c));
    const float gv[4] = {gt.x, gt.y, gt.z, gt.w};
    auto body = [&](int t, float (&v)[4]) {
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = __half2float(__float2half_rn(fmaxf(v[q] * gv[q], 0.f)));
      store4h(ys + static_cast<size_t>(t) * C, v);
#pragma unroll
      for (int q = 0; q < 4; ++q) { s1[q] += v[q]; s2[q] = fmaf(v[q], v[q], s2[q]); }
    };
    int t = slice;
    for (; t + 3 * kTsSlices < T; t += 4 * kTsSlices) {
      float v0[4], v1[4], v2[4], v3[4];
      load4h(xs + static_cast<size_t>(t) * C, v0);
      load4h(xs + static_cast<size_t>(t + kTsSlices) * C, v1);
      load4h(xs + static_cast<size_t>(t + 2 * kTsSlices) * C, v2);
      load4h(xs + static_cast<size_t>(t + 3 * kTsSlices) * C, v3);
      body(t, v0); body(t + kTsSlices, v1); body(t + 2 * kTsSlices, v2); body(t + 3 * kTsSlices, v3);
    }
    for (; t < T; t += kTsSlices) {
      float v[4];
      load4h(xs + static_cast<size_t>(t) * C, v);
      body(t, v);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) { s_1[slice][g][q] = s1[q]; s_2[slice][g][q] = s2[q]; }
  __syncthreads();
  if (slice == 0 && active) {
    const float inv = 1.f / static_cast<float>(T);
    float mean[4], sd[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int sl = 0; sl < kTsSlices; ++sl) { a += s_1[sl][g][q]; b += s_2[sl][g][q]; }
      mean[q] = a * inv;
      sd[q] = sqrtf(fmaxf(b * inv - mean[q] * mean[q], 1e-10f));
    }
    store4h(stats16 + static_cast<size_t>(seg) * 2 * C + c, mean);
    store4h(stats16 + static_cast<size_t>(seg) * 2 * C + C + c, sd);
  }
}

// ------------------------------------------------------------------------------------ attentive pooling
// alpha = softmax over time of e (per channel); mu = sum alpha x; sg = sqrt(clamp(sum alpha (x - mu)^2, 1e-10)).
// Single pass with an online (running-max) softmax per time slice; the kTsSlices partial states (max, Z, S1, S2) are
// merged in a fixed order through shared memory.
__global__ void __launch_bounds__(64 * kTsSlices) attn_pool_kernel(const __half* __restrict__ x, const __half* __restrict__ e, int T, int C,
                                                                   __half* __restrict__ out16) {
  __shared__ float s_m[kTsSlices][64][4], s_z[kTsSlices][64][4], s_1[kTsSlices][64][4], s_2[kTsSlices][64][4];
  const int seg = blockIdx.x;
  const int g = threadIdx.x & 63, slice = threadIdx.x >> 6;
  const int c = (blockIdx.y * 64 + g) * 4;
  const bool active = c < C;
  const __half* xs = x + static_cast<size_t>(seg) * T * C + c;
  const __half* es = e + static_cast<size_t>(seg) * T * C + c;
  float m[4], z[4], s1[4], s2[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) { m[q] = -INFINITY; z[q] = 0.f; s1[q] = 0.f; s2[q] = 0.f; }
  if (active) {
    auto body = [&](const float (&xv)[4], const float (&ev)[4]) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float mn = fmaxf(m[q], ev[q]);
        const float r = __expf(m[q] - mn);  // exp(-inf) = 0 on the first frame
        const float wgt = __expf(ev[q] - mn);
        z[q] = z[q] * r + wgt;
        s1[q] = s1[q] * r + wgt * xv[q];
        s2[q] = s2[q] * r + wgt * xv[q] * xv[q];
        m[q] = mn;
      }
    };
    int t = slice;
    for (; t + 3 * kTsSlices < T; t += 4 * kTsSlices) {  // eight independent row requests in flight per thread
      float x0[4], x1[4], x2[4], x3[4], e0[4], e1[4], e2[4], e3[4];
      load4h(xs + static_cast<size_t>(t) * C, x0); load4h(es + static_cast<size_t>(t) * C, e0);
      load4h(xs + static_cast<size_t>(t + kTsSlices) * C, x1); load4h(es + static_cast<size_t>(t + kTsSlices) * C, e1);
      load4h(xs + static_cast<size_t>(t + 2 * kTsSlices) * C, x2); load4h(es + static_cast<size_t>(t + 2 * kTsSlices) * C, e2);
      load4h(xs + static_cast<size_t>(t + 3 * kTsSlices) * C, x3); load4h(es + static_cast<size_t>(t + 3 * kTsSlices) * C, e3);
      body(x0, e0); body(x1, e1); body(x2, e2); body(x3, e3);
    }
    for (; t < T; t += kTsSlices) {
      float xv[4], ev[4];
      load4h(xs + static_cast<size_t>(t) * C, xv);
      load4h(es + static_cast<size_t>(t) * C, ev);
      body(xv, ev);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) { s_m[slice][g][q] = m[q]; s_z[slice][g][q] = z[q]; s_1[slice][g][q] = s1[q]; s_2[slice][g][q] = s2[q]; }
  __syncthreads();
  if (slice == 0 && active) {
    float mu[4], sg[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float mm = s_m[0][g][q];
#pragma unroll
      for (int sl = 1; sl < kTsSlices; ++sl) mm = fmaxf(mm, s_m[sl][g][q]);
      float zz = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int sl = 0; sl < kTsSlices; ++sl) {
        const float r = (s_m[sl][g][q] == -INFINITY) ? 0.f : __expf(s_m[sl][g][q] - mm);
        zz += s_z[sl][g][q] * r;
        a1 += s_1[sl][g][q] * r;
        a2 += s_2[sl][g][q] * r;
      }
      const float iz = 1.f / zz;
      mu[q] = a1 * iz;
      sg[q] = sqrtf(fmaxf(a2 * iz - mu[q] * mu[q], 1e-10f));
    }
    store4h(out16 + static_cast<size_t>(seg) * 2 * C + c, mu);
    store4h(out16 + static_cast<size_t>(seg) * 2 * C + C + c, sg);
  }
}

}  // namespace b200d

using namespace b200d;

extern "C" int b200d_depthwise_conv(const void* x, void* y, const float* w, int32_t n_seg, int32_t T, int32_t C, int32_t ksize,
                                    void* stream) {
  B200D_CHECK_ARG(x && y && w && x != y);
  B200D_CHECK_ARG(n_seg > 0 && T > 0 && C > 0 && C % 4 == 0);
  const int strips = (T + kDwTT - 1) / kDwTT;
  const int threads = (C / 2) < 256 ? ((C / 2 + 31) / 32) * 32 : 256;
  B200D_CHECK_ARG(static_cast<long long>(n_seg) * strips < 2147483647LL);
  dim3 grid(static_cast<unsigned>(n_seg) * strips, (C / 2 + threads - 1) / threads);
  const __half* xi = reinterpret_cast<const __half*>(x);
  __half* yo = reinterpret_cast<__half*>(y);
  cudaStream_t st = as_stream(stream);
  // development knob for tools/concurrency_check.py: dynamic shared memory that only lowers this kernel's residency per SM
  static const int dw_smem = getenv("B200D_EXP_DW_SMEM") ? atoi(getenv("B200D_EXP_DW_SMEM")) : 0;
#define B200D_DW(KS)                                                                                            \
  if (C == 1024) {                                                                                              \
    if (dw_smem > 48 * 1024) cudaFuncSetAttribute(depthwise_kernel<KS, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, dw_smem); \
    depthwise_kernel<KS, 1024><<<grid, threads, dw_smem, st>>>(xi, yo, w, n_seg, T, C, strips);                 \
  }                                                                                                             \
  else if (C == 128) depthwise_kernel<KS, 128><<<grid, threads, dw_smem, st>>>(xi, yo, w, n_seg, T, C, strips); \
  else depthwise_kernel<KS, 0><<<grid, threads, dw_smem, st>>>(xi, yo, w, n_seg, T, C, strips)
  switch (ksize) {
    case 3: B200D_DW(3); break;
    case 7: B200D_DW(7); break;
    case 11: B200D_DW(11); break;
    case 15: B200D_DW(15); break;
    default: return b200d::set_error(B200D_EINVAL, "%s: unsupported kernel size (3, 7, 11, 15)%s", "b200d_depthwise_conv");
  }
#undef B200D_DW
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_time_stats(const void* x, int32_t n_seg, int32_t T, int32_t C, int32_t with_std, void* out16, void* stream) {
  B200D_CHECK_ARG(x && out16 && n_seg > 0 && T > 0 && C % 4 == 0);
  const dim3 grid(n_seg, (C / 4 + 63) / 64);
  const int threads = 64 * kTsSlices;
  if (with_std)
    time_stats_kernel<true><<<grid, threads, 0, as_stream(stream)>>>(reinterpret_cast<const __half*>(x), T, C, reinterpret_cast<__half*>(out16));
  else
    time_stats_kernel<false><<<grid, threads, 0, as_stream(stream)>>>(reinterpret_cast<const __half*>(x), T, C, reinterpret_cast<__half*>(out16));
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_se_mean_from_colsum(const float* colsum, int32_t n_seg, int32_t T, int32_t C, void* mean16, void* stream) {
  B200D_CHECK_ARG(colsum && mean16 && n_seg > 0 && T >= 32 && C % 4 == 0);
  const int threads = 128;
  se_mean_from_colsum_kernel<<<dim3(n_seg, (C / 4 + threads - 1) / threads), threads, 0, as_stream(stream)>>>(colsum, n_seg, T, C,
                                                                                                              reinterpret_cast<__half*>(mean16));
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_se_apply_relu(const void* x, const float* gate, void* y, int32_t n_seg, int32_t T, int32_t C, void* stream) {
  B200D_CHECK_ARG(x && gate && y && n_seg > 0 && T > 0 && C % 4 == 0);
  const size_t total4 = static_cast<size_t>(n_seg) * T * (C / 4);
  const int blocks = static_cast<int>(total4 / 256 + 1 < static_cast<size_t>(kNumSMs) * 16 ? total4 / 256 + 1 : kNumSMs * 16);
  se_apply_relu_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const __half*>(x), gate, reinterpret_cast<__half*>(y), total4, T, C);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_attn_pool(const void* x, const void* e, int32_t n_seg, int32_t T, int32_t C, void* out16, void* stream) {
  B200D_CHECK_ARG(x && e && out16 && n_seg > 0 && T > 0 && C % 4 == 0);
  attn_pool_kernel<<<dim3(n_seg, (C / 4 + 63) / 64), 64 * kTsSlices, 0, as_stream(stream)>>>(reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(e), T, C,
                                                            reinterpret_cast<__half*>(out16));
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_se_apply_relu_stats(const void* x, const float* gate, void* y, int32_t n_seg, int32_t T, int32_t C, void* stats16,
                                         void* stream) {
  B200D_CHECK_ARG(x && gate && y && stats16 && n_seg > 0 && T > 0 && C % 4 == 0);
  se_apply_stats_kernel<<<dim3(n_seg, (C / 4 + 63) / 64), 64 * kTsSlices, 0, as_stream(stream)>>>(
      reinterpret_cast<const __half*>(x), gate, reinterpret_cast<__half*>(y), T, C, reinterpret_cast<__half*>(stats16));
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}
