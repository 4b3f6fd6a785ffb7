// Multi-scale cosine affinity (offline_clustering.cos_similarity / getCosAffinityMatrix /
// ScalerMinMax / getMultiScaleCosAffinityMatrix of upstream NeMo), fp32 throughout so that the
// fused matrix stays within 1e-4 absolute of the CPU path.
//
//  l2_normalize   xn = x / (||x|| + eps)                          one warp per row
//  cos_affinity   cos = xn xn^T (diag := 1) + global min / max     64x64 tiles, 4x4 per thread
//  fuse_scales    fused[i][j] = sum_s w_s (cos_s[m_s(i)][m_s(j)] - min_s) / (max_s - min_s)
//                 one pass over the N x N output, the per-scale N x N expansions of upstream's
//                 repeat_interleave are never materialised.
#include "common.cuh"

namespace b200d {

__global__ void l2_normalize_kernel(const float* __restrict__ x, float* __restrict__ xn, int n, int d, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* xr = x + static_cast<size_t>(row) * d;
  float ss = 0.f;
  for (int k = lane; k < d; k += 32) ss = fmaf(xr[k], xr[k], ss);
  ss = warp_sum(ss);
  const float denom = sqrtf(ss) + eps;
  for (int k = lane; k < d; k += 32) xn[static_cast<size_t>(row) * d + k] = xr[k] / denom;
}

__device__ __forceinline__ unsigned enc_ordered(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

__global__ void minmax_init_kernel(unsigned* mm) {
  mm[0] = 0xFFFFFFFFu;
  mm[1] = 0u;
}
__global__ void minmax_decode_kernel(unsigned* mm) {
  float* f = reinterpret_cast<float*>(mm);
  const float mn = dec_ordered(mm[0]), mx = dec_ordered(mm[1]);
  f[0] = mn;
  f[1] = mx;
}

constexpr int CT = 128;  // tile edge
constexpr int CK = 16;   // k slab

// 128 x 128 tile per CTA, 8 x 8 outputs per thread (64 FMAs per 16 shared-memory floats; the first version's 64 x 64 / 4 x 4
// tiling ran at 8.7 TFLOP/s, a quarter of what the fp32 pipe gives).  Every output is ONE fused multiply-add chain over k in
// ascending order, whatever the tiling: cos[i][j] and cos[j][i] are the same chain (the symmetry sym_combine_rows relies on),
// and the row-sharded launches give the bits of the full one.
// row_lo / row_hi: the rows this launch computes (the whole matrix: 0 / n); cosm holds those rows only, row i at (i - row_lo).
__global__ void __launch_bounds__(256) cos_affinity_kernel(const float* __restrict__ xn, int n, int d, float* __restrict__ cosm,
                                                           unsigned* __restrict__ mm, int row_lo, int row_hi) {
  __shared__ __align__(16) float sa[CK][CT + 4];
  __shared__ __align__(16) float sb[CK][CT + 4];
  __shared__ float s_min[8], s_max[8];
  const int bi = row_lo + blockIdx.y * CT, bj = blockIdx.x * CT;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
  const bool vec = (d & 3) == 0;
  for (int k0 = 0; k0 < d; k0 += CK) {
    // 128 rows x 16 k per operand: thread -> (row r, 4 consecutive k), two passes of 64 rows
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int r = (threadIdx.x >> 2) + 64 * q, k4 = (threadIdx.x & 3) * 4;
      const int gi = bi + r, gj = bj + r, gk = k0 + k4;
      float va[4] = {0.f, 0.f, 0.f, 0.f}, vb[4] = {0.f, 0.f, 0.f, 0.f};
      if (vec && gk + 3 < d) {
        if (gi < row_hi) *reinterpret_cast<float4*>(va) = *reinterpret_cast<const float4*>(xn + static_cast<size_t>(gi) * d + gk);
        if (gj < n) *reinterpret_cast<float4*>(vb) = *reinterpret_cast<const float4*>(xn + static_cast<size_t>(gj) * d + gk);
      } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (gi < row_hi && gk + t < d) va[t] = xn[static_cast<size_t>(gi) * d + gk + t];
          if (gj < n && gk + t < d) vb[t] = xn[static_cast<size_t>(gj) * d + gk + t];
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        sa[k4 + t][r] = va[t];
        sb[k4 + t][r] = vb[t];
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CK; ++k) {
      float av[8], bv[8];
      // rows ty*4 + {0..3} and 64 + ty*4 + {0..3}; columns tx*4 + {0..3} and 64 + tx*4 + {0..3}: conflict-free float4 reads
      *reinterpret_cast<float4*>(av) = *reinterpret_cast<const float4*>(&sa[k][ty * 4]);
      *reinterpret_cast<float4*>(av + 4) = *reinterpret_cast<const float4*>(&sa[k][64 + ty * 4]);
      *reinterpret_cast<float4*>(bv) = *reinterpret_cast<const float4*>(&sb[k][tx * 4]);
      *reinterpret_cast<float4*>(bv + 4) = *reinterpret_cast<const float4*>(&sb[k][64 + tx * 4]);
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
  float mn = INFINITY, mx = -INFINITY;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int i = bi + (a >> 2) * 64 + ty * 4 + (a & 3);
    if (i >= row_hi) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j0 = bj + h * 64 + tx * 4;
      float v[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        v[b] = (i == j0 + b) ? 1.f : acc[a][h * 4 + b];
        if (j0 + b < n) {
          mn = fminf(mn, v[b]);
          mx = fmaxf(mx, v[b]);
        }
      }
      float* o = cosm + static_cast<size_t>(i - row_lo) * n + j0;
      if (j0 + 3 < n && (n & 3) == 0) {
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (j0 + b < n) o[b] = v[b];
      }
    }
  }
  mn = warp_min(mn);
  mx = warp_max(mx);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_min[warp] = mn; s_max[warp] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { mn = fminf(mn, s_min[w]); mx = fmaxf(mx, s_max[w]); }
    if (mn <= mx) {
      atomicMin(&mm[0], enc_ordered(mn));
      atomicMax(&mm[1], enc_ordered(mx));
    }
  }
}

constexpr int kMaxScales = 8;
struct FuseParams {
  int S;
  const float* cosm[kMaxScales];
  int ns[kMaxScales];
  const int* map[kMaxScales];
  const float* minmax[kMaxScales];
  float w[kMaxScales];
  float* fused;
  int n;
  int row_lo, row_hi;            // rows of the output this launch writes (fused holds those rows only)
  int cos_row0[kMaxScales];      // first row of scale s's matrix present in cosm[s]
};

__global__ void __launch_bounds__(256) fuse_scales_kernel(const FuseParams p) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p.n) return;
  for (int i = p.row_lo + blockIdx.y; i < p.row_hi; i += gridDim.y) {
    float acc = 0.f;
#pragma unroll 1
    for (int s = 0; s < p.S; ++s) {
      const float mn = __ldg(p.minmax[s]), mx = __ldg(p.minmax[s] + 1);
      const int mi = __ldg(p.map[s] + i) - p.cos_row0[s], mj = __ldg(p.map[s] + j);
      const float c = __ldg(p.cosm[s] + static_cast<size_t>(mi) * p.ns[s] + mj);
      acc += p.w[s] * ((c - mn) / (mx - mn));
    }
    p.fused[static_cast<size_t>(i - p.row_lo) * p.n + j] = acc;
  }
}


// Scale-interpolated embeddings of the long-form path (offline_clustering.get_scale_interpolated_embs):
// out[i] = sum_s w_s * emb_s[map_s[i]]
struct InterpParams {
  int S;
  const float* emb[kMaxScales];
  const int* map[kMaxScales];
  float w[kMaxScales];
  float* out;
  int n, d;
};

__global__ void __launch_bounds__(256) interp_scales_kernel(const InterpParams p) {
  const size_t total = static_cast<size_t>(p.n) * p.d;
  for (size_t e = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; e < total; e += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(e / p.d), c = static_cast<int>(e - static_cast<size_t>(i) * p.d);
    float acc = 0.f;
#pragma unroll 1
    for (int s = 0; s < p.S; ++s) acc += p.w[s] * __ldg(p.emb[s] + static_cast<size_t>(__ldg(p.map[s] + i)) * p.d + c);
    p.out[e] = acc;
  }
}

// out[j] = sum_i [labels[i] == labels[j]] mat[j][i]: the within-cluster affinity mass of every point, which ranks
// the merge candidates of online_clustering.run_reducer / get_closest_embeddings for all clusters in one pass.
__global__ void __launch_bounds__(256) masked_rowsum_kernel(const float* __restrict__ mat, int n, const int* __restrict__ labels,
                                                            float* __restrict__ out) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const int lab = __ldg(labels + row);
  const float* r = mat + static_cast<size_t>(row) * n;
  float acc = 0.f;
  for (int i = lane; i < n; i += 32)
    if (__ldg(labels + i) == lab) acc += __ldg(r + i);
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc;
}


// out[s] = mean over i in [seg_off[s], seg_off[s+1]) of x[idx[i]]  (fixed summation order: deterministic).
// The merged embedding of online_clustering.merge_vectors for every cluster of a long-form chunk in one launch.
__global__ void __launch_bounds__(256) gather_segment_mean_kernel(const float* __restrict__ x, int d, const int* __restrict__ idx,
                                                                  const int* __restrict__ seg_off, float* __restrict__ out) {
  const int s = blockIdx.x;
  const int a = __ldg(seg_off + s), b = __ldg(seg_off + s + 1);
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float acc = 0.f;
    for (int i = a; i < b; ++i) acc += __ldg(x + static_cast<size_t>(__ldg(idx + i)) * d + c);
    out[static_cast<size_t>(s) * d + c] = b > a ? acc / static_cast<float>(b - a) : 0.f;
  }
}

}  // namespace b200d

using namespace b200d;

extern "C" int b200d_l2_normalize(const float* x, float* xn, int32_t n, int32_t d, float eps, void* stream) {
  B200D_CHECK_ARG(x && xn && n > 0 && d > 0);
  l2_normalize_kernel<<<(n + 7) / 8, 256, 0, as_stream(stream)>>>(x, xn, n, d, eps);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_cos_affinity(const float* xn, int32_t n, int32_t d, float* cosm, float* minmax, void* stream) {
  B200D_CHECK_ARG(xn && cosm && minmax && n > 1 && d > 0);
  unsigned* mm = reinterpret_cast<unsigned*>(minmax);
  minmax_init_kernel<<<1, 1, 0, as_stream(stream)>>>(mm);
  dim3 grid((n + CT - 1) / CT, (n + CT - 1) / CT);
  cos_affinity_kernel<<<grid, 256, 0, as_stream(stream)>>>(xn, n, d, cosm, mm, 0, n);
  minmax_decode_kernel<<<1, 1, 0, as_stream(stream)>>>(mm);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_cos_affinity_rows(const float* xn, int32_t n, int32_t d, int32_t row_lo, int32_t row_hi, float* cos_rows, float* minmax,
                                       void* stream) {
  B200D_CHECK_ARG(xn && cos_rows && minmax && n > 1 && d > 0 && row_lo >= 0 && row_lo < row_hi && row_hi <= n);
  unsigned* mm = reinterpret_cast<unsigned*>(minmax);
  minmax_init_kernel<<<1, 1, 0, as_stream(stream)>>>(mm);
  dim3 grid((n + CT - 1) / CT, (row_hi - row_lo + CT - 1) / CT);
  cos_affinity_kernel<<<grid, 256, 0, as_stream(stream)>>>(xn, n, d, cos_rows, mm, row_lo, row_hi);
  minmax_decode_kernel<<<1, 1, 0, as_stream(stream)>>>(mm);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_fuse_scales(int32_t n_scales, const float* const* cos_host, const int32_t* ns_host, const int32_t* const* map_host,
                                 const float* const* minmax_host, const float* weights_host, float* fused, int32_t n_base,
                                 void* stream) {
  B200D_CHECK_ARG(n_scales > 0 && n_scales <= kMaxScales && cos_host && ns_host && map_host && minmax_host && weights_host && fused);
  B200D_CHECK_ARG(n_base > 0);
  FuseParams p;
  p.S = n_scales;
  for (int s = 0; s < n_scales; ++s) {
    B200D_CHECK_ARG(cos_host[s] && map_host[s] && minmax_host[s] && ns_host[s] > 0);
    p.cosm[s] = cos_host[s]; p.ns[s] = ns_host[s]; p.map[s] = map_host[s]; p.minmax[s] = minmax_host[s]; p.w[s] = weights_host[s];
  }
  p.fused = fused;
  p.n = n_base;
  p.row_lo = 0;
  p.row_hi = n_base;
  for (int s = 0; s < kMaxScales; ++s) p.cos_row0[s] = 0;
  dim3 grid((n_base + 255) / 256, n_base < 65535 ? n_base : 65535);
  fuse_scales_kernel<<<grid, 256, 0, as_stream(stream)>>>(p);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_fuse_scales_rows(int32_t n_scales, const float* const* cos_host, const int32_t* ns_host, const int32_t* cos_row0_host,
                                      const int32_t* const* map_host, const float* const* minmax_host, const float* weights_host,
                                      float* fused_rows, int32_t n_base, int32_t row_lo, int32_t row_hi, void* stream) {
  B200D_CHECK_ARG(n_scales > 0 && n_scales <= kMaxScales && cos_host && ns_host && cos_row0_host && map_host && minmax_host && weights_host && fused_rows);
  B200D_CHECK_ARG(n_base > 0 && row_lo >= 0 && row_lo < row_hi && row_hi <= n_base);
  FuseParams p;
  p.S = n_scales;
  for (int s = 0; s < kMaxScales; ++s) p.cos_row0[s] = 0;
  for (int s = 0; s < n_scales; ++s) {
    B200D_CHECK_ARG(cos_host[s] && map_host[s] && minmax_host[s] && ns_host[s] > 0 && cos_row0_host[s] >= 0);
    p.cosm[s] = cos_host[s]; p.ns[s] = ns_host[s]; p.map[s] = map_host[s]; p.minmax[s] = minmax_host[s]; p.w[s] = weights_host[s];
    p.cos_row0[s] = cos_row0_host[s];
  }
  p.fused = fused_rows;
  p.n = n_base;
  p.row_lo = row_lo;
  p.row_hi = row_hi;
  const int m = row_hi - row_lo;
  dim3 grid((n_base + 255) / 256, m < 65535 ? m : 65535);
  fuse_scales_kernel<<<grid, 256, 0, as_stream(stream)>>>(p);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_interp_scales(int32_t n_scales, const float* const* emb_host, const int32_t* const* map_host,
                                   const float* weights_host, float* out, int32_t n_base, int32_t d, void* stream) {
  B200D_CHECK_ARG(n_scales > 0 && n_scales <= kMaxScales && emb_host && map_host && weights_host && out && n_base > 0 && d > 0);
  InterpParams p;
  p.S = n_scales;
  for (int s = 0; s < n_scales; ++s) {
    B200D_CHECK_ARG(emb_host[s] && map_host[s]);
    p.emb[s] = emb_host[s]; p.map[s] = map_host[s]; p.w[s] = weights_host[s];
  }
  p.out = out; p.n = n_base; p.d = d;
  const size_t total = static_cast<size_t>(n_base) * d;
  const int blocks = static_cast<int>((total + 255) / 256 < static_cast<size_t>(kNumSMs) * 8 ? (total + 255) / 256 : kNumSMs * 8);
  interp_scales_kernel<<<blocks, 256, 0, as_stream(stream)>>>(p);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_masked_rowsum(const float* mat, int32_t n, const int32_t* labels, float* out, void* stream) {
  B200D_CHECK_ARG(mat && labels && out && n > 0);
  masked_rowsum_kernel<<<(n + 7) / 8, 256, 0, as_stream(stream)>>>(mat, n, labels, out);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_gather_segment_mean(const float* x, int32_t d, const int32_t* idx, const int32_t* seg_off, int32_t n_seg, float* out,
                                         void* stream) {
  B200D_CHECK_ARG(x && idx && seg_off && out && d > 0 && n_seg > 0);
  gather_segment_mean_kernel<<<n_seg, 256, 0, as_stream(stream)>>>(x, d, idx, seg_off, out);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}
