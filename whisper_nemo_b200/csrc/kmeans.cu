// k-means on the spectral embedding (upstream offline_clustering.kmeans_plusplus_torch / kmeans_torch:
// k-means++ seeding with 30 local trials, then Lloyd iterations, <= 15, until the squared summed
// centre shift drops below 1e-4; an empty cluster is re-seeded with a random point).
//
// Upstream draws its random numbers from the CPU torch generator under torch.manual_seed(0); to
// reproduce its labels the host makes exactly those draws (first-centre index, rand(30) per further
// centre, then one randint per empty-cluster event) and hands them to the kernel.
//
// The problem is tiny (N <= 60 k points of dimension <= 64, k <= 64), so one CTA of 1024 threads
// does all of it without leaving the SM: points are streamed from L2, centres / candidates / partial
// sums live in shared memory, and every reduction has a fixed order (results are deterministic).
#include "common.cuh"

namespace b200d {

constexpr int kKmThreads = 1024;
constexpr int kKmWarps = kKmThreads / 32;
constexpr int kKmMaxDim = 128;
constexpr int kKmMaxK = 128;
constexpr int kKmMaxTrials = 32;

struct KmeansParams {
  const float* x;
  int n, dim, k;
  int first_center;
  const float* rand_vals;  // [k-1][n_trials]
  int n_trials;
  const int* fallback;
  int n_fallback;
  int iter_limit;
  float threshold;
  int* labels;
  float* closest;  // [n]
  float* cum;      // [n]
};

__device__ __forceinline__ float block_sum_1024(float v, float* s_red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float t = s_red[lane];
  return warp_sum(t);
}

__device__ __forceinline__ float sqdist(const float* __restrict__ a, const float* __restrict__ b, int dim) {
  float s = 0.f;
  for (int d = 0; d < dim; ++d) {
    const float t = a[d] - b[d];
    s = fmaf(t, t, s);
  }
  return s;
}

__global__ void __launch_bounds__(kKmThreads, 1) kmeans_kernel(const KmeansParams p) {
  extern __shared__ float km_smem[];  // centres [k][dim] | previous centres [k][dim] | candidates [n_trials][dim]
  __shared__ int s_cand_id[kKmMaxTrials];
  __shared__ float s_pot[kKmMaxTrials];
  __shared__ float s_part[kKmWarps][kKmMaxTrials];
  __shared__ float s_red[kKmWarps];
  __shared__ float s_carry, s_curpot;
  __shared__ int s_best, s_count[kKmMaxK], s_fb_used, s_stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = p.n, dim = p.dim, k = p.k;
  const float* X = p.x;
  float* const s_cent_base = km_smem;
  float* const s_prev_base = km_smem + k * dim;
  float* const s_cand_base = km_smem + 2 * k * dim;
  auto s_cent = [&](int c) -> float* { return s_cent_base + c * dim; };
  auto s_prev = [&](int c) -> float* { return s_prev_base + c * dim; };
  auto s_cand = [&](int t) -> float* { return s_cand_base + t * dim; };

  // ---------------- k-means++ ----------------
  for (int d = tid; d < dim; d += kKmThreads) s_cent(0)[d] = X[static_cast<size_t>(p.first_center) * dim + d];
  if (tid == 0) s_fb_used = 0;
  __syncthreads();
  float part = 0.f;
  for (int i = tid; i < n; i += kKmThreads) {
    const float dsq = sqdist(X + static_cast<size_t>(i) * dim, s_cent(0), dim);
    p.closest[i] = dsq;
    part += dsq;
  }
  {
    const float tot = block_sum_1024(part, s_red);
    if (tid == 0) s_curpot = tot;
  }
  __syncthreads();

  for (int c = 1; c < k; ++c) {
    // inclusive prefix sum of closest (chunks of 1024 with a running carry)
    if (tid == 0) s_carry = 0.f;
    __syncthreads();
    for (int base = 0; base < n; base += kKmThreads) {
      const int i = base + tid;
      float v = (i < n) ? p.closest[i] : 0.f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      if (lane == 31) s_red[warp] = v;
      __syncthreads();
      if (warp == 0) {
        float w = s_red[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float t = __shfl_up_sync(0xffffffffu, w, o);
          if (lane >= o) w += t;
        }
        s_red[lane] = w;
      }
      __syncthreads();
      const float carry = s_carry;
      const float pre = (warp > 0 ? s_red[warp - 1] : 0.f) + carry;
      if (i < n) p.cum[i] = v + pre;
      __syncthreads();
      if (tid == 0) s_carry = carry + s_red[kKmWarps - 1];
      __syncthreads();
    }
    // candidates: searchsorted(cum, rand * pot), left side, clamped to n - 1
    if (tid < p.n_trials) {
      const float val = p.rand_vals[(c - 1) * p.n_trials + tid] * s_curpot;
      int lo = 0, hi = n;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (p.cum[mid] < val) lo = mid + 1; else hi = mid;
      }
      s_cand_id[tid] = lo < n - 1 ? lo : n - 1;
    }
    __syncthreads();
    for (int e = tid; e < p.n_trials * dim; e += kKmThreads) {
      const int t = e / dim, d = e - t * dim;
      s_cand(t)[d] = X[static_cast<size_t>(s_cand_id[t]) * dim + d];
    }
    __syncthreads();
    // potential of every candidate: sum_i min(closest_i, |x_i - cand_t|^2)
    float pot[kKmMaxTrials];
#pragma unroll
    for (int t = 0; t < kKmMaxTrials; ++t) pot[t] = 0.f;
    for (int i = tid; i < n; i += kKmThreads) {
      const float* xi = X + static_cast<size_t>(i) * dim;
      const float cl = p.closest[i];
#pragma unroll
      for (int t = 0; t < kKmMaxTrials; ++t) {
        if (t < p.n_trials) pot[t] += fminf(cl, sqdist(xi, s_cand(t), dim));
      }
    }
#pragma unroll
    for (int t = 0; t < kKmMaxTrials; ++t) {
      const float v = warp_sum(pot[t]);
      if (lane == 0) s_part[warp][t] = v;
    }
    __syncthreads();
    if (tid < p.n_trials) {
      float s = 0.f;
      for (int w = 0; w < kKmWarps; ++w) s += s_part[w][tid];
      s_pot[tid] = s;
    }
    __syncthreads();
    if (tid == 0) {
      int best = 0;
      for (int t = 1; t < p.n_trials; ++t)
        if (s_pot[t] < s_pot[best]) best = t;
      s_best = best;
      s_curpot = s_pot[best];
    }
    __syncthreads();
    const int best = s_best;
    for (int d = tid; d < dim; d += kKmThreads) s_cent(c)[d] = s_cand(best)[d];
    for (int i = tid; i < n; i += kKmThreads)
      p.closest[i] = fminf(p.closest[i], sqdist(X + static_cast<size_t>(i) * dim, s_cand(best), dim));
    __syncthreads();
  }

  // ---------------- Lloyd ----------------
  for (int i = tid; i < n; i += kKmThreads) p.labels[i] = 0;
  __syncthreads();
  for (int it = 0; it < p.iter_limit; ++it) {
    for (int i = tid; i < n; i += kKmThreads) {
      const float* xi = X + static_cast<size_t>(i) * dim;
      float bd = sqdist(xi, s_cent(0), dim);
      int bc = 0;
      for (int c = 1; c < k; ++c) {
        const float dsq = sqdist(xi, s_cent(c), dim);
        if (dsq < bd) { bd = dsq; bc = c; }
      }
      p.labels[i] = bc;
    }
    for (int e = tid; e < k * dim; e += kKmThreads) s_prev_base[e] = s_cent_base[e];
    __syncthreads();
    // centroid update: warp w owns clusters w, w + 32, ...; fixed summation order
    for (int c = warp; c < k; c += kKmWarps) {
      int cnt = 0;
      for (int i = lane; i < n; i += 32) cnt += (p.labels[i] == c);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (lane == 0) s_count[c] = cnt;
      if (cnt > 0) {
        for (int d0 = 0; d0 < dim; d0 += 8) {
          float acc[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[q] = 0.f;
          for (int i = lane; i < n; i += 32) {
            if (p.labels[i] == c) {
              const float* xi = X + static_cast<size_t>(i) * dim + d0;
#pragma unroll
              for (int q = 0; q < 8; ++q)
                if (d0 + q < dim) acc[q] += xi[q];
            }
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float v = warp_sum(acc[q]);
            if (lane == 0 && d0 + q < dim) s_cent(c)[d0 + q] = v / static_cast<float>(cnt);
          }
        }
      }
    }
    __syncthreads();
    // empty clusters take a random point, consumed in cluster order like upstream's loop
    if (tid == 0) {
      for (int c = 0; c < k; ++c) {
        if (s_count[c] == 0) {
          const int idx = (s_fb_used < p.n_fallback) ? p.fallback[s_fb_used] : 0;
          ++s_fb_used;
          for (int d = 0; d < dim; ++d) s_cent(c)[d] = X[static_cast<size_t>(idx) * dim + d];
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
      float shift = 0.f;
      for (int c = lane; c < k; c += 32) {
        float s = 0.f;
        for (int d = 0; d < dim; ++d) {
          const float t = s_cent(c)[d] - s_prev(c)[d];
          s = fmaf(t, t, s);
        }
        shift += sqrtf(s);
      }
      shift = warp_sum(shift);
      if (lane == 0) s_stop = (shift * shift < p.threshold) ? 1 : 0;
    }
    __syncthreads();
    if (s_stop) break;
  }
}

}  // namespace b200d

using namespace b200d;

extern "C" size_t b200d_kmeans_workspace_bytes(int32_t n, int32_t dim, int32_t n_clusters, int32_t n_trials) {
  (void)dim; (void)n_clusters; (void)n_trials;
  return n > 0 ? static_cast<size_t>(n) * 2 * sizeof(float) : 0;
}

extern "C" int b200d_kmeans(const float* x, int32_t n, int32_t dim, int32_t n_clusters, int32_t first_center, const float* rand_vals,
                            int32_t n_trials, const int32_t* fallback_idx, int32_t n_fallback, int32_t iter_limit, float threshold,
                            int32_t* labels, void* ws, size_t ws_bytes, void* stream) {
  B200D_CHECK_ARG(x && labels && ws && n > 0 && dim > 0 && dim <= kKmMaxDim && n_clusters >= 1 && n_clusters <= kKmMaxK);
  B200D_CHECK_ARG(first_center >= 0 && first_center < n && n_trials >= 1 && n_trials <= kKmMaxTrials && iter_limit >= 0);
  B200D_CHECK_ARG(n_clusters == 1 || rand_vals);
  B200D_CHECK_ARG(n_fallback == 0 || fallback_idx);
  if (ws_bytes < b200d_kmeans_workspace_bytes(n, dim, n_clusters, n_trials))
    return b200d::set_error(B200D_EWORKSPACE, "%s: workspace too small%s", "b200d_kmeans");
  KmeansParams p;
  p.x = x; p.n = n; p.dim = dim; p.k = n_clusters; p.first_center = first_center; p.rand_vals = rand_vals; p.n_trials = n_trials;
  p.fallback = fallback_idx; p.n_fallback = n_fallback; p.iter_limit = iter_limit; p.threshold = threshold; p.labels = labels;
  p.closest = reinterpret_cast<float*>(ws);
  p.cum = p.closest + n;
  const size_t smem = static_cast<size_t>(2 * n_clusters + n_trials) * dim * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    B200D_CHECK_CUDA(cudaFuncSetAttribute(kmeans_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (2 * kKmMaxK + kKmMaxTrials) * kKmMaxDim * sizeof(float)));
    attr_set = true;
  }
  kmeans_kernel<<<1, kKmThreads, smem, as_stream(stream)>>>(p);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}
