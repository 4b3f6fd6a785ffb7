// k-means on the spectral embedding (upstream offline_clustering.kmeans_plusplus_torch / kmeans_torch:
// k-means++ seeding with 30 local trials, then Lloyd iterations, <= 15, until the squared summed
// centre shift drops below 1e-4; an empty cluster is re-seeded with a random point).
//
// Upstream draws its random numbers from the CPU torch generator under torch.manual_seed(0); to
// reproduce its labels the host makes exactly those draws (first-centre index, rand(30) per further
// centre, then one randint per empty-cluster event) and hands them to the kernel.
//
// One cooperative kernel spans the chip: the points are dealt to the CTAs in contiguous chunks, every
// CTA keeps the centres / candidates in shared memory, and the few global quantities (prefix sums of
// the k-means++ potential, candidate potentials, centroid sums) are combined through small per-CTA
// partial arrays with a grid barrier and a FIXED summation order, so labels are deterministic.  The
// long-form path (k = dim = 50 on 10 000 points, three times per hour of audio) made the earlier
// single-CTA version the largest item of the step; spread over 148 SMs it is noise.
#include "common.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace b200d {

constexpr int kKmThreads = 256;
constexpr int kKmWarps = kKmThreads / 32;
constexpr int kKmMaxDim = 128;
constexpr int kKmMaxK = 128;
constexpr int kKmMaxTrials = 32;
constexpr int kKmSub = 128;  // points per sub-chunk staged in shared memory

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
  // workspace
  float* closest;    // [n]
  float* cum;        // [n]
  float* chunk_tot;  // [G]
  float* part_pot;   // [G][32]
  float* part_sum;   // [G][k][dim]
  int* part_cnt;     // [G][k]
  float* centers;    // [k][dim]
  int* counts;       // [k]
};

__device__ __forceinline__ float sqdist(const float* __restrict__ a, const float* __restrict__ b, int dim) {
  float s = 0.f;
  for (int d = 0; d < dim; ++d) {
    const float t = a[d] - b[d];
    s = fmaf(t, t, s);
  }
  return s;
}

__global__ void __launch_bounds__(kKmThreads, 1) kmeans_kernel(const KmeansParams p) {
  extern __shared__ float km_smem[];  // centres [k][dim] | previous centres [k][dim] | candidates [32][dim] | dist [128][32]
  __shared__ int s_cand_id[kKmMaxTrials];
  __shared__ float s_pot[kKmMaxTrials];
  __shared__ float s_red[kKmWarps];
  __shared__ float s_carry, s_curpot;
  __shared__ int s_best, s_fb_used, s_stop;
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = p.n, dim = p.dim, k = p.k, G = gridDim.x, g = blockIdx.x;
  const float* X = p.x;
  float* const s_cent = km_smem;
  float* const s_prev = km_smem + k * dim;
  float* const s_cand = km_smem + 2 * k * dim;
  float* const s_dist = s_cand + kKmMaxTrials * dim;  // [kKmSub][32]
  const int per = (n + G - 1) / G;
  const int i0 = min(n, g * per), i1 = min(n, i0 + per);  // this CTA's points

  // ---------------- k-means++ ----------------
  for (int d = tid; d < dim; d += kKmThreads) s_cent[d] = X[static_cast<size_t>(p.first_center) * dim + d];
  if (tid == 0) s_fb_used = 0;
  __syncthreads();
  for (int i = i0 + tid; i < i1; i += kKmThreads) p.closest[i] = sqdist(X + static_cast<size_t>(i) * dim, s_cent, dim);
  __syncthreads();

  for (int c = 1; c < k; ++c) {
    // chunk-local inclusive prefix sum of `closest`, chunk total -> global
    if (tid == 0) s_carry = 0.f;
    __syncthreads();
    for (int base = i0; base < i1; base += kKmThreads) {
      const int i = base + tid;
      float v = (i < i1) ? p.closest[i] : 0.f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      if (lane == 31) s_red[warp] = v;
      __syncthreads();
      if (warp == 0) {
        float w = lane < kKmWarps ? s_red[lane] : 0.f;
#pragma unroll
        for (int o = 1; o < kKmWarps; o <<= 1) {
          const float t = __shfl_up_sync(0xffffffffu, w, o);
          if (lane >= o) w += t;
        }
        if (lane < kKmWarps) s_red[lane] = w;
      }
      __syncthreads();
      const float carry = s_carry;
      if (i < i1) p.cum[i] = v + (warp > 0 ? s_red[warp - 1] : 0.f) + carry;
      __syncthreads();
      if (tid == 0) s_carry = carry + s_red[kKmWarps - 1];
      __syncthreads();
    }
    if (tid == 0) p.chunk_tot[g] = s_carry;
    __threadfence();
    grid.sync();
    // current potential = sum of all chunk totals (fixed order)
    if (tid == 0) {
      float tot = 0.f;
      for (int q = 0; q < G; ++q) tot += __ldcg(p.chunk_tot + q);
      s_curpot = tot;
    }
    __syncthreads();
    // candidates: searchsorted(cumsum(closest), rand * pot), left side, clamped to n - 1
    if (tid < p.n_trials) {
      const float val = p.rand_vals[(c - 1) * p.n_trials + tid] * s_curpot;
      // locate the chunk, then the element inside it
      float before = 0.f;
      int q = 0;
      for (; q < G; ++q) {
        const float t = __ldcg(p.chunk_tot + q);
        const int qa = min(n, q * per), qb = min(n, qa + per);
        if (qb > qa && before + t >= val) break;
        before += t;
      }
      int id = n - 1;
      if (q < G) {
        int lo = min(n, q * per), hi = min(n, lo + per);
        const int end = hi;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (before + __ldcg(p.cum + mid) < val) lo = mid + 1; else hi = mid;
        }
        id = lo < end ? lo : end - 1;
      }
      s_cand_id[tid] = id < n - 1 ? id : n - 1;
    }
    __syncthreads();
    for (int e = tid; e < p.n_trials * dim; e += kKmThreads) {
      const int t = e / dim, d = e - t * dim;
      s_cand[t * dim + d] = X[static_cast<size_t>(s_cand_id[t]) * dim + d];
    }
    __syncthreads();
    // potential of every candidate over this CTA's points: sum_i min(closest_i, |x_i - cand_t|^2)
    float run = 0.f;  // thread t < n_trials accumulates candidate t
    for (int base = i0; base < i1; base += kKmSub) {
      const int cnt = min(kKmSub, i1 - base);
      for (int e = tid; e < cnt * p.n_trials; e += kKmThreads) {
        const int pi = e / p.n_trials, t = e - pi * p.n_trials;
        const int i = base + pi;
        s_dist[pi * kKmMaxTrials + t] = fminf(p.closest[i], sqdist(X + static_cast<size_t>(i) * dim, s_cand + t * dim, dim));
      }
      __syncthreads();
      if (tid < p.n_trials)
        for (int pi = 0; pi < cnt; ++pi) run += s_dist[pi * kKmMaxTrials + tid];
      __syncthreads();
    }
    if (tid < p.n_trials) p.part_pot[g * kKmMaxTrials + tid] = run;
    __threadfence();
    grid.sync();
    if (tid < p.n_trials) {
      float s = 0.f;
      for (int q = 0; q < G; ++q) s += __ldcg(p.part_pot + q * kKmMaxTrials + tid);
      s_pot[tid] = s;
    }
    __syncthreads();
    if (tid == 0) {
      int best = 0;
      for (int t = 1; t < p.n_trials; ++t)
        if (s_pot[t] < s_pot[best]) best = t;
      s_best = best;
    }
    __syncthreads();
    const int best = s_best;
    for (int d = tid; d < dim; d += kKmThreads) s_cent[c * dim + d] = s_cand[best * dim + d];
    for (int i = i0 + tid; i < i1; i += kKmThreads)
      p.closest[i] = fminf(p.closest[i], sqdist(X + static_cast<size_t>(i) * dim, s_cand + best * dim, dim));
    __syncthreads();
  }

  // ---------------- Lloyd ----------------
  for (int i = i0 + tid; i < i1; i += kKmThreads) p.labels[i] = 0;
  for (int it = 0; it < p.iter_limit; ++it) {
    // assignment: 8 lanes per point, each lane scans every 8th centre; first minimum wins
    for (int base = i0; base < i1; base += kKmThreads / 8) {
      const int i = base + (tid >> 3), l8 = tid & 7;
      float bd = INFINITY;
      int bc = 0x7fffffff;
      if (i < i1) {
        const float* xi = X + static_cast<size_t>(i) * dim;
        for (int c = l8; c < k; c += 8) {
          const float dsq = sqdist(xi, s_cent + c * dim, dim);
          if (dsq < bd) { bd = dsq; bc = c; }
        }
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, o);
        const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
        if (od < bd || (od == bd && oc < bc)) { bd = od; bc = oc; }
      }
      if (i < i1 && l8 == 0) p.labels[i] = bc;
    }
    for (int e = tid; e < k * dim; e += kKmThreads) s_prev[e] = s_cent[e];
    __syncthreads();
    // per-CTA partial centroid sums: warp w owns clusters w, w + 8, ...
    for (int c = warp; c < k; c += kKmWarps) {
      int cnt = 0;
      for (int i = i0 + lane; i < i1; i += 32) cnt += (p.labels[i] == c);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      if (lane == 0) p.part_cnt[g * k + c] = cnt;
      for (int d0 = 0; d0 < dim; d0 += 8) {
        float acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = 0.f;
        if (cnt > 0) {
          for (int i = i0 + lane; i < i1; i += 32) {
            if (p.labels[i] == c) {
              const float* xi = X + static_cast<size_t>(i) * dim + d0;
#pragma unroll
              for (int q = 0; q < 8; ++q)
                if (d0 + q < dim) acc[q] += xi[q];
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float v = warp_sum(acc[q]);
          if (lane == 0 && d0 + q < dim) p.part_sum[(static_cast<size_t>(g) * k + c) * dim + d0 + q] = v;
        }
      }
    }
    __threadfence();
    grid.sync();
    // combine the partials in CTA order; entry e = (cluster, dimension) is owned by one thread of the grid
    for (int e = g * kKmThreads + tid; e < k * dim; e += G * kKmThreads) {
      const int c = e / dim;
      float s = 0.f;
      int cnt = 0;
      for (int q = 0; q < G; ++q) {
        s += __ldcg(p.part_sum + static_cast<size_t>(q) * k * dim + e);
        cnt += __ldcg(p.part_cnt + q * k + c);
      }
      p.centers[e] = cnt > 0 ? s / static_cast<float>(cnt) : 0.f;
      if (e == c * dim) p.counts[c] = cnt;
    }
    __threadfence();
    grid.sync();
    for (int e = tid; e < k * dim; e += kKmThreads) s_cent[e] = __ldcg(p.centers + e);
    __syncthreads();
    // empty clusters take a random point, consumed in cluster order like upstream's loop (same in every CTA)
    if (tid == 0) {
      for (int c = 0; c < k; ++c) {
        if (__ldcg(p.counts + c) == 0) {
          const int idx = (s_fb_used < p.n_fallback) ? p.fallback[s_fb_used] : 0;
          ++s_fb_used;
          for (int d = 0; d < dim; ++d) s_cent[c * dim + d] = X[static_cast<size_t>(idx) * dim + d];
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
      float shift = 0.f;
      for (int c = lane; c < k; c += 32) {
        float s = 0.f;
        for (int d = 0; d < dim; ++d) {
          const float t = s_cent[c * dim + d] - s_prev[c * dim + d];
          s = fmaf(t, t, s);
        }
        shift += sqrtf(s);
      }
      shift = warp_sum(shift);
      if (lane == 0) s_stop = (shift * shift < p.threshold) ? 1 : 0;
    }
    __syncthreads();
    if (s_stop) break;  // identical in every CTA: no barrier is skipped by a subset of the grid
  }
}

static int kmeans_grid(int n) {
  int g = (n + 63) / 64;
  if (g < 1) g = 1;
  if (g > kNumSMs) g = kNumSMs;
  return g;
}

}  // namespace b200d

using namespace b200d;

extern "C" size_t b200d_kmeans_workspace_bytes(int32_t n, int32_t dim, int32_t n_clusters, int32_t n_trials) {
  (void)n_trials;
  if (n <= 0 || dim <= 0 || n_clusters <= 0) return 0;
  const size_t G = static_cast<size_t>(kmeans_grid(n));
  const size_t kd = static_cast<size_t>(n_clusters) * dim;
  return sizeof(float) * (2 * static_cast<size_t>(n) + G + G * kKmMaxTrials + G * kd + G * n_clusters + kd + n_clusters) + 64;
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
  const int G = kmeans_grid(n);
  const size_t kd = static_cast<size_t>(n_clusters) * dim;
  KmeansParams p;
  p.x = x; p.n = n; p.dim = dim; p.k = n_clusters; p.first_center = first_center; p.rand_vals = rand_vals; p.n_trials = n_trials;
  p.fallback = fallback_idx; p.n_fallback = n_fallback; p.iter_limit = iter_limit; p.threshold = threshold; p.labels = labels;
  float* w = reinterpret_cast<float*>(ws);
  p.closest = w; w += n;
  p.cum = w; w += n;
  p.chunk_tot = w; w += G;
  p.part_pot = w; w += static_cast<size_t>(G) * kKmMaxTrials;
  p.part_sum = w; w += static_cast<size_t>(G) * kd;
  p.part_cnt = reinterpret_cast<int*>(w); w += static_cast<size_t>(G) * n_clusters;
  p.centers = w; w += kd;
  p.counts = reinterpret_cast<int*>(w);
  const size_t smem = (2 * kd + static_cast<size_t>(kKmMaxTrials) * dim + static_cast<size_t>(kKmSub) * kKmMaxTrials) * sizeof(float);
  static bool attr_set_dev[kMaxDevices] = {};  // kernel attributes are per device
  const int attr_dev = current_device();
  const bool attr_known = attr_dev >= 0 && attr_dev < kMaxDevices;
  if (!attr_known || !attr_set_dev[attr_dev]) {
    B200D_CHECK_CUDA(cudaFuncSetAttribute(kmeans_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (2 * kKmMaxK * kKmMaxDim + kKmMaxTrials * kKmMaxDim + kKmSub * kKmMaxTrials) * sizeof(float)));
    if (attr_known) attr_set_dev[attr_dev] = true;
  }
  void* args[] = {const_cast<KmeansParams*>(&p)};
  B200D_CHECK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kmeans_kernel), dim3(G), dim3(kKmThreads), args, smem, as_stream(stream)));
  return B200D_OK;
}
