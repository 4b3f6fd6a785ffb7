// Batched symmetric eigenvalues for the NME-SC p-sweep (upstream offline_clustering.NMESC.getEigRatio
// -> estimateNumofSpeakers -> eigDecompose: one torch.linalg.eigh per p value, eigenvalues only).
//
// All Laplacians of the sweep (<= 64 matrices of n <= 1024) are reduced to tridiagonal form
// together by ONE cooperative kernel that spans the whole chip: matrix b is owned by `cpm`
// CTAs, rows of the trailing block are dealt cyclically to their warps, and every Householder
// step is a single fused pass over the trailing block (apply the previous step's rank-2 update
// while accumulating the symmetric mat-vec of the current step), followed by one grid barrier.
// The sweep's working set (30 x 600^2 fp32 = 43 MB) lives in the 126 MB L2.  Eigenvalues are
// then isolated by 32-way multisection on the Sturm sequence in fp64, one warp per wanted
// eigenvalue: only the max_num_speakers + 1 smallest and the largest are needed by
// getLamdaGaplist / getEigRatio.
#include "common.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace b200d {

constexpr int kTriThreads = 512;
constexpr int kTriWarps = kTriThreads / 32;
constexpr int kMaxEigN = 2048;

__device__ __forceinline__ float block_sum_512(float v, float* s_red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // protect s_red reuse
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float t = (lane < kTriWarps) ? s_red[lane] : 0.f;
  t = warp_sum(t);
  return t;  // every thread holds the total
}

// a [batch][n][n] (destroyed), d [batch][n], e [batch][n], pbuf [batch][2][n]
__global__ void __launch_bounds__(kTriThreads, 2) tridiag_kernel(float* __restrict__ a, int n, int cpm, float* __restrict__ d,
                                                                  float* __restrict__ e, float* __restrict__ pbuf, unsigned* __restrict__ bar) {
  extern __shared__ float sm[];
  float* s_vp = sm;           // v' (previous step's Householder vector), indexed by absolute row
  float* s_wp = sm + n;       // w'
  float* s_v = sm + 2 * n;    // v  (current step)
  float* s_r = sm + 3 * n;    // updated pivot row / scratch
  __shared__ float s_red[kTriWarps];

  const int b = blockIdx.x / cpm, c = blockIdx.x % cpm;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* A = a + static_cast<size_t>(b) * n * n;
  float* pb = pbuf + static_cast<size_t>(b) * 2 * n;
  const int slots = cpm * kTriWarps;
  const int slot = c * kTriWarps + warp;

  for (int i = tid; i < n; i += kTriThreads) { s_vp[i] = 0.f; s_wp[i] = 0.f; s_v[i] = 0.f; }
  __syncthreads();
  float vpj = 0.f, wpj = 0.f;  // v'_j, w'_j of the pending update

  for (int j = 0; j < n - 1; ++j) {
    // ---- pivot row j with the pending rank-2 update applied (redundantly in every CTA of the matrix)
    vpj = s_vp[j];
    wpj = s_wp[j];
    const float* rowj = A + static_cast<size_t>(j) * n;
    float ss = 0.f;
    for (int k = j + tid; k < n; k += kTriThreads) {
      const float r = __ldcg(rowj + k) - vpj * s_wp[k] - wpj * s_vp[k];  // L2 only: other CTAs write A
      s_r[k] = r;
      if (k > j + 1) ss = fmaf(r, r, ss);
    }
    ss = block_sum_512(ss, s_red);  // sum of squares of x[1:]
    __syncthreads();
    const float x0 = s_r[j + 1];
    const float xnorm2 = ss + x0 * x0;
    float alpha, beta;
    if (ss == 0.f) {  // nothing to annihilate (also the last step): H = I
      alpha = x0;
      beta = 0.f;
    } else {
      const float xn = sqrtf(xnorm2);
      alpha = -copysignf(xn, x0);
      beta = 1.f / (xn * (xn + fabsf(x0)));
    }
    if (c == 0 && tid == 0) {
      d[static_cast<size_t>(b) * n + j] = s_r[j];
      e[static_cast<size_t>(b) * n + j] = alpha;
    }
    for (int k = j + 1 + tid; k < n; k += kTriThreads) s_v[k] = (beta == 0.f) ? 0.f : ((k == j + 1) ? (x0 - alpha) : s_r[k]);
    __syncthreads();

    // ---- fused pass over the owned rows of the trailing block (rows / cols j+1 .. n-1)
    float* pw = pb + static_cast<size_t>(j & 1) * n;
    // two rows of the warp are processed together: eight independent 128-byte requests in flight per warp (the pass is
    // L2-latency bound, not bandwidth bound)
    for (int i = j + 1 + slot; i < n; i += 2 * slots) {
      const int i2 = i + slots;
      const bool two = i2 < n;
      float* row = A + static_cast<size_t>(i) * n;
      float* row2 = A + static_cast<size_t>(two ? i2 : i) * n;
      const float vpi = s_vp[i], wpi = s_wp[i];
      const float vpi2 = two ? s_vp[i2] : 0.f, wpi2 = two ? s_wp[i2] : 0.f;
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
      float bcc0 = 0.f, bcc1 = 0.f, bcc2 = 0.f, bcc3 = 0.f;
      int k = j + 1 + lane;
      for (; k + 96 < n; k += 128) {
        const float a0 = __ldcg(row + k), a1 = __ldcg(row + k + 32), a2 = __ldcg(row + k + 64), a3 = __ldcg(row + k + 96);
        float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
        if (two) { b0 = __ldcg(row2 + k); b1 = __ldcg(row2 + k + 32); b2 = __ldcg(row2 + k + 64); b3 = __ldcg(row2 + k + 96); }
        const float w0 = s_wp[k], w1 = s_wp[k + 32], w2 = s_wp[k + 64], w3 = s_wp[k + 96];
        const float p0 = s_vp[k], p1 = s_vp[k + 32], p2 = s_vp[k + 64], p3 = s_vp[k + 96];
        const float u0 = s_v[k], u1 = s_v[k + 32], u2 = s_v[k + 64], u3 = s_v[k + 96];
        const float v0 = a0 - vpi * w0 - wpi * p0, v1 = a1 - vpi * w1 - wpi * p1;
        const float v2 = a2 - vpi * w2 - wpi * p2, v3 = a3 - vpi * w3 - wpi * p3;
        __stcg(row + k, v0); __stcg(row + k + 32, v1); __stcg(row + k + 64, v2); __stcg(row + k + 96, v3);
        acc0 = fmaf(v0, u0, acc0); acc1 = fmaf(v1, u1, acc1); acc2 = fmaf(v2, u2, acc2); acc3 = fmaf(v3, u3, acc3);
        if (two) {
          const float y0 = b0 - vpi2 * w0 - wpi2 * p0, y1 = b1 - vpi2 * w1 - wpi2 * p1;
          const float y2 = b2 - vpi2 * w2 - wpi2 * p2, y3 = b3 - vpi2 * w3 - wpi2 * p3;
          __stcg(row2 + k, y0); __stcg(row2 + k + 32, y1); __stcg(row2 + k + 64, y2); __stcg(row2 + k + 96, y3);
          bcc0 = fmaf(y0, u0, bcc0); bcc1 = fmaf(y1, u1, bcc1); bcc2 = fmaf(y2, u2, bcc2); bcc3 = fmaf(y3, u3, bcc3);
        }
      }
      for (; k < n; k += 32) {
        const float v = __ldcg(row + k) - vpi * s_wp[k] - wpi * s_vp[k];
        __stcg(row + k, v);
        acc0 = fmaf(v, s_v[k], acc0);
        if (two) {
          const float y = __ldcg(row2 + k) - vpi2 * s_wp[k] - wpi2 * s_vp[k];
          __stcg(row2 + k, y);
          bcc0 = fmaf(y, s_v[k], bcc0);
        }
      }
      const float acc = warp_sum((acc0 + acc1) + (acc2 + acc3));
      const float bcc = warp_sum((bcc0 + bcc1) + (bcc2 + bcc3));
      if (lane == 0) {
        pw[i] = beta * acc;
        if (two) pw[i2] = beta * bcc;
      }
    }
    // barrier among the cpm CTAs of THIS matrix only (the matrices are independent; all CTAs are co-resident under the
    // cooperative launch): monotone arrival counter in global memory
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(bar + b, 1u);
      const unsigned target = static_cast<unsigned>(j + 1) * static_cast<unsigned>(cpm);
      while (*reinterpret_cast<volatile unsigned*>(bar + b) < target) __nanosleep(32);
      __threadfence();
    }
    __syncthreads();

    // ---- w = p - (beta/2 v^T p) v ; becomes the pending update
    float kk = 0.f;
    for (int k = j + 1 + tid; k < n; k += kTriThreads) {
      const float pv = __ldcg(pw + k);
      s_r[k] = pv;
      kk = fmaf(s_v[k], pv, kk);
    }
    kk = block_sum_512(kk, s_red) * 0.5f * beta;
    __syncthreads();
    for (int k = j + 1 + tid; k < n; k += kTriThreads) {
      const float v = s_v[k];
      s_vp[k] = v;
      s_wp[k] = s_r[k] - kk * v;
    }
    __syncthreads();
  }
  // last diagonal entry: A[n-1][n-1] received its final update in the pass of step n-2 (made visible by
  // that step's grid barrier) and the pending update of that step is zero (beta == 0).
  if (c == 0 && tid == 0) {
    d[static_cast<size_t>(b) * n + n - 1] = __ldcg(A + static_cast<size_t>(n - 1) * n + n - 1);
    e[static_cast<size_t>(b) * n + n - 1] = 0.f;
  }
}

// Number of eigenvalues of the symmetric tridiagonal (d, e) strictly below x = number of sign changes in the sequence of
// leading principal minors p_0 = 1, p_1 = d_0 - x, p_{i+1} = (d_i - x) p_i - e_{i-1}^2 p_{i-1}.  The division-free form is a
// chain of two dependent multiply-adds per row instead of an fp64 division (~10x shorter); the pair (p_i, p_{i-1}) is
// rescaled every 8 rows so that it stays inside the fp64 range (growth per row is bounded by the Gershgorin width).
__device__ __forceinline__ int sturm_count(const double* __restrict__ d, const double* __restrict__ e2, int n, double x, double tiny) {
  double pm = 1.0, p = d[0] - x;
  if (p == 0.0) p = -tiny;
  int cnt = p < 0.0 ? 1 : 0;
  for (int i = 1; i < n; ++i) {
    double pn = fma(d[i] - x, p, -e2[i - 1] * pm);
    if (pn == 0.0) pn = (p < 0.0) ? tiny : -tiny;  // a zero minor counts as a sign change, like q = -tiny in the ratio form
    cnt += ((pn < 0.0) != (p < 0.0)) ? 1 : 0;
    pm = p;
    p = pn;
    if ((i & 7) == 0) {
      const double a = fabs(p);
      if (a > 1e120) { p *= 1e-120; pm *= 1e-120; }
      else if (a < 1e-120) { p *= 1e120; pm *= 1e120; }
    }
  }
  return cnt;
}

// evals [batch][n_low + 1]: the n_low smallest eigenvalues ascending, then the largest.
__global__ void __launch_bounds__(1024) bisect_kernel(const float* __restrict__ d, const float* __restrict__ e, int n, int n_low,
                                                       float* __restrict__ evals) {
  extern __shared__ double smd[];
  double* s_d = smd;
  double* s_e2 = smd + n;
  __shared__ double s_lo, s_hi;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const float* db = d + static_cast<size_t>(b) * n;
  const float* eb = e + static_cast<size_t>(b) * n;
  for (int i = tid; i < n; i += blockDim.x) {
    s_d[i] = static_cast<double>(db[i]);
    const double ev = static_cast<double>(eb[i]);
    s_e2[i] = ev * ev;
  }
  __syncthreads();
  if (warp == 0) {  // Gershgorin interval
    double lo = 1e300, hi = -1e300;
    for (int i = lane; i < n; i += 32) {
      const double r = (i > 0 ? sqrt(s_e2[i - 1]) : 0.0) + (i < n - 1 ? sqrt(s_e2[i]) : 0.0);
      lo = fmin(lo, s_d[i] - r);
      hi = fmax(hi, s_d[i] + r);
    }
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) {
      const double pad = 1e-6 * fmax(fabs(lo), fabs(hi)) + 1e-12;
      s_lo = lo - pad;
      s_hi = hi + pad;
    }
  }
  __syncthreads();
  const double glo = s_lo, ghi = s_hi;
  const double tiny = 1e-290 + 1e-30 * fmax(fabs(glo), fabs(ghi));
  const int wanted = n_low + 1;
  for (int w = warp; w < wanted; w += nwarps) {
    const int k = (w < n_low) ? w : (n - 1);  // 0-based index of the eigenvalue, ascending
    double lo = glo, hi = ghi;                // count(lo) <= k < count(hi)
    for (int round = 0; round < 7; ++round) {  // 33^7 = 4e10: below fp32 resolution of the Gershgorin width
      const double step = (hi - lo) / 33.0;
      const double x = lo + step * (lane + 1);
      const int cnt = sturm_count(s_d, s_e2, n, x, tiny);
      const unsigned above = __ballot_sync(0xffffffffu, cnt >= k + 1);  // lanes whose x is an upper bound
      const int first = above ? (__ffs(above) - 1) : 32;
      const double nlo = (first == 0) ? lo : lo + step * first;
      const double nhi = (first == 32) ? hi : lo + step * (first + 1);
      lo = nlo;
      hi = nhi;
      if (hi - lo <= 1e-13 * fmax(fabs(lo), fabs(hi))) break;
    }
    if (lane == 0) evals[static_cast<size_t>(b) * wanted + w] = static_cast<float>(0.5 * (lo + hi));
  }
}

}  // namespace b200d

using namespace b200d;

static int tridiag_cpm(int batch, int blocks_per_sm) {
  int cpm = (kNumSMs * blocks_per_sm) / batch;  // CTAs per matrix; the whole grid must be co-resident (cooperative launch)
  if (cpm < 1) cpm = 1;
  if (cpm > 16) cpm = 16;
  return cpm;
}

extern "C" size_t b200d_eigvals_workspace_bytes(int32_t batch, int32_t n) {
  if (batch <= 0 || n <= 0) return 0;
  return static_cast<size_t>(batch) * n * 4 * sizeof(float) + static_cast<size_t>(batch) * sizeof(unsigned);  // d, e, p[2], barrier counters
}

extern "C" int b200d_eigvals_batched(float* a, int32_t batch, int32_t n, int32_t n_low, float* evals, void* ws, size_t ws_bytes,
                                     void* stream) {
  return b200d_eigvals_batched_layout(a, batch, batch, n, n_low, evals, ws, ws_bytes, stream);
}

// layout_batch >= batch: every matrix is tridiagonalised by as many CTAs -- i.e. with the reduction order -- as in a launch of
// layout_batch matrices, so a SUBSET of a batch computed alone (the p values of the NME sweep dealt to the ranks of a
// row-sharded recording) gives bit for bit the eigenvalues the full batch gives.
extern "C" int b200d_eigvals_batched_layout(float* a, int32_t batch, int32_t layout_batch, int32_t n, int32_t n_low, float* evals, void* ws,
                                            size_t ws_bytes, void* stream) {
  B200D_CHECK_ARG(a && evals && ws);
  B200D_CHECK_ARG(batch > 0 && layout_batch >= batch && layout_batch <= kNumSMs && n >= 2 && n <= kMaxEigN && n_low >= 1 && n_low <= n);
  static int blocks_per_sm = 0;  // occupancy of tridiag_kernel: a property of the kernel and sm_100, the same on every device
  if (ws_bytes < b200d_eigvals_workspace_bytes(batch, n))
    return b200d::set_error(B200D_EWORKSPACE, "%s: workspace too small%s", "b200d_eigvals_batched");
  cudaStream_t s = as_stream(stream);
  float* d = reinterpret_cast<float*>(ws);
  float* e = d + static_cast<size_t>(batch) * n;
  float* pbuf = e + static_cast<size_t>(batch) * n;
  unsigned* bar = reinterpret_cast<unsigned*>(pbuf + static_cast<size_t>(batch) * 2 * n);
  B200D_CHECK_CUDA(cudaMemsetAsync(bar, 0, static_cast<size_t>(batch) * sizeof(unsigned), s));
  int n_arg = n;
  const size_t smem = static_cast<size_t>(4) * n * sizeof(float);
  static bool attr_set_dev[kMaxDevices] = {};  // kernel attributes are per device
  const int attr_dev = current_device();
  const bool attr_known = attr_dev >= 0 && attr_dev < kMaxDevices;
  if (!attr_known || !attr_set_dev[attr_dev]) {
    B200D_CHECK_CUDA(cudaFuncSetAttribute(tridiag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kMaxEigN * sizeof(float)));
    B200D_CHECK_CUDA(cudaFuncSetAttribute(bisect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kMaxEigN * sizeof(double)));
    if (attr_known) attr_set_dev[attr_dev] = true;
  }
  if (blocks_per_sm == 0) {
    int occ = 1;
    B200D_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tridiag_kernel, kTriThreads, 4 * kMaxEigN * sizeof(float)));
    blocks_per_sm = occ < 1 ? 1 : (occ > 2 ? 2 : occ);
  }
  int cpm = tridiag_cpm(layout_batch, blocks_per_sm);
  void* args[] = {&a, &n_arg, &cpm, &d, &e, &pbuf, &bar};
  B200D_CHECK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(tridiag_kernel), dim3(batch * cpm), dim3(kTriThreads), args, smem, s));
  const int wanted = n_low + 1;
  int threads = 32 * (wanted < 32 ? wanted : 32);
  bisect_kernel<<<batch, threads, 2 * static_cast<size_t>(n) * sizeof(double), s>>>(d, e, n, n_low, evals);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}
