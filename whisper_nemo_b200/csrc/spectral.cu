// Small dense helpers of the Chebyshev-filtered subspace iteration that replaces the full
// torch.linalg.eigh(N x N) of upstream SpectralClustering.getSpectralEmbeddings
// (nemo/collections/asr/parts/utils/offline_clustering.py).  The O(N^2 b) work -- A*V on the
// binarised affinity -- runs on the tcgen05 GEMM (gemm_tcgen05.cu, B200D_EPI_CHEB); what is left
// is O(N b^2) and O(b^3):
//   gram        G = X^T Y                      (block partials in fp32, fixed-order fp64 combine)
//   small_eig   b x b symmetric eigenproblem   (parallel cyclic Jacobi, fp64, one CTA)
//               or scaled Cholesky inverse     (CholQR orthonormalisation)
//   right_mul   Y = X Q  (+ the 3-way bf16 split of Y, transposed: the next GEMM's W operand)
//   resid       r_j = || W_j - theta_j X_j ||^2
#include "common.cuh"

#include <cuda_bf16.h>

namespace b200d {

constexpr int kGramRows = 256;  // rows per CTA
constexpr int kMaxB = 64;       // block width of the subspace iteration (gram / right_mul / resid)
constexpr int kMaxJacobi = 128;  // largest dense eigenproblem of small_eig: A fp64 + V fp64 up to 96 (147 KB of shared memory),
                                // A fp64 + V fp32 above (128: 196 KB)

// ------------------------------------------------------------------------------------ gram
template <int B>
__global__ void __launch_bounds__(256) gram_partial_kernel(const float* __restrict__ x, const float* __restrict__ y, int n, int ld,
                                                            float* __restrict__ part) {
  constexpr int R = B / 16;  // each thread owns an R x R block of G
  __shared__ float sx[32][B + 1];
  __shared__ float sy[32][B + 1];
  const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
  const int row0 = blockIdx.x * kGramRows;
  float acc[R][R];
#pragma unroll
  for (int a = 0; a < R; ++a)
#pragma unroll
    for (int c = 0; c < R; ++c) acc[a][c] = 0.f;
  for (int r0 = 0; r0 < kGramRows; r0 += 32) {
    for (int e = tid; e < 32 * B; e += 256) {
      const int r = e / B, c = e - r * B;
      const int gr = row0 + r0 + r;
      sx[r][c] = (gr < n) ? x[static_cast<size_t>(gr) * ld + c] : 0.f;
      sy[r][c] = (gr < n) ? y[static_cast<size_t>(gr) * ld + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      float xv[R], yv[R];
#pragma unroll
      for (int a = 0; a < R; ++a) xv[a] = sx[r][ti * R + a];
#pragma unroll
      for (int c = 0; c < R; ++c) yv[c] = sy[r][tj * R + c];
#pragma unroll
      for (int a = 0; a < R; ++a)
#pragma unroll
        for (int c = 0; c < R; ++c) acc[a][c] = fmaf(xv[a], yv[c], acc[a][c]);
    }
    __syncthreads();
  }
  float* out = part + static_cast<size_t>(blockIdx.x) * B * B;
#pragma unroll
  for (int a = 0; a < R; ++a)
#pragma unroll
    for (int c = 0; c < R; ++c) out[(ti * R + a) * B + tj * R + c] = acc[a][c];
}

__global__ void gram_combine_kernel(const float* __restrict__ part, int nparts, int bb, float* __restrict__ g) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= bb) return;
  double s = 0.0;
  for (int p = 0; p < nparts; ++p) s += static_cast<double>(part[static_cast<size_t>(p) * bb + e]);
  g[e] = static_cast<float>(s);
}

// ------------------------------------------------------------------------------------ small_eig
// mode 0: eigen-decomposition by parallel cyclic Jacobi (round-robin pairing) in fp64.
// mode 1: scaled Cholesky G = S^-1 R^T R S^-1, returns Q = S R^-1 so that (Y Q)^T (Y Q) = I.
// BT: the matrix order when it is known at compile time (32 / 64, the subspace blocks: every index division becomes a
// shift), 0 = run-time order.
template <typename VT, int BT>
__global__ void __launch_bounds__(512) small_eig_kernel(const float* __restrict__ g, int b_rt, float* __restrict__ evals,
                                                         float* __restrict__ evecs, int mode) {
  const int b = BT > 0 ? BT : b_rt;
  // rows are b + 1 apart: the column phase of a rotation walks DOWN two columns (lane = row), and with an odd pitch
  // those 8-byte accesses are conflict-free.  (With lane = pair the 32 scattered columns of one row cost 2-4 wavefronts
  // per access and the shared-memory pipe, not the fp64 units, bounded the round: ~3700 clk for 64 x 64.)
  const int pitch = b + 1;
  extern __shared__ double sd[];
  double* A = sd;                                   // [b][pitch]
  VT* V = reinterpret_cast<VT*>(sd + b * pitch);    // [b][pitch] accumulated rotations (fp32 only for b > 96)
  __shared__ double s_c[kMaxJacobi / 2], s_s[kMaxJacobi / 2];
  __shared__ int s_p[kMaxJacobi / 2], s_q[kMaxJacobi / 2];
  __shared__ double s_scale[kMaxJacobi];
  __shared__ double s_red[16];
  __shared__ int s_done;
  const int tid = threadIdx.x, nth = blockDim.x;
  for (int e = tid; e < b * b; e += nth) {
    const int i = e / b, j = e - i * b;
    // symmetrise the input (the Gram partials are symmetric only up to rounding)
    A[i * pitch + j] = 0.5 * (static_cast<double>(g[i * b + j]) + static_cast<double>(g[j * b + i]));
    V[i * pitch + j] = static_cast<VT>((i == j) ? 1.0 : 0.0);
  }
  __syncthreads();

  if (mode == 1) {
    if (tid < b) {
      const double dg = A[tid * pitch + tid];
      s_scale[tid] = dg > 0.0 ? rsqrt(dg) : 1.0;
    }
    __syncthreads();
    for (int e = tid; e < b * b; e += nth) {
      const int i = e / b, j = e - i * b;
      A[i * pitch + j] *= s_scale[i] * s_scale[j];
    }
    __syncthreads();
    // right-looking Cholesky, upper factor R stored in A's upper triangle
    for (int j = 0; j < b; ++j) {
      if (tid == 0) {
        const double piv = A[j * pitch + j];
        A[j * pitch + j] = sqrt(piv > 1e-14 ? piv : 1e-14);  // scaled diagonal is 1: 1e-14 ~ fully dependent column
      }
      __syncthreads();
      const double rjj = A[j * pitch + j];
      for (int i = j + 1 + tid; i < b; i += nth) A[j * pitch + i] /= rjj;
      __syncthreads();
      const int m = b - j - 1;
      for (int e = tid; e < m * m; e += nth) {
        const int i = j + 1 + e / m, k = j + 1 + e % m;
        if (k >= i) A[i * pitch + k] -= A[j * pitch + i] * A[j * pitch + k];
      }
      __syncthreads();
    }
    // V = R^-1 (upper triangular), one column per thread by back substitution: R V[:,c] = e_c
    if (tid < b) {
      const int c = tid;
      for (int i = b - 1; i >= 0; --i) {
        double s = (i == c) ? 1.0 : 0.0;
        if (i > c) {
          V[i * pitch + c] = static_cast<VT>(0.0);
          continue;
        }
        for (int k = i + 1; k <= c; ++k) s -= A[i * pitch + k] * static_cast<double>(V[k * pitch + c]);
        V[i * pitch + c] = static_cast<VT>(s / A[i * pitch + i]);
      }
    }
    __syncthreads();
    for (int e = tid; e < b * b; e += nth) {
      const int i = e / b, j = e - i * b;
      evecs[e] = static_cast<float>(s_scale[i] * static_cast<double>(V[i * pitch + j]));
    }
    if (tid < b && evals) evals[tid] = static_cast<float>(A[tid * pitch + tid]);
    return;
  }

  const int half = b / 2;
  for (int sweep = 0; sweep < 30; ++sweep) {
    // convergence: off-diagonal Frobenius norm against the diagonal
    double off = 0.0, dia = 0.0;
    for (int e = tid; e < b * b; e += nth) {
      const int i = e / b, j = e - i * b;
      const double v = A[i * pitch + j] * A[i * pitch + j];
      if (i == j) dia += v; else off += v;
    }
    off = warp_sum(off);
    dia = warp_sum(dia);
    if ((tid & 31) == 0) { s_red[tid >> 5] = off; }
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < (nth >> 5); ++w) t += s_red[w];
      s_red[0] = t;
    }
    __syncthreads();
    const double off_total = s_red[0];
    __syncthreads();
    if ((tid & 31) == 0) { s_red[tid >> 5] = dia; }
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < (nth >> 5); ++w) t += s_red[w];
      s_done = (off_total <= 1e-24 * t) ? 1 : 0;
    }
    __syncthreads();
    if (s_done) break;

    for (int round = 0; round < b - 1; ++round) {
      if (tid < half) {
        // round-robin tournament: slot 0 is fixed, the other b-1 slots rotate
        const int k = tid;
        auto player = [&](int slot) -> int { return slot == 0 ? 0 : 1 + (slot - 1 + round) % (b - 1); };
        int p = player(k), q = player(b - 1 - k);
        if (p > q) { const int t = p; p = q; q = t; }
        const double apq = A[p * pitch + q], app = A[p * pitch + p], aqq = A[q * pitch + q];
        double c = 1.0, s = 0.0;
        if (fabs(apq) > 1e-300 && fabs(apq) > 1e-18 * (fabs(app) + fabs(aqq))) {
          // t = sign(tau) / (|tau| + sqrt(1 + tau^2)), tau = (aqq - app) / (2 apq), written with one sqrt, one division
          // and one rsqrt: this warp's dependent fp64 chain is the serial part of every round
          const double da = aqq - app, b2 = 2.0 * apq;
          double t = fabs(b2) / (fabs(da) + sqrt(fma(da, da, b2 * b2)));
          if (da != 0.0 && ((da < 0.0) != (b2 < 0.0))) t = -t;
          c = rsqrt(fma(t, t, 1.0));
          s = t * c;
        }
        s_c[k] = c; s_s[k] = s; s_p[k] = p; s_q[k] = q;
      }
      __syncthreads();
      // columns: A <- A J, V <- V J
      for (int e = tid; e < b * half; e += nth) {
        const int k = e / b, r = e - k * b;  // a warp: 32 consecutive rows of one pair
        const int p = s_p[k], q = s_q[k];
        const double c = s_c[k], s = s_s[k];
        const double ap = A[r * pitch + p], aq = A[r * pitch + q];
        A[r * pitch + p] = c * ap - s * aq;
        A[r * pitch + q] = s * ap + c * aq;
        const double vp = static_cast<double>(V[r * pitch + p]), vq = static_cast<double>(V[r * pitch + q]);
        V[r * pitch + p] = static_cast<VT>(c * vp - s * vq);
        V[r * pitch + q] = static_cast<VT>(s * vp + c * vq);
      }
      __syncthreads();
      // rows: A <- J^T A  (a warp walks 32 consecutive columns of one pair: conflict-free shared-memory rows)
      for (int e = tid; e < b * half; e += nth) {
        const int k = e / b, col = e - k * b;
        const int p = s_p[k], q = s_q[k];
        const double c = s_c[k], s = s_s[k];
        const double ap = A[p * pitch + col], aq = A[q * pitch + col];
        A[p * pitch + col] = c * ap - s * aq;
        A[q * pitch + col] = s * ap + c * aq;
      }
      __syncthreads();
    }
  }
  // ascending order; column `rank` of evecs is the eigenvector of the rank-th smallest eigenvalue
  if (tid < b) {
    const double li = A[tid * pitch + tid];
    int rank = 0;
    for (int j = 0; j < b; ++j) {
      const double lj = A[j * pitch + j];
      if (lj < li || (lj == li && j < tid)) ++rank;
    }
    evals[rank] = static_cast<float>(li);
    reinterpret_cast<int*>(s_scale)[tid] = rank;
  }
  __syncthreads();
  const int* ranks = reinterpret_cast<const int*>(s_scale);
  for (int e = tid; e < b * b; e += nth) {
    const int i = e / b, j = e - i * b;
    evecs[i * b + ranks[j]] = static_cast<float>(V[i * pitch + j]);
  }
}

// ------------------------------------------------------------------------------------ right_mul
__device__ __forceinline__ void split3(float v, __nv_bfloat16& h, __nv_bfloat16& m, __nv_bfloat16& l) {
  h = __float2bfloat16_rn(v);
  float r = v - __bfloat162float(h);
  m = __float2bfloat16_rn(r);
  r -= __bfloat162float(m);
  l = __float2bfloat16_rn(r);
}

// y[n][b] = x[n][b] q[b][b]  (y may alias x).  vt (optional): bf16 [3*b][ldvt] = [hi | mid | lo] of y^T, part q of
// column j in row q * b + j (the W operand layout of the CHEB GEMM; for b = 32 the buffer has 32 more, zero, rows).
// q == nullptr: y = x (only the split is produced).
template <int B>
__global__ void __launch_bounds__(256) right_mul_kernel(const float* __restrict__ x, int n, int ld, const float* __restrict__ q,
                                                         float* __restrict__ y, __nv_bfloat16* __restrict__ vt, int ldvt) {
  __shared__ float sq[B][B + 1];
  __shared__ float sx[32][B + 1];
  __shared__ float so[32][B + 1];
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * 32;
  if (q) {
    for (int e = tid; e < B * B; e += 256) sq[e / B][e % B] = q[e];
  }
  for (int e = tid; e < 32 * B; e += 256) {
    const int r = e / B, c = e - r * B;
    const int gr = row0 + r;
    sx[r][c] = (gr < n) ? x[static_cast<size_t>(gr) * ld + c] : 0.f;
  }
  __syncthreads();
  constexpr int CPT = B / 8;  // columns per thread
  const int r = tid >> 3, c0 = (tid & 7) * CPT;
  float acc[CPT];
  if (q) {
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[c] = 0.f;
    for (int k = 0; k < B; ++k) {
      const float xv = sx[r][k];
#pragma unroll
      for (int c = 0; c < CPT; ++c) acc[c] = fmaf(xv, sq[k][c0 + c], acc[c]);
    }
  } else {
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[c] = sx[r][c0 + c];
  }
#pragma unroll
  for (int c = 0; c < CPT; ++c) so[r][c0 + c] = acc[c];
  if (y && row0 + r < n) {
#pragma unroll
    for (int c = 0; c < CPT; ++c) y[static_cast<size_t>(row0 + r) * ld + c0 + c] = acc[c];
  }
  if (vt) {
    __syncthreads();
    // transposed write: thread -> (column, row) with rows fastest
    for (int e = tid; e < 32 * B; e += 256) {
      const int col = e >> 5, rr = e & 31;
      if (row0 + rr < n) {
        __nv_bfloat16 h, m, l;
        split3(so[rr][col], h, m, l);
        const size_t vrow = static_cast<size_t>(col);  // part q of vector j in row q * B + j: [hi | mid | lo]
        vt[vrow * ldvt + row0 + rr] = h;
        vt[(vrow + B) * ldvt + row0 + rr] = m;
        vt[(vrow + 2 * B) * ldvt + row0 + rr] = l;
      }
    }
  }
}

// ------------------------------------------------------------------------------------ residuals
// out[j] = sum_r (w[r][j] - theta[j] x[r][j])^2: per-CTA partials over 256-row slabs, then a fixed-order combine (the
// value steers the convergence test of the subspace iteration, so it must not depend on the order of atomics: a flipped
// iteration count would make two runs on the same input differ in the last bits of the eigenvectors).
__global__ void __launch_bounds__(256) resid_partial_kernel(const float* __restrict__ w, const float* __restrict__ x,
                                                             const float* __restrict__ theta, int n, int b, int ld, float* __restrict__ part) {
  __shared__ float s_acc[4][kMaxB];
  const int col = threadIdx.x & 63, rg = threadIdx.x >> 6;
  const int r0 = blockIdx.x * kGramRows;
  float acc = 0.f;
  if (col < b) {
    const float th = theta[col];
    const int r1 = min(n, r0 + kGramRows);
    for (int r = r0 + rg; r < r1; r += 4) {
      const float d = w[static_cast<size_t>(r) * ld + col] - th * x[static_cast<size_t>(r) * ld + col];
      acc = fmaf(d, d, acc);
    }
  }
  s_acc[rg][col] = acc;
  __syncthreads();
  if (rg == 0) part[static_cast<size_t>(blockIdx.x) * kMaxB + col] = (s_acc[0][col] + s_acc[1][col]) + (s_acc[2][col] + s_acc[3][col]);
}

__global__ void resid_combine_kernel(const float* __restrict__ part, int nparts, int b, float* __restrict__ out) {
  const int col = threadIdx.x;
  if (col >= b) return;
  float t = 0.f;
  for (int p = 0; p < nparts; ++p) t += part[static_cast<size_t>(p) * kMaxB + col];
  out[col] = t;
}

}  // namespace b200d

using namespace b200d;

extern "C" size_t b200d_gram_workspace_bytes(int32_t n, int32_t b) {
  if (n <= 0 || b <= 0) return 0;
  return static_cast<size_t>((n + kGramRows - 1) / kGramRows) * b * b * sizeof(float);
}

extern "C" int b200d_gram(const float* x, const float* y, int32_t n, int32_t b, int32_t ld, float* g, void* ws, size_t ws_bytes,
                          void* stream) {
  B200D_CHECK_ARG(x && y && g && ws && n > 0 && (b == 32 || b == 64) && ld >= b);
  if (ws_bytes < b200d_gram_workspace_bytes(n, b)) return b200d::set_error(B200D_EWORKSPACE, "%s: workspace too small%s", "b200d_gram");
  const int parts = (n + kGramRows - 1) / kGramRows;
  float* part = reinterpret_cast<float*>(ws);
  cudaStream_t s = as_stream(stream);
  if (b == 32) gram_partial_kernel<32><<<parts, 256, 0, s>>>(x, y, n, ld, part);
  else gram_partial_kernel<64><<<parts, 256, 0, s>>>(x, y, n, ld, part);
  gram_combine_kernel<<<(b * b + 255) / 256, 256, 0, s>>>(part, parts, b * b, g);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_small_eig(const float* g, int32_t b, float* evals, float* evecs, int32_t mode, void* stream) {
  B200D_CHECK_ARG(g && evecs && b >= 2 && b <= kMaxJacobi && b % 2 == 0 && (mode == 0 || mode == 1) && (mode == 1 || evals));
  static bool attr_set_dev[kMaxDevices] = {};  // kernel attributes are per device
  const int attr_dev = current_device();
  const bool attr_known = attr_dev >= 0 && attr_dev < kMaxDevices;
  if (!attr_known || !attr_set_dev[attr_dev]) {
    const int big = 2 * 96 * 97 * sizeof(double);
    B200D_CHECK_CUDA(cudaFuncSetAttribute(small_eig_kernel<double, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    B200D_CHECK_CUDA(cudaFuncSetAttribute(small_eig_kernel<double, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    B200D_CHECK_CUDA(cudaFuncSetAttribute(small_eig_kernel<float, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          kMaxJacobi * (kMaxJacobi + 1) * (sizeof(double) + sizeof(float))));
    if (attr_known) attr_set_dev[attr_dev] = true;
  }
  cudaStream_t st = as_stream(stream);
  const size_t smem64 = static_cast<size_t>(2) * b * (b + 1) * sizeof(double);
  if (b == 64) small_eig_kernel<double, 64><<<1, 512, smem64, st>>>(g, b, evals, evecs, mode);
  else if (b == 32) small_eig_kernel<double, 32><<<1, 512, smem64, st>>>(g, b, evals, evecs, mode);
  else if (b <= 96) small_eig_kernel<double, 0><<<1, 512, smem64, st>>>(g, b, evals, evecs, mode);
  else small_eig_kernel<float, 0><<<1, 512, static_cast<size_t>(b) * (b + 1) * (sizeof(double) + sizeof(float)), st>>>(g, b, evals, evecs, mode);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_right_mul(const float* x, int32_t n, int32_t b, int32_t ld, const float* q, float* y, void* vt_bf16, int32_t ldvt,
                               void* stream) {
  B200D_CHECK_ARG(x && n > 0 && (b == 32 || b == 64) && ld >= b && (y || vt_bf16));
  B200D_CHECK_ARG(vt_bf16 == nullptr || ldvt >= n);
  cudaStream_t s = as_stream(stream);
  const int grid = (n + 31) / 32;
  __nv_bfloat16* vt = reinterpret_cast<__nv_bfloat16*>(vt_bf16);
  if (b == 32) right_mul_kernel<32><<<grid, 256, 0, s>>>(x, n, ld, q, y, vt, ldvt);
  else right_mul_kernel<64><<<grid, 256, 0, s>>>(x, n, ld, q, y, vt, ldvt);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_resid_norms(const float* w, const float* x, const float* theta, int32_t n, int32_t b, int32_t ld, float* out,
                                 void* ws, size_t ws_bytes, void* stream) {
  B200D_CHECK_ARG(w && x && theta && out && ws && n > 0 && b > 0 && b <= kMaxB && ld >= b);
  const int parts = (n + kGramRows - 1) / kGramRows;
  if (ws_bytes < static_cast<size_t>(parts) * kMaxB * sizeof(float))
    return b200d::set_error(B200D_EWORKSPACE, "%s: workspace too small%s", "b200d_resid_norms");
  float* part = reinterpret_cast<float*>(ws);
  resid_partial_kernel<<<parts, 256, 0, as_stream(stream)>>>(w, x, theta, n, b, ld, part);
  resid_combine_kernel<<<1, 64, 0, as_stream(stream)>>>(part, parts, b, out);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}
