// Sparse form of the binarised p-neighbour graph and the Chebyshev step on it.
//
// getAffinityGraphMat keeps p neighbours per row, so A = 0.5 (B + B^T) has at most 2 p non-zeros per row; for the
// long-form chunks (n = 10 000, p ~ 10) that is > 99.7 % zeros, and the dense tcgen05 product (L2-bandwidth bound on
// the 200 MB operand) is replaced by a row-gather over the CSR lists: fp32 throughout, fixed (ascending column)
// summation order, ~n * 2p * b * 4 bytes of L2 reads per product.
#include "common.cuh"

#include <cuda_bf16.h>

namespace b200d {

constexpr unsigned kOneBit = 0x80000000u;  // entry = column | kOneBit when a[i][column] == 1 (else 0.5)

__device__ __forceinline__ int nonzero_halves(const uint4& v) {
  int c = 0;
  c += (v.x & 0xffffu) != 0; c += (v.x >> 16) != 0;
  c += (v.y & 0xffffu) != 0; c += (v.y >> 16) != 0;
  c += (v.z & 0xffffu) != 0; c += (v.z >> 16) != 0;
  c += (v.w & 0xffffu) != 0; c += (v.w >> 16) != 0;
  return c;
}

// one warp per row: cnt[row] = non-zeros of a[row][0 .. lda)   (columns >= n hold zeros)
__global__ void __launch_bounds__(256) csr_count_kernel(const __nv_bfloat16* __restrict__ a, int n, int lda, int* __restrict__ cnt) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const uint4* r = reinterpret_cast<const uint4*>(a + static_cast<size_t>(row) * lda);
  const int nv = lda >> 3;
  int c = 0;
  for (int v = lane; v < nv; v += 32) c += nonzero_halves(__ldg(r + v));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) cnt[row] = c;
}

// in-place exclusive scan of rowptr[0 .. n) by one block; rowptr[n] = total
__global__ void __launch_bounds__(1024) csr_scan_kernel(int* __restrict__ rowptr, int n) {
  __shared__ int s_warp[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (n + 1023) / 1024;
  const int lo = min(tid * per, n), hi = min(lo + per, n);
  int local = 0;
  for (int i = lo; i < hi; ++i) local += rowptr[i];
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    s_warp[lane] = w;
  }
  __syncthreads();
  int run = incl - local + (warp > 0 ? s_warp[warp - 1] : 0);
  for (int i = lo; i < hi; ++i) {
    const int c = rowptr[i];
    rowptr[i] = run;
    run += c;
  }
  if (tid == 1023) rowptr[n] = s_warp[31];
}

// one warp per row: entries in ascending column order
__global__ void __launch_bounds__(256) csr_fill_kernel(const __nv_bfloat16* __restrict__ a, int n, int lda, const int* __restrict__ rowptr,
                                                       long long capacity, unsigned* __restrict__ colw) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const uint4* r = reinterpret_cast<const uint4*>(a + static_cast<size_t>(row) * lda);
  const int nv = lda >> 3;
  long long base = rowptr[row];
  for (int v0 = 0; v0 < nv; v0 += 32) {
    const int v = v0 + lane;
    uint4 q = make_uint4(0, 0, 0, 0);
    if (v < nv) q = __ldg(r + v);
    const int c = nonzero_halves(q);
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (c) {
      long long pos = base + incl - c;
      if (pos + c > capacity) __trap();  // capacity >= 2 n p is a bound on the non-zeros: cannot happen
      const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const unsigned h = (w[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
        if (h) colw[pos++] = static_cast<unsigned>(v * 8 + e) | (h == 0x3f80u ? kOneBit : 0u);
      }
    }
    base += total;
  }
}

// out = ca * (deg .* x - A x) + cb * x + cc * xprev, one warp per row, B columns (B / 32 per lane)
template <int B>
__global__ void __launch_bounds__(256) spmm_cheb_kernel(const int* __restrict__ rowptr, const unsigned* __restrict__ colw, int n,
                                                        const float* __restrict__ deg, const float* __restrict__ x,
                                                        const float* __restrict__ xprev, int ldx, float ca, float cb, float cc,
                                                        float* __restrict__ out, int ldo) {
  constexpr int V = B / 32;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const int beg = rowptr[row], end = rowptr[row + 1];
  float acc[V];
#pragma unroll
  for (int c = 0; c < V; ++c) acc[c] = 0.f;
  for (int e0 = beg; e0 < end; e0 += 32) {
    const int m = min(32, end - e0);
    const unsigned mine = (lane < m) ? __ldg(colw + e0 + lane) : 0u;
#pragma unroll 4
    for (int j = 0; j < m; ++j) {
      const unsigned ent = __shfl_sync(0xffffffffu, mine, j);
      const float w = (ent & kOneBit) ? 1.0f : 0.5f;
      const float* xr = x + static_cast<size_t>(ent & ~kOneBit) * ldx + lane * V;
      if constexpr (V == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(xr));
        acc[0] = fmaf(w, t.x, acc[0]);
        acc[1] = fmaf(w, t.y, acc[1]);
      } else {
        acc[0] = fmaf(w, __ldg(xr), acc[0]);
      }
    }
  }
  const float dg = __ldg(deg + row);
  const float* xs = x + static_cast<size_t>(row) * ldx + lane * V;
  const float* xp = xprev ? xprev + static_cast<size_t>(row) * ldx + lane * V : nullptr;
  float* o = out + static_cast<size_t>(row) * ldo + lane * V;
#pragma unroll
  for (int c = 0; c < V; ++c) {
    const float xv = xs[c];
    float y = ca * (dg * xv - acc[c]) + cb * xv;
    if (xp) y += cc * xp[c];
    o[c] = y;
  }
}

}  // namespace b200d

using namespace b200d;

extern "C" int b200d_csr_from_dense(const void* a_bf16, int32_t n, int32_t lda, int32_t* rowptr, uint32_t* colw, int64_t capacity,
                                    void* stream) {
  B200D_CHECK_ARG(a_bf16 && rowptr && colw && n > 0 && lda >= n && lda % 8 == 0 && capacity > 0);
  B200D_CHECK_ARG((reinterpret_cast<uintptr_t>(a_bf16) & 15) == 0);
  cudaStream_t s = as_stream(stream);
  const __nv_bfloat16* a = reinterpret_cast<const __nv_bfloat16*>(a_bf16);
  const int grid = (n + 7) / 8;
  csr_count_kernel<<<grid, 256, 0, s>>>(a, n, lda, rowptr);
  csr_scan_kernel<<<1, 1024, 0, s>>>(rowptr, n);
  csr_fill_kernel<<<grid, 256, 0, s>>>(a, n, lda, rowptr, capacity, colw);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_spmm_cheb(const int32_t* rowptr, const uint32_t* colw, int32_t n, int32_t b, const float* deg, const float* x,
                               const float* xprev, int32_t ldx, float ca, float cb, float cc, float* out, int32_t ldo, void* stream) {
  B200D_CHECK_ARG(rowptr && colw && deg && x && out && n > 0 && (b == 32 || b == 64));
  B200D_CHECK_ARG(ldx >= b && ldo >= b && ldx % 2 == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0);
  B200D_CHECK_ARG(out != x && out != xprev);
  cudaStream_t s = as_stream(stream);
  const int grid = (n + 7) / 8;
  if (b == 64) spmm_cheb_kernel<64><<<grid, 256, 0, s>>>(rowptr, colw, n, deg, x, xprev, ldx, ca, cb, cc, out, ldo);
  else spmm_cheb_kernel<32><<<grid, 256, 0, s>>>(rowptr, colw, n, deg, x, xprev, ldx, ca, cb, cc, out, ldo);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}
