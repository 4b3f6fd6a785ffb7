// p-neighbour graph construction for NME-SC (upstream offline_clustering.getKneighborsConnections,
// getAffinityGraphMat, getLaplacian, isGraphFullyConnected / getTheLargestComponent).
//
// Upstream sorts every row with argsort(descending) for every p of the sweep.  Here the sweep
// matrix (<= 1024 x 1024 strided subsample) is ranked ONCE (bitonic sort per row in shared
// memory -> rank[i][j] and its transpose), after which every p-neighbour graph, its Laplacian
// and its connectivity are element-wise functions of the two rank matrices.  The final full
// N x N binarisation uses a 4-pass radix select per row (threshold = p-th largest, ties by
// lower column index) instead of a full sort.
#include "common.cuh"

#include <cuda_bf16.h>

namespace b200d {

__device__ __forceinline__ unsigned enc_desc(float f) {  // larger float -> larger unsigned
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// ------------------------------------------------------------------------------------ row rank
__global__ void __launch_bounds__(1024) row_rank_kernel(const float* __restrict__ mat, long long ld, int stride, int n,
                                                         uint16_t* __restrict__ rank, uint16_t* __restrict__ rankT) {
  __shared__ unsigned long long key[1024];
  const int i = blockIdx.x, tid = threadIdx.x;
  unsigned long long k = ~0ull;
  if (tid < n) {
    const float v = mat[static_cast<size_t>(i) * stride * ld + static_cast<size_t>(tid) * stride];
    k = (static_cast<unsigned long long>(~enc_desc(v)) << 32) | static_cast<unsigned>(tid);
  }
  key[tid] = k;
  __syncthreads();
  for (int kk = 2; kk <= 1024; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      const int ixj = tid ^ j;
      if (ixj > tid) {
        const bool asc = (tid & kk) == 0;
        const unsigned long long a = key[tid], b = key[ixj];
        if ((a > b) == asc) {
          key[tid] = b;
          key[ixj] = a;
        }
      }
      __syncthreads();
    }
  }
  if (tid < n) {
    const int col = static_cast<int>(key[tid] & 0xFFFFFFFFull);
    rank[static_cast<size_t>(i) * n + col] = static_cast<uint16_t>(tid);
    rankT[static_cast<size_t>(col) * n + i] = static_cast<uint16_t>(tid);
  }
}

// ------------------------------------------------------------------------------------ Laplacians of the sweep
struct PList {
  int np;
  int p[64];
};

__global__ void __launch_bounds__(256) laplacian_from_rank_kernel(const uint16_t* __restrict__ rank, const uint16_t* __restrict__ rankT,
                                                                   int n, const PList pl, float* __restrict__ lap) {
  __shared__ float s_part[8];
  const int i = blockIdx.x, b = blockIdx.y;
  const int p = pl.p[b];
  float* out = lap + (static_cast<size_t>(b) * n + i) * n;
  const uint16_t* r0 = rank + static_cast<size_t>(i) * n;
  const uint16_t* r1 = rankT + static_cast<size_t>(i) * n;
  float deg = 0.f;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    float a = 0.f;
    if (j != i) a = 0.5f * (static_cast<float>(r0[j] < p) + static_cast<float>(r1[j] < p));
    out[j] = -a;
    deg += a;
  }
  deg = warp_sum(deg);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = deg;
  __syncthreads();
  if (threadIdx.x == 0) {
    float d = 0.f;
    for (int w = 0; w < 8; ++w) d += s_part[w];
    out[i] = __half2float(__float2half_rn(d));  // upstream builds the graph in fp16: degree is fp16-rounded
  }
}

// ------------------------------------------------------------------------------------ connectivity on the rank form
// One CTA per p of the list: nodes reachable from node 0 in the p-neighbour graph (frontier expansion).
__global__ void __launch_bounds__(1024) graph_reach_rank_kernel(const uint16_t* __restrict__ rank, const uint16_t* __restrict__ rankT,
                                                                 int n, const PList pl, int* __restrict__ reach) {
  __shared__ unsigned char visited[1024], frontier[1024], nxt[1024];
  __shared__ int changed, count;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int p = pl.p[blockIdx.x];
  visited[tid] = (tid == 0);
  frontier[tid] = (tid == 0);
  nxt[tid] = 0;
  if (tid == 0) { changed = 1; count = 0; }
  __syncthreads();
  while (true) {
    if (changed == 0) break;
    __syncthreads();
    if (tid == 0) changed = 0;
    __syncthreads();
    for (int i = warp; i < n; i += 32) {
      if (!frontier[i]) continue;
      const uint16_t* r0 = rank + static_cast<size_t>(i) * n;
      const uint16_t* r1 = rankT + static_cast<size_t>(i) * n;
      for (int j = lane; j < n; j += 32) {
        if (j != i && !visited[j] && (r0[j] < p || r1[j] < p)) {
          nxt[j] = 1;
          changed = 1;
        }
      }
    }
    __syncthreads();
    frontier[tid] = nxt[tid] && !visited[tid];
    visited[tid] = visited[tid] || nxt[tid];
    nxt[tid] = 0;
    __syncthreads();
  }
  if (tid < n && visited[tid]) atomicAdd(&count, 1);
  __syncthreads();
  if (tid == 0) reach[blockIdx.x] = count;
}

// ------------------------------------------------------------------------------------ full-matrix top-p select
// thr_out / cut_out (both or neither): the threshold code of the row and the column of the last selected entry equal to it --
// all a row-sharded caller needs to evaluate [i in top_p(row j)] for rows j it does not hold (sym_combine_rows_kernel).
__global__ void __launch_bounds__(256) topp_select_kernel(const float* __restrict__ mat, int n, int p, unsigned char* __restrict__ sel,
                                                          unsigned* __restrict__ thr_out, int* __restrict__ cut_out) {
  __shared__ int hist[256];
  __shared__ unsigned s_prefix, s_mask;
  __shared__ int s_remaining;
  __shared__ int s_warp_cnt[8];
  __shared__ int s_base;
  const int i = blockIdx.x, tid = threadIdx.x;
  const float* row = mat + static_cast<size_t>(i) * n;
  if (tid == 0) { s_prefix = 0; s_mask = 0; s_remaining = p; }
  __syncthreads();
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[tid] = 0;
    __syncthreads();
    const unsigned prefix = s_prefix, mask = s_mask;
    for (int j = tid; j < n; j += 256) {
      const unsigned e = enc_desc(__ldg(row + j));
      if ((e & mask) == prefix) atomicAdd(&hist[(e >> shift) & 255], 1);
    }
    __syncthreads();
    if (tid == 0) {
      int rem = s_remaining, bkt = 255;
      for (; bkt > 0; --bkt) {
        if (hist[bkt] >= rem) break;
        rem -= hist[bkt];
      }
      s_remaining = rem;
      s_prefix = prefix | (static_cast<unsigned>(bkt) << shift);
      s_mask = mask | (255u << shift);
    }
    __syncthreads();
  }
  const unsigned thr = s_prefix;
  const int need_eq = s_remaining;  // how many elements equal to the threshold are selected (lowest columns first)
  if (tid == 0) s_base = 0;
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int j0 = 0; j0 < n; j0 += 256) {
    const int j = j0 + tid;
    unsigned e = 0;
    bool eq = false;
    if (j < n) {
      e = enc_desc(__ldg(row + j));
      eq = (e == thr);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) s_warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_warp_cnt[w];
    before += __popc(bal & ((1u << lane) - 1u));
    if (j < n) sel[static_cast<size_t>(i) * n + j] = (e > thr || (eq && before < need_eq)) ? 1 : 0;
    if (cut_out != nullptr && eq && before == need_eq - 1) {
      cut_out[i] = j;
      thr_out[i] = thr;
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += s_warp_cnt[w];
      s_base += tot;
    }
    __syncthreads();
  }
}

// a[i][j] = (i == j) ? 0 : 0.5 (sel[i][j] + sel[j][i]) as bf16;  deg[i] = row sum
__global__ void __launch_bounds__(1024) sym_combine_kernel(const unsigned char* __restrict__ sel, int n, __nv_bfloat16* __restrict__ a,
                                                           int lda, float* __restrict__ deg) {
  __shared__ unsigned char t[32][33];
  const int bi = blockIdx.y * 32, bj = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  // transposed tile: t[ty][tx] = sel[bj + ty][bi + tx]
  {
    const int r = bj + ty, c = bi + tx;
    t[ty][tx] = (r < n && c < n) ? sel[static_cast<size_t>(r) * n + c] : 0;
  }
  __syncthreads();
  const int i = bi + ty, j = bj + tx;
  float v = 0.f;
  if (i < n && j < n && i != j) v = 0.5f * (static_cast<float>(sel[static_cast<size_t>(i) * n + j]) + static_cast<float>(t[tx][ty]));
  if (i < n && j < lda) a[static_cast<size_t>(i) * lda + j] = __float2bfloat16_rn(j < n ? v : 0.f);
  const float s = warp_sum(v);
  if (tx == 0 && i < n && s != 0.f) atomicAdd(deg + i, s);  // multiples of 0.5: exact, order independent
}

// Row shard of sym_combine: local row r is global row i = row_lo + r.  The transposed term [i in top_p(row j)] comes from the
// symmetric entry mat[i][j] (== mat[j][i]) and row j's threshold / tie cut-off instead of from sel[j][i].
__global__ void __launch_bounds__(256) sym_combine_rows_kernel(const float* __restrict__ mat_rows, const unsigned char* __restrict__ sel,
                                                               const unsigned* __restrict__ thr_all, const int* __restrict__ cut_all,
                                                               int row_lo, int n, __nv_bfloat16* __restrict__ a, int lda,
                                                               float* __restrict__ deg) {
  __shared__ float s_part[8];
  const int r = blockIdx.x, i = row_lo + r;
  const float* row = mat_rows + static_cast<size_t>(r) * n;
  const unsigned char* srow = sel + static_cast<size_t>(r) * n;
  float acc = 0.f;
  for (int j = threadIdx.x; j < lda; j += 256) {
    float v = 0.f;
    if (j < n && j != i) {
      const unsigned e = enc_desc(__ldg(row + j));
      const unsigned tj = __ldg(thr_all + j);
      const bool in_j = e > tj || (e == tj && i <= __ldg(cut_all + j));
      v = 0.5f * (static_cast<float>(srow[j]) + (in_j ? 1.f : 0.f));
    }
    a[static_cast<size_t>(r) * lda + j] = __float2bfloat16_rn(v);
    acc += v;
  }
  acc = warp_sum(acc);  // multiples of 0.5 below 2^24: exact in any order
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float d = 0.f;
    for (int w = 0; w < 8; ++w) d += s_part[w];
    deg[r] = __half2float(__float2half_rn(d));
  }
}

__global__ void deg_round_kernel(float* deg, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) deg[i] = __half2float(__float2half_rn(deg[i]));
}

}  // namespace b200d

using namespace b200d;

extern "C" int b200d_row_rank(const float* mat, int64_t ld, int32_t stride, int32_t n, void* rank_u16, void* rankT_u16, void* stream) {
  B200D_CHECK_ARG(mat && rank_u16 && rankT_u16 && n > 0 && n <= 1024 && stride >= 1 && ld > 0);
  row_rank_kernel<<<n, 1024, 0, as_stream(stream)>>>(mat, ld, stride, n, reinterpret_cast<uint16_t*>(rank_u16),
                                                     reinterpret_cast<uint16_t*>(rankT_u16));
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_laplacian_from_rank(const void* rank_u16, const void* rankT_u16, int32_t n, const int32_t* p_list_host, int32_t np,
                                         float* lap, void* stream) {
  B200D_CHECK_ARG(rank_u16 && rankT_u16 && p_list_host && lap && n > 0 && n <= 1024 && np > 0 && np <= 64);
  PList pl;
  pl.np = np;
  for (int b = 0; b < np; ++b) pl.p[b] = p_list_host[b];
  dim3 grid(n, np);
  laplacian_from_rank_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint16_t*>(rank_u16),
                                                                 reinterpret_cast<const uint16_t*>(rankT_u16), n, pl, lap);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_graph_reach_rank(const void* rank_u16, const void* rankT_u16, int32_t n, const int32_t* p_list_host, int32_t np,
                                      int32_t* reach, void* stream) {
  B200D_CHECK_ARG(rank_u16 && rankT_u16 && reach && p_list_host && n > 0 && n <= 1024 && np > 0 && np <= 64);
  PList pl;
  pl.np = np;
  for (int b = 0; b < np; ++b) pl.p[b] = p_list_host[b];
  graph_reach_rank_kernel<<<np, 1024, 0, as_stream(stream)>>>(reinterpret_cast<const uint16_t*>(rank_u16),
                                                              reinterpret_cast<const uint16_t*>(rankT_u16), n, pl, reach);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_topp_binarize(const float* mat, int32_t n, int32_t p, void* a_bf16, int32_t lda, float* deg, void* sel_u8,
                                   void* stream) {
  B200D_CHECK_ARG(mat && a_bf16 && deg && sel_u8 && n > 0 && p > 0 && p <= n && lda >= n && lda % 8 == 0);
  cudaStream_t s = as_stream(stream);
  B200D_CHECK_CUDA(cudaMemsetAsync(deg, 0, sizeof(float) * n, s));
  topp_select_kernel<<<n, 256, 0, s>>>(mat, n, p, reinterpret_cast<unsigned char*>(sel_u8), nullptr, nullptr);
  dim3 grid((lda + 31) / 32, (n + 31) / 32);
  sym_combine_kernel<<<grid, 1024, 0, s>>>(reinterpret_cast<const unsigned char*>(sel_u8), n, reinterpret_cast<__nv_bfloat16*>(a_bf16), lda, deg);
  deg_round_kernel<<<(n + 255) / 256, 256, 0, s>>>(deg, n);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_topp_select_rows(const float* mat_rows, int32_t m, int32_t n, int32_t p, void* sel_u8, uint32_t* thr, int32_t* cut,
                                      void* stream) {
  B200D_CHECK_ARG(mat_rows && sel_u8 && thr && cut && m > 0 && n > 0 && p > 0 && p <= n);
  topp_select_kernel<<<m, 256, 0, as_stream(stream)>>>(mat_rows, n, p, reinterpret_cast<unsigned char*>(sel_u8), thr, cut);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_sym_combine_rows(const float* mat_rows, const void* sel_u8, const uint32_t* thr_all, const int32_t* cut_all,
                                      int32_t row_lo, int32_t m, int32_t n, void* a_rows_bf16, int32_t lda, float* deg_rows, void* stream) {
  B200D_CHECK_ARG(mat_rows && sel_u8 && thr_all && cut_all && a_rows_bf16 && deg_rows && m > 0 && n > 0 && row_lo >= 0 && row_lo + m <= n);
  B200D_CHECK_ARG(lda >= n && lda % 8 == 0);
  sym_combine_rows_kernel<<<m, 256, 0, as_stream(stream)>>>(mat_rows, reinterpret_cast<const unsigned char*>(sel_u8), thr_all, cut_all, row_lo, n,
                                                            reinterpret_cast<__nv_bfloat16*>(a_rows_bf16), lda, deg_rows);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}
