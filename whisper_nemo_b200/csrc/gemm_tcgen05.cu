// Pointwise-conv / projection GEMM on Blackwell 5th-gen tensor cores.
//
//   acc[m][n] = sum_k A[m][k] * W[n][k]       A [M][lda], W [N][ldw], 16-bit, both K-major
//
// One persistent CTA per SM, warp-specialised:
//   warp 0   TMA producer  (cp.async.bulk.tensor.2d, 128B swizzle, 64-wide K slabs)
//   warp 1   MMA issuer    (tcgen05.mma cta_group::1 kind::f16, M=128, N=BLOCK_N, K=16 per instruction)
//   warp 2   TMEM allocator (2 accumulator buffers -> epilogue of tile i overlaps MMA of tile i+1)
//   warps 4-7 epilogue     (tcgen05.ld 32x32b, fused epilogue in fp32, 16-byte global stores)
// smem ring: STAGES x (A 128x64 + W BLOCK_Nx64); mbarriers full/empty per stage, tmem_full/empty
// per accumulator.  Replaces the cuDNN/cuBLAS 1x1 conv1d calls under JasperBlock / SpeakerDecoder
// (nemo/collections/asr/parts/submodules/jasper.py, modules/conv_asr.py) and the A*V products of
// the spectral solver (offline_clustering.SpectralClustering.getSpectralEmbeddings).
#include "common.cuh"

#include <cuda.h>
#include <cuda_bf16.h>
#include <atomic>
#include <cstdlib>
#include <mutex>

namespace b200d {

// ------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major operand, 128-byte swizzle, 64-element (128 B) rows:
// start>>4 [0,14) | LBO>>4 [16,30) (ignored for swizzled K-major) | SBO>>4 = 1024>>4 [32,46) |
// version 1 [46,48) | layout SWIZZLE_128B = 2 [61,64)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// ------------------------------------------------------------------------------------ kernel
struct GemmParams {
  int M, N, K;
  void* out;
  int ldo;
  int kseg;  // EPI_PARTIAL: K blocks per segment (a tile of the launch is one (row tile, K segment)); else 0
  b200d_gemm_epilogue epi;
};

// Internal epilogue of the split-K Chebyshev product: raw fp32 accumulators of one K segment, out32[seg][M][N].
constexpr int EPI_PARTIAL = 100;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : (BLOCK_N == 192) ? 5 : 6;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int TMEM_COLS = (BLOCK_N > 128) ? 512 : 256;  // two accumulators, allocation a power of two
  static constexpr int ACC_STRIDE = (BLOCK_N == 96) ? 128 : BLOCK_N;  // TMEM columns between the two accumulator buffers
  static constexpr int SMEM_BYTES = 1024 + STAGES * (A_BYTES + B_BYTES) + 256;
};

__device__ __forceinline__ uint4 pack8_f16(const float* v) {
  __half2 h0 = __floats2half2_rn(v[0], v[1]);
  __half2 h1 = __floats2half2_rn(v[2], v[3]);
  __half2 h2 = __floats2half2_rn(v[4], v[5]);
  __half2 h3 = __floats2half2_rn(v[6], v[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  u.z = *reinterpret_cast<uint32_t*>(&h2);
  u.w = *reinterpret_cast<uint32_t*>(&h3);
  return u;
}
__device__ __forceinline__ void unpack8_f16(uint4 u, float* v) {
  __half2 h0 = *reinterpret_cast<__half2*>(&u.x), h1 = *reinterpret_cast<__half2*>(&u.y);
  __half2 h2 = *reinterpret_cast<__half2*>(&u.z), h3 = *reinterpret_cast<__half2*>(&u.w);
  float2 f;
  f = __half22float2(h0); v[0] = f.x; v[1] = f.y;
  f = __half22float2(h1); v[2] = f.x; v[3] = f.y;
  f = __half22float2(h2); v[4] = f.x; v[5] = f.y;
  f = __half22float2(h3); v[6] = f.x; v[7] = f.y;
}

// Epilogue on 32 consecutive columns [n0, n0+32) of one row.
template <int MODE, bool STORE = true>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, int row, int n0, float* acc) {
  const b200d_gemm_epilogue& e = p.epi;
  if constexpr (MODE == B200D_EPI_BIAS || MODE == B200D_EPI_BIAS_RELU || MODE == B200D_EPI_BIAS_F32) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + n0 + j));
      acc[j] += b.x; acc[j + 1] += b.y; acc[j + 2] += b.z; acc[j + 3] += b.w;
    }
    if constexpr (MODE == B200D_EPI_BIAS_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = fmaxf(acc[j], 0.f);
    }
  } else if constexpr (MODE == B200D_EPI_SE_RES) {
    const int seg = row / e.rows_per_seg;
    const float* gate = e.rowvec + static_cast<size_t>(seg) * p.N + n0;
    const __half* aux = reinterpret_cast<const __half*>(e.aux16) + static_cast<size_t>(row) * p.ldo + n0;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      float x[8];
      unpack8_f16(__ldg(reinterpret_cast<const uint4*>(aux + j)), x);
      float4 g0 = __ldg(reinterpret_cast<const float4*>(gate + j)), g1 = __ldg(reinterpret_cast<const float4*>(gate + j + 4));
      float4 b0 = __ldg(reinterpret_cast<const float4*>(e.bias + n0 + j)), b1 = __ldg(reinterpret_cast<const float4*>(e.bias + n0 + j + 4));
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[j + q] = fmaxf(fmaf(x[q], g[q], acc[j + q] + b[q]), 0.f);
    }
  } else if constexpr (MODE == B200D_EPI_TDNN) {
    const int seg = row / e.rows_per_seg;
    const float* rb = e.rowvec + static_cast<size_t>(seg) * p.N + n0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float z = fmaxf(acc[j] + __ldg(rb + j), 0.f);
      acc[j] = tanhf(fmaf(__ldg(e.scale + n0 + j), z, __ldg(e.shift + n0 + j)));
    }
  }
  if constexpr (MODE == B200D_EPI_SIGMOID_F32) {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 1.f / (1.f + __expf(-acc[j]));
  }
  if constexpr (!STORE) {
    return;
  } else if constexpr (MODE == B200D_EPI_BIAS_F32 || MODE == B200D_EPI_SIGMOID_F32 || MODE == EPI_PARTIAL) {
    float* o = reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + n0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
  } else {
    __half* o = reinterpret_cast<__half*>(p.out) + static_cast<size_t>(row) * p.ldo + n0;
#pragma unroll
    for (int j = 0; j < 32; j += 8) *reinterpret_cast<uint4*>(o + j) = pack8_f16(acc + j);
  }
}

// Chebyshev epilogue.  The W operand holds V^T (b = 32 or 64 block vectors) split into three bf16 parts, part q of
// vector j in row q * b + j: [hi | mid | lo] (b = 64: N = 192; b = 32: N = 128 with 32 zero rows of padding), so the
// accumulator columns (j, b + j, 2 b + j) sum to (A V)[row][j] at ~fp32 accuracy on the exactly representable graph.
__device__ __forceinline__ void split3_bf16(float v, __nv_bfloat16& h, __nv_bfloat16& m, __nv_bfloat16& l) {
  h = __float2bfloat16_rn(v);
  float r = v - __bfloat162float(h);
  m = __float2bfloat16_rn(r);
  r -= __bfloat162float(m);
  l = __float2bfloat16_rn(r);
}

// 32 block vectors [c0, c0 + 32) of one accumulator row: the Chebyshev step y = ca (deg x - A x) + cb x + cc xprev and
// the same 3-way split of y for the next step's W operand.
// TS: distance (TMEM columns) between the hi / mid / lo accumulators of one vector; VS: the same in rows of the W operand
// (the block size b).  c_tmem: first accumulator column of the 32 vectors; c_out: their first column in x / out / V.
template <int TS, int VS>
__device__ __forceinline__ void cheb_epilogue_cols(const GemmParams& p, int row, uint32_t t_addr, int c_tmem, int c_out) {
  const b200d_gemm_epilogue& e = p.epi;
  uint32_t r0[32], r1[32], r2[32];
  tmem_ld32(t_addr + c_tmem, r0);
  tmem_ld32(t_addr + TS + c_tmem, r1);
  tmem_ld32(t_addr + 2 * TS + c_tmem, r2);
  tmem_ld_wait();
  if (row < p.M) {
    const float dg = __ldg(e.deg + row);
    const float* x = e.x32 + static_cast<size_t>(row) * e.ldx + c_out;
    const float* xp = e.xprev32 ? e.xprev32 + static_cast<size_t>(row) * e.ldx + c_out : nullptr;
    float* o = reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + c_out;
    __nv_bfloat16* vh = reinterpret_cast<__nv_bfloat16*>(e.vt);
    if (e.n_peers <= 0) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float av = (__uint_as_float(r0[j]) + __uint_as_float(r1[j])) + __uint_as_float(r2[j]);
        const float xv = x[j];
        float y = e.ca * (dg * xv - av) + e.cb * xv;
        if (xp) y += e.cc * xp[j];
        o[j] = y;
        if (vh) {
          __nv_bfloat16 h, m, l;
          split3_bf16(y, h, m, l);
          const size_t vrow = static_cast<size_t>(c_out + j);
          vh[(vrow) * e.ldvt + row] = h;
          vh[(vrow + VS) * e.ldvt + row] = m;
          vh[(vrow + 2 * VS) * e.ldvt + row] = l;
        }
      }
    } else {
      // Row-sharded product: this rank owns the rows of the launch; the bf16 split of y (the next product's W operand, needed
      // in full by every rank) goes straight into every peer's copy over NVLink (lanes = consecutive rows: 64-byte segments),
      // the fp32 rows only when the caller needs them everywhere (B200D_GEMM_PEER_OUT32), as 16-byte stores.
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float av = (__uint_as_float(r0[j]) + __uint_as_float(r1[j])) + __uint_as_float(r2[j]);
        const float xv = x[j];
        float y = e.ca * (dg * xv - av) + e.cb * xv;
        if (xp) y += e.cc * xp[j];
        r0[j] = __float_as_uint(y);
      }
      const bool out_everywhere = (e.flags & B200D_GEMM_PEER_OUT32) != 0;
#pragma unroll 1
      for (int r = 0; r < e.n_peers; ++r) {
        const long long delta = e.peer_delta[r];
        if (out_everywhere || delta == 0) {
          float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<char*>(o) + delta);
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            o4[j >> 2] = make_float4(__uint_as_float(r0[j]), __uint_as_float(r0[j + 1]), __uint_as_float(r0[j + 2]), __uint_as_float(r0[j + 3]));
        }
        if (vh) {
          __nv_bfloat16* vr = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(vh) + delta);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            __nv_bfloat16 h, m, l;
            split3_bf16(__uint_as_float(r0[j]), h, m, l);
            const size_t vrow = static_cast<size_t>(c_out + j);
            vr[(vrow) * e.ldvt + row] = h;
            vr[(vrow + VS) * e.ldvt + row] = m;
            vr[(vrow + 2 * VS) * e.ldvt + row] = l;
          }
        }
      }
    }
  }
}

template <int BLOCK_N, int MODE, bool BF16>
__global__ void __launch_bounds__(256, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::B_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // BLOCK_N == 96: the 64-vector Chebyshev product split over TWO CTAs per row tile (32 vectors x [hi | mid | lo] each), so
  // that a 10 000-row graph gives 158 tiles instead of 79 on the 148 SMs; the second read of the A rows comes from L2
  constexpr bool SPLIT = BLOCK_N == 96;
  const int num_m = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int kblocks = (p.K + BLOCK_K - 1) / BLOCK_K;
  // EPI_PARTIAL: one column tile (N == BLOCK_N); the second tile coordinate is the K segment
  const int kseg = (MODE == EPI_PARTIAL) ? p.kseg : kblocks;
  const int num_n = (MODE == EPI_PARTIAL) ? (kblocks + kseg - 1) / kseg : SPLIT ? 2 : p.N / BLOCK_N;
  const int total = num_m * num_n;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int m_blk = tile / num_n, n_blk = (MODE == EPI_PARTIAL) ? 0 : tile % num_n;
        const int kb0 = (MODE == EPI_PARTIAL) ? (tile % num_n) * kseg : 0;
        const int kb1 = (MODE == EPI_PARTIAL) ? (kb0 + kseg < kblocks ? kb0 + kseg : kblocks) : kblocks;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], A_BYTES + Cfg::B_BYTES);
          tma_load_2d(sA + stage * A_BYTES, &tmA, &full[stage], kb * BLOCK_K, m_blk * BLOCK_M);
          if constexpr (SPLIT) {  // rows [32 h, 32 h + 32) of the hi, mid and lo parts (64 rows apart) of V^T: three 32-row boxes
#pragma unroll
            for (int q = 0; q < 3; ++q)
              tma_load_2d(sB + stage * Cfg::B_BYTES + q * 32 * BLOCK_K * 2, &tmB, &full[stage], kb * BLOCK_K, q * 64 + n_blk * 32);
          } else {
            tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, &full[stage], kb * BLOCK_K, n_blk * BLOCK_N);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D fp32 (1<<4), A/B format (f16 = 0, bf16 = 1) at bits 7 / 10, both K-major,
      // N>>3 at bit 17, M>>4 at bit 24
      constexpr uint32_t fmt = BF16 ? 1u : 0u;
      constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(BLOCK_N >> 3) << 17) |
                                 (static_cast<uint32_t>(BLOCK_M >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_STRIDE;
        const int kb0 = (MODE == EPI_PARTIAL) ? (tile % num_n) * kseg : 0;
        const int kb1 = (MODE == EPI_PARTIAL) ? (kb0 + kseg < kblocks ? kb0 + kseg : kblocks) : kblocks;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = make_sw128_desc(smem_u32(sA + stage * A_BYTES));
          const uint64_t b_desc = make_sw128_desc(smem_u32(sB + stage * Cfg::B_BYTES));
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k)
            umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int wq = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * BLOCK_M + wq * 32 + lane;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * Cfg::ACC_STRIDE;
      if constexpr (MODE == EPI_PARTIAL) {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(t_row + c * 32, r);
          tmem_ld_wait();
          if (row < p.M) {
            float acc_f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc_f[j] = __uint_as_float(r[j]);
            // segment n_blk of the partial buffer: rows [n_blk * M, (n_blk + 1) * M)
            epilogue_chunk<MODE>(p, n_blk * p.M + row, c * 32, acc_f);
          }
        }
      } else if constexpr (MODE == B200D_EPI_CHEB && SPLIT) {
        cheb_epilogue_cols<32, 64>(p, row, t_row, 0, n_blk * 32);
      } else if constexpr (MODE == B200D_EPI_CHEB) {
        constexpr int B = (BLOCK_N == 192) ? 64 : 32;  // one block of vectors per launch (N == BLOCK_N)
#pragma unroll 1
        for (int c0 = 0; c0 < B; c0 += 32) cheb_epilogue_cols<B, B>(p, row, t_row, c0, c0);
      } else {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(t_row + c * 32, r);
          tmem_ld_wait();
          if (row < p.M) {
            float acc_f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc_f[j] = __uint_as_float(r[j]);
            epilogue_chunk<MODE>(p, row, n_blk * BLOCK_N + c * 32, acc_f);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------ 2-CTA kernel
// CTA-pair variant (cta_group::2) for the large pointwise convs.  The 1-CTA kernel above moves 48 KB from L2 per
// 128x256x64 MMA block and saturates the L2 -> SM path (~6.3 KB/clk chip-wide) at about half of the tensor peak.
// Here the two CTAs of a cluster (the two SMs of a TPC) share one 256 x 256 output tile: each loads only ITS 128
// rows of A and ITS 128 rows of W (32 KB per block instead of 48 KB), the leader issues tcgen05.mma.cta_group::2
// with M = 256 reading both halves, and each CTA's TMEM ends up with its own 128 accumulator rows.
//   TMA loads of both CTAs signal the LEADER's full barrier (peer bit cleared in the barrier address);
//   tcgen05.commit multicasts the "stage free" / "accumulator ready" arrivals to both CTAs;
//   the epilogue warps of both CTAs arrive on the leader's tmem-empty barrier (remote arrive for the peer).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
constexpr int B2_BYTES = 128 * BLOCK_K * 2;             // this CTA's half of the 256-row W tile
constexpr int STAGES2 = 6;
// fp16 epilogues stage 32 rows x 64 columns per warp in shared memory and write them out as full 128-byte lines: the row-per-thread 16-byte stores of the first version reached only
// 1.6 TB/s on the store-bound K = 128 GEMMs (ncu r01 v3: attention projection 583 MB in 359 us)
// Eight epilogue warps: two per TMEM lane quarter, each taking one 128-column half of the accumulator.  The epilogue of
// a 32-column chunk is a latency chain (tcgen05.ld -> wait -> operand loads -> math -> staged store), eight of them in a
// row took about as long as the tile's MMAs with one warp per quarter (se_res 0.68, bias 0.79 of the bias_relu rate).
constexpr int EPI_WARPS2 = 8;
constexpr int THREADS2 = (4 + EPI_WARPS2) * 32;
// staging rows are 128 bytes with the 16-byte chunk index XOR-ed with (row & 7): conflict-free for the row-per-lane
// writes and for the line-per-8-lanes reads, without padding (8 x 4 KB fit next to six pipeline stages)
constexpr int STAGE_PITCH = 128;
constexpr int STAGE_BYTES = EPI_WARPS2 * 32 * STAGE_PITCH;
constexpr int SMEM2_BYTES = 1024 + STAGES2 * (A_BYTES + B2_BYTES) + 256 + STAGE_BYTES;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {  // arrive on the even CTA's copy of `bar`
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}

template <int MODE, bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS2, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  constexpr int BN = (MODE == B200D_EPI_CHEB) ? 192 : 256;  // Chebyshev: 64 vectors x [hi | mid | lo]
  constexpr int B_HALF_BYTES = (BN / 2) * BLOCK_K * 2;      // bytes this CTA loads per stage (the slot stays B2_BYTES)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES2 * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES2 * B2_BYTES);
  uint64_t* empty = full + STAGES2;
  uint64_t* tfull = empty + STAGES2;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* stage_base = reinterpret_cast<uint8_t*>(full) + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES2; ++i) {
      mbar_init(&full[i], 2);   // leader's arrive.expect_tx + the peer's remote arrive
      mbar_init(&empty[i], 1);  // multicast tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);   // multicast tcgen05.commit
      mbar_init(&tempty[i], 2 * EPI_WARPS2);  // epilogue warps of both CTAs (leader's copy only)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int num_n = p.N / BN;
  const int total = num_m * num_n;
  const int kblocks = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cid; tile < total; tile += ncl) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        const int row_a = m_blk * 2 * BLOCK_M + static_cast<int>(rank) * BLOCK_M;
        const int row_b = n_blk * BN + static_cast<int>(rank) * (BN / 2);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          tma_load_2d_2sm(sA + stage * A_BYTES, &tmA, &full[stage], kb * BLOCK_K, row_a);
          tma_load_2d_2sm(sB + stage * B2_BYTES, &tmB, &full[stage], kb * BLOCK_K, row_b);
          if (leader) mbar_expect_tx(&full[stage], 2 * (A_BYTES + B_HALF_BYTES));
          else mbar_arrive_leader(&full[stage]);
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      // D fp32, A/B fp16 K-major, N = BN, M = 256 (both CTAs)
      constexpr uint32_t fmt = BF16 ? 1u : 0u;
      constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
                                 (static_cast<uint32_t>((2 * BLOCK_M) >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cid; tile < total; tile += ncl) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = make_sw128_desc(smem_u32(sA + stage * A_BYTES));
          const uint64_t b_desc = make_sw128_desc(smem_u32(sB + stage * B2_BYTES));
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k)
            umma_f16_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&empty[stage]);
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int wq = warp & 3;           // TMEM lane quarter
    const int half = (warp - 4) >> 2;  // which 128 accumulator columns
    int acc = 0;
    uint32_t acc_phase = 0;
    // SE-residual epilogue: the residual operand (256 bytes of this thread's row per tile) does not depend on the
    // accumulator, and read on demand it costs one DRAM latency per 32-column chunk -- eight in a row, about as long as
    // the tile's MMAs.  Its lines are pulled into L2 one tile ahead instead.
    auto prefetch_aux = [&](int t) {
      if constexpr (MODE == B200D_EPI_SE_RES) {
        if (t < total) {
          const int prow = (t / num_n) * 2 * BLOCK_M + static_cast<int>(rank) * BLOCK_M + wq * 32 + lane;
          if (prow < p.M) {
            const char* a = reinterpret_cast<const char*>(reinterpret_cast<const __half*>(p.epi.aux16) +
                                                          static_cast<size_t>(prow) * p.ldo + (t % num_n) * BN + half * 128);
#pragma unroll
            for (int i = 0; i < 2; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + i * 128));
          }
        }
      }
    };
    prefetch_aux(cid);
    for (int tile = cid; tile < total; tile += ncl) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      prefetch_aux(tile + ncl);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * 2 * BLOCK_M + static_cast<int>(rank) * BLOCK_M + wq * 32 + lane;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * BN;
      if constexpr (MODE == B200D_EPI_CHEB) {
        cheb_epilogue_cols<64, 64>(p, row, t_row, half * 32, half * 32);
      } else if constexpr (MODE == B200D_EPI_BIAS_F32 || MODE == B200D_EPI_SIGMOID_F32) {
#pragma unroll 1
        for (int c = half * 4; c < half * 4 + 4; ++c) {
          uint32_t r[32];
          tmem_ld32(t_row + c * 32, r);
          tmem_ld_wait();
          if (row < p.M) {
            float acc_f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc_f[j] = __uint_as_float(r[j]);
            epilogue_chunk<MODE>(p, row, n_blk * BN + c * 32, acc_f);
          }
        }
      } else {
        uint8_t* stage = stage_base + (warp - 4) * 32 * STAGE_PITCH;
        const int row_w0 = m_blk * 2 * BLOCK_M + static_cast<int>(rank) * BLOCK_M + wq * 32;  // first row of this warp
        __half* out16 = reinterpret_cast<__half*>(p.out);
#pragma unroll 1
        for (int c = half * 4; c < half * 4 + 4; ++c) {
          uint32_t r[32];
          tmem_ld32(t_row + c * 32, r);
          tmem_ld_wait();
          float acc_f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) acc_f[j] = __uint_as_float(r[j]);
          if (row < p.M) epilogue_chunk<MODE, false>(p, row, n_blk * BN + c * 32, acc_f);
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            *reinterpret_cast<uint4*>(stage + lane * STAGE_PITCH + ((((c & 1) * 4 + (j >> 3)) ^ (lane & 7)) << 4)) = pack8_f16(acc_f + j);
          if (c & 1) {
            __syncwarp();
            const int col0 = n_blk * BN + (c - 1) * 32;  // 64 columns = 128 bytes per row
            // B200D_EPI_BIAS with epi.colsum: per-window column sums of the (fp16-rounded) output for SqueezeExcite's time
            // mean, taken from the staged tile on its way out -- lane (seg, rows seg-group) already holds 8 columns of 8
            // rows; rows before / from the window boundary `br` go to two partial sums, combined over the four lanes that
            // share a column segment.  Replaces a full re-read of the activation by time_stats_kernel.
            float sa[8], sb[8];
            bool want_sums = false;
            int br = 32;
            if constexpr (MODE == B200D_EPI_BIAS) {
              want_sums = p.epi.colsum != nullptr;
              if (want_sums) {
                const int T = p.epi.rows_per_seg;
                br = (row_w0 / T + 1) * T - row_w0;
#pragma unroll
                for (int q = 0; q < 8; ++q) sa[q] = sb[q] = 0.f;
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = i * 4 + (lane >> 3), seg = lane & 7;
              const uint4 v = *reinterpret_cast<const uint4*>(stage + rr * STAGE_PITCH + ((seg ^ (rr & 7)) << 4));
              if (row_w0 + rr < p.M) {
                *reinterpret_cast<uint4*>(out16 + static_cast<size_t>(row_w0 + rr) * p.ldo + col0 + seg * 8) = v;
                if constexpr (MODE == B200D_EPI_BIAS) {
                  if (want_sums) {
                    float f[8];
                    unpack8_f16(v, f);
                    if (rr < br) {
#pragma unroll
                      for (int q = 0; q < 8; ++q) sa[q] += f[q];
                    } else {
#pragma unroll
                      for (int q = 0; q < 8; ++q) sb[q] += f[q];
                    }
                  }
                }
              }
            }
            if constexpr (MODE == B200D_EPI_BIAS) {
              if (want_sums) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  sa[q] += __shfl_xor_sync(0xffffffffu, sa[q], 8);
                  sa[q] += __shfl_xor_sync(0xffffffffu, sa[q], 16);
                  sb[q] += __shfl_xor_sync(0xffffffffu, sb[q], 8);
                  sb[q] += __shfl_xor_sync(0xffffffffu, sb[q], 16);
                }
                if (lane < 8 && row_w0 < p.M) {
                  float* o = p.epi.colsum + (static_cast<size_t>(row_w0 >> 5) * 2) * p.N + col0 + lane * 8;
                  *reinterpret_cast<float4*>(o) = make_float4(sa[0], sa[1], sa[2], sa[3]);
                  *reinterpret_cast<float4*>(o + 4) = make_float4(sa[4], sa[5], sa[6], sa[7]);
                  *reinterpret_cast<float4*>(o + p.N) = make_float4(sb[0], sb[1], sb[2], sb[3]);
                  *reinterpret_cast<float4*>(o + p.N + 4) = make_float4(sb[4], sb[5], sb[6], sb[7]);
                }
              }
            }
            __syncwarp();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tempty[acc]);
        else mbar_arrive_leader(&tempty[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;

static EncodeTiledFn get_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  });
  return g_encode;
}

// 2-D K-major map: dim0 = K (contiguous), dim1 = rows; box = 64 x box_rows, 128B swizzle, zero OOB fill.
static int make_map(CUtensorMap* map, const void* base, bool bf16, int64_t rows, int64_t k, int64_t ld_elems, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(B200D_ELAUNCH, "%s: cuTensorMapEncodeTiled entry point not found%s", "make_map");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld_elems) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BLOCK_K), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    static thread_local char buf[32];
    snprintf(buf, sizeof(buf), "%d", static_cast<int>(r));
    return set_error(B200D_ELAUNCH, "%s: cuTensorMapEncodeTiled failed with CUresult %s", "make_map", buf);
  }
  return B200D_OK;
}


// ------------------------------------------------------------------------------------ split-K Chebyshev product, second half
// The products of the spectral solver are skinny (M = K = 10 000 .. 57 600 graph rows, N = 128 / 192): one 128-row tile per CTA
// leaves the chip idle or quantised (79 tiles on 148 SMs: 2 rounds for 10 of them) and re-reads the whole W operand per tile.
// With a workspace (epi.splitk_ws) the launch is cut into (row tile, K segment) units of kSplitKBlocks K blocks -- a fixed
// length, so that the summation order of an output row depends on nothing but K: the same bits on any number of GPUs / rows --
// each unit's raw accumulators go to partial[seg][M][N] (EPI_PARTIAL), and this kernel sums the segments in ascending order and
// applies the Chebyshev epilogue: y = ca (deg x - (hi + mid + lo)) + cb x + cc xprev, the fp32 rows and the 3-way bf16 split
// of y^T for the next product, locally or into every peer (b200d_gemm_epilogue.n_peers).
constexpr int kSplitKBlocks = 24;

// 1024 threads per 32-row block: the kernel is a pure stream (56 MB of partials per 10 000-row product), and with 256 threads a
// block the ~2 resident blocks per SM had too few requests in flight (ncu r02: 25 us, 2.4 TB/s).
constexpr int kFixupThreads = 1024;

template <int B>
__global__ void __launch_bounds__(kFixupThreads) cheb_fixup_kernel(const float* __restrict__ partial, int nseg, int M, const b200d_gemm_epilogue e,
                                                         float* __restrict__ out, int ldo) {
  constexpr int NW = (B == 64) ? 192 : 128;
  constexpr int ROWS = 32;
  __shared__ float ys[ROWS][B + 1];
  const int row0 = blockIdx.x * ROWS;
  const int j = threadIdx.x % B;
  constexpr int RSTEP = kFixupThreads / B;
  const size_t seg_stride = static_cast<size_t>(M) * NW;
  const bool out_everywhere = e.n_peers > 0 && (e.flags & B200D_GEMM_PEER_OUT32) != 0;
  for (int r = threadIdx.x / B; r < ROWS; r += RSTEP) {
    const int row = row0 + r;
    float y = 0.f;
    if (row < M) {
      const float* p0 = partial + static_cast<size_t>(row) * NW + j;
      // ascending-segment sums; four segments' loads (12 independent requests) are issued before they are added -- one load per
      // add made the kernel a latency chain (ncu r02: 24 us for 56 MB)
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
      int s = 0;
      for (; s + 4 <= nseg; s += 4) {
        float t0[4], t1[4], t2[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float* q = p0 + (s + u) * seg_stride;
          t0[u] = __ldcs(q);
          t1[u] = __ldcs(q + B);
          t2[u] = __ldcs(q + 2 * B);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          a0 += t0[u];
          a1 += t1[u];
          a2 += t2[u];
        }
      }
      for (; s < nseg; ++s) {
        a0 += __ldcs(p0 + s * seg_stride);
        a1 += __ldcs(p0 + s * seg_stride + B);
        a2 += __ldcs(p0 + s * seg_stride + 2 * B);
      }
      const float av = (a0 + a1) + a2;
      const float xv = e.x32[static_cast<size_t>(row) * e.ldx + j];
      y = e.ca * (__ldg(e.deg + row) * xv - av) + e.cb * xv;
      if (e.xprev32) y += e.cc * e.xprev32[static_cast<size_t>(row) * e.ldx + j];
      float* o = out + static_cast<size_t>(row) * ldo + j;
      if (e.n_peers <= 0) {
        *o = y;
      } else {
        for (int q = 0; q < e.n_peers; ++q)
          if (out_everywhere || e.peer_delta[q] == 0) *reinterpret_cast<float*>(reinterpret_cast<char*>(o) + e.peer_delta[q]) = y;
      }
    }
    ys[r][j] = y;
  }
  if (e.vt == nullptr) return;
  __syncthreads();
  // transposed bf16 split: lane = row (32 consecutive rows = 64 contiguous bytes per part and vector), warps over the vectors
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = row0 + lane;
  if (row >= M) return;
  __nv_bfloat16* vh = reinterpret_cast<__nv_bfloat16*>(e.vt);
  const int n_peers = e.n_peers > 0 ? e.n_peers : 1;
  for (int q = 0; q < n_peers; ++q) {
    __nv_bfloat16* vr = e.n_peers > 0 ? reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(vh) + e.peer_delta[q]) : vh;
    for (int c = warp; c < B; c += kFixupThreads / 32) {
      __nv_bfloat16 h, m, l;
      split3_bf16(ys[lane][c], h, m, l);
      vr[static_cast<size_t>(c) * e.ldvt + row] = h;
      vr[static_cast<size_t>(c + B) * e.ldvt + row] = m;
      vr[static_cast<size_t>(c + 2 * B) * e.ldvt + row] = l;
    }
  }
}

template <int BLOCK_N, int MODE, bool BF16>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  auto kern = gemm_tcgen05_kernel<BLOCK_N, MODE, BF16>;
  static bool attr_set[kMaxDevices] = {};  // the attribute is per device: a process may drive more than one GPU
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices || !attr_set[dev]) {
    B200D_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    if (dev >= 0 && dev < kMaxDevices) attr_set[dev] = true;
  }
  const int kblocks = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int tiles = ((p.M + BLOCK_M - 1) / BLOCK_M) *
                    (MODE == EPI_PARTIAL ? (kblocks + p.kseg - 1) / p.kseg : BLOCK_N == 96 ? 2 : p.N / BLOCK_N);
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  kern<<<grid, 256, Cfg::SMEM_BYTES, stream>>>(ta, tb, p);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

template <int MODE, bool BF16 = false>
static int launch_2cta(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  auto kern = gemm_tcgen05_2cta_kernel<MODE, BF16>;
  static bool attr_set[kMaxDevices] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices || !attr_set[dev]) {
    B200D_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
    if (dev >= 0 && dev < kMaxDevices) attr_set[dev] = true;
  }
  constexpr int BN = (MODE == B200D_EPI_CHEB) ? 192 : 256;
  const int tiles = ((p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * (p.N / BN);
  int pairs = tiles < kNumSMs / 2 ? tiles : kNumSMs / 2;
  // development knobs for tools/concurrency_check.py (the pair kernel next to another stream's kernels, see b200d.h)
  static const int exp_pairs = getenv("B200D_EXP_PAIRS") ? atoi(getenv("B200D_EXP_PAIRS")) : 0;
  static const int exp_policy = getenv("B200D_EXP_POLICY") ? atoi(getenv("B200D_EXP_POLICY")) : 0;
  if (exp_pairs > 0 && pairs > exp_pairs) pairs = exp_pairs;
  if (exp_policy != 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(THREADS2);
    cfg.dynamicSmemBytes = SMEM2_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterSchedulingPolicyPreference;
    attr[0].val.clusterSchedulingPolicyPreference = exp_policy == 1 ? cudaClusterSchedulingPolicySpread : cudaClusterSchedulingPolicyLoadBalancing;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B200D_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, p));
    return B200D_OK;
  }
  kern<<<2 * pairs, THREADS2, SMEM2_BYTES, stream>>>(ta, tb, p);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

static int dispatch_mode_2cta(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t s) {
  switch (p.epi.mode) {
    case B200D_EPI_BIAS: return launch_2cta<B200D_EPI_BIAS>(ta, tb, p, s);
    case B200D_EPI_BIAS_RELU: return launch_2cta<B200D_EPI_BIAS_RELU>(ta, tb, p, s);
    case B200D_EPI_SE_RES: return launch_2cta<B200D_EPI_SE_RES>(ta, tb, p, s);
    case B200D_EPI_TDNN: return launch_2cta<B200D_EPI_TDNN>(ta, tb, p, s);
    case B200D_EPI_BIAS_F32: return launch_2cta<B200D_EPI_BIAS_F32>(ta, tb, p, s);
    case B200D_EPI_SIGMOID_F32: return launch_2cta<B200D_EPI_SIGMOID_F32>(ta, tb, p, s);
    case B200D_EPI_CHEB: return launch_2cta<B200D_EPI_CHEB, true>(ta, tb, p, s);
  }
  return set_error(B200D_EINVAL, "%s: unknown epilogue mode%s", "b200d_gemm_f16");
}

template <int BLOCK_N>
static int dispatch_mode(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t s) {
  switch (p.epi.mode) {
    case B200D_EPI_BIAS: return launch<BLOCK_N, B200D_EPI_BIAS, false>(ta, tb, p, s);
    case B200D_EPI_BIAS_RELU: return launch<BLOCK_N, B200D_EPI_BIAS_RELU, false>(ta, tb, p, s);
    case B200D_EPI_SE_RES: return launch<BLOCK_N, B200D_EPI_SE_RES, false>(ta, tb, p, s);
    case B200D_EPI_TDNN: return launch<BLOCK_N, B200D_EPI_TDNN, false>(ta, tb, p, s);
    case B200D_EPI_BIAS_F32: return launch<BLOCK_N, B200D_EPI_BIAS_F32, false>(ta, tb, p, s);
    case B200D_EPI_SIGMOID_F32: return launch<BLOCK_N, B200D_EPI_SIGMOID_F32, false>(ta, tb, p, s);
    default: break;  // B200D_EPI_CHEB is launched by b200d_gemm_f16 directly (its tile width is the block's N)
  }
  return set_error(B200D_EINVAL, "%s: unknown epilogue mode%s", "b200d_gemm_f16");
}

// Which launches take the CTA-pair kernel: at least one 256 x 256 tile per TPC (the large pointwise convs), and the
// Chebyshev products of 64-vector blocks on large graphs (L2-traffic bound: the pair halves the W bytes per SM) -- unless
// the caller asks for the 1-CTA kernel on this call (B200D_GEMM_NO_PAIR: it runs several streams at once, see b200d.h).
bool gemm_uses_pair_kernel(int M, int N, int mode, int flags) {
  static const bool env_allow_2cta = getenv("B200D_GEMM_1CTA") == nullptr;
  if (!env_allow_2cta || (flags & B200D_GEMM_NO_PAIR)) return false;
  if (mode == B200D_EPI_CHEB) return N == 192 && M >= 4096;
  const long long tiles2 = static_cast<long long>((M + 255) / 256) * (N / 256);
  return N % 256 == 0 && tiles2 >= kNumSMs / 2;
}

}  // namespace b200d

using namespace b200d;

extern "C" size_t b200d_gemm_cheb_splitk_bytes(int32_t M, int32_t N, int32_t K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int kblocks = (K + BLOCK_K - 1) / BLOCK_K;
  const int nseg = (kblocks + kSplitKBlocks - 1) / kSplitKBlocks;
  return static_cast<size_t>(nseg) * M * N * sizeof(float);
}

extern "C" int b200d_gemm_f16(const void* A, int32_t lda, const void* W, int32_t ldw, int32_t M, int32_t N, int32_t K, void* out,
                              int32_t ldo, const b200d_gemm_epilogue* epi, void* stream) {
  B200D_CHECK_ARG(A && W && out && epi);
  B200D_CHECK_ARG(M > 0 && N > 0 && K > 0);
  B200D_CHECK_ARG(N % 128 == 0 || (epi->mode == B200D_EPI_CHEB && N == 192));
  B200D_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0);
  B200D_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const int mode = epi->mode;
  if (mode == B200D_EPI_CHEB) {
    B200D_CHECK_ARG(N == 128 || N == 192);  // block of 32 vectors ([hi | mid | lo | pad]) or of 64 ([hi | mid | lo])
    B200D_CHECK_ARG(epi->deg && epi->x32);
    B200D_CHECK_ARG(epi->vt == nullptr || epi->ldvt >= M);
  } else {
    B200D_CHECK_ARG(ldo % 8 == 0 && ldo >= N);
    if (mode == B200D_EPI_SE_RES) B200D_CHECK_ARG(epi->bias && epi->rowvec && epi->aux16 && epi->rows_per_seg > 0);
    if (mode == B200D_EPI_TDNN) B200D_CHECK_ARG(epi->scale && epi->shift && epi->rowvec && epi->rows_per_seg > 0);
    if (mode == B200D_EPI_BIAS || mode == B200D_EPI_BIAS_RELU || mode == B200D_EPI_BIAS_F32) B200D_CHECK_ARG(epi->bias);
  }
  const bool bf16 = mode == B200D_EPI_CHEB;
  const int block_n = (mode == B200D_EPI_CHEB) ? N : (N % 256 == 0) ? 256 : 128;
  const bool use_2cta = gemm_uses_pair_kernel(M, N, mode, epi->flags);
  if (epi->colsum != nullptr) {  // per-window column sums: an extra of the CTA-pair kernel's staged fp16 epilogue
    if (mode != B200D_EPI_BIAS || !use_2cta || epi->rows_per_seg < 32)
      return set_error(B200D_EINVAL, "%s: epi.colsum needs B200D_EPI_BIAS on a launch that takes the CTA-pair kernel and rows_per_seg >= 32%s", "b200d_gemm_f16");
  }
  // 64-vector Chebyshev products that leave SMs idle as one tile per 128 rows: two CTAs per row tile (BLOCK_N 96 kernel)
  static const bool env_no_split = getenv("B200D_CHEB_NO_SPLIT") != nullptr;
  const bool cheb_split = mode == B200D_EPI_CHEB && N == 192 && !use_2cta && !env_no_split && (M + BLOCK_M - 1) / BLOCK_M < kNumSMs;
  CUtensorMap ta, tb;
  int rc = make_map(&ta, A, bf16, M, K, lda, BLOCK_M);
  if (rc) return rc;
  rc = make_map(&tb, W, bf16, N, K, ldw, cheb_split ? 32 : use_2cta ? (mode == B200D_EPI_CHEB ? 96 : 128) : block_n);
  if (rc) return rc;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.out = out; p.ldo = ldo; p.kseg = 0; p.epi = *epi;
  if (mode == B200D_EPI_CHEB && epi->splitk_ws != nullptr) {
    // split-K: (row tile, K segment) units -> partial[seg][M][N], then the fix-up kernel (see cheb_fixup_kernel)
    const size_t need = b200d_gemm_cheb_splitk_bytes(M, N, K);
    if (need > epi->splitk_ws_bytes)
      return set_error(B200D_EWORKSPACE, "%s: epi.splitk_ws too small (b200d_gemm_cheb_splitk_bytes)%s", "b200d_gemm_f16");
    B200D_CHECK_ARG((reinterpret_cast<uintptr_t>(epi->splitk_ws) & 15) == 0);
    rc = make_map(&tb, W, true, N, K, ldw, N);
    if (rc) return rc;
    p.out = epi->splitk_ws;
    p.ldo = N;
    p.kseg = kSplitKBlocks;
    // (measured and dropped in round 2: 256-row units -- two MMA tiles per W stage, 352 instead of 503 MB through L2 per
    // 10 000^2 x 192 product -- gave the same bits and no speed-up, 56.0 against 52.8 ms per step: the product is bound by the
    // MMA's shared-memory operand reads at N = 192 per 128 rows, which only the CTA-pair form halves)
    rc = (N == 192) ? launch<192, EPI_PARTIAL, true>(ta, tb, p, as_stream(stream)) : launch<128, EPI_PARTIAL, true>(ta, tb, p, as_stream(stream));
    if (rc) return rc;
    const int kblocks = (K + BLOCK_K - 1) / BLOCK_K;
    const int nseg = (kblocks + kSplitKBlocks - 1) / kSplitKBlocks;
    const int blocks = (M + 31) / 32;
    if (N == 192)
      cheb_fixup_kernel<64><<<blocks, kFixupThreads, 0, as_stream(stream)>>>(reinterpret_cast<const float*>(epi->splitk_ws), nseg, M, *epi, reinterpret_cast<float*>(out), ldo);
    else
      cheb_fixup_kernel<32><<<blocks, kFixupThreads, 0, as_stream(stream)>>>(reinterpret_cast<const float*>(epi->splitk_ws), nseg, M, *epi, reinterpret_cast<float*>(out), ldo);
    B200D_CHECK_LAUNCH();
    return B200D_OK;
  }
  if (use_2cta) return dispatch_mode_2cta(ta, tb, p, as_stream(stream));
  if (mode == B200D_EPI_CHEB) {
    if (cheb_split) return launch<96, B200D_EPI_CHEB, true>(ta, tb, p, as_stream(stream));
    if (N == 192) return launch<192, B200D_EPI_CHEB, true>(ta, tb, p, as_stream(stream));
    return launch<128, B200D_EPI_CHEB, true>(ta, tb, p, as_stream(stream));
  }
  if (block_n == 256) return dispatch_mode<256>(ta, tb, p, as_stream(stream));
  return dispatch_mode<128>(ta, tb, p, as_stream(stream));
}
