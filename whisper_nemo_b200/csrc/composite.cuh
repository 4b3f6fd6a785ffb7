// Shared by the composite entry points (titanet.cu, eig_bottomk.cu): kernel accounting / optional per-kernel event spans.
#pragma once
#include "common.cuh"

namespace b200d {

struct Span {
  char name[48];
  double work;
  cudaEvent_t e0, e1;
};

// Brackets the kernel(s) one inner call launches: always counts them (b200d_launch_count); records a cudaEvent pair on the
// stream while b200d_profile_start() is active.
struct ProfScope {
  bool on;
  cudaStream_t st;
  Span sp;
  ProfScope(const char* name, double work, cudaStream_t stream, int kernels = 1);
  ~ProfScope();
};

bool gemm_uses_pair_kernel(int M, int N, int mode, int flags);  // gemm_tcgen05.cu

#define RC(expr)                       \
  do {                                 \
    const int rc__ = (expr);           \
    if (rc__ != B200D_OK) return rc__; \
  } while (0)

}  // namespace b200d
