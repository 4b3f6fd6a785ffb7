/* A host that is NOT Python: plain C over include/b200d.h + the CUDA runtime, driving the clustering half of the path
 * (cosine affinity -> min-max / fusion -> top-p binarisation -> bottom-k Laplacian eigenvectors -> k-means) on planted
 * clusters, the way a C / C++ / Go-cgo / JNI caller of libb200d.so would.
 *
 *   gcc -O2 -I include tools/cabi_host_example.c -o /tmp/cabi_host -L whisper_nemo_b200 -lb200d \
 *       -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/whisper_nemo_b200 -Wl,-rpath,/usr/local/cuda/lib64
 *   /tmp/cabi_host            # exit code 0 and "labels match the planted clusters" on a B200
 *
 * Every buffer is allocated here (cudaMalloc); the library only computes.  tests/test_gpu_kernels.py builds and runs it.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "b200d.h"

#define CHECK_CUDA(x)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (x);                                                              \
    if (e__ != cudaSuccess) {                                                           \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e__));                         \
      return 2;                                                                         \
    }                                                                                   \
  } while (0)
#define CHECK_B200D(x)                                                                  \
  do {                                                                                  \
    int rc__ = (x);                                                                     \
    if (rc__ != B200D_OK) {                                                             \
      fprintf(stderr, "%s failed (%d): %s\n", #x, rc__, b200d_last_error());            \
      return 3;                                                                         \
    }                                                                                   \
  } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static double uniform01(void) { /* xorshift64*: the draws only have to be reproducible */
  rng_state ^= rng_state >> 12;
  rng_state ^= rng_state << 25;
  rng_state ^= rng_state >> 27;
  return (double)((rng_state * 0x2545F4914F6CDD1Dull) >> 11) / 9007199254740992.0;
}
static float gauss(void) { return (float)(sqrt(-2.0 * log(uniform01() + 1e-300)) * cos(6.283185307179586 * uniform01())); }

int main(void) {
  const int n = 3000, d = 192, k = 4, p = 300, n_trials = 30;
  printf("%s\n", b200d_version());
  CHECK_B200D(b200d_check_device());
  cudaStream_t stream;
  CHECK_CUDA(cudaStreamCreate(&stream));

  /* planted clusters: k centres, unit noise */
  float* x_h = (float*)malloc(sizeof(float) * n * d);
  int* truth = (int*)malloc(sizeof(int) * n);
  float* centres = (float*)malloc(sizeof(float) * k * d);
  for (int i = 0; i < k * d; ++i) centres[i] = 2.5f * gauss();
  for (int i = 0; i < n; ++i) {
    truth[i] = (int)(uniform01() * k) % k;
    for (int c = 0; c < d; ++c) x_h[(size_t)i * d + c] = centres[truth[i] * d + c] + gauss();
  }

  float *x, *xn, *cosm, *minmax, *fused, *deg, *ev;
  int32_t *ident, *labels;
  void *a16, *sel;
  const int lda = (n + 7) / 8 * 8;
  CHECK_CUDA(cudaMalloc((void**)&x, sizeof(float) * n * d));
  CHECK_CUDA(cudaMalloc((void**)&xn, sizeof(float) * n * d));
  CHECK_CUDA(cudaMalloc((void**)&cosm, sizeof(float) * (size_t)n * n));
  CHECK_CUDA(cudaMalloc((void**)&fused, sizeof(float) * (size_t)n * n));
  CHECK_CUDA(cudaMalloc((void**)&minmax, sizeof(float) * 2));
  CHECK_CUDA(cudaMalloc((void**)&ident, sizeof(int32_t) * n));
  CHECK_CUDA(cudaMalloc((void**)&deg, sizeof(float) * n));
  CHECK_CUDA(cudaMalloc(&a16, (size_t)2 * n * lda));
  CHECK_CUDA(cudaMalloc(&sel, (size_t)n * n));
  CHECK_CUDA(cudaMalloc((void**)&labels, sizeof(int32_t) * n));
  CHECK_CUDA(cudaMemcpyAsync(x, x_h, sizeof(float) * n * d, cudaMemcpyHostToDevice, stream));
  int32_t* ident_h = (int32_t*)malloc(sizeof(int32_t) * n);
  for (int i = 0; i < n; ++i) ident_h[i] = i;
  CHECK_CUDA(cudaMemcpyAsync(ident, ident_h, sizeof(int32_t) * n, cudaMemcpyHostToDevice, stream));

  /* getCosAffinityMatrix: cos_similarity (eps 3.5e-4) + ScalerMinMax, as one fused-scale pass with an identity mapping */
  CHECK_B200D(b200d_l2_normalize(x, xn, n, d, 3.5e-4f, stream));
  CHECK_B200D(b200d_cos_affinity(xn, n, d, cosm, minmax, stream));
  const float* cos_list[1] = {cosm};
  const int32_t ns_list[1] = {n};
  const int32_t* map_list[1] = {ident};
  const float* mm_list[1] = {minmax};
  const float w_list[1] = {1.0f};
  CHECK_B200D(b200d_fuse_scales(1, cos_list, ns_list, map_list, mm_list, w_list, fused, n, stream));

  /* getAffinityGraphMat with a fixed p (the NME sweep that picks p needs host logic and is left out of this example) */
  CHECK_B200D(b200d_topp_binarize(fused, n, p, a16, lda, deg, sel, stream));

  /* SpectralClustering.getSpectralEmbeddings: the k lowest eigenvectors of D - A */
  const int b = b200d_eig_bottomk_block(k);
  b200d_eig_options opt;
  memset(&opt, 0, sizeof(opt));
  opt.tol = 2e-6;
  opt.max_outer = 40;
  opt.sparse_max_row_nnz = 32;
  opt.sparse_max_density = 1.0 / 64.0;
  b200d_eig_stats stats;
  memset(&stats, 0, sizeof(stats));
  const size_t ws_bytes = b200d_eig_bottomk_workspace_bytes(n, k, p, &opt);
  void* ws;
  CHECK_CUDA(cudaMalloc(&ws, ws_bytes));
  float* ev_h = (float*)malloc(sizeof(float) * n * b);
  for (int i = 0; i < n * b; ++i) ev_h[i] = gauss();
  CHECK_CUDA(cudaMalloc((void**)&ev, sizeof(float) * n * b));
  CHECK_CUDA(cudaMemcpyAsync(ev, ev_h, sizeof(float) * n * b, cudaMemcpyHostToDevice, stream));
  CHECK_B200D(b200d_eig_bottomk(a16, lda, deg, n, k, p, ev, b, &opt, &stats, ws, ws_bytes, stream));
  printf("eig_bottomk: block %d, %d outer iterations, %d products, residual %.2e, converged %d, sparse products %d\n", stats.block, stats.outer,
         stats.gemms, stats.max_resid, stats.converged, stats.sparse);

  /* kmeans_torch on the first k columns; the host supplies the random draws */
  CHECK_CUDA(cudaMemcpyAsync(ev_h, ev, sizeof(float) * n * b, cudaMemcpyDeviceToHost, stream));
  CHECK_CUDA(cudaStreamSynchronize(stream));
  float* emb_h = (float*)malloc(sizeof(float) * n * k);
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < k; ++c) emb_h[i * k + c] = ev_h[(size_t)i * b + c];
  float *emb, *rands;
  int32_t* fallback;
  CHECK_CUDA(cudaMalloc((void**)&emb, sizeof(float) * n * k));
  CHECK_CUDA(cudaMemcpyAsync(emb, emb_h, sizeof(float) * n * k, cudaMemcpyHostToDevice, stream));
  float* rands_h = (float*)malloc(sizeof(float) * (k - 1) * n_trials);
  for (int i = 0; i < (k - 1) * n_trials; ++i) rands_h[i] = (float)uniform01();
  int32_t fallback_h[16];
  for (int i = 0; i < 16; ++i) fallback_h[i] = (int32_t)(uniform01() * n) % n;
  CHECK_CUDA(cudaMalloc((void**)&rands, sizeof(float) * (k - 1) * n_trials));
  CHECK_CUDA(cudaMalloc((void**)&fallback, sizeof(fallback_h)));
  CHECK_CUDA(cudaMemcpyAsync(rands, rands_h, sizeof(float) * (k - 1) * n_trials, cudaMemcpyHostToDevice, stream));
  CHECK_CUDA(cudaMemcpyAsync(fallback, fallback_h, sizeof(fallback_h), cudaMemcpyHostToDevice, stream));
  const size_t km_bytes = b200d_kmeans_workspace_bytes(n, k, k, n_trials);
  void* km_ws;
  CHECK_CUDA(cudaMalloc(&km_ws, km_bytes));
  CHECK_B200D(b200d_kmeans(emb, n, k, k, (int32_t)(uniform01() * n) % n, rands, n_trials, fallback, 16, 15, 1e-4f, labels, km_ws, km_bytes, stream));
  int32_t* labels_h = (int32_t*)malloc(sizeof(int32_t) * n);
  CHECK_CUDA(cudaMemcpyAsync(labels_h, labels, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, stream));
  CHECK_CUDA(cudaStreamSynchronize(stream));

  /* every planted cluster must map to exactly one label and vice versa */
  int table[4][4];
  memset(table, 0, sizeof(table));
  for (int i = 0; i < n; ++i) {
    if (labels_h[i] < 0 || labels_h[i] >= k) {
      fprintf(stderr, "label %d out of range at %d\n", labels_h[i], i);
      return 4;
    }
    table[truth[i]][labels_h[i]] += 1;
  }
  int used[4] = {0, 0, 0, 0}, agree = 0;
  for (int t = 0; t < k; ++t) {
    int best = 0;
    for (int l = 1; l < k; ++l)
      if (table[t][l] > table[t][best]) best = l;
    if (used[best]) {
      fprintf(stderr, "two planted clusters share label %d\n", best);
      return 5;
    }
    used[best] = 1;
    agree += table[t][best];
  }
  printf("labels match the planted clusters on %d of %d points\n", agree, n);
  return agree == n ? 0 : 6;
}
