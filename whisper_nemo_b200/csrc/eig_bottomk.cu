// b200d_eig_bottomk: the k lowest eigenvectors of L = diag(deg) - A for the binarised affinity graph -- the device work behind
// upstream's SpectralClustering.getSpectralEmbeddings (offline_clustering.py: torch.linalg.eigh of the full N x N Laplacian,
// k columns kept), as ONE composite C-ABI call.
//
// Chebyshev-filtered subspace iteration: a block of b = 32 / 64 vectors is repeatedly pushed through a Chebyshev polynomial
// of L that damps [theta_b, lambda_max] (each term one tcgen05 GEMM A*V with V split into three bf16 parts, i.e.
// fp32-accurate products on an exactly representable A -- or an fp32 row-gather product over the graph's CSR lists when it
// has few neighbours), re-orthonormalised by CholQR2 and rotated to Ritz vectors.  k-means only sees the k-dimensional
// invariant subspace, which is what converges here.  The polynomial degree and the stopping test are decided on the host
// from 2 x b floats read back per outer iteration, so this entry point synchronises `stream` (~10-15 times per call).
#include "common.cuh"
#include "composite.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace b200d {

static inline int block_of(int k) { return k + 8 <= 32 ? 32 : (k + 8 <= 64 ? 64 : 0); }
static inline int operand_rows(int b) { return b == 64 ? 192 : 128; }  // [hi | mid | lo] parts of b rows each (b = 32: 32 zero rows)

static bool use_csr(int n, int p, int max_row_nnz, double max_density) {
  if (p <= 0) return false;
  const long long nnz = 2LL * p;
  return nnz <= max_row_nnz || static_cast<double>(nnz) <= max_density * n;
}

struct EigWs {
  float *W, *Y[3], *G, *Q, *theta, *resid;
  void* vt[2];
  void* gws;
  size_t gws_bytes;
  int32_t* rowptr;
  uint32_t* colw;
  long long capacity;
  void* splitk;          // partial accumulators of the split-K products (dense path)
  size_t splitk_bytes;
};

static size_t carve_eig(uint8_t* base, int n, int b, bool sparse, int p, EigWs* out) {
  size_t pos = 0;
  auto take = [&](size_t bytes) {
    uint8_t* q = base ? base + pos : nullptr;
    pos = (pos + bytes + 255) & ~static_cast<size_t>(255);
    return q;
  };
  EigWs w{};
  const size_t nb = static_cast<size_t>(n) * b * 4;
  w.W = reinterpret_cast<float*>(take(nb));
  for (int i = 0; i < 3; ++i) w.Y[i] = reinterpret_cast<float*>(take(nb));
  w.G = reinterpret_cast<float*>(take(static_cast<size_t>(b) * b * 4));
  w.Q = reinterpret_cast<float*>(take(static_cast<size_t>(b) * b * 4));
  w.theta = reinterpret_cast<float*>(take(b * 4));
  w.resid = reinterpret_cast<float*>(take(b * 4));
  w.gws_bytes = b200d_gram_workspace_bytes(n, b);
  w.gws = take(w.gws_bytes);
  const size_t ldvt = (static_cast<size_t>(n) + 7) / 8 * 8;
  if (sparse) {
    w.capacity = static_cast<long long>(std::min(2LL * p, static_cast<long long>(n))) * n;
    w.rowptr = reinterpret_cast<int32_t*>(take((static_cast<size_t>(n) + 1) * 4));
    w.colw = reinterpret_cast<uint32_t*>(take(static_cast<size_t>(w.capacity) * 4));
  } else {
    for (int i = 0; i < 2; ++i) w.vt[i] = take(static_cast<size_t>(operand_rows(b)) * ldvt * 2);
    w.splitk_bytes = b200d_gemm_cheb_splitk_bytes(n, operand_rows(b), n);
    w.splitk = take(w.splitk_bytes);
  }
  if (out) *out = w;
  return pos;
}

}  // namespace b200d

using namespace b200d;

extern "C" int32_t b200d_eig_bottomk_block(int32_t k) { return block_of(k); }

extern "C" size_t b200d_eig_bottomk_workspace_bytes(int32_t n, int32_t k, int32_t p, const b200d_eig_options* opt) {
  const int b = block_of(k);
  if (n <= 0 || b == 0) return 0;
  const int max_row_nnz = opt ? opt->sparse_max_row_nnz : 32;
  const double max_density = opt ? opt->sparse_max_density : 1.0 / 64.0;
  return carve_eig(nullptr, n, b, use_csr(n, p, max_row_nnz, max_density), p, nullptr);
}

// The iteration.  grp == NULL: the whole graph on this GPU (a_bf16 = all n rows).  grp != NULL: a_bf16 holds the rows
// [row_lo, row_hi) of the graph only; every product computes those rows and stores the bf16 split of the result (and, for the
// two products per outer iteration whose fp32 result the small dense steps need in full, the fp32 rows) into EVERY rank's
// workspace over NVLink, followed by a device-side barrier; the O(n b^2) steps around the products run replicated on every
// rank from identical data, so all ranks take the same host decisions and end with the same block, bit for bit the one the
// single-GPU call computes.
static int eig_impl(const void* a_bf16, int32_t lda, const float* deg, int32_t n, int32_t row_lo, int32_t row_hi, int32_t k, int32_t p,
                    float* x, int32_t ldx, const b200d_eig_options* opt, b200d_eig_stats* stats, void* ws, size_t ws_bytes,
                    b200d_peer_group* grp, void* stream) {
  B200D_CHECK_ARG(a_bf16 && deg && x && ws && n > 0 && k > 0 && lda >= n && lda % 8 == 0);
  const int b = block_of(k);
  if (b == 0) return set_error(B200D_EINVAL, "%s: the subspace block is limited to 64 vectors (k <= 56)%s", "b200d_eig_bottomk");
  B200D_CHECK_ARG(ldx == b);
  if (n < 2 * b) return set_error(B200D_EINVAL, "%s: needs n >= 2 * block (use b200d_small_eig on the dense Laplacian for tiny graphs)%s", "b200d_eig_bottomk");
  const double tol = opt && opt->tol > 0 ? opt->tol : 2e-6;
  const int max_outer = opt && opt->max_outer > 0 ? opt->max_outer : 40;
  const int flags = opt ? opt->gemm_flags : 0;
  const bool sparse = grp == nullptr && use_csr(n, p, opt ? opt->sparse_max_row_nnz : 32, opt ? opt->sparse_max_density : 1.0 / 64.0);
  if (carve_eig(nullptr, n, b, sparse, p, nullptr) > ws_bytes)
    return set_error(B200D_EWORKSPACE, "%s: workspace too small (b200d_eig_bottomk_workspace_bytes)%s", "b200d_eig_bottomk");
  B200D_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 255) == 0);
  EigWs w;
  carve_eig(reinterpret_cast<uint8_t*>(ws), n, b, sparse, p, &w);
  cudaStream_t st = as_stream(stream);
  const int nw = operand_rows(b);
  const int ldvt = (n + 7) / 8 * 8;
  float* X = x;

  // Gershgorin bound on lambda_max(L) = 2 max deg
  std::vector<float> host(std::max(n, 2 * b));
  B200D_CHECK_CUDA(cudaMemcpyAsync(host.data(), deg, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, st));
  B200D_CHECK_CUDA(cudaStreamSynchronize(st));
  const double up = 2.0 * static_cast<double>(*std::max_element(host.begin(), host.begin() + n)) * 1.01 + 1e-3;

  if (sparse) {
    ProfScope ps("csr_from_dense", 0, st, 3);
    RC(b200d_csr_from_dense(a_bf16, n, lda, w.rowptr, w.colw, w.capacity, st));
  } else {
    for (int i = 0; i < 2; ++i) B200D_CHECK_CUDA(cudaMemsetAsync(w.vt[i], 0, static_cast<size_t>(nw) * ldvt * 2, st));
  }
  int gemms = 0;
  // one operator application: out = ca (deg x - A x) + cb x + cc xprev; vin / vout are the bf16 splits of x / out (dense path)
  // `everywhere`: the fp32 result is needed in full on every rank (sharded call only; the bf16 split always is)
  auto step = [&](const void* vin, float* out, const float* xx, const float* xprev, double ca, double cb, double cc, void* vout,
                  bool everywhere) -> int {
    ++gemms;
    if (sparse) {
      ProfScope ps("spmm_cheb", 0, st);
      return b200d_spmm_cheb(w.rowptr, w.colw, n, b, deg, xx, xprev, b, static_cast<float>(ca), static_cast<float>(cb), static_cast<float>(cc), out, b, st);
    }
    b200d_gemm_epilogue e{};
    e.mode = B200D_EPI_CHEB;
    e.ca = static_cast<float>(ca); e.cb = static_cast<float>(cb); e.cc = static_cast<float>(cc);
    e.ldx = b;
    e.ldvt = ldvt;
    e.flags = flags;
    e.splitk_ws = w.splitk;  // always the split-K form: an output row's bits then do not depend on how many rows the launch has
    e.splitk_ws_bytes = w.splitk_bytes;
    if (grp == nullptr) {
      e.deg = deg;
      e.x32 = xx;
      e.xprev32 = xprev;
      e.vt = vout;
      ProfScope ps("gemm[cheb|splitk]", 2.0 * n * nw * static_cast<double>(n), st, 2);
      return b200d_gemm_f16(a_bf16, lda, vin, ldvt, n, nw, n, out, b, &e, st);
    }
    const int m_rows = row_hi - row_lo;
    if (m_rows > 0) {
      const size_t off = static_cast<size_t>(row_lo) * b;
      e.deg = deg + row_lo;
      e.x32 = xx + off;
      e.xprev32 = xprev ? xprev + off : nullptr;
      e.vt = vout ? reinterpret_cast<uint16_t*>(vout) + row_lo : nullptr;  // column row_lo + r of V^T for local row r
      e.n_peers = grp->world;
      for (int r = 0; r < grp->world; ++r)
        e.peer_delta[r] = reinterpret_cast<const char*>(grp->base[r]) - reinterpret_cast<const char*>(grp->base[grp->rank]);
      if (everywhere) e.flags |= B200D_GEMM_PEER_OUT32;
      ProfScope ps("gemm[cheb|splitk|rows]", 2.0 * m_rows * nw * static_cast<double>(n), st, 2);
      RC(b200d_gemm_f16(a_bf16, lda, vin, ldvt, m_rows, nw, n, out + off, b, &e, st));
    }
    ProfScope ps("peer_barrier", 0, st);
    return b200d_peer_barrier(grp, st);
  };
  auto gram = [&](const float* a, const float* c, float* out) -> int {
    ProfScope ps("gram", 0, st, 2);
    return b200d_gram(a, c, n, b, b, out, w.gws, w.gws_bytes, st);
  };
  auto cholqr = [&](float* V, void* want_vt) -> int {
    RC(gram(V, V, w.G));
    { ProfScope ps("small_eig[cholesky]", 0, st); RC(b200d_small_eig(w.G, b, nullptr, w.Q, 1, st)); }
    ProfScope ps("right_mul", 0, st);
    return b200d_right_mul(V, n, b, b, w.Q, V, want_vt, ldvt, st);
  };
  RC(cholqr(X, nullptr));
  RC(cholqr(X, w.vt[0]));
  std::vector<double> history;
  double max_resid = NAN;
  bool converged = false;
  int outer = 0;
  for (; outer < max_outer;) {
    ++outer;
    // Rayleigh-Ritz on span(X): W = L X, H = X^T W, X <- X Q, W <- W Q
    RC(step(w.vt[0], w.W, X, nullptr, 1.0, 0.0, 0.0, nullptr, true));
    RC(gram(X, w.W, w.G));
    { ProfScope ps("small_eig[jacobi]", 0, st); RC(b200d_small_eig(w.G, b, w.theta, w.Q, 0, st)); }
    { ProfScope ps("right_mul", 0, st); RC(b200d_right_mul(X, n, b, b, w.Q, X, w.vt[0], ldvt, st)); }
    { ProfScope ps("right_mul", 0, st); RC(b200d_right_mul(w.W, n, b, b, w.Q, w.W, nullptr, ldvt, st)); }
    { ProfScope ps("resid_norms", 0, st, 2); RC(b200d_resid_norms(w.W, X, w.theta, n, b, b, w.resid, w.gws, w.gws_bytes, st)); }
    B200D_CHECK_CUDA(cudaMemcpyAsync(host.data(), w.theta, b * 4, cudaMemcpyDeviceToHost, st));
    B200D_CHECK_CUDA(cudaMemcpyAsync(host.data() + b, w.resid, b * 4, cudaMemcpyDeviceToHost, st));
    B200D_CHECK_CUDA(cudaStreamSynchronize(st));
    if (grp) RC(b200d_peer_status(grp, st));
    double rmax = 0.0;
    for (int j = 0; j < k; ++j) rmax = std::max(rmax, std::sqrt(std::max(static_cast<double>(host[b + j]), 0.0)));
    max_resid = rmax / up;
    history.push_back(max_resid);
    if (stats && outer <= B200D_EIG_HISTORY) stats->history[outer - 1] = static_cast<float>(max_resid);
    if (max_resid <= tol) { converged = true; break; }
    if (history.size() >= 8 && max_resid < 5e-5 && max_resid > 0.5 * history[history.size() - 4]) { converged = true; break; }  // fp32 floor of the products
    // filter interval [a, up]: damp everything above the block's largest Ritz value
    const double th_b = host[b - 1], th_k = host[k - 1];
    const double a = std::min(std::max(th_b, 1e-3 * up), 0.95 * up);
    const double e = (up - a) / 2.0, c = (up + a) / 2.0;
    const double t_k = (c - std::max(th_k, 0.0)) / e, t_0 = c / e;
    int m = static_cast<int>(std::nearbyint(std::acosh(1e3) / std::max(std::acosh(std::max(t_k, 1.0 + 1e-9)), 1e-6)));
    const int m_cap = static_cast<int>(std::acosh(1e7) / std::acosh(t_0));
    m = std::max(3, std::min(std::min(m, std::max(m_cap, 3)), 48));
    // scaled three-term recurrence (gain 1 at lambda = 0):  Y1 = (s1/e)(L - c) X,  Y_{i+1} = (2 s_{i+1}/e)(L - c) Y_i - s_i s_{i+1} Y_{i-1}
    const double sigma1 = e / (0.0 - c);
    double sigma = sigma1;
    const double tau = 2.0 / sigma1;
    RC(step(w.vt[0], w.Y[0], X, nullptr, sigma1 / e, -c * sigma1 / e, 0.0, w.vt[1], m == 1));
    const float* prev = X;
    float* cur = w.Y[0];
    int vin = 1;
    for (int i = 2; i <= m; ++i) {
      float* nxt = w.Y[(i - 1) % 3];
      const double sigma_new = 1.0 / (tau - sigma);
      const double ca = 2.0 * sigma_new / e;
      RC(step(w.vt[vin], nxt, cur, prev, ca, -c * ca, -sigma * sigma_new, w.vt[1 - vin], i == m));
      sigma = sigma_new;
      prev = cur;
      cur = nxt;
      vin = 1 - vin;
    }
    B200D_CHECK_CUDA(cudaMemcpyAsync(X, cur, static_cast<size_t>(n) * b * 4, cudaMemcpyDeviceToDevice, st));
    RC(cholqr(X, nullptr));
    RC(cholqr(X, w.vt[0]));
  }
  if (stats) {
    stats->block = b;
    stats->outer = outer;
    stats->gemms = gemms;
    stats->max_resid = static_cast<float>(max_resid);
    stats->converged = converged ? 1 : 0;
    stats->sparse = sparse ? 1 : 0;
  }
  return B200D_OK;
}

extern "C" int b200d_eig_bottomk(const void* a_bf16, int32_t lda, const float* deg, int32_t n, int32_t k, int32_t p, float* x, int32_t ldx,
                                 const b200d_eig_options* opt, b200d_eig_stats* stats, void* ws, size_t ws_bytes, void* stream) {
  return eig_impl(a_bf16, lda, deg, n, 0, n, k, p, x, ldx, opt, stats, ws, ws_bytes, nullptr, stream);
}

extern "C" size_t b200d_eig_bottomk_sharded_peer_bytes(int32_t n, int32_t k) {
  const int b = block_of(k);
  if (n <= 0 || b == 0) return 0;
  return B200D_PEER_HEADER_BYTES + carve_eig(nullptr, n, b, false, 0, nullptr);
}

extern "C" int b200d_eig_bottomk_sharded(const void* a_rows_bf16, int32_t lda, const float* deg, int32_t n, int32_t row_lo, int32_t row_hi,
                                         int32_t k, float* x, int32_t ldx, const b200d_eig_options* opt, b200d_eig_stats* stats,
                                         b200d_peer_group* grp, void* stream) {
  B200D_CHECK_ARG(grp && grp->world >= 1 && grp->world <= B200D_MAX_PEERS && grp->rank >= 0 && grp->rank < grp->world);
  B200D_CHECK_ARG(row_lo >= 0 && row_lo <= row_hi && row_hi <= n);
  B200D_CHECK_ARG(row_hi == row_lo || a_rows_bf16 != nullptr);
  const size_t need = b200d_eig_bottomk_sharded_peer_bytes(n, k);
  if (need == 0 || need > grp->bytes)
    return set_error(B200D_EWORKSPACE, "%s: peer buffers too small (b200d_eig_bottomk_sharded_peer_bytes)%s", "b200d_eig_bottomk_sharded");
  for (int r = 0; r < grp->world; ++r) B200D_CHECK_ARG(grp->base[r] != nullptr);
  uint8_t* ws = reinterpret_cast<uint8_t*>(grp->base[grp->rank]) + B200D_PEER_HEADER_BYTES;
  // entry barrier: no rank may store into a peer's workspace while that peer still runs the previous user of it
  RC(b200d_peer_barrier(grp, stream));
  const void* a = a_rows_bf16 ? a_rows_bf16 : static_cast<const void*>(ws);  // (an empty shard never dereferences it)
  return eig_impl(a, lda, deg, n, row_lo, row_hi, k, 0, x, ldx, opt, stats, ws, need - B200D_PEER_HEADER_BYTES, grp, stream);
}
