// Peer memory of the GPUs of one NVSwitch box: the plumbing under the row-sharded spectral solver (b200d_eig_bottomk_sharded).
//
// Every rank (one process per GPU) owns ONE buffer of identical size and layout; the other ranks map it (CUDA IPC) and
// store into it directly over NVLink from their kernels' epilogues.  The first B200D_PEER_HEADER_BYTES of a buffer hold the
// barrier flags: flags[s] is written by rank s only (its barrier epoch), read by the owner only.
//
// The reference has no multi-GPU code; this replaces what would otherwise be an NCCL all-gather after
// every product of the iterative eigensolver ("final spectral embedding").
#include "common.cuh"

namespace b200d {

struct PeerBases {
  void* base[B200D_MAX_PEERS];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// One CTA, lane t talks to rank t: publish this rank's epoch in rank t's flags[rank] (after a system-scope fence, so the
// peer stores of the kernels before this one on the stream are visible first), then wait until rank t has published the
// same epoch here.  A rank that never arrives (crashed peer, diverged call sequence) ends the wait after timeout_ns with
// status = 1 instead of hanging the GPU.
__global__ void peer_barrier_kernel(const PeerBases pb, int rank, int world, uint32_t epoch, unsigned long long timeout_ns) {
  const int t = threadIdx.x;
  uint32_t* mine = reinterpret_cast<uint32_t*>(pb.base[rank]);
  if (t < world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(pb.base[t]) + rank, epoch);
    const unsigned long long t0 = globaltimer_ns();
    while (static_cast<int32_t>(ld_acquire_sys(mine + t) - epoch) < 0) {
      if (globaltimer_ns() - t0 > timeout_ns) {
        mine[B200D_PEER_STATUS_WORD] = 1u;
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
}

}  // namespace b200d

using namespace b200d;

extern "C" int b200d_peer_alloc(size_t bytes, void** ptr, void* handle_host) {
  B200D_CHECK_ARG(ptr && bytes >= B200D_PEER_HEADER_BYTES);
  void* p = nullptr;
  B200D_CHECK_CUDA(cudaMalloc(&p, bytes));
  B200D_CHECK_CUDA(cudaMemset(p, 0, B200D_PEER_HEADER_BYTES));
  B200D_CHECK_CUDA(cudaDeviceSynchronize());
  if (handle_host) {
    static_assert(sizeof(cudaIpcMemHandle_t) == B200D_PEER_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
      cudaFree(p);
      return set_error(B200D_ELAUNCH, "%s: cudaIpcGetMemHandle: %s", "b200d_peer_alloc", cudaGetErrorString(e));
    }
    memcpy(handle_host, &h, sizeof(h));
  }
  *ptr = p;
  return B200D_OK;
}

extern "C" int b200d_peer_open(const void* handle_host, void** ptr) {
  B200D_CHECK_ARG(handle_host && ptr);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  B200D_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return B200D_OK;
}

extern "C" int b200d_peer_close(void* ptr) {
  B200D_CHECK_ARG(ptr);
  B200D_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return B200D_OK;
}

extern "C" int b200d_peer_free(void* ptr) {
  B200D_CHECK_ARG(ptr);
  B200D_CHECK_CUDA(cudaFree(ptr));
  return B200D_OK;
}

static int check_group(const b200d_peer_group* g, const char* who) {
  if (!g || g->world < 1 || g->world > B200D_MAX_PEERS || g->rank < 0 || g->rank >= g->world)
    return set_error(B200D_EINVAL, "%s: bad peer group (rank / world)%s", who);
  for (int r = 0; r < g->world; ++r)
    if (!g->base[r]) return set_error(B200D_EINVAL, "%s: peer group with an unmapped rank%s", who);
  return B200D_OK;
}

extern "C" int b200d_peer_barrier(b200d_peer_group* grp, void* stream) {
  int rc = check_group(grp, "b200d_peer_barrier");
  if (rc) return rc;
  PeerBases pb{};
  for (int r = 0; r < grp->world; ++r) pb.base[r] = grp->base[r];
  grp->epoch += 1;
  const unsigned long long timeout_ns = 1000000ull * (grp->timeout_ms > 0 ? grp->timeout_ms : 10000u);
  peer_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(pb, grp->rank, grp->world, grp->epoch, timeout_ns);
  B200D_CHECK_LAUNCH();
  return B200D_OK;
}

extern "C" int b200d_peer_status(const b200d_peer_group* grp, void* stream) {
  int rc = check_group(grp, "b200d_peer_status");
  if (rc) return rc;
  uint32_t st = 0;
  B200D_CHECK_CUDA(cudaMemcpyAsync(&st, reinterpret_cast<const uint32_t*>(grp->base[grp->rank]) + B200D_PEER_STATUS_WORD, 4,
                                   cudaMemcpyDeviceToHost, as_stream(stream)));
  B200D_CHECK_CUDA(cudaStreamSynchronize(as_stream(stream)));
  if (st != 0) return set_error(B200D_ELAUNCH, "%s: a peer barrier timed out (a rank did not arrive: crashed peer or diverged call sequence)%s", "b200d_peer_status");
  return B200D_OK;
}
