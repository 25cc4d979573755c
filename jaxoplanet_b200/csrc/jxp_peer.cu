// Peer-memory all-gather of per-set values: the only exchange of the set-sharded path is 8 bytes per
// parameter set and step.  At that size a collective is pure latency -- ncclAllGather through torch.distributed
// costs ~25 us per step (its own stream, two cross-stream event waits, the LL-protocol kernel) next to a 90-200 us
// step -- so each rank STORES its block straight into every rank's exchange buffer over NVLink (buffers shared
// between the processes of the box by CUDA IPC), publishes an epoch flag there, and waits for the flags of the
// others in its own buffer: one launch of one CTA, no host synchronisation, no extra stream.
//
// Exchange buffer of a rank (jxp_peer_buffer_bytes):
//   [0, 128)    u64 flag[16]   flag[r] = the last epoch rank r has delivered into this buffer
//   [128, 136)  u64 status     != 0: a wait timed out (the peer named in the low bits never delivered)
//   [136, 144)  u64 epoch      the last epoch this rank has started (epoch argument 0: the kernel counts itself,
//                              so that the launch has no argument that changes from call to call and a CUDA graph
//                              that holds it can be replayed)
//   [256, ...)  2 x n_total doubles: the gathered values of even / odd epochs (double-buffered: a rank that is one
//               step ahead writes the other half, and cannot get two steps ahead without this rank's next flag)
#include <cstdint>
#include <cstring>

#include <cuda_runtime.h>

#include "../../include/jaxoplanet_b200.h"
#include "jxp_host.h"

namespace {
constexpr int PEER_MAX_WORLD = 16;
constexpr size_t PEER_HDR = 256;

struct PeerArgs {
  unsigned long long bufs[PEER_MAX_WORLD];  // every rank's exchange buffer as mapped into THIS process
  const double* src;
  double* out;  // optional: the gathered values copied out of the exchange buffer (n_total doubles), or nullptr
  long long n_local, offset, n_total;
  int rank, world;
  unsigned long long epoch;
  unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(1024) k_peer_allgather(const __grid_constant__ PeerArgs a) {
  __shared__ unsigned long long s_epoch;
  if (threadIdx.x == 0) {
    unsigned long long* ctr = reinterpret_cast<unsigned long long*>(a.bufs[a.rank]) + 17;
    const unsigned long long e = a.epoch ? a.epoch : *ctr + 1ull;
    *ctr = e;
    s_epoch = e;
  }
  __syncthreads();
  const unsigned long long epoch = s_epoch;
  const size_t slot = PEER_HDR + (size_t)(epoch & 1ull) * (size_t)a.n_total * sizeof(double);
  for (long long i = threadIdx.x; i < a.n_local; i += blockDim.x) {
    const double v = a.src[i];
    for (int r = 0; r < a.world; ++r)
      reinterpret_cast<double*>(a.bufs[r] + slot)[a.offset + i] = v;  // (peer stores: NVLink; own buffer: local)
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < a.world) {
    const int r = (int)threadIdx.x;
    __threadfence_system();  // (cumulative: orders the stores of the whole CTA, observed through the barrier, before the flag)
    unsigned long long* theirs = reinterpret_cast<unsigned long long*>(a.bufs[r]) + a.rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(a.bufs[a.rank]) + r;
    const unsigned long long t0 = globaltimer_ns();
    for (;;) {
      unsigned long long f;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(mine) : "memory");
      if (f >= epoch) break;
      if (globaltimer_ns() - t0 > a.timeout_ns) {  // never hang the GPU on a peer that died
        reinterpret_cast<unsigned long long*>(a.bufs[a.rank])[16] = 0x100ull | (unsigned long long)r;
        break;
      }
      __nanosleep(100);
    }
  }
  __syncthreads();
  if (a.out != nullptr) {  // (the flags were read with acquire semantics; the values bypass L1)
    const double* got = reinterpret_cast<const double*>(a.bufs[a.rank] + slot);
    for (long long i = threadIdx.x; i < a.n_total; i += blockDim.x) a.out[i] = __ldcg(got + i);
  }
}
}  // namespace

using namespace jxph;

extern "C" size_t jxp_peer_buffer_bytes(int64_t n_total) {
  return n_total < 0 ? 0 : PEER_HDR + 2 * (size_t)n_total * sizeof(double);
}

extern "C" int jxp_peer_alloc(size_t bytes, void** ptr, void* handle_out) {
  if (!ptr || !handle_out) return fail(JXP_ERR_NULL_POINTER, "jxp_peer_alloc: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes < PEER_HDR ? PEER_HDR : bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, bytes < PEER_HDR ? PEER_HDR : bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    if (p) cudaFree(p);
    return fail(JXP_ERR_CUDA, "jxp_peer_alloc: %s", cudaGetErrorString(e));
  }
  std::memcpy(handle_out, &h, sizeof(h));
  *ptr = p;
  return JXP_OK;
}

extern "C" int jxp_peer_open(const void* handle, void** ptr) {
  if (!handle || !ptr) return fail(JXP_ERR_NULL_POINTER, "jxp_peer_open: null pointer");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail(JXP_ERR_CUDA, "jxp_peer_open: %s", cudaGetErrorString(e));
  *ptr = p;
  return JXP_OK;
}

extern "C" int jxp_peer_close(void* ptr) {
  cudaError_t e = ptr ? cudaIpcCloseMemHandle(ptr) : cudaSuccess;
  return e == cudaSuccess ? JXP_OK : fail(JXP_ERR_CUDA, "jxp_peer_close: %s", cudaGetErrorString(e));
}

extern "C" int jxp_peer_free(void* ptr) {
  cudaError_t e = ptr ? cudaFree(ptr) : cudaSuccess;
  return e == cudaSuccess ? JXP_OK : fail(JXP_ERR_CUDA, "jxp_peer_free: %s", cudaGetErrorString(e));
}

extern "C" int jxp_peer_allgather_f64(const double* src, int64_t n_local, int64_t offset, int64_t n_total,
                                      const uint64_t* bufs, int32_t rank, int32_t world, uint64_t epoch,
                                      double* out, void* stream) {
  if (!src || !bufs) return fail(JXP_ERR_NULL_POINTER, "jxp_peer_allgather: null pointer");
  if (world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world)
    return fail(JXP_ERR_BAD_SHAPE, "jxp_peer_allgather: rank %d of %d (at most %d ranks)", rank, world, PEER_MAX_WORLD);
  if (n_local < 0 || offset < 0 || offset + n_local > n_total)
    return fail(JXP_ERR_BAD_SHAPE, "jxp_peer_allgather: block [%lld, %lld) outside [0, %lld)", (long long)offset,
                (long long)(offset + n_local), (long long)n_total);
  // (epoch 0: the kernel counts the epochs itself, in its own buffer -- every rank must then do so)
  PeerArgs a;
  for (int r = 0; r < PEER_MAX_WORLD; ++r) a.bufs[r] = r < world ? bufs[r] : 0ull;
  for (int r = 0; r < world; ++r)
    if (!a.bufs[r]) return fail(JXP_ERR_NULL_POINTER, "jxp_peer_allgather: no buffer for rank %d", r);
  a.src = src;
  a.out = out;
  a.n_local = n_local;
  a.offset = offset;
  a.n_total = n_total;
  a.rank = rank;
  a.world = world;
  a.epoch = epoch;
  a.timeout_ns = 5000000000ull;  // 5 s
  k_peer_allgather<<<1, 1024, 0, (cudaStream_t)stream>>>(a);
  return check_launch("jxp_peer_allgather");
}

extern "C" int jxp_peer_status(const void* own_buf, uint64_t* status_host, void* stream) {
  if (!own_buf || !status_host) return fail(JXP_ERR_NULL_POINTER, "jxp_peer_status: null pointer");
  // status_host[0] = status word, status_host[1] = the last epoch started (two words)
  cudaError_t e = cudaMemcpyAsync(status_host, static_cast<const unsigned char*>(own_buf) + 128, 2 * sizeof(uint64_t),
                                  cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
  return e == cudaSuccess ? JXP_OK : fail(JXP_ERR_CUDA, "jxp_peer_status: %s", cudaGetErrorString(e));
}
