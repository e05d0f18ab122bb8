// vsb200 -- Dynamic Sequence Parallelism reshard as direct 128-bit stores into the peers' receive windows
// (NVLink 5 / NVSwitch P2P), replacing tensor_split+contiguous -> NCCL all_to_all -> cat+contiguous (+pad / narrow).
//
// The sender walks its local tensor once; every 16-byte vector goes straight to its final position in the
// destination rank's output tensor (zero padding of the scattered axis is synthesised, padding of the gathered
// axis is never sent).  Completion: every CTA fences (system scope); the last CTA to finish publishes the epoch
// into each peer's flag slot; vsb_dsp_wait spins (acquire, system scope) on the receiver's stream.
#include "vsb_common.cuh"
#include "vsb_host.h"

namespace vsb {

constexpr int kMaxWorld = 16;
struct DspPeers {
  bf16* recv[kMaxWorld];
  unsigned* flags[kMaxWorld];
};

__device__ unsigned g_dsp_done_ctas = 0;

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// dir 0: local [B, T, Sl, C] -> rank d receives frames [d*Tl, (d+1)*Tl) into [B, Tl, S, C] at columns rank*Sl + sl
// dir 1: local [B, Tl, S, C] -> rank d receives columns [d*Sl, (d+1)*Sl) into [B, T, Sl, C] at frames rank*Tl + tl
__global__ void __launch_bounds__(256) dsp_scatter_kernel(const bf16* __restrict__ local, DspPeers peers, int rank,
                                                          int world, int dir, int B, int T, int S, int C,
                                                          unsigned epoch) {
  const int Tp = ((T + world - 1) / world) * world, Sp = ((S + world - 1) / world) * world;
  const int Tl = Tp / world, Sl = Sp / world;
  const int cv = C >> 3;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  // 4 independent 16-byte vectors per thread per trip: all loads are issued before the (posted) peer stores so
  // that enough bytes are in flight to cover the NVLink round trip.
  constexpr int U = 4;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = dir == 0 ? (long long)B * Tp * Sl * cv : (long long)B * Tl * Sp * cv;
  for (long long i0 = tid; i0 < total; i0 += nthreads * U) {
    uint4 v[U];
    uint4* dst[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + (long long)u * nthreads;
      dst[u] = nullptr;
      v[u] = zero;
      if (i >= total) continue;
      const int c = int(i % cv);
      long long r = i / cv;
      if (dir == 0) {
        const int sl = int(r % Sl);
        r /= Sl;
        const int t = int(r % Tp);
        const int b = int(r / Tp);
        const int col = rank * Sl + sl;
        if (col >= S) continue;  // gathered-axis padding is narrowed away at the receiver: never sent
        const int d = t / Tl, tl = t - d * Tl;
        if (t < T) v[u] = *reinterpret_cast<const uint4*>(local + (((size_t)b * T + t) * Sl + sl) * C + c * 8);
        dst[u] = reinterpret_cast<uint4*>(peers.recv[d] + (((size_t)b * Tl + tl) * S + col) * C + c * 8);
      } else {
        const int col = int(r % Sp);
        r /= Sp;
        const int tl = int(r % Tl);
        const int b = int(r / Tl);
        const int tg = rank * Tl + tl;
        if (tg >= T) continue;
        const int d = col / Sl, sl = col - d * Sl;
        if (col < S) v[u] = *reinterpret_cast<const uint4*>(local + (((size_t)b * Tl + tl) * S + col) * C + c * 8);
        dst[u] = reinterpret_cast<uint4*>(peers.recv[d] + (((size_t)b * T + tg) * Sl + sl) * C + c * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (dst[u] != nullptr) *dst[u] = v[u];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(&g_dsp_done_ctas, 1u) + 1u;
    if (done == gridDim.x) {
      g_dsp_done_ctas = 0;
      __threadfence_system();
      for (int d = 0; d < world; ++d) st_release_sys(peers.flags[d] + rank, epoch);
    }
  }
}

__global__ void dsp_wait_kernel(const unsigned* flags, int world, unsigned epoch) {
  const int src = threadIdx.x;
  if (src < world) {
    long long t0 = clock64();
    while ((int)(ld_acquire_sys(flags + src) - epoch) < 0) {
      if (clock64() - t0 > 20000000000ll) {
        printf("vsb200: dsp_wait watchdog src=%d epoch=%u have=%u\n", src, epoch, ld_acquire_sys(flags + src));
        __trap();
      }
    }
  }
}

}  // namespace vsb

using namespace vsb;

extern "C" int vsb_dsp_scatter(const vsb_bf16* local, void* const* host_peer_recv, void* const* host_peer_flags,
                               int rank, int world, int to_spatial_shard, int B, int T, int S, int C, unsigned epoch,
                               void* stream) {
  if (!local || !host_peer_recv || !host_peer_flags || world < 1 || rank < 0 || rank >= world || B <= 0 || T <= 0 ||
      S <= 0 || C <= 0)
    return fail(VSB_ERR_INVALID, "dsp_scatter: bad args");
  if (world > kMaxWorld || C % 8 || !aligned16(local)) return fail(VSB_ERR_UNSUPPORTED, "dsp_scatter: world <= 16, C %% 8 == 0");
  DspPeers peers;
  for (int i = 0; i < world; ++i) {
    peers.recv[i] = (bf16*)host_peer_recv[i];
    peers.flags[i] = (unsigned*)host_peer_flags[i];
    if (!peers.recv[i] || !peers.flags[i] || !aligned16(peers.recv[i])) return fail(VSB_ERR_INVALID, "dsp_scatter: peer %d window", i);
  }
  const int grid = num_sms() * 4;
  dsp_scatter_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)local, peers, rank, world, to_spatial_shard,
                                                              B, T, S, C, epoch);
  return check_launch("dsp_scatter");
}

extern "C" int vsb_dsp_wait(const void* my_flags, int world, unsigned epoch, void* stream) {
  if (!my_flags || world < 1 || world > kMaxWorld) return fail(VSB_ERR_INVALID, "dsp_wait: bad args");
  dsp_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const unsigned*)my_flags, world, epoch);
  return check_launch("dsp_wait");
}

// ---- symmetric-window plumbing (CUDA IPC); windows are cudaMalloc'ed here so the handle maps the exact base ----
extern "C" int vsb_dsp_alloc(void** out, size_t bytes) {
  if (!out || bytes == 0) return fail(VSB_ERR_INVALID, "dsp_alloc: bad args");
  cudaError_t e = cudaMalloc(out, bytes);
  if (e != cudaSuccess) return fail(VSB_ERR_CUDA, "dsp_alloc: %s", cudaGetErrorString(e));
  e = cudaMemset(*out, 0, bytes);
  if (e != cudaSuccess) return fail(VSB_ERR_CUDA, "dsp_alloc memset: %s", cudaGetErrorString(e));
  return VSB_OK;
}
extern "C" int vsb_dsp_free(void* p) {
  cudaError_t e = cudaFree(p);
  return e == cudaSuccess ? VSB_OK : fail(VSB_ERR_CUDA, "dsp_free: %s", cudaGetErrorString(e));
}
extern "C" int vsb_ipc_get_handle(void* devptr, void* out_handle64) {
  if (!devptr || !out_handle64) return fail(VSB_ERR_INVALID, "ipc_get_handle: bad args");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, devptr);
  if (e != cudaSuccess) return fail(VSB_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  memcpy(out_handle64, &h, sizeof(h));
  return VSB_OK;
}
extern "C" int vsb_ipc_open_handle(const void* handle64, void** out_devptr) {
  if (!handle64 || !out_devptr) return fail(VSB_ERR_INVALID, "ipc_open_handle: bad args");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(out_devptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail(VSB_ERR_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  return VSB_OK;
}
extern "C" int vsb_ipc_close_handle(void* devptr) {
  cudaError_t e = cudaIpcCloseMemHandle(devptr);
  return e == cudaSuccess ? VSB_OK : fail(VSB_ERR_CUDA, "cudaIpcCloseMemHandle: %s", cudaGetErrorString(e));
}
