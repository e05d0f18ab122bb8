// tcgen05.mma execution cost on one SM (design aid for the attention kernels; not part of libvsb200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I videosys_b200/csrc tools/mma_microbench.cu -o tools/_bin/mma_microbench
// One CTA, one issuing warp (converged, elected lane).  Every experiment is a fully unrolled compile-time pattern of
// 16 "steps" repeated 4 times, then one commit + wait: cycles per step, both at issue and at completion.
// Operand contents are irrelevant (zeroed smem / whatever TMEM holds).
#include <cstdio>
#include <vector>

#include "vsb_common.cuh"

using namespace vsb;

__host__ __device__ constexpr uint32_t idesc_of(int n, int mn_b) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(mn_b) << 16) | (uint32_t(n >> 3) << 17) | (uint32_t(128 >> 4) << 24);
}

// PATTERN: 0 = SS chain (N), ACC accumulators round-robin
//          1 = TS chain (N, B MN-major)
//          2 = TS N=64 + TS N=16 pairs (the head_dim-72 P V step)
//          3 = the kt64 attention sequence for ONE query tile: 4 x (TS64 + TS16) into O, then 5 x SS64 into S   (13 MMAs / step-group)
//          4 = the same for TWO query tiles, A then B (26 MMAs)
//          5 = two query tiles interleaved MMA by MMA (26 MMAs)
template <int PATTERN, int N, int ACC>
__global__ void __launch_bounds__(128, 1) mma_bench(long long* out) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_ptr);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (warp == 0) {
    const uint32_t elected = elect_one();
    constexpr uint32_t hi128 = umma_desc_hi(1024, 2);
    constexpr uint32_t hi32 = umma_desc_hi(256, 6);
    const uint32_t a_lo = umma_desc_lo(smem_u32(smem), 16);                    // Q-like: 2 x 20 KB
    const uint32_t b_lo = umma_desc_lo(smem_u32(smem + 40960), 16);            // K-like
    const uint32_t v_lo = umma_desc_lo(smem_u32(smem + 40960 + 10240), 16) - (1u << 16) + ((8192u >> 4) << 16);
    const uint32_t v16_lo = umma_desc_lo(smem_u32(smem + 40960 + 18432), 16) - (1u << 16) + ((2048u >> 4) << 16);
    constexpr uint32_t id_n = idesc_of(N, 0), id_n_mn = idesc_of(N, 1), id64 = idesc_of(64, 0), id64_mn = idesc_of(64, 1),
                       id16_mn = idesc_of(16, 1);
    auto pv = [&](int x) {  // O_x += P_x V : 4 k-steps x (N=64, N=16)
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        umma_ts_w(elected, tb + 256 + x * 80, tb + x * 128 + ks * 8, desc_pack(v_lo + ks * 128, hi128), id64_mn, 1u);
        umma_ts_w(elected, tb + 256 + x * 80 + 64, tb + x * 128 + ks * 8, desc_pack(v16_lo + ks * 32, hi32), id16_mn, 1u);
      }
    };
    auto qk = [&](int x) {  // S_x = Q_x K^T : 4 + 1 k-steps, N = 64
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_ss_w(elected, tb + x * 128 + 64, desc_pack(a_lo + x * 1280 + 2 * k, hi128), desc_pack(b_lo + 2 * k, hi128), id64, k > 0 ? 1u : 0u);
      umma_ss_w(elected, tb + x * 128 + 64, desc_pack(a_lo + x * 1280 + 1024, hi32), desc_pack(b_lo + 512, hi32), id64, 1u);
    };
    __syncwarp();
    const long long t0 = clock64();
#pragma unroll 1
    for (int rep = 0; rep < 4; ++rep) {
      if constexpr (PATTERN == 0) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
          umma_ss_w(elected, tb + (r % ACC) * (N > 64 ? 256 : 96), desc_pack(a_lo + 2 * (r & 3), hi128), desc_pack(b_lo + 2 * (r & 3), hi128), id_n, 1u);
      } else if constexpr (PATTERN == 1) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
          umma_ts_w(elected, tb + (r % ACC) * 96, tb + 448 + (r & 3) * 8, desc_pack(v_lo + (r & 3) * 128, hi128), id_n_mn, 1u);
      } else if constexpr (PATTERN == 2) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          umma_ts_w(elected, tb + (r % ACC) * 96, tb + 448 + (r & 3) * 8, desc_pack(v_lo + (r & 3) * 128, hi128), id64_mn, 1u);
          umma_ts_w(elected, tb + (r % ACC) * 96 + 64, tb + 448 + (r & 3) * 8, desc_pack(v16_lo + (r & 3) * 32, hi32), id16_mn, 1u);
        }
      } else if constexpr (PATTERN == 3) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          pv(0);
          qk(0);
        }
      } else if constexpr (PATTERN == 4) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          pv(0);
          qk(0);
          pv(1);
          qk(1);
        }
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int x = 0; x < 2; ++x) {
              umma_ts_w(elected, tb + 256 + x * 80, tb + x * 128 + ks * 8, desc_pack(v_lo + ks * 128, hi128), id64_mn, 1u);
              umma_ts_w(elected, tb + 256 + x * 80 + 64, tb + x * 128 + ks * 8, desc_pack(v16_lo + ks * 32, hi32), id16_mn, 1u);
            }
          }
#pragma unroll
          for (int k = 0; k < 5; ++k) {
#pragma unroll
            for (int x = 0; x < 2; ++x) {
              if (k < 4)
                umma_ss_w(elected, tb + x * 128 + 64, desc_pack(a_lo + x * 1280 + 2 * k, hi128), desc_pack(b_lo + 2 * k, hi128), id64, k > 0 ? 1u : 0u);
              else
                umma_ss_w(elected, tb + x * 128 + 64, desc_pack(a_lo + x * 1280 + 1024, hi32), desc_pack(b_lo + 512, hi32), id64, 1u);
            }
          }
        }
      }
    }
    const long long t1 = clock64();
    umma_commit_w(elected, &bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (lane == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tb);
}


// One query tile's per-key-tile work with knobs, to find what makes the mixed sequence 4x slower than its parts:
//   PV_TS: P V step takes A from TMEM (1) or from smem (0);  P_COL: TMEM column of the P operand;  S_COL: column of the S
//   accumulator;  S_OVERWRITE: first S MMA has accumulate = 0;  O_COL: column of the O accumulator;  FINE: alternate MMA by MMA
template <int PV_TS, int P_COL, int S_COL, int S_OVERWRITE, int O_COL, int FINE>
__global__ void __launch_bounds__(128, 1) mix_bench(long long* out) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_ptr);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (warp == 0) {
    const uint32_t elected = elect_one();
    constexpr uint32_t hi128 = umma_desc_hi(1024, 2);
    constexpr uint32_t hi32 = umma_desc_hi(256, 6);
    const uint32_t a_lo = umma_desc_lo(smem_u32(smem), 16);
    const uint32_t b_lo = umma_desc_lo(smem_u32(smem + 40960), 16);
    const uint32_t v_lo = umma_desc_lo(smem_u32(smem + 40960 + 10240), 16) - (1u << 16) + ((8192u >> 4) << 16);
    const uint32_t v16_lo = umma_desc_lo(smem_u32(smem + 40960 + 18432), 16) - (1u << 16) + ((2048u >> 4) << 16);
    constexpr uint32_t id64 = idesc_of(64, 0), id64_mn = idesc_of(64, 1), id16_mn = idesc_of(16, 1);
    auto pv64 = [&](int ks) {
      if (PV_TS) umma_ts_w(elected, tb + O_COL, tb + P_COL + ks * 8, desc_pack(v_lo + ks * 128, hi128), id64_mn, 1u);
      else umma_ss_w(elected, tb + O_COL, desc_pack(a_lo + 2 * ks, hi128), desc_pack(v_lo + ks * 128, hi128), id64_mn, 1u);
    };
    auto pv16 = [&](int ks) {
      if (PV_TS) umma_ts_w(elected, tb + O_COL + 64, tb + P_COL + ks * 8, desc_pack(v16_lo + ks * 32, hi32), id16_mn, 1u);
      else umma_ss_w(elected, tb + O_COL + 64, desc_pack(a_lo + 2 * ks, hi128), desc_pack(v16_lo + ks * 32, hi32), id16_mn, 1u);
    };
    auto sk = [&](int k) {
      if (k < 4) umma_ss_w(elected, tb + S_COL, desc_pack(a_lo + 2 * k, hi128), desc_pack(b_lo + 2 * k, hi128), id64, (k > 0 || !S_OVERWRITE) ? 1u : 0u);
      else umma_ss_w(elected, tb + S_COL, desc_pack(a_lo + 1024, hi32), desc_pack(b_lo + 512, hi32), id64, 1u);
    };
    __syncwarp();
    const long long t0 = clock64();
#pragma unroll 1
    for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (FINE) {
          pv64(0); sk(0); pv16(0); sk(1); pv64(1); sk(2); pv16(1); sk(3); pv64(2); sk(4); pv16(2); pv64(3); pv16(3);
        } else {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            pv64(ks);
            pv16(ks);
          }
#pragma unroll
          for (int k = 0; k < 5; ++k) sk(k);
        }
      }
    }
    const long long t1 = clock64();
    umma_commit_w(elected, &bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (lane == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tb);
}


// Does a back-pressured tcgen05.mma issuer slow down the OTHER warps of its SM sub-partition?  Warp 0 (SMSP 0) issues
// MMAs back to back (MODE 1) or idles (MODE 0) or waits for every batch of 4 to complete before the next (MODE 2); warps
// 4 (SMSP 0) and 5 (SMSP 1) run the same exp2-style loop and report their elapsed cycles.
template <int MODE>
__global__ void __launch_bounds__(256, 1) smsp_bench(long long* out, float* sink) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    done = 0;
  }
  if (warp == 1) tmem_alloc<512>(&tmem_ptr);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (warp == 0 && MODE != 0) {
    const uint32_t elected = elect_one();
    constexpr uint32_t hi128 = umma_desc_hi(1024, 2);
    const uint32_t a_lo = umma_desc_lo(smem_u32(smem), 16);
    const uint32_t b_lo = umma_desc_lo(smem_u32(smem + 40960), 16);
    constexpr uint32_t id64 = idesc_of(64, 0);
    uint32_t phase = 0;
    long long n = 0;
    while (!done) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
        umma_ss_w(elected, tb + r * 64, desc_pack(a_lo + 2 * r, hi128), desc_pack(b_lo + 2 * r, hi128), id64, 1u);
      n += 4;
      if (MODE == 2) {
        umma_commit_w(elected, &bar);
        mbar_wait(&bar, phase);
        phase ^= 1;
      }
    }
    umma_commit_w(elected, &bar);
    mbar_wait(&bar, phase);
    if (lane == 0) out[2] = n;
  } else if (warp == 4 || warp == 5) {
    float a0 = 0.001f * lane, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, s = 0.f;
    __syncwarp();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 4000; ++it) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a0 = fast_exp2(fmaf(a0, 0.999f, -0.5f));
        a1 = fast_exp2(fmaf(a1, 0.999f, -0.5f));
        a2 = fast_exp2(fmaf(a2, 0.999f, -0.5f));
        a3 = fast_exp2(fmaf(a3, 0.999f, -0.5f));
        s += a0 + a1 + a2 + a3;
      }
    }
    const long long t1 = clock64();
    if (lane == 0) out[warp - 4] = t1 - t0;
    sink[threadIdx.x] = s;
    __syncwarp();
    if (warp == 4 && lane == 0) done = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tb);
}

static long long* d_out;

template <int PATTERN, int N, int ACC>
void run(const char* what, int mmas_per_rep) {
  cudaFuncSetAttribute(mma_bench<PATTERN, N, ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98 * 1024);
  long long best[2] = {1ll << 60, 1ll << 60};
  for (int pass = 0; pass < 3; ++pass) {
    mma_bench<PATTERN, N, ACC><<<1, 128, 98 * 1024>>>(d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%-44s CUDA error: %s\n", what, cudaGetErrorString(e));
      exit(1);
    }
    long long o[2];
    cudaMemcpy(o, d_out, sizeof(o), cudaMemcpyDeviceToHost);
    if (o[1] < best[1]) best[0] = o[0], best[1] = o[1];
  }
  const int total = 4 * mmas_per_rep;
  printf("%-44s N=%3d acc=%d | %4d MMAs: issue %6lld cyc (%5.1f/MMA)  complete %6lld cyc (%5.1f/MMA)\n", what, N, ACC, total, best[0],
         double(best[0]) / total, best[1], double(best[1]) / total);
}


template <int PV_TS, int P_COL, int S_COL, int S_OVERWRITE, int O_COL, int FINE>
void run_mix(const char* what) {
  cudaFuncSetAttribute(mix_bench<PV_TS, P_COL, S_COL, S_OVERWRITE, O_COL, FINE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98 * 1024);
  long long best[2] = {1ll << 60, 1ll << 60};
  for (int pass = 0; pass < 3; ++pass) {
    mix_bench<PV_TS, P_COL, S_COL, S_OVERWRITE, O_COL, FINE><<<1, 128, 98 * 1024>>>(d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%-60s CUDA error: %s\n", what, cudaGetErrorString(e));
      exit(1);
    }
    long long o[2];
    cudaMemcpy(o, d_out, sizeof(o), cudaMemcpyDeviceToHost);
    if (o[1] < best[1]) best[0] = o[0], best[1] = o[1];
  }
  printf("%-60s ts=%d P@%3d S@%3d ow=%d O@%3d fine=%d | per (PV,S) group of 13: issue %6.0f  complete %6.0f cyc\n", what, PV_TS, P_COL,
         S_COL, S_OVERWRITE, O_COL, FINE, best[0] / 16.0, best[1] / 16.0);
}


template <int MODE>
void run_smsp(const char* what) {
  cudaFuncSetAttribute(smsp_bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98 * 1024);
  float* sink;
  cudaMalloc(&sink, 4096);
  cudaMemset(d_out, 0, 32);
  smsp_bench<MODE><<<1, 256, 98 * 1024>>>(d_out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-40s CUDA error: %s\n", what, cudaGetErrorString(e));
    exit(1);
  }
  long long o[3];
  cudaMemcpy(o, d_out, sizeof(o), cudaMemcpyDeviceToHost);
  printf("%-58s exp loop (16000 MUFU/warp): warp 4 (same SMSP as issuer) %7lld cyc, warp 5 (other SMSP) %7lld cyc; MMAs issued %lld\n", what,
         o[0], o[1], o[2]);
  cudaFree(sink);
}

int main() {
  cudaMalloc(&d_out, 32);
  run_smsp<0>("issuer idle");
  run_smsp<1>("issuer back to back (blocked on the MMA queue)");
  run_smsp<2>("issuer: 4 MMAs, commit, sleep-wait, repeat");
  run_mix<1, 0, 64, 1, 256, 0>("kernel as is");
  run_mix<1, 0, 64, 0, 256, 0>("S never overwrites (accumulate=1 throughout)");
  run_mix<1, 0, 384, 1, 256, 0>("S accumulator far from P (col 384)");
  run_mix<1, 448, 64, 1, 256, 0>("P operand far from S (col 448)");
  run_mix<1, 448, 0, 0, 256, 0>("P far, S at 0, no overwrite");
  run_mix<0, 0, 64, 1, 256, 0>("P V in SS mode (A from smem)");
  run_mix<0, 0, 64, 0, 256, 0>("P V in SS mode, no overwrite");
  run_mix<1, 0, 64, 1, 256, 1>("fine-grained alternation PV / S");
  run_mix<1, 0, 64, 0, 256, 1>("fine-grained alternation, no overwrite");
  run_mix<1, 0, 64, 1, 128, 0>("O next to S (col 128)");
  run<0, 16, 1>("SS chain", 16);
  run<0, 16, 4>("SS round-robin accumulators", 16);
  run<0, 64, 1>("SS chain", 16);
  run<0, 64, 2>("SS round-robin accumulators", 16);
  run<0, 64, 4>("SS round-robin accumulators", 16);
  run<0, 128, 1>("SS chain", 16);
  run<0, 128, 2>("SS round-robin accumulators", 16);
  run<0, 256, 1>("SS chain", 16);
  run<1, 16, 1>("TS chain (B MN-major)", 16);
  run<1, 16, 4>("TS round-robin accumulators", 16);
  run<1, 64, 1>("TS chain (B MN-major)", 16);
  run<1, 64, 2>("TS round-robin accumulators", 16);
  run<1, 64, 4>("TS round-robin accumulators", 16);
  run<2, 64, 1>("TS 64+16 pairs, one O", 32);
  run<2, 64, 2>("TS 64+16 pairs, two O", 32);
  run<3, 64, 1>("kt64 sequence, one query tile (PV, S)", 13);
  run<4, 64, 1>("kt64 sequence, A then B", 26);
  run<5, 64, 1>("kt64 sequence, A/B interleaved per MMA", 26);
  return 0;
}
