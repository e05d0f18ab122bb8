// vsb200 -- bf16 TN GEMM, CTA-pair variant: tcgen05.mma.cta_group::2 (M = 256 across two SMs).
//
// Why: with one CTA per tile every SM pulls its own copy of the W tile through L2; at 128x192 the arithmetic
// intensity is 38 MAC/B and the L2->SM fabric (not the tensor pipe) bounds the kernel.  A CTA pair shares the W tile:
// each CTA stages 128 rows of A and HALF of the W tile (BN/2 rows); the leader's single thread issues
// M=256 x N=BN x K=16 MMAs that read both CTAs' shared memory and write both CTAs' TMEM -> 55 (BN=192) / 64 (BN=256)
// MAC per L2 byte and half the shared-memory operand traffic per SM.
//
// Protocol (cluster of 2, rank 0 = leader):
//   full[s]    lives in the LEADER; the leader's producer arms it with the bytes of BOTH CTAs, both producers'
//              TMA loads (.cta_group::2) complete_tx on it.
//   empty[s]   one per CTA; the leader's MMA thread releases a stage in both CTAs with a multicast tcgen05.commit.
//   tfull[a]   one per CTA (multicast commit), each CTA's epilogue drains its own 128 accumulator rows.
//   tempty[a]  lives in the LEADER; 8 arrivals (4 epilogue warps x 2 CTAs, the peer's arrive remotely).
#include "vsb_common.cuh"
#include "vsb_host.h"

namespace vsb {

constexpr int k2BM = 128;  // rows per CTA (256 per pair)
constexpr int k2BK = 64;
constexpr int k2Threads = 384;  // 4 control warps + 8 epilogue warps (2 per TMEM lane quarter, split by columns)

template <int BN>
struct Gemm2Cfg {
  static constexpr int kABytes = k2BM * k2BK * 2;
  static constexpr int kBBytes = (BN / 2) * k2BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kCBytes = k2BM * BN * 2;
  static constexpr int kStages = (BN >= 256) ? 5 : 6;
  static constexpr int kTmemCols = 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + kCBytes + 1024 + 256 + 512;  // + bias tile
};

__device__ __forceinline__ float gelu_tanh_f2(float x) {
  // torch GELU(approximate='tanh') = 0.5 x (1 + tanh(u)), u = sqrt(2/pi) (x + 0.044715 x^3), rewritten as
  //   x * sigmoid(2u) = x / (1 + 2^(-2 u log2 e)):
  // 5 ALU + 2 MUFU per element instead of ~14 + 2 (the fc1 epilogue, not its mainloop, set the pace of that GEMM:
  // tensor pipe 66 % active against 82-94 % for the other shapes, profiles/r01_gemm2_ncu_full_final.txt), and no
  // cancellation in the negative tail.  |difference| to the tanh form <= 5e-7; 99.6 % of all bf16 inputs in [-12, 12]
  // round to the same bf16 output as torch's fp32 evaluation, the rest differ below 1e-6 absolute (x < -3).
  const float kC0 = -2.f * 0.7978845608028654f * 1.4426950408889634f, kC1 = kC0 * 0.044715f;
  const float w = x * fmaf(x * x, kC1, kC0);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fast_exp2(w) + 1.f));
  return x * r;
}

// ACT == 2: out = resid + g, g = bf16(gate * y) (gate_row >= 0, per-frame t/t0 select) or g = y (plain residual add),
// y = bf16(acc + bias): the eager chain of open_sora_transformer_3d.py:219-228 / :240 / :270-284 in the epilogue.
struct EpiArgs {
  const bf16* resid;       // [M, N], may alias the output
  const bf16* mod;         // [2, B, 6, N] or null
  const uint8_t* x_mask;   // [B, T] or null
  int gate_row, B, T, S;
};

template <int BN, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2Threads, 1)
gemm2_bf16_tn_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                     const __grid_constant__ CUtensorMap tm_c, const __grid_constant__ CUtensorMap tm_r,
                     const bf16* __restrict__ bias, int M, int N, int K, const EpiArgs ep) {
  using Cfg = Gemm2Cfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* smem_ab = smem;
  unsigned char* smem_c = smem + kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + Cfg::kCBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tfull = bars + 2 * kStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* rfull = tempty + 2;  // ACT == 2: the residual tile has landed in the output staging buffer
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(rfull + 1);
  bf16* sbias = reinterpret_cast<bf16*>(reinterpret_cast<unsigned char*>(bars) + 256);  // [BN] bias of the current tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = (cta == 0);
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int tiles_m = (M + 2 * k2BM - 1) / (2 * k2BM), tiles_n = (N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + k2BK - 1) / k2BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_c);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 16);  // 8 epilogue warps x 2 CTAs
    }
    mbar_init(rfull, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm<Cfg::kTmemCols>(tmem_ptr);
  tc_fence_before();
  cluster_sync_all();  // barriers of BOTH CTAs are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; whole warp converged, one elected lane issues) ==============
    {
      const uint32_t elected = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m0 = (tile / tiles_n) * (2 * k2BM) + int(cta) * k2BM;
        const int n0 = (tile % tiles_n) * BN + int(cta) * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* sa = smem_ab + stage * Cfg::kStageBytes;
          if (leader) mbar_arrive_expect_tx_w(elected, &full[stage], 2 * Cfg::kStageBytes);
          const uint32_t fb = mapa_u32(&full[stage], 0);
          tma_load_2d_2sm_to_w(elected, &tm_a, fb, sa, kb * k2BK, m0);
          tma_load_2d_2sm_to_w(elected, &tm_w, fb, sa + Cfg::kABytes, kb * k2BK, n0);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only; whole warp converged, one elected lane issues) ==========
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * k2BM, BN, 0, 0);
      constexpr uint32_t dhi = umma_desc_hi(1024, 2);  // K-major SWIZZLE_128B: 8-row groups 1024 B apart
      const uint32_t elected = elect_one();
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t lo0 = umma_desc_lo(smem_u32(smem_ab), 16);  // stage 0, A tile; +2 per K=16 step (32 B)
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tb + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t la = lo0 + stage * (Cfg::kStageBytes >> 4);
          const uint32_t lb = la + (Cfg::kABytes >> 4);
#pragma unroll
          for (int k = 0; k < k2BK / 16; ++k)
            umma_ss_2sm_w(elected, d_tmem, desc_pack(la + 2 * k, dhi), desc_pack(lb + 2 * k, dhi), idesc,
                          (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_2sm_mcast_w(elected, &empty[stage], 0x3);  // frees the stage in both CTAs
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm_mcast_w(elected, &tfull[acc], 0x3);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows; 8 warps: lane quarter x column half) ==========
    const int ew = (warp - 4) & 3;
    const int half = (warp - 4) >> 2;
    constexpr int kChunks = BN / 64;
    const int c_begin = half ? (kChunks + 1) / 2 : 0, c_end = half ? kChunks : (kChunks + 1) / 2;
    const int row = ew * 32 + lane;
    const int et = threadIdx.x - 128;  // 0..255 among the epilogue threads
    int acc = 0;
    uint32_t acc_phase = 0;
    // ACT == 2: the residual tile is TMA-prefetched INTO the output staging buffer (same 128B-swizzled chunks the result
    // is stored from) as soon as the previous tile's store has released it -- while this tile's MMAs are still running --
    // and every thread reads x from the very 16 bytes it then overwrites with x + g.  The first version read the
    // residual per thread from global memory (32 rows = 32 sectors per load instruction): the epilogue, not the
    // mainloop, paced the GEMM (+41 ms per step for -28 ms of elementwise passes).
    uint32_t r_phase = 0;
    auto prefetch_resid = [&](int tile) {
      const int m0 = (tile / tiles_n) * (2 * k2BM) + int(cta) * k2BM;
      const int n0 = (tile % tiles_n) * BN;
      int nch = 0;
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c)
        if (n0 + c * 64 < N) ++nch;
      mbar_arrive_expect_tx(rfull, nch * (k2BM * 128));
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c)
        if (n0 + c * 64 < N) tma_load_2d(&tm_r, rfull, smem_c + c * (k2BM * 128), n0 + c * 64, m0);
    };
    if (ACT == 2 && threadIdx.x == 128 && pair < num_tiles) prefetch_resid(pair);
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m0 = (tile / tiles_n) * (2 * k2BM) + int(cta) * k2BM;
      const int n0 = (tile % tiles_n) * BN;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if (ACT != 2 && threadIdx.x == 128) tma_store_wait_read0();
      if (ACT == 2) {
        mbar_wait(rfull, r_phase);
        r_phase ^= 1;
      }
      if (et < BN / 8) {  // stage the tile's bias once (every thread needs all BN values)
        uint4 bv = make_uint4(0, 0, 0, 0);
        if (bias != nullptr && n0 + et * 8 < N) bv = __ldg(reinterpret_cast<const uint4*>(bias + n0 + et * 8));
        *reinterpret_cast<uint4*>(sbias + et * 8) = bv;
      }
      named_bar_sync(1, 256);
      const uint32_t t_row = tmem_base + (uint32_t(ew * 32) << 16) + acc * BN;
      const long long grow = (long long)m0 + row;
      const bf16* gate_vec = nullptr;
      if (ACT == 2 && grow < M) {
        if (ep.gate_row >= 0) {
          const long long bt = grow / ep.S;
          const int bb = int(bt / ep.T);
          const int sel = (ep.x_mask != nullptr && ep.x_mask[bt] == 0) ? 1 : 0;
          gate_vec = ep.mod + ((size_t)(sel * ep.B + bb) * 6 + ep.gate_row) * N;
        }
      }
#pragma unroll 1
      for (int c = c_begin; c < c_end; ++c) {
        uint32_t r0[32], r1[32];
        tmem_ld32(t_row + c * 64, r0);
        tmem_ld32(t_row + c * 64 + 32, r1);
        tmem_wait_ld();
        if (c == c_end - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(mapa_u32(&tempty[acc], 0));
        }
        unsigned char* crow = smem_c + c * (k2BM * 128) + row * 128;
        const int ncol0 = n0 + c * 64;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v[8];
          float xr[8], gt[8];
          if (ACT == 2) {
            const int gc = ncol0 + j * 8;
            // x: the residual chunk sits where this thread's result goes (rows past M were zero-filled by the TMA)
            const uint4 ux = *reinterpret_cast<const uint4*>(crow + ((j ^ (row & 7)) << 4));
            uint4 ug = make_uint4(0, 0, 0, 0);
            if (gate_vec != nullptr && gc < N) ug = __ldg(reinterpret_cast<const uint4*>(gate_vec + gc));
            const uint32_t wx[4] = {ux.x, ux.y, ux.z, ux.w}, wg[4] = {ug.x, ug.y, ug.z, ug.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 fx = unpack_bf16x2(wx[q]), fg = unpack_bf16x2(wg[q]);
              xr[2 * q] = fx.x;
              xr[2 * q + 1] = fx.y;
              gt[2 * q] = fg.x;
              gt[2 * q + 1] = fg.y;
            }
          }
          const uint4 bq = *reinterpret_cast<const uint4*>(sbias + c * 64 + j * 8);  // smem broadcast
          const uint32_t bw[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int col = j * 8 + e;
            const float a = __uint_as_float(col < 32 ? r0[col] : r1[col - 32]);
            const float2 b2 = unpack_bf16x2(bw[e >> 1]);
            float x = a + ((e & 1) ? b2.y : b2.x);
            if (ACT == 1) x = gelu_tanh_f2(rbf(x));
            if (ACT == 2) {
              float y = rbf(x);                                  // the Linear's bf16 output
              if (gate_vec != nullptr) y = rbf(gt[e] * y);       // bf16(gate * y)
              x = xr[e] + y;                                     // residual add (rounded by the pack below)
            }
            v[e] = x;
          }
          uint4 u;
          u.x = pack_bf16x2(v[0], v[1]);
          u.y = pack_bf16x2(v[2], v[3]);
          u.z = pack_bf16x2(v[4], v[5]);
          u.w = pack_bf16x2(v[6], v[7]);
          *reinterpret_cast<uint4*>(crow + ((j ^ (row & 7)) << 4)) = u;
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 256);
      if (threadIdx.x == 128) {
        if (m0 < M) {
#pragma unroll 1
          for (int c = 0; c < BN / 64; ++c)
            if (n0 + c * 64 < N) tma_store_2d(&tm_c, smem_c + c * (k2BM * 128), n0 + c * 64, m0);
        }
        tma_store_commit();
        if (ACT == 2) {
          tma_store_wait_read0();  // the staging buffer is free again: fetch the next tile's residual into it
          if (tile + num_pairs < num_tiles) prefetch_resid(tile + num_pairs);
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (threadIdx.x == 128) tma_store_wait0();
  }

  tc_fence_before();
  cluster_sync_all();  // nobody leaves while the peer may still read my smem / signal my barriers
  if (warp == 2) tmem_dealloc_2sm<Cfg::kTmemCols>(tmem_base);
}

template <int BN, int ACT>
static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tc, const bf16* bias, int M,
                        int N, int K, cudaStream_t st, const EpiArgs& ep = EpiArgs{nullptr, nullptr, nullptr, -1, 1, 1, 1},
                        const CUtensorMap* tr = nullptr) {
  using Cfg = Gemm2Cfg<BN>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_bf16_tn_kernel<BN, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return fail(VSB_ERR_CUDA, "gemm2: smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  const int tiles = ((M + 2 * k2BM - 1) / (2 * k2BM)) * ((N + BN - 1) / BN);
  int pairs = num_sms() / 2;
  if (pairs > tiles) pairs = tiles;
  gemm2_bf16_tn_kernel<BN, ACT><<<pairs * 2, k2Threads, Cfg::kSmemBytes, st>>>(ta, tw, tc, tr ? *tr : tc, bias, M, N, K, ep);
  return check_launch("gemm2_bf16_tn");
}

// Called by vsb_gemm_bias_act / vsb_gemm_bias_residual (gemm_tcgen05.cu).  Returns 1 if this variant does not apply.
// act: 0 none, 1 gelu, 2 fused residual (ep != nullptr).
int gemm2_dispatch(const void* A, const void* W, const void* bias, void* out, int M, int N, int K, int act,
                   cudaStream_t st, const EpiArgs* ep) {
  int BN;
  if (N % 256 == 0)
    BN = 256;
  else if (N % 192 == 0)
    BN = 192;
  else
    return 1;
  CUtensorMap ta, tw, tc;
  unsigned long long da[2] = {(unsigned long long)K, (unsigned long long)M};
  unsigned long long sa[1] = {(unsigned long long)K * 2};
  unsigned ba[2] = {k2BK, k2BM};
  int rc = make_tmap_bf16(&ta, A, 2, da, sa, ba, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  unsigned long long dw[2] = {(unsigned long long)K, (unsigned long long)N};
  unsigned bw[2] = {k2BK, (unsigned)(BN / 2)};
  rc = make_tmap_bf16(&tw, W, 2, dw, sa, bw, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  unsigned long long dc[2] = {(unsigned long long)N, (unsigned long long)M};
  unsigned long long sc[1] = {(unsigned long long)N * 2};
  unsigned bc[2] = {64, k2BM};
  rc = make_tmap_bf16(&tc, out, 2, dc, sc, bc, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  const bf16* b = (const bf16*)bias;
  if (act == 2) {
    CUtensorMap tr;  // the residual, tiled exactly like the output
    rc = make_tmap_bf16(&tr, ep->resid, 2, dc, sc, bc, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if (BN == 256) return launch_gemm2<256, 2>(ta, tw, tc, b, M, N, K, st, *ep, &tr);
    return launch_gemm2<192, 2>(ta, tw, tc, b, M, N, K, st, *ep, &tr);
  }
  if (BN == 256)
    return act ? launch_gemm2<256, 1>(ta, tw, tc, b, M, N, K, st) : launch_gemm2<256, 0>(ta, tw, tc, b, M, N, K, st);
  return act ? launch_gemm2<192, 1>(ta, tw, tc, b, M, N, K, st) : launch_gemm2<192, 0>(ta, tw, tc, b, M, N, K, st);
}

int gemm2_residual(const void* A, const void* W, const void* bias, const void* resid, void* out, const void* mod,
                   const unsigned char* x_mask, int gate_row, int M, int N, int K, int B, int T, int S,
                   cudaStream_t st) {
  EpiArgs ep{(const bf16*)resid, (const bf16*)mod, x_mask, gate_row, B, T, S};
  return gemm2_dispatch(A, W, bias, out, M, N, K, 2, st, &ep);
}

}  // namespace vsb
