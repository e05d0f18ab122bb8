// vsb200 -- bf16 TN GEMM on the 5th-gen tensor cores: out[M,N] = act(A[M,K] @ W[N,K]^T + bias).
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer   (A tile 128x64, W tile BNx64, SWIZZLE_128B, kStages-deep mbarrier ring)
//   warp 1      MMA issuer     (one elected thread: tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, K=16)
//   warp 2      TMEM allocator (2 accumulator buffers of BN fp32 columns -> epilogue overlaps the next mainloop)
//   warps 4..7  epilogue       (tcgen05.ld 32x32b -> +bias -> [bf16 round -> tanh-GELU] -> bf16 -> swizzled smem
//                               -> TMA store)
// Tiles are walked N-fastest so that the 128-row A panel stays L2-resident across its N tiles and W (<= 10.6 MB)
// stays L2-resident for the whole GEMM.  TMA zero-fills K/M/N tails on load and clips them on store, so any
// M, N % 8 == 0, K % 8 == 0 is legal.
#include "vsb_common.cuh"
#include "vsb_host.h"

namespace vsb {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 256;

template <int BN>
struct GemmCfg {
  static constexpr int kStageBytes = (kBM + BN) * kBK * 2;
  static constexpr int kCBytes = kBM * BN * 2;
  static constexpr int kStages = (BN >= 256) ? 3 : 4;
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  // + 1024 for manual alignment, + 256 for barriers / tmem pointer
  static constexpr int kSmemBytes = kStages * kStageBytes + kCBytes + 1024 + 256;
};

__device__ __forceinline__ float gelu_tanh_f(float x) {
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

template <int BN, int ACT>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                    const __grid_constant__ CUtensorMap tm_c, const bf16* __restrict__ bias, int M, int N, int K) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* smem_ab = smem;                                   // [kStages][A 16 KB | W BN*128 B]
  unsigned char* smem_c = smem + kStages * Cfg::kStageBytes;       // [BN/64][128 rows][128 B] swizzled
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + Cfg::kCBytes);
  uint64_t* full = bars;                  // [kStages]
  uint64_t* empty = bars + kStages;       // [kStages]
  uint64_t* tfull = bars + 2 * kStages;   // [2]
  uint64_t* tempty = tfull + 2;           // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (M + kBM - 1) / kBM, tiles_n = (N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + kBK - 1) / kBK;

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
      mbar_init(&tempty[i], 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (whole warp converged, one elected lane issues) =====================
    const uint32_t elected = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * kBM, n0 = (tile % tiles_n) * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        unsigned char* sa = smem_ab + stage * Cfg::kStageBytes;
        mbar_arrive_expect_tx_w(elected, &full[stage], Cfg::kStageBytes);
        tma_load_2d_w(elected, &tm_a, &full[stage], sa, kb * kBK, m0);
        tma_load_2d_w(elected, &tm_w, &full[stage], sa + kBM * kBK * 2, kb * kBK, n0);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
    constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, 0, 0);
    constexpr uint32_t dhi = umma_desc_hi(1024, 2);  // K-major SWIZZLE_128B: 8-row groups 1024 B apart
    const uint32_t elected = elect_one();
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t lo0 = umma_desc_lo(smem_u32(smem_ab), 16);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tb + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t la = lo0 + stage * (Cfg::kStageBytes >> 4);
        const uint32_t lb = la + ((kBM * kBK * 2) >> 4);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k)  // +32 B (2 x 16-byte units) per K=16 step inside the swizzle atom
          umma_ss_w(elected, d_tmem, desc_pack(la + 2 * k, dhi), desc_pack(lb + 2 * k, dhi), idesc,
                    (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit_w(elected, &empty[stage]);  // smem slot reusable once these MMAs retire
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit_w(elected, &tfull[acc]);  // accumulator complete
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;            // == warp % 4 -> TMEM lanes 32*ew .. 32*ew+31
    const int row = ew * 32 + lane;     // row inside the 128-row tile
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * kBM, n0 = (tile % tiles_n) * BN;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      // the previous tile's TMA store must have finished READING the staging buffer
      if (threadIdx.x == 128) tma_store_wait_read0();
      named_bar_sync(1, 128);
      const uint32_t t_row = tmem_base + (uint32_t(ew * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        uint32_t r0[32], r1[32];
        tmem_ld32(t_row + c * 64, r0);
        tmem_ld32(t_row + c * 64 + 32, r1);
        tmem_wait_ld();
        if (c == BN / 64 - 1) {  // all TMEM reads of this accumulator are done -> hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        unsigned char* crow = smem_c + c * (kBM * 128) + row * 128;
        const int ncol0 = n0 + c * 64;
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // 8 columns (16 bytes) per step
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int col = j * 8 + e;
            const float a = __uint_as_float(col < 32 ? r0[col] : r1[col - 32]);
            const int gcol = ncol0 + col;
            const float b = (bias != nullptr && gcol < N) ? e_to_float(__ldg(bias + gcol)) : 0.f;
            float x = a + b;
            if (ACT == 1) x = gelu_tanh_f(rbf(x));
            v[e] = x;
          }
          uint4 u;
          u.x = pack_bf16x2(v[0], v[1]);
          u.y = pack_bf16x2(v[2], v[3]);
          u.z = pack_bf16x2(v[4], v[5]);
          u.w = pack_bf16x2(v[6], v[7]);
          *reinterpret_cast<uint4*>(crow + ((j ^ (row & 7)) << 4)) = u;  // 128B swizzle, conflict-free
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (threadIdx.x == 128) {
#pragma unroll 1
        for (int c = 0; c < BN / 64; ++c)
          if (n0 + c * 64 < N) tma_store_2d(&tm_c, smem_c + c * (kBM * 128), n0 + c * 64, m0);
        tma_store_commit();
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (threadIdx.x == 128) tma_store_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

template <int BN, int ACT>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tc, const bf16* bias, int M,
                       int N, int K, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return fail(VSB_ERR_CUDA, "gemm: smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  const int tiles = ((M + kBM - 1) / kBM) * ((N + BN - 1) / BN);
  int grid = num_sms();
  if (grid > tiles) grid = tiles;
  gemm_bf16_tn_kernel<BN, ACT><<<grid, kGemmThreads, Cfg::kSmemBytes, st>>>(ta, tw, tc, bias, M, N, K);
  return check_launch("gemm_bf16_tn");
}

struct EpiArgs;
int gemm2_dispatch(const void* A, const void* W, const void* bias, void* out, int M, int N, int K, int act,
                   cudaStream_t st, const EpiArgs* ep);  // gemm_tcgen05_2sm.cu
int gemm2_residual(const void* A, const void* W, const void* bias, const void* resid, void* out, const void* mod,
                   const unsigned char* x_mask, int gate_row, int M, int N, int K, int B, int T, int S,
                   cudaStream_t st);

}  // namespace vsb

using namespace vsb;

extern "C" int VSB_API(vsb_gemm_bias_act)(const vsb_bf16* A, const vsb_bf16* W, const vsb_bf16* bias, vsb_bf16* out, int M,
                                 int N, int K, int act, void* stream) {
  if (!A || !W || !out || M <= 0 || N <= 0 || K <= 0) return fail(VSB_ERR_INVALID, "gemm: bad args");
  if (K % 8 || N % 8 || !aligned16(A) || !aligned16(W) || !aligned16(out))
    return fail(VSB_ERR_UNSUPPORTED, "gemm: need K %% 8 == 0, N %% 8 == 0, 16B-aligned pointers (M=%d N=%d K=%d)", M, N, K);
  if (act != 0 && act != 1) return fail(VSB_ERR_INVALID, "gemm: act=%d", act);
  if (g_opt_gemm_2sm && M >= 1024) {
    const int rc2 = gemm2_dispatch(A, W, bias, out, M, N, K, act, (cudaStream_t)stream, nullptr);
    if (rc2 <= 0) return rc2;  // 1 = not applicable -> single-CTA kernel below
  }
  // tile width: 192 divides every STDiT3 width (1152, 2304, 3456, 4608); narrow outputs use 64/128
  const int BN = (N % 192 == 0) ? 192 : (N >= 256 && N % 256 == 0) ? 256 : (N >= 192 ? 192 : (N > 64 ? 128 : 64));
  CUtensorMap ta, tw, tc;
  unsigned long long da[2] = {(unsigned long long)K, (unsigned long long)M};
  unsigned long long sa[1] = {(unsigned long long)K * 2};
  unsigned ba[2] = {kBK, kBM};
  int rc = make_tmap_bf16(&ta, A, 2, da, sa, ba, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  unsigned long long dw[2] = {(unsigned long long)K, (unsigned long long)N};
  unsigned bw[2] = {kBK, (unsigned)BN};
  rc = make_tmap_bf16(&tw, W, 2, dw, sa, bw, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  unsigned long long dc[2] = {(unsigned long long)N, (unsigned long long)M};
  unsigned long long sc[1] = {(unsigned long long)N * 2};
  unsigned bc[2] = {64, kBM};
  rc = make_tmap_bf16(&tc, out, 2, dc, sc, bc, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bf16* b = (const bf16*)bias;
#define VSB_GEMM_CASE(bn)                                                         \
  case bn:                                                                        \
    return act ? launch_gemm<bn, 1>(ta, tw, tc, b, M, N, K, st) : launch_gemm<bn, 0>(ta, tw, tc, b, M, N, K, st);
  switch (BN) {
    VSB_GEMM_CASE(64)
    VSB_GEMM_CASE(128)
    VSB_GEMM_CASE(192)
    VSB_GEMM_CASE(256)
  }
#undef VSB_GEMM_CASE
  return fail(VSB_ERR_UNSUPPORTED, "gemm: no tile config");
}

// Fused epilogue: out = resid + [gate *] (A @ W^T + bias), see include/vsb200.h.  Returns 1 (nothing launched) when
// the CTA-pair kernel does not take this shape; the caller then runs vsb_gemm_bias_act + vsb_gate_residual.
extern "C" int VSB_API(vsb_gemm_bias_residual)(const vsb_bf16* A, const vsb_bf16* W, const vsb_bf16* bias, const vsb_bf16* resid,
                                      vsb_bf16* out, const vsb_bf16* mod, const uint8_t* x_mask, int gate_row, int M,
                                      int N, int K, int B, int T, int S, void* stream) {
  if (!A || !W || !resid || !out || M <= 0 || N <= 0 || K <= 0) return fail(VSB_ERR_INVALID, "gemm_residual: bad args");
  if (K % 8 || N % 8 || !aligned16(A) || !aligned16(W) || !aligned16(out) || !aligned16(resid))
    return fail(VSB_ERR_UNSUPPORTED, "gemm_residual: need K %% 8 == 0, N %% 8 == 0, 16B-aligned pointers");
  if (gate_row >= 0) {
    if (!mod || gate_row > 5 || B <= 0 || T <= 0 || S <= 0 || (long long)B * T * S != M || !aligned16(mod))
      return fail(VSB_ERR_INVALID, "gemm_residual: gate needs mod[2,B,6,N] and B*T*S == M");
  }
  if (!g_opt_gemm_2sm || M < 1024) return 1;
  return gemm2_residual(A, W, bias, resid, out, mod, x_mask, gate_row, M, N, K, B, T, S, (cudaStream_t)stream);
}
