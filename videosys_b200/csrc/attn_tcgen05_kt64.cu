// vsb200 -- flash attention forward on tcgen05, 64-key tiles with a double-buffered S (attn_variant = 2).
//
// Why a second schedule.  At head_dim 72 the kernel is bound by the exp2 (MUFU, 16 lanes/clk/SM) of the softmax more
// than by the tensor pipe: a pair of 128x128 score tiles costs >= 2048 MUFU cycles against ~1430 tensor cycles.  The 128-key kernel
// (attn_tcgen05.cu) keeps P aliased on the only S buffer of a query tile, so S(j+1) cannot start before PV(j) and each
// softmax warpgroup sits idle for a full PV + S round trip per key tile (measured: 1250 idle cycles of a 3750-cycle
// period).  Here every query tile owns TWO 64-column S buffers in TMEM:
//
//     softmax WG x, tile t   : wait S_x(t) in buf t&1 -> max / exp2 / bf16 P_x(t) back into the head of buf t&1
//     MMA warp, iteration t  : wait P_x(t) -> O_x += P_x(t) V(t) -> S_x(t+2) = Q_x K(t+2)^T into buf t&1
//
// so while warpgroup x exponentiates tile t+1 (buf (t+1)&1, filled one iteration earlier) the tensor pipe retires
// PV_x(t) and S_x(t+2): the softmax warps do not wait on the tensor core as long as it keeps up.  Measured costs per
// 64-key step of both query tiles (tools/mma_microbench.cu, per-warp clock64 trace): 868 tensor cycles (the SS MMAs of
// S are shared-memory-bound at N = 64: 50.7 cycles each, not 32; the TS 64+16 pair of P V 47.5) against ~1150 MUFU
// cycles (two softmax warps per SM sub-partition: ~9 cycles per MUFU.EX2 warp instruction, F2FP packs included).
// Both warpgroups exponentiate concurrently.
//
//   warps 0..3    MMA issuers, one per SM sub-partition: warp (x, parity) owns S buffer `parity` of query tile x, i.e.
//                 PV_x(t) and S_x(t+2) for t = parity (mod 2).  Four issuers instead of one because (per-warp clock64
//                 trace + tools/mma_microbench.cu): (1) a tcgen05.mma that finds the tensor queue full (4-5 MMAs deep)
//                 stalls in the issue stage and slows the OTHER warps of its sub-partition -- next to a single issuer
//                 two softmax warps ran 1.65x slower than the other six and set the pace; (2) every mbarrier wait costs
//                 >= 90 cycles even when already complete, and one warp doing 5 of them per key tile between blocked
//                 issue bursts left the tensor pipe idle half of the time.  The four issuers need no ordering among
//                 themselves: S buffers are private, O_x is only ever accumulated into (the softmax warps zero it).
//                 Warp 0 also owns the TMEM allocation: S buffers at columns x*128 + buf*64, O_x at 256 + x*80.
//   warps 4..7    softmax warpgroup A (thread = query row), warps 8..11 warpgroup B
//   warp 12       TMA producer: Q tiles once, K/V 64-key tiles through a 6-deep mbarrier ring (K leads V by 2 tiles)
//
// Lazy rescale of O needs O quiescent: PV_x(t-1) may still be in flight when tile t finds a much larger max, so that
// (rare) path first waits for the commit the MMA warp posts after PV_x(t-1) [+ S_x(t+1)] on s_full[x][(t+1)&1].
// A CTA whose second query tile starts past nq (the ragged tail of a sequence) runs warpgroup A only.
#include "attn_params.cuh"

namespace vsb {

constexpr int kT64Threads = 416;  // 4 MMA issuer warps, 8 softmax warps, 1 TMA producer warp
constexpr int kT64Stages = 6;      // Q tiles staged in shared memory (SS-mode S)
constexpr int kT64StagesQT = 8;    // Q resident in TMEM (TS-mode S): its 40 KB of smem become two more K/V stages
constexpr int kQA = 128 * 128;  // Q: 128 rows x 64 bf16, SWIZZLE_128B
constexpr int kQB = 128 * 32;   // Q: 128 rows x 16 bf16, SWIZZLE_32B
constexpr int kQT = kQA + kQB;
constexpr int kKA = 64 * 128;   // K / V: 64 keys x 64 bf16, SWIZZLE_128B
constexpr int kKB = 64 * 32;    // K / V: 64 keys x 16 bf16, SWIZZLE_32B
constexpr int kKT = kKA + kKB;
constexpr int kStage = 2 * kKT;  // K_A | K_B | V_A | V_B
constexpr int kT64Smem = 2 * kQT + kT64Stages * kStage + 1024 + 512;
constexpr int kT64SmemQT = kT64StagesQT * kStage + 1024 + 512;

__host__ __device__ constexpr uint32_t c_s(int x, int buf) { return uint32_t(x) * 128u + uint32_t(buf) * 64u; }
__host__ __device__ constexpr uint32_t c_o(int x) { return 256u + uint32_t(x) * 80u; }
__host__ __device__ constexpr uint32_t c_q(int x) { return 416u + uint32_t(x) * 40u; }  // Q_x as bf16 pairs: 80 / 2 columns

// kSumMMA (head_dim 72 only): the softmax row sum comes out of the tensor core.  The V tile's zero padding column d = 72
// is overwritten with ones by the issuer warp, so O[:, 72] accumulates sum_j bf16(p_j) -- exactly the weights that
// multiply V -- and the 64 FADDs per thread per key tile (a quarter of the softmax warps' FMA-pipe work, which becomes
// the co-limiter once part of the exp2 moves off the MUFU pipe) disappear.
template <int D, int kPoly, bool kQTm, bool kSumMMA>
__global__ void __launch_bounds__(kT64Threads, 1)
attn_flash_kt64_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_qb,
                       const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_kb,
                       const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_vb,
                       const __grid_constant__ AttnParams p) {
  constexpr bool kHasB = (D > 64);
  static_assert(!kSumMMA || D == 72, "the ones column lives in the padding of head_dim 72");
  constexpr int kQTx = kHasB ? kQT : kQA;
  constexpr int kKTx = kHasB ? kKT : kKA;
  constexpr int ST = kQTm ? kT64StagesQT : kT64Stages;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;                            // [2][Q_A | Q_B] (absent when Q lives in TMEM)
  unsigned char* sKV = smem + (kQTm ? 0 : 2 * kQT);    // [stages][K_A | K_B | V_A | V_B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + ST * kStage);
  uint64_t* q_full = bars;              // [2]: smem Q: [0] armed by the TMA; TMEM Q: [x] gets 4 softmax-warp arrivals
  uint64_t* k_full = bars + 2;          // [ST]
  uint64_t* v_full = k_full + ST;       // [ST]
  uint64_t* kv_empty = v_full + ST;     // [ST]
  uint64_t* s_full = kv_empty + ST;     // [x][buf]
  uint64_t* p_full = s_full + 4;        // [x][buf]
  uint64_t* o_full = p_full + 4;        // [x]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
  const int kv_len = p.has_lens ? p.lens[b] : p.nk;
  const int n_tiles = (kv_len + 63) / 64;
  const int nx = (q0 + 128 < p.nq) ? 2 : 1;  // query tiles of this CTA that hold real rows

  if (warp == 12 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    if (kHasB) {
      tma_prefetch_desc(&tm_qb);
      tma_prefetch_desc(&tm_kb);
      tma_prefetch_desc(&tm_vb);
    }
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&q_full[0], kQTm ? 4 : 1);
    mbar_init(&q_full[1], 4);
    for (int i = 0; i < ST; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], nx);  // one commit per active query tile (after its P V of the tile)
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);  // one arrival per softmax warp
    }
    mbar_init(&o_full[0], 2);  // both issuers of the query tile
    mbar_init(&o_full[1], 2);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 12) {
    // =============================== TMA producer ===============================
    const uint32_t elected = elect_one();
    if (!kQTm) {
      mbar_arrive_expect_tx_w(elected, q_full, nx * kQTx);
      for (int x = 0; x < nx; ++x) {
        tma_load_4d_w(elected, &tm_q, q_full, sQ + x * kQT, 0, h, q0 + x * 128, b);
        if (kHasB) tma_load_4d_w(elected, &tm_qb, q_full, sQ + x * kQT + kQA, 64, h, q0 + x * 128, b);
      }
    }
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j % ST;
      const uint32_t ph = (j / ST) & 1;
      mbar_wait(&kv_empty[s], ph ^ 1);
      unsigned char* st = sKV + s * kStage;
      mbar_arrive_expect_tx_w(elected, &k_full[s], kKTx);
      tma_load_4d_w(elected, &tm_k, &k_full[s], st, 0, h, j * 64, b);
      if (kHasB) tma_load_4d_w(elected, &tm_kb, &k_full[s], st + kKA, 64, h, j * 64, b);
      mbar_arrive_expect_tx_w(elected, &v_full[s], kKTx);
      tma_load_4d_w(elected, &tm_v, &v_full[s], st + kKT, 0, h, j * 64, b);
      if (kHasB) tma_load_4d_w(elected, &tm_vb, &v_full[s], st + kKT + kKA, 64, h, j * 64, b);
    }
  } else if (warp < 4 && (warp >> 1) < nx) {
    // =============================== MMA issuers: warp = (query tile x, S buffer / key-tile parity) ===============
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);    // S = Q K^T (128 x 64), both K-major
    constexpr uint32_t idesc_o64 = umma_idesc_bf16(128, 64, 0, 1);  // O[:, 0:64]  += P V, V MN-major
    constexpr uint32_t idesc_o16 = umma_idesc_bf16(128, 16, 0, 1);  // O[:, 64:80] += P V
    constexpr uint32_t hi128 = umma_desc_hi(1024, 2);               // SWIZZLE_128B, 8-row / 8-key groups 1024 B apart
    constexpr uint32_t hi32 = umma_desc_hi(256, 6);                 // SWIZZLE_32B, groups 256 B apart
    const uint32_t elected = elect_one();
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t q_lo = umma_desc_lo(smem_u32(sQ), 16);
    const uint32_t kv_lo = umma_desc_lo(smem_u32(sKV), 16);
    auto issue_S = [&](int x, int buf, int stage) {
      const uint32_t ka = kv_lo + stage * (kStage >> 4);
      const uint32_t d = tb + c_s(x, buf);
      if (kQTm) {
        // A = Q_x from TMEM (bf16 pairs, 8 columns per 16-wide K step): an SS MMA re-reads its 4 KB Q slice from shared
        // memory every K step and is smem-bound at N = 64 (50.7 cycles against 34 in TS mode, tools/mma_microbench.cu)
        const uint32_t qt = tb + c_q(x);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ts_w(elected, d, qt + 8 * k, desc_pack(ka + 2 * k, hi128), idesc_s, k > 0 ? 1u : 0u);
        if (kHasB) umma_ts_w(elected, d, qt + 32, desc_pack(ka + (kKA >> 4), hi32), idesc_s, 1u);
      } else {
        const uint32_t qa = q_lo + x * (kQT >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss_w(elected, d, desc_pack(qa + 2 * k, hi128), desc_pack(ka + 2 * k, hi128), idesc_s, k > 0 ? 1u : 0u);
        if (kHasB)
          umma_ss_w(elected, d, desc_pack(qa + (kQA >> 4), hi32), desc_pack(ka + (kKA >> 4), hi32), idesc_s, 1u);
      }
    };
    auto issue_PV = [&](int x, int buf, int stage) {
      // V tiles are MN-major (d contiguous): the LBO field is the stride between d atoms, unused with a single atom
      const uint32_t va = kv_lo + stage * (kStage >> 4) + (kKT >> 4) - (1u << 16) + ((uint32_t(kKA) >> 4) << 16);
      const uint32_t vb = kv_lo + stage * (kStage >> 4) + ((kKT + kKA) >> 4) - (1u << 16) + ((uint32_t(kKB) >> 4) << 16);
      const uint32_t pt = tb + c_s(x, buf);  // bf16 P: 32 columns at the head of the S buffer
      const uint32_t d = tb + c_o(x);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {  // 16 keys per step: +2048 B in the 128B-swizzled chunk, +512 B in the 32B one
        umma_ts_w(elected, d, pt + ks * 8, desc_pack(va + ks * 128, hi128), idesc_o64, 1u);
        if (kHasB) umma_ts_w(elected, d + 64, pt + ks * 8, desc_pack(vb + ks * 32, hi32), idesc_o16, 1u);
      }
    };
    const int x = warp >> 1, par = warp & 1;
    mbar_wait(&q_full[kQTm ? x : 0], 0);
    if (kQTm) tc_fence_after();
    if (par < n_tiles) {  // prologue: S_x(par) -> my buffer
      mbar_wait(&k_full[par], 0);
      tc_fence_after();
      issue_S(x, par, par);
      umma_commit_w(elected, &s_full[x * 2 + par]);
    }
    for (int t = par; t < n_tiles; t += 2) {
      const int s = t % ST;
      const int s2 = (t + 2) % ST;
      // operands first (they land long before P is ready: these waits are off the critical path) ...
      mbar_wait(&v_full[s], (t / ST) & 1);
      if (kSumMMA) {
        // V_B chunk: 64 keys x 16 bf16 (32-byte rows, SWIZZLE_32B: the 16-byte half is XORed with address bit 7 = key
        // bit 2): element d = 72 is the first of the upper half.  Both query tiles' issuers write it (same value).
        unsigned char* vb = sKV + s * kStage + kKT + kKA;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int r = lane + rr * 32;
          *reinterpret_cast<uint16_t*>(vb + r * 32 + ((1 ^ ((r >> 2) & 1)) << 4)) = VSB_ONE_BITS;  // 1.0 in the element type
        }
        fence_proxy_async_smem();
        __syncwarp();
      }
      if (t + 2 < n_tiles) mbar_wait(&k_full[s2], ((t + 2) / ST) & 1);
      // ... then the softmax warps' P_x(t)
      mbar_wait(&p_full[x * 2 + par], (t >> 1) & 1);
      tc_fence_after();
      VSB_TRACE_W(t, x * 2);
      issue_PV(x, par, s);
      // K(t) (read by my S_x(t) two tiles ago, retired before this P V) and V(t) are consumed by query tile x
      umma_commit_w(elected, &kv_empty[s]);
      if (t + 2 < n_tiles) issue_S(x, par, s2);
      // also posted without a new S: the softmax warps use it as "PV_x(t) has completed" on their rare rescale path
      umma_commit_w(elected, &s_full[x * 2 + par]);
      VSB_TRACE_W(t, x * 2 + 1);
    }
    umma_commit_w(elected, &o_full[x]);  // all my P V MMAs (immediate if I had no key tile)
  } else if (warp >= 4 && ((warp - 4) >> 2) < nx) {
    // =============================== softmax warpgroups ===============================
    const int x = (warp - 4) >> 2;  // query tile 0/1
    const int ew = warp & 3;        // TMEM lane quarter
    const int row = ew * 32 + lane;
    const uint32_t lane_off = uint32_t(ew * 32) << 16;
    const uint32_t tO = tmem_base + lane_off + c_o(x);
    const float sl2 = p.scale_log2;
    float l_run = 0.f, m_run = -INFINITY;
    if (kQTm) {  // my query row -> TMEM (A operand of S = Q K^T): D bf16 = D/2 columns, zero-padded to the next K step
      const int qr = q0 + x * 128 + row;
      const uint4* src = reinterpret_cast<const uint4*>(p.q + (size_t)b * p.q_batch_stride + (size_t)qr * p.q_row_stride +
                                                         (size_t)h * D);
      const uint32_t tQ = tmem_base + lane_off + c_q(x);
      constexpr int NV = D / 8;  // 16-byte vectors per row
#pragma unroll
      for (int c = 0; c < (kHasB ? 5 : 4); ++c) {
        uint32_t w[8];
#pragma unroll
        for (int v2 = 0; v2 < 2; ++v2) {
          const int vi = c * 2 + v2;
          uint4 u = make_uint4(0, 0, 0, 0);
          if (vi < NV && qr < p.nq) u = __ldg(src + vi);
          w[4 * v2] = u.x;
          w[4 * v2 + 1] = u.y;
          w[4 * v2 + 2] = u.z;
          w[4 * v2 + 3] = u.w;
        }
        tmem_st8(tQ + c * 8, w);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_full[x]);
    }
    {  // O_x starts at zero: both issuers of this query tile only ever accumulate into it
      uint32_t z[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
      for (int c = 0; c < (kHasB ? 5 : 4); ++c) tmem_st16(tO + c * 16, z);
    }
    for (int t = 0; t < n_tiles; ++t) {
      const int buf = t & 1;
      const uint32_t tS = tmem_base + lane_off + c_s(x, buf);
      mbar_wait(&s_full[x * 2 + buf], (t >> 1) & 1);
      tc_fence_after();
      VSB_TRACE_W(t, 0);
      const int valid = kv_len - t * 64;  // >= 64: full tile; columns >= valid are masked
      uint32_t a[2][32];
      tmem_ld32(tS, a[0]);
      tmem_ld32(tS + 32, a[1]);
      tmem_wait_ld();
      VSB_TRACE_W(t, 1);
      float mx;
      if (valid >= 64) {
        float m0 = fmax3(__uint_as_float(a[0][0]), __uint_as_float(a[0][1]), __uint_as_float(a[0][2]));
        float m1 = fmax3(__uint_as_float(a[1][0]), __uint_as_float(a[1][1]), __uint_as_float(a[1][2]));
        float m2 = fmax3(__uint_as_float(a[0][3]), __uint_as_float(a[0][4]), __uint_as_float(a[0][5]));
        float m3 = fmax3(__uint_as_float(a[1][3]), __uint_as_float(a[1][4]), __uint_as_float(a[1][5]));
#pragma unroll
        for (int i = 6; i < 30; i += 4) {
          m0 = fmax3(m0, __uint_as_float(a[0][i]), __uint_as_float(a[0][i + 1]));
          m1 = fmax3(m1, __uint_as_float(a[1][i]), __uint_as_float(a[1][i + 1]));
          m2 = fmax3(m2, __uint_as_float(a[0][i + 2]), __uint_as_float(a[0][i + 3]));
          m3 = fmax3(m3, __uint_as_float(a[1][i + 2]), __uint_as_float(a[1][i + 3]));
        }
        m0 = fmax3(m0, __uint_as_float(a[0][30]), __uint_as_float(a[0][31]));
        m1 = fmax3(m1, __uint_as_float(a[1][30]), __uint_as_float(a[1][31]));
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      } else {  // ragged last tile: masked columns never win the max and get p = 0 below
        mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (c * 32 + i >= valid) a[c][i] = 0xff800000u;  // -inf
            mx = fmaxf(mx, __uint_as_float(a[c][i]));
          }
      }
      // ---- lazy rescale: keep the stale running max unless it grew by more than 2^8 (p stays <= 256) ----
      const float m_new = fmaxf(m_run, mx);
      const bool grow = (m_new - m_run) * sl2 > 8.f;  // first tile: m_run = -inf -> true
      const float alpha = grow ? fast_exp2((m_run - m_new) * sl2) : 1.f;
      if (t > 0 && __any_sync(0xffffffffu, grow)) {
        // PV_x(t-1) may still be running: the MMA warp's iteration t-1 ends with a commit on s_full[x][(t+1)&1]
        mbar_wait(&s_full[x * 2 + (buf ^ 1)], ((t + 1) >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < (kHasB ? 5 : 4); ++c) {  // one 16-column chunk at a time (register budget)
          uint32_t o[16];
          tmem_ld16(tO + c * 16, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st16(tO + c * 16, o);
        }
      }
      if (grow) m_run = m_new;
      const float mb = m_run * sl2;
      // ---- p = exp2(s*sl2 - m*sl2); row sum; bf16 P into the head of this S buffer ----
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (c < 2) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = fmaf(__uint_as_float(a[c][i]), sl2, -mb);
            // kPoly: 2 / 3 / 4 of every 8 exponentials run as a polynomial on the FMA pipe (relieves the MUFU pipe)
            constexpr uint32_t kMask = kPoly == 1 ? 0x88u : kPoly == 2 ? 0xA8u : kPoly == 3 ? 0xAAu : 0u;
            a[c][i] = __float_as_uint(((kMask >> (i & 7)) & 1u) ? exp2_poly(e) : fast_exp2(e));
          }
        }
        if (c > 0) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float p0 = __uint_as_float(a[c - 1][i]), p1 = __uint_as_float(a[c - 1][i + 1]);
            const float p2 = __uint_as_float(a[c - 1][i + 2]), p3 = __uint_as_float(a[c - 1][i + 3]);
            if (!kSumMMA) {
              s0 += p0;
              s1 += p1;
              s2 += p2;
              s3 += p3;
            }
            pk[i >> 1] = pack_bf16x2(p0, p1);
            pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
          }
          // P chunk c-1 overwrites S columns [16(c-1), 16c): all 64 score columns already sit in registers
          tmem_st16(tS + (c - 1) * 16, pk);
        }
      }
      l_run = l_run * alpha + ((s0 + s1) + (s2 + s3));
      VSB_TRACE_W(t, 2);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[x * 2 + buf]);
      VSB_TRACE_W(t, 3);
    }
    // ---- epilogue: O / l -> bf16 -> global ----
    mbar_wait(&o_full[x], 0);
    tc_fence_after();
    const int qrow = q0 + x * 128 + row;
    if (kSumMMA) {  // the row sum accumulated by the tensor core in O's column 72
      uint32_t r8[8];
      tmem_ld8(tO + 72, r8);
      tmem_wait_ld();
      l_run = __uint_as_float(r8[0]);
    }
    const float inv = 1.f / l_run;
    bf16* dst = p.out + (size_t)b * p.out_batch_stride + (size_t)(qrow < p.nq ? qrow : 0) * p.out_row_stride + (size_t)h * D;
#pragma unroll 1
    for (int c = 0; c < D / 8; ++c) {
      uint32_t r[8];
      tmem_ld8(tO + c * 8, r);
      tmem_wait_ld();
      if (qrow < p.nq) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv);
        u.y = pack_bf16x2(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv);
        u.z = pack_bf16x2(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv);
        u.w = pack_bf16x2(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv);
        *reinterpret_cast<uint4*>(dst + c * 8) = u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

template <int D, int kPoly, bool kQTm, bool kSumMMA = false>
static int launch_kt64(const CUtensorMap* tm, const AttnParams& prm, cudaStream_t st) {
  constexpr int smem = kQTm ? kT64SmemQT : kT64Smem;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_flash_kt64_kernel<D, kPoly, kQTm, kSumMMA>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(VSB_ERR_CUDA, "attn_flash(kt64): smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  dim3 grid((prm.nq + 255) / 256, prm.H, prm.nb);
  attn_flash_kt64_kernel<D, kPoly, kQTm, kSumMMA><<<grid, kT64Threads, smem, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], prm);
  return check_launch("attn_flash(kt64)");
}

template <bool kQTm>
static int dispatch_kt64(const CUtensorMap* tm, const AttnParams& prm, int D, int poly, cudaStream_t st) {
  if (D == 72) {
    switch (poly) {
      case 1: return launch_kt64<72, 1, kQTm>(tm, prm, st);
      case 2: return launch_kt64<72, 2, kQTm>(tm, prm, st);
      case 3: return launch_kt64<72, 3, kQTm>(tm, prm, st);
      default: return launch_kt64<72, 0, kQTm>(tm, prm, st);
    }
  }
  return poly ? launch_kt64<64, 1, kQTm>(tm, prm, st) : launch_kt64<64, 0, kQTm>(tm, prm, st);
}

int attn_flash_kt64_launch(const CUtensorMap* tm, const AttnParams& prm, int D, int poly, int q_tmem, cudaStream_t st) {
  if (q_tmem == 2 && D == 72) {  // variant 5: Q in TMEM + row sum on the tensor core
    switch (poly) {
      case 1: return launch_kt64<72, 1, true, true>(tm, prm, st);
      case 2: return launch_kt64<72, 2, true, true>(tm, prm, st);
      case 3: return launch_kt64<72, 3, true, true>(tm, prm, st);
      default: return launch_kt64<72, 0, true, true>(tm, prm, st);
    }
  }
  return q_tmem ? dispatch_kt64<true>(tm, prm, D, poly, st) : dispatch_kt64<false>(tm, prm, D, poly, st);
}

}  // namespace vsb
