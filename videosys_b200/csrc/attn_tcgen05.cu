// vsb200 -- flash attention forward on tcgen05 for head_dim 72 (STDiT3) and 64 (CogVideoX).
//
// One CTA per SM-slot handles TWO 128-row query tiles (A, B) of one (batch, head) against all key tiles of 128:
//   warp 0        TMA producer: Q tiles once, K/V tiles through a 3-deep mbarrier ring
//   warp 1        MMA issuer  : S_X = Q_X K^T  (SS, fp32 in TMEM),  O_X += P_X V (TS: P read from TMEM, V MN-major)
//   warp 2        TMEM allocator (512 columns: S_A|S_B 128 each, O_A|O_B 80 each; bf16 P aliases the head of S)
//   warps 4..7    softmax warpgroup A (thread = query row; tcgen05.ld 32x32b, no shuffles needed)
//   warps 8..11   softmax warpgroup B
// The two query tiles ping-pong: while warpgroup A runs exp2 on S_A(j) the tensor core does PV_B(j-1) / S_B(j).
//
// head_dim 72 is not a multiple of the 128-byte swizzle span: every operand tile is staged as a 64-wide
// SWIZZLE_128B chunk plus a 16-wide SWIZZLE_32B chunk (columns 64..79; the tensor map's inner extent is 72, so
// TMA zero-fills 72..79).  QK^T runs 4+1 K-steps, PV runs two N-slices (64 and 16) per K-step.
#include "attn_params.cuh"

namespace vsb {

constexpr int kAttnThreads = 384;
constexpr int kKvStages = 3;
constexpr int kTileA = 128 * 128;  // bytes: 128 rows x 64 bf16 (SWIZZLE_128B)
constexpr int kTileB = 128 * 32;   // bytes: 128 rows x 16 bf16 (SWIZZLE_32B)
constexpr int kQBytes = kTileA + kTileB;
constexpr int kKvStageBytes = 2 * (kTileA + kTileB);
constexpr int kAttnSmem = 2 * kQBytes + kKvStages * kKvStageBytes + 1024 + 256;

// TMEM columns
__host__ __device__ constexpr uint32_t col_s(int x) { return uint32_t(x) * 128u; }        // S_A, S_B
__host__ __device__ constexpr uint32_t col_o(int x) { return 256u + uint32_t(x) * 80u; }  // O_A, O_B: 64 + 16 columns

template <int D>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_flash_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_qb,
                  const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_kb,
                  const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_vb,
                  const __grid_constant__ AttnParams p) {
  constexpr bool kHasB = (D > 64);
  constexpr int kQTx = kHasB ? kQBytes : kTileA;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;                       // [2][Q_A | Q_B]
  unsigned char* sKV = smem + 2 * kQBytes;        // [stages][K_A | K_B | V_A | V_B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + kKvStages * kKvStageBytes);
  uint64_t* q_full = bars;                 // [1]
  uint64_t* k_full = bars + 1;             // [stages]
  uint64_t* v_full = k_full + kKvStages;   // [stages]
  uint64_t* kv_empty = v_full + kKvStages; // [stages]
  uint64_t* s_full = kv_empty + kKvStages; // [2]
  uint64_t* p_full = s_full + 2;           // [2]
  uint64_t* o_full = p_full + 2;           // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
  const int kv_len = p.has_lens ? p.lens[b] : p.nk;
  const int n_tiles = (kv_len + 127) / 128;

  if (warp == 0 && lane == 0) {
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
    mbar_init(q_full, 1);
    for (int i = 0; i < kKvStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);  // one arrival per softmax warp
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // =============================== TMA producer (whole warp converged, one elected lane issues) ==========
    const uint32_t elected = elect_one();
    mbar_arrive_expect_tx_w(elected, q_full, 2 * kQTx);
    for (int x = 0; x < 2; ++x) {
      tma_load_4d_w(elected, &tm_q, q_full, sQ + x * kQBytes, 0, h, q0 + x * 128, b);
      if (kHasB) tma_load_4d_w(elected, &tm_qb, q_full, sQ + x * kQBytes + kTileA, 64, h, q0 + x * 128, b);
    }
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j % kKvStages;
      const uint32_t ph = (j / kKvStages) & 1;
      mbar_wait(&kv_empty[s], ph ^ 1);
      unsigned char* st = sKV + s * kKvStageBytes;
      mbar_arrive_expect_tx_w(elected, &k_full[s], kQTx);
      tma_load_4d_w(elected, &tm_k, &k_full[s], st, 0, h, j * 128, b);
      if (kHasB) tma_load_4d_w(elected, &tm_kb, &k_full[s], st + kTileA, 64, h, j * 128, b);
      mbar_arrive_expect_tx_w(elected, &v_full[s], kQTx);
      tma_load_4d_w(elected, &tm_v, &v_full[s], st + kQBytes, 0, h, j * 128, b);
      if (kHasB) tma_load_4d_w(elected, &tm_vb, &v_full[s], st + kQBytes + kTileA, 64, h, j * 128, b);
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (whole warp converged, one elected lane issues) ============
    // Operands are warp-uniform and the descriptors are (constant high word, low word + small immediate), so each
    // tcgen05.mma costs a handful of instructions: with 21 small MMAs per key tile the issue rate, not the tensor
    // pipe, was the bottleneck when they were built inside `if (lane == 0)` (measured 83 cycles per MMA).
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // S = Q K^T, both K-major
    constexpr uint32_t idesc_o64 = umma_idesc_bf16(128, 64, 0, 1);  // O[:, 0:64]  += P V, V MN-major
    constexpr uint32_t idesc_o16 = umma_idesc_bf16(128, 16, 0, 1);  // O[:, 64:80] += P V
    constexpr uint32_t hi128 = umma_desc_hi(1024, 2);               // SWIZZLE_128B, 8-row / 8-key groups 1024 B apart
    constexpr uint32_t hi32 = umma_desc_hi(256, 6);                 // SWIZZLE_32B, groups 256 B apart
    const uint32_t elected = elect_one();
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t q_lo = umma_desc_lo(smem_u32(sQ), 16);
    const uint32_t kv_lo = umma_desc_lo(smem_u32(sKV), 16);
    auto issue_S = [&](int x, int stage) {
      const uint32_t qa = q_lo + x * (kQBytes >> 4);
      const uint32_t ka = kv_lo + stage * (kKvStageBytes >> 4);
      const uint32_t d = tb + col_s(x);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_ss_w(elected, d, desc_pack(qa + 2 * k, hi128), desc_pack(ka + 2 * k, hi128), idesc_s, k > 0 ? 1u : 0u);
      if (kHasB)
        umma_ss_w(elected, d, desc_pack(qa + (kTileA >> 4), hi32), desc_pack(ka + (kTileA >> 4), hi32), idesc_s, 1u);
      umma_commit_w(elected, &s_full[x]);
    };
    auto issue_PV = [&](int x, int stage, bool accumulate) {
      // V tiles are MN-major: LBO = stride between 64-wide (16-wide) d atoms, unused with a single atom
      const uint32_t va = kv_lo + stage * (kKvStageBytes >> 4) + (kQBytes >> 4) - (1u << 16) + ((16384u >> 4) << 16);
      const uint32_t vb = kv_lo + stage * (kKvStageBytes >> 4) + ((kQBytes + kTileA) >> 4) - (1u << 16) + ((4096u >> 4) << 16);
      const uint32_t pt = tb + col_s(x);  // bf16 P aliases the first 64 columns of S_x
      const uint32_t d = tb + col_o(x);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {  // 16 keys per step: +2048 B in the 128B-swizzled chunk, +512 B in the 32B one
        const uint32_t acc = (accumulate || ks > 0) ? 1u : 0u;
        umma_ts_w(elected, d, pt + ks * 8, desc_pack(va + ks * 128, hi128), idesc_o64, acc);
        if (kHasB) umma_ts_w(elected, d + 64, pt + ks * 8, desc_pack(vb + ks * 32, hi32), idesc_o16, acc);
      }
    };
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    issue_S(0, 0);
    issue_S(1, 0);
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j % kKvStages;
      const uint32_t ph = (j / kKvStages) & 1;
      const int s1 = (j + 1) % kKvStages;
      const uint32_t ph1 = ((j + 1) / kKvStages) & 1;
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        mbar_wait(&p_full[x], j & 1);
        if (x == 0) mbar_wait(&v_full[s], ph);
        tc_fence_after();
        VSB_TRACE(0, j, x * 2);
        issue_PV(x, s, j > 0);
        if (x == 1) umma_commit_w(elected, &kv_empty[s]);  // K_j and V_j fully consumed by both query tiles
        if (j + 1 < n_tiles) {
          if (x == 0) {
            mbar_wait(&k_full[s1], ph1);
            tc_fence_after();
          }
          issue_S(x, s1);
          VSB_TRACE(0, j, x * 2 + 1);
        } else {
          umma_commit_w(elected, &o_full[x]);
        }
      }
    }
  } else if (warp >= 4) {
    // =============================== softmax warpgroups ===============================
    const int x = (warp - 4) >> 2;          // query tile 0/1
    const int ew = warp & 3;                // TMEM lane quarter
    const int row = ew * 32 + lane;
    const uint32_t lane_off = uint32_t(ew * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + col_s(x);
    const uint32_t tO = tmem_base + lane_off + col_o(x);
    const float sl2 = p.scale_log2;
    float l_run = 0.f;
    if (p.pingpong && x == 1) named_bar_arrive(2, 256);  // warpgroup A goes first
    {
    float m_run = -INFINITY;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&s_full[x], j & 1);
      tc_fence_after();
      VSB_TRACE(1 + x, j, 0);
      const int valid = kv_len - j * 128;  // >= 128: full tile; columns >= valid are masked
      // ---- all 128 scores of my row -> registers (4 loads in flight, one wait) ----
      uint32_t a[4][32];
      tmem_ld32(tS, a[0]);
      tmem_ld32(tS + 32, a[1]);
      tmem_ld32(tS + 64, a[2]);
      tmem_ld32(tS + 96, a[3]);
      tmem_wait_ld();
      VSB_TRACE(1 + x, j, 1);
      float mx;
      if (valid >= 128) {
        float m0 = fmax3(__uint_as_float(a[0][0]), __uint_as_float(a[0][1]), __uint_as_float(a[0][2]));
        float m1 = fmax3(__uint_as_float(a[1][0]), __uint_as_float(a[1][1]), __uint_as_float(a[1][2]));
        float m2 = fmax3(__uint_as_float(a[2][0]), __uint_as_float(a[2][1]), __uint_as_float(a[2][2]));
        float m3 = fmax3(__uint_as_float(a[3][0]), __uint_as_float(a[3][1]), __uint_as_float(a[3][2]));
#pragma unroll
        for (int i = 3; i < 31; i += 2) {
          m0 = fmax3(m0, __uint_as_float(a[0][i]), __uint_as_float(a[0][i + 1]));
          m1 = fmax3(m1, __uint_as_float(a[1][i]), __uint_as_float(a[1][i + 1]));
          m2 = fmax3(m2, __uint_as_float(a[2][i]), __uint_as_float(a[2][i + 1]));
          m3 = fmax3(m3, __uint_as_float(a[3][i]), __uint_as_float(a[3][i + 1]));
        }
        m0 = fmaxf(m0, __uint_as_float(a[0][31]));
        m1 = fmaxf(m1, __uint_as_float(a[1][31]));
        m2 = fmaxf(m2, __uint_as_float(a[2][31]));
        m3 = fmaxf(m3, __uint_as_float(a[3][31]));
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      } else {  // ragged last tile: masked columns never win the max and get p = 0 below
        mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c)
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
      if (j > 0 && __any_sync(0xffffffffu, grow)) {
        // O is quiescent here: S_x(j) complete implies PV_x(j-1) complete (in-order tensor pipe)
#pragma unroll 1
        for (int c = 0; c < (kHasB ? 5 : 4); ++c) {  // rare path: one 16-column chunk at a time (register budget)
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
      if (p.pingpong) named_bar_sync(2 + x, 256);  // my turn on the MUFU pipe (the other warpgroup arrived)
      // ---- p = exp2(s*sl2 - m*sl2) in place; row sum and bf16 packing of chunk c-1 are interleaved with the
      //      exponentials of chunk c, so the MUFU results are consumed ~32 instructions after they were issued and a
      //      single warp can keep the MUFU pipe (8 cycles per warp instruction) busy back to back ----
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        if (c < 4) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            a[c][i] = __float_as_uint(fast_exp2(fmaf(__uint_as_float(a[c][i]), sl2, -mb)));
        }
        if (c > 0) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float p0 = __uint_as_float(a[c - 1][i]), p1 = __uint_as_float(a[c - 1][i + 1]);
            const float p2 = __uint_as_float(a[c - 1][i + 2]), p3 = __uint_as_float(a[c - 1][i + 3]);
            s0 += p0;
            s1 += p1;
            s2 += p2;
            s3 += p3;
            pk[i >> 1] = pack_bf16x2(p0, p1);
            pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
          }
          tmem_st16(tS + (c - 1) * 16, pk);
        }
      }
      l_run = l_run * alpha + ((s0 + s1) + (s2 + s3));
      if (p.pingpong) named_bar_arrive(2 + (x ^ 1), 256);  // hand the MUFU pipe to the other warpgroup
      VSB_TRACE(1 + x, j, 2);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[x]);
      VSB_TRACE(1 + x, j, 3);
    }
    }
    // ---- epilogue: O / l -> bf16 -> global ----
    mbar_wait(&o_full[x], 0);
    tc_fence_after();
    const int qrow = q0 + x * 128 + row;
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
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// g_attn_trace and the attn_* options live in api.cu (namespace vsbs, shared with the fp16 twin)

}  // namespace vsb

using namespace vsb;

#ifndef VSB_HALF
extern "C" int vsb_debug_attn_trace(void* device_buffer) {
  g_attn_trace = (long long*)device_buffer;
  return VSB_OK;
}
#endif

extern "C" int VSB_API(vsb_attn_flash)(const vsb_bf16* q, const vsb_bf16* k, const vsb_bf16* v, vsb_bf16* out, int nb, int nq,
                              int nk, int H, int D, long long q_row_stride, long long q_batch_stride,
                              long long kv_row_stride, long long kv_batch_stride, const int* host_kv_lens, float scale,
                              void* stream) {
  return VSB_API(vsb_attn_flash_strided)(q, k, v, out, nb, nq, nk, H, D, q_row_stride, q_batch_stride, kv_row_stride,
                                kv_batch_stride, (long long)H * D, (long long)nq * H * D, host_kv_lens, scale, stream);
}

extern "C" int VSB_API(vsb_attn_flash_strided)(const vsb_bf16* q, const vsb_bf16* k, const vsb_bf16* v, vsb_bf16* out, int nb,
                                      int nq, int nk, int H, int D, long long q_row_stride, long long q_batch_stride,
                                      long long kv_row_stride, long long kv_batch_stride, long long out_row_stride,
                                      long long out_batch_stride, const int* host_kv_lens, float scale, void* stream) {
  if (!q || !k || !v || !out || nb <= 0 || nq <= 0 || nk <= 0 || H <= 0) return fail(VSB_ERR_INVALID, "attn_flash: bad args");
  if ((out_row_stride % 8) || (out_batch_stride % 8) || out_row_stride < (long long)H * D)
    return fail(VSB_ERR_UNSUPPORTED, "attn_flash: output strides must be multiples of 8 elements, rows >= H*D apart");
  if (D % 16 != 0 && D != 72) return fail(VSB_ERR_UNSUPPORTED, "attn_flash: head_dim %d", D);
  if ((q_row_stride % 8) || (q_batch_stride % 8) || (kv_row_stride % 8) || (kv_batch_stride % 8) || !aligned16(q) ||
      !aligned16(k) || !aligned16(v) || !aligned16(out))
    return fail(VSB_ERR_UNSUPPORTED, "attn_flash: strides must be multiples of 8 elements and pointers 16B-aligned");
  if (host_kv_lens && nb > kAttnMaxLens)
    return fail(VSB_ERR_UNSUPPORTED, "attn_flash: per-batch key lengths need nb <= %d", kAttnMaxLens);
  if (nb > 65535 || H > 65535) return fail(VSB_ERR_UNSUPPORTED, "attn_flash: grid too large");
  AttnParams prm;
  prm.trace = g_attn_trace;
  prm.pingpong = g_opt_attn_pingpong;
  prm.poly_exp = g_opt_attn_poly;
  prm.out = (bf16*)out;
  prm.out_row_stride = out_row_stride;
  prm.out_batch_stride = out_batch_stride;
  prm.q = (const bf16*)q;
  prm.q_row_stride = q_row_stride;
  prm.q_batch_stride = q_batch_stride;
  prm.nb = nb;
  prm.nq = nq;
  prm.nk = nk;
  prm.H = H;
  prm.scale_log2 = scale * 1.4426950408889634f;
  prm.has_lens = host_kv_lens ? 1 : 0;
  for (int i = 0; i < kAttnMaxLens; ++i) prm.lens[i] = 0;
  if (host_kv_lens)
    for (int i = 0; i < nb; ++i) {
      if (host_kv_lens[i] < 1 || host_kv_lens[i] > nk) return fail(VSB_ERR_INVALID, "attn_flash: kv_lens[%d]=%d", i, host_kv_lens[i]);
      prm.lens[i] = host_kv_lens[i];
    }
  if (D != 72 && D != 64)  // no tcgen05 layout for this head_dim: the mma.sync kernel (attn_mma.cu)
    return attn_mma_launch((const bf16*)q, (const bf16*)k, (const bf16*)v, kv_row_stride, kv_batch_stride, prm, D,
                           (cudaStream_t)stream);
  // Two tensor maps per operand: the 64-wide SWIZZLE_128B chunk and (head_dim 72 only) the 16-wide SWIZZLE_32B
  // chunk at d = 64..79.  The inner extent is D, so TMA zero-fills 72..79 and rows past nq / nk.
  // auto: text cross-attention (a handful of key tiles per query pair) is dominated by per-CTA fixed costs
  // ... long key sequences: Q resident in TMEM (+8 % on the 720p spatial shape), head_dim 72 also takes the row sum from
  // the tensor core (interleaved microbenchmark, profiles/r02_kernel_bench.json: 670 -> 726 / 729 TFLOP/s)
  // ... and up to 320 keys (text cross-attention): K/V of a (batch, head) resident in shared memory (variant 6)
  int variant = g_opt_attn_variant >= 0 ? g_opt_attn_variant : (nk <= 320 ? 6 : (nk <= 1024 ? 3 : (D == 72 ? 5 : 4)));
  if (variant == 6 && nk > 320) variant = 3;
  CUtensorMap tm[6];
  const vsb_bf16* base[3] = {q, k, v};
  for (int i = 0; i < 3; ++i) {
    const long long rs = i == 0 ? q_row_stride : kv_row_stride, bs = i == 0 ? q_batch_stride : kv_batch_stride;
    unsigned long long dims[4] = {(unsigned long long)D, (unsigned long long)H, (unsigned long long)(i == 0 ? nq : nk),
                                  (unsigned long long)nb};
    unsigned long long str[3] = {(unsigned long long)D * 2, (unsigned long long)rs * 2, (unsigned long long)bs * 2};
    const unsigned rows = i == 0 ? 128u : (variant == 6 ? 160u : (variant >= 2 ? 64u : 128u));  // key tile of the schedule
    unsigned boxA[4] = {64, 1, rows, 1}, boxB[4] = {16, 1, rows, 1};
    int rc = make_tmap_bf16(&tm[2 * i], base[i], 4, dims, str, boxA, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if (D == 72) {
      rc = make_tmap_bf16(&tm[2 * i + 1], base[i], 4, dims, str, boxB, CU_TENSOR_MAP_SWIZZLE_32B);
      if (rc) return rc;
    } else {
      tm[2 * i + 1] = tm[2 * i];  // unused by the head_dim-64 instantiation
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (variant == 6) return attn_flash_kvres_launch(tm, prm, D, st);
  if (variant == 3) return attn_flash_kt64p_launch(tm, prm, D, g_opt_attn_poly, st);
  if (variant == 2 || variant == 4 || variant == 5)
    return attn_flash_kt64_launch(tm, prm, D, g_opt_attn_poly, variant == 5 ? 2 : (variant == 4 ? 1 : 0), st);
  dim3 grid((nq + 255) / 256, H, nb);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_flash_kernel<72>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_flash_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e != cudaSuccess) return fail(VSB_ERR_CUDA, "attn_flash: smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  if (D == 72)
    attn_flash_kernel<72><<<grid, kAttnThreads, kAttnSmem, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], prm);
  else
    attn_flash_kernel<64><<<grid, kAttnThreads, kAttnSmem, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], prm);
  return check_launch("attn_flash");
}
