// vsb200 -- flash attention forward on tcgen05, 64-key tiles, double-buffered S, PERSISTENT CTAs (attn_variant = 3).
//
// Same per-tile machinery as attn_tcgen05_kt64.cu (read that header first); what changes is the lifetime of a CTA.
// There a CTA serves one pair of query tiles and pays ~12k cycles of fixed cost around its key loop (launch, barrier
// init, TMEM allocation, the first Q / K round trip, draining the tensor pipe, the TMEM -> global epilogue, teardown):
// 63 % of a text cross-attention CTA (5 key tiles = 7k cycles of work) and ~12 % of a 720p spatial one (57 key tiles).
// Here the grid is one CTA per SM; each walks a contiguous range of (batch, head, query-pair) items and every role runs
// ahead across item boundaries on one global key-tile counter g:
//   * the TMA producer streams K/V tiles of item i+1 into the ring while item i is still being consumed, and Q tiles
//     into the second of two Q buffers (released by the issuers when the S MMAs of the item two back have retired);
//   * issuer (x, parity) owns the global tiles g = parity (mod 2) of query tile x: after PV(g) it issues S for its next
//     own tile even when that tile belongs to the next item, so the softmax warps find S ready when they come back from
//     the epilogue;
//   * the softmax warps of x zero O_x, run the item's tiles, wait for both issuers' last P V, write O / l, and go on.
// Barrier phases are counted per barrier (uses so far), not derived from the tile index: the two query tiles of a CTA
// no longer see the same tiles (the ragged last pair of a sequence has one live query tile).
#include "attn_params.cuh"

namespace vsb {

constexpr int kPThreads = 416;  // 4 MMA issuer warps, 8 softmax warps, 1 TMA producer warp
constexpr int kPStages = 6;
constexpr int kQA = 128 * 128;  // Q: 128 rows x 64 bf16, SWIZZLE_128B
constexpr int kQB = 128 * 32;   // Q: 128 rows x 16 bf16, SWIZZLE_32B
constexpr int kQT = kQA + kQB;
constexpr int kKA = 64 * 128;   // K / V: 64 keys x 64 bf16, SWIZZLE_128B
constexpr int kKB = 64 * 32;    // K / V: 64 keys x 16 bf16, SWIZZLE_32B
constexpr int kKT = kKA + kKB;
constexpr int kStage = 2 * kKT;  // K_A | K_B | V_A | V_B
constexpr int kPSmem = 4 * kQT + kPStages * kStage + 1024 + 512;  // two Q buffers of two query tiles each

__host__ __device__ constexpr uint32_t c_s(int x, int buf) { return uint32_t(x) * 128u + uint32_t(buf) * 64u; }
__host__ __device__ constexpr uint32_t c_o(int x) { return 256u + uint32_t(x) * 80u; }

struct Item {
  int b, h, q0, kv_len, n_tiles, nx;
};
__device__ __forceinline__ Item item_of(const AttnParams& p, int item, int n_pairs) {
  Item it;
  const int bh = item / n_pairs;
  it.q0 = (item - bh * n_pairs) * 256;
  it.b = bh / p.H;
  it.h = bh - it.b * p.H;
  it.kv_len = p.has_lens ? p.lens[it.b] : p.nk;
  it.n_tiles = (it.kv_len + 63) / 64;
  it.nx = (it.q0 + 128 < p.nq) ? 2 : 1;  // query tiles of the pair that hold real rows
  return it;
}

template <int D, int kPoly>
__global__ void __launch_bounds__(kPThreads, 1)
attn_flash_kt64p_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_qb,
                        const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_kb,
                        const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_vb,
                        const __grid_constant__ AttnParams p) {
  constexpr bool kHasB = (D > 64);
  constexpr int kQTx = kHasB ? kQT : kQA;
  constexpr int kKTx = kHasB ? kKT : kKA;
  constexpr int ST = kPStages;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* sQ = smem;               // [q buffer][x][Q_A | Q_B]
  unsigned char* sKV = smem + 4 * kQT;    // [stages][K_A | K_B | V_A | V_B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + ST * kStage);
  uint64_t* q_full = bars;              // [2]
  uint64_t* q_empty = bars + 2;         // [2]  8 arrivals (softmax warps, after the item's last MMA has retired)
  uint64_t* k_full = bars + 4;          // [ST]
  uint64_t* v_full = k_full + ST;       // [ST]
  uint64_t* kv_empty = v_full + ST;     // [ST] 2 commits: P V of both query tiles (one issuer commits twice if nx = 1)
  uint64_t* s_full = kv_empty + ST;     // [x][buf]
  uint64_t* p_full = s_full + 4;        // [x][buf]
  uint64_t* o_full = p_full + 4;        // [x][issuer parity]: that issuer's last P V of the item has retired
  uint64_t* pv_done = o_full + 4;       // [x][issuer parity]: its latest P V has retired (rare O-rescale path)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_pairs = (p.nq + 255) / 256;
  const int total = p.nb * p.H * n_pairs;
  // Items are dealt round-robin (CTA c takes c, c + G, c + 2G, ...): at any moment the SMs work on neighbouring query
  // pairs of a few (batch, head)s, whose K/V stay L2-resident.  (Contiguous ranges put every SM on its own (batch,
  // head): 148 x 1 MB of K/V at 720p thrashes the 126 MB L2 -- measured 5 % slower than one CTA per pair.)
  const int G = gridDim.x;
  const int i0 = blockIdx.x;   // first item; local item li is item i0 + li * G
  const int i1 = total;        // end of the item space

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
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 8);
    }
    for (int i = 0; i < ST; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 2);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);  // one arrival per softmax warp
      mbar_init(&o_full[i], 1);
      mbar_init(&pv_done[i], 1);
    }
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
    int g = 0;
    for (int item = i0, li = 0; item < i1; item += G, ++li) {
      const Item it = item_of(p, item, n_pairs);
      const int qb = li & 1;
      mbar_wait(&q_empty[qb], ((li >> 1) & 1) ^ 1);
      mbar_arrive_expect_tx_w(elected, &q_full[qb], it.nx * kQTx);
      for (int x = 0; x < it.nx; ++x) {
        unsigned char* dst = sQ + (qb * 2 + x) * kQT;
        tma_load_4d_w(elected, &tm_q, &q_full[qb], dst, 0, it.h, it.q0 + x * 128, it.b);
        if (kHasB) tma_load_4d_w(elected, &tm_qb, &q_full[qb], dst + kQA, 64, it.h, it.q0 + x * 128, it.b);
      }
      for (int t = 0; t < it.n_tiles; ++t, ++g) {
        const int s = g % ST;
        mbar_wait(&kv_empty[s], ((g / ST) & 1) ^ 1);
        unsigned char* st = sKV + s * kStage;
        mbar_arrive_expect_tx_w(elected, &k_full[s], kKTx);
        tma_load_4d_w(elected, &tm_k, &k_full[s], st, 0, it.h, t * 64, it.b);
        if (kHasB) tma_load_4d_w(elected, &tm_kb, &k_full[s], st + kKA, 64, it.h, t * 64, it.b);
        mbar_arrive_expect_tx_w(elected, &v_full[s], kKTx);
        tma_load_4d_w(elected, &tm_v, &v_full[s], st + kKT, 0, it.h, t * 64, it.b);
        if (kHasB) tma_load_4d_w(elected, &tm_vb, &v_full[s], st + kKT + kKA, 64, it.h, t * 64, it.b);
      }
    }
  } else if (warp < 4) {
    // =============================== MMA issuers: warp = (query tile x, S buffer = global key-tile parity) ==========
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);    // S = Q K^T (128 x 64), both K-major
    constexpr uint32_t idesc_o64 = umma_idesc_bf16(128, 64, 0, 1);  // O[:, 0:64]  += P V, V MN-major
    constexpr uint32_t idesc_o16 = umma_idesc_bf16(128, 16, 0, 1);  // O[:, 64:80] += P V
    constexpr uint32_t hi128 = umma_desc_hi(1024, 2);               // SWIZZLE_128B, 8-row / 8-key groups 1024 B apart
    constexpr uint32_t hi32 = umma_desc_hi(256, 6);                 // SWIZZLE_32B, groups 256 B apart
    const uint32_t elected = elect_one();
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t q_lo = umma_desc_lo(smem_u32(sQ), 16);
    const uint32_t kv_lo = umma_desc_lo(smem_u32(sKV), 16);
    auto issue_S = [&](int x, int buf, int stage, int qb) {
      const uint32_t qa = q_lo + (qb * 2 + x) * (kQT >> 4);
      const uint32_t ka = kv_lo + stage * (kStage >> 4);
      const uint32_t d = tb + c_s(x, buf);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_ss_w(elected, d, desc_pack(qa + 2 * k, hi128), desc_pack(ka + 2 * k, hi128), idesc_s, k > 0 ? 1u : 0u);
      if (kHasB)
        umma_ss_w(elected, d, desc_pack(qa + (kQA >> 4), hi32), desc_pack(ka + (kKA >> 4), hi32), idesc_s, 1u);
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
    struct Cur {
      int item, li, t, g;
      Item it;
    };
    auto step = [&](Cur& c) {  // next key tile in the CTA's global order
      ++c.t;
      ++c.g;
      if (c.t >= c.it.n_tiles) {
        c.item += G;
        ++c.li;
        c.t = 0;
        if (c.item < i1) c.it = item_of(p, c.item, n_pairs);
      }
    };
    auto seek = [&](Cur& c) {  // ... that is mine: my parity, and my query tile is live in its item
      while (c.item < i1 && !(((c.g & 1) == par) && c.it.nx > x)) step(c);
    };
    Cur cur;
    cur.item = i0;
    cur.li = 0;
    cur.t = 0;
    cur.g = 0;
    if (i0 < i1) cur.it = item_of(p, i0, n_pairs);
    seek(cur);
    uint32_t cnt = 0;  // P tiles consumed from my buffer
    if (cur.item < i1) {  // prologue: S for my first tile
      const int li = cur.li;
      mbar_wait(&q_full[li & 1], (li >> 1) & 1);
      mbar_wait(&k_full[cur.g % ST], (cur.g / ST) & 1);
      tc_fence_after();
      issue_S(x, par, cur.g % ST, li & 1);
      umma_commit_w(elected, &s_full[x * 2 + par]);
    }
    while (cur.item < i1) {
      Cur nxt = cur;
      step(nxt);
      seek(nxt);
      const bool has_next = nxt.item < i1;
      const bool last_in_item = nxt.item != cur.item;
      // The operands of my next tile can be awaited BEFORE P(cur) only if they cannot depend on work I still owe:
      // the next item's Q sits in the other Q buffer and a K tile at most two ahead has had its ring slot free for
      // a while.  Anything further (I skip an item whose query tile 1 is empty) needs this item's Q buffer or ring
      // slots released -- by retiring, among others, the very P V below.
      const bool near = has_next && nxt.li <= cur.li + 1 && nxt.g <= cur.g + 2;
      const int s = cur.g % ST;
      mbar_wait(&v_full[s], (cur.g / ST) & 1);
      if (near) {
        const int li = nxt.li;
        mbar_wait(&q_full[li & 1], (li >> 1) & 1);
        mbar_wait(&k_full[nxt.g % ST], (nxt.g / ST) & 1);
      }
      mbar_wait(&p_full[x * 2 + par], cnt & 1);
      ++cnt;
      tc_fence_after();
      VSB_TRACE_W(cur.g, x * 2);
      issue_PV(x, par, s);
      umma_commit_w(elected, &kv_empty[s]);
      if (cur.it.nx == 1) umma_commit_w(elected, &kv_empty[s]);  // no second query tile to wait for
      umma_commit_w(elected, &pv_done[x * 2 + par]);
      if (last_in_item) umma_commit_w(elected, &o_full[x * 2 + par]);
      if (has_next) {
        if (!near) {
          const int li = nxt.li;
          mbar_wait(&q_full[li & 1], (li >> 1) & 1);
          mbar_wait(&k_full[nxt.g % ST], (nxt.g / ST) & 1);
          tc_fence_after();
        }
        issue_S(x, par, nxt.g % ST, nxt.li & 1);
        umma_commit_w(elected, &s_full[x * 2 + par]);
      }
      VSB_TRACE_W(cur.g, x * 2 + 1);
      cur = nxt;
    }
  } else if (warp >= 4 && warp < 12) {
    // =============================== softmax warpgroups ===============================
    const int x = (warp - 4) >> 2;  // query tile 0/1
    const int ew = warp & 3;        // TMEM lane quarter
    const int row = ew * 32 + lane;
    const uint32_t lane_off = uint32_t(ew * 32) << 16;
    const uint32_t tO = tmem_base + lane_off + c_o(x);
    const float sl2 = p.scale_log2;
    uint32_t cnt[2] = {0u, 0u};    // S tiles consumed per buffer
    uint32_t odone[2] = {0u, 0u};  // o_full phases consumed per issuer
    int g = 0;
    for (int item = i0, li = 0; item < i1; item += G, ++li) {
      const Item it = item_of(p, item, n_pairs);
      if (it.nx <= x) {  // the ragged last pair of a sequence: query tile 1 holds no rows
        g += it.n_tiles;
        continue;
      }
      float l_run = 0.f, m_run = -INFINITY;
      {  // O_x starts at zero: both issuers of this query tile only ever accumulate into it
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
        for (int c = 0; c < (kHasB ? 5 : 4); ++c) tmem_st16(tO + c * 16, z);
      }
      for (int t = 0; t < it.n_tiles; ++t, ++g) {
      const int buf = g & 1;
      const uint32_t tS = tmem_base + lane_off + c_s(x, buf);
      mbar_wait(&s_full[x * 2 + buf], cnt[buf] & 1);
      ++cnt[buf];
      tc_fence_after();
      VSB_TRACE_W(g, 0);
      const int valid = it.kv_len - t * 64;  // >= 64: full tile; columns >= valid are masked
      uint32_t a[2][32];
      tmem_ld32(tS, a[0]);
      tmem_ld32(tS + 32, a[1]);
      tmem_wait_ld();
      VSB_TRACE_W(g, 1);
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
        // PV_x(t-1) may still be running: wait for the other buffer's issuer to report its latest P V retired
        mbar_wait(&pv_done[x * 2 + (buf ^ 1)], (cnt[buf ^ 1] - 1) & 1);
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
            s0 += p0;
            s1 += p1;
            s2 += p2;
            s3 += p3;
            pk[i >> 1] = pack_bf16x2(p0, p1);
            pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
          }
          // P chunk c-1 overwrites S columns [16(c-1), 16c): all 64 score columns already sit in registers
          tmem_st16(tS + (c - 1) * 16, pk);
        }
      }
      l_run = l_run * alpha + ((s0 + s1) + (s2 + s3));
      VSB_TRACE_W(g, 2);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[x * 2 + buf]);
      VSB_TRACE_W(g, 3);
      }
      // ---- epilogue: the last P V of each issuer that had a tile in this item has retired; O / l -> bf16 -> global ----
      {
        const int g0 = g - it.n_tiles;  // global index of the item's first key tile
#pragma unroll
        for (int par = 0; par < 2; ++par) {
          if (it.n_tiles >= 2 || (g0 & 1) == par) {
            mbar_wait(&o_full[x * 2 + par], odone[par] & 1);
            ++odone[par];
          }
        }
      }
      tc_fence_after();
      const int qrow = it.q0 + x * 128 + row;
      const float inv = 1.f / l_run;
      bf16* dst = p.out + (size_t)it.b * p.out_batch_stride + (size_t)(qrow < p.nq ? qrow : 0) * p.out_row_stride + (size_t)it.h * D;
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
      tc_fence_before();  // the zeroing of O for the next item is ordered after these reads
      // every MMA of this item that reads the Q buffer retired before the o_full waits above returned (the S MMAs
      // precede the P V in the tensor pipe): release it.  Eight arrivals per item; warpgroup A stands in for B when
      // the pair has one live query tile.
      if (lane == 0) {
        mbar_arrive(&q_empty[li & 1]);
        if (it.nx == 1) mbar_arrive(&q_empty[li & 1]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

template <int D, int kPoly>
static int launch_kt64p(const CUtensorMap* tm, const AttnParams& prm, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_flash_kt64p_kernel<D, kPoly>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmem);
    if (e != cudaSuccess) return fail(VSB_ERR_CUDA, "attn_flash(kt64p): smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  const long long total = (long long)prm.nb * prm.H * ((prm.nq + 255) / 256);
  if (total >= (1ll << 31)) return fail(VSB_ERR_UNSUPPORTED, "attn_flash(kt64p): too many query pairs");
  const int grid = total < num_sms() ? (int)total : num_sms();
  attn_flash_kt64p_kernel<D, kPoly><<<grid, kPThreads, kPSmem, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], prm);
  return check_launch("attn_flash(kt64p)");
}

int attn_flash_kt64p_launch(const CUtensorMap* tm, const AttnParams& prm, int D, int poly, cudaStream_t st) {
  if (D == 72) {
    switch (poly) {
      case 1: return launch_kt64p<72, 1>(tm, prm, st);
      case 2: return launch_kt64p<72, 2>(tm, prm, st);
      case 3: return launch_kt64p<72, 3>(tm, prm, st);
      default: return launch_kt64p<72, 0>(tm, prm, st);
    }
  }
  return poly ? launch_kt64p<64, 1>(tm, prm, st) : launch_kt64p<64, 0>(tm, prm, st);
}

}  // namespace vsb
