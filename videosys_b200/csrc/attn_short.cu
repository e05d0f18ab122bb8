// vsb200 -- short-sequence attention (n < 30 tokens): the temporal self-attention of STDiT3 (n = T = 15..20).
//
// FLOPs are negligible (3.7e11 per step at 720p); the kernel is HBM-bound on one read of q,k,v and one write
// of o.  One warp owns one (sequence, head): it stages q,k,v rows in shared memory as bf16, applies the
// per-head RMSNorm and RoPE in place, then lane i computes query row i with keys/values broadcast from smem.
// Rounding follows the reference's eager op order (attentions.py:111-120): bf16(q*scale), bf16(q@k^T),
// fp32 softmax, bf16(probs), bf16(probs@v).
#include "vsb_common.cuh"
#include "vsb_host.h"

namespace vsb {

constexpr int kMaxN = 32;
constexpr int kWarpsPerBlock = 4;

template <int D>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) attn_short_kernel(
    const bf16* __restrict__ qkv, bf16* __restrict__ out, const bf16* __restrict__ wq, const bf16* __restrict__ wk,
    const float* __restrict__ rope_cos, const float* __restrict__ rope_sin, int n_outer, int n_inner,
    long long outer_stride, long long inner_stride, long long tok_stride, int n, int H, float eps, float scale) {
  constexpr int VPR = D / 8;  // 16-byte vectors per head row
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  bf16* sq = reinterpret_cast<bf16*>(smem_raw) + (size_t)warp * 3 * kMaxN * D;
  bf16* sk = sq + kMaxN * D;
  bf16* sv = sk + kMaxN * D;
  const long long total = (long long)n_outer * n_inner * H;
  const int C = H * D;

  for (long long item = (long long)blockIdx.x * kWarpsPerBlock + warp; item < total;
       item += (long long)gridDim.x * kWarpsPerBlock) {
    const int h = int(item % H);
    const long long seq = item / H;
    const long long row0 = (seq / n_inner) * outer_stride + (seq % n_inner) * inner_stride;

    // ---- stage q,k,v rows (n x D each) ----
    for (int idx = lane; idx < 3 * n * VPR; idx += 32) {
      const int which = idx / (n * VPR);
      const int rem = idx - which * n * VPR;
      const int j = rem / VPR, c = rem - j * VPR;
      const bf16* src = qkv + ((size_t)(row0 + (long long)j * tok_stride) * 3 + which) * C + (size_t)h * D + c * 8;
      *reinterpret_cast<uint4*>(sq + (size_t)which * kMaxN * D + j * D + c * 8) =
          __ldg(reinterpret_cast<const uint4*>(src));
    }
    __syncwarp();

    if (n == 1) {  // attentions.py:65-66: x = v
      for (int c = lane; c < VPR; c += 32)
        *reinterpret_cast<uint4*>(out + (size_t)row0 * C + (size_t)h * D + c * 8) =
            *reinterpret_cast<const uint4*>(sv + c * 8);
      __syncwarp();
      continue;
    }

    // ---- RMSNorm(q), RMSNorm(k) (+RoPE) in place: lane i < n handles row i of q, lane 16+... would idle, so
    //      do q rows then k rows with the same lanes ----
    if (lane < n) {
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        bf16* r = (which ? sk : sq) + lane * D;
        const bf16* w = which ? wk : wq;
        float ss = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          float f = __bfloat162float(r[d]);
          ss += f * f;
        }
        const float rs = rsqrtf(ss / (float)D + eps);
#pragma unroll
        for (int d = 0; d < D; d += 2) {
          // normalization.py:28-33: h = bf16(x*rstd); y = bf16(w*h)
          float y0 = rbf(__bfloat162float(w[d]) * rbf(__bfloat162float(r[d]) * rs));
          float y1 = rbf(__bfloat162float(w[d + 1]) * rbf(__bfloat162float(r[d + 1]) * rs));
          if (rope_cos != nullptr) {
            // rotate_queries_or_keys: out = t*cos + rotate_half(t)*sin in fp32, pairs (2i, 2i+1): rot = (-x2, x1)
            const float c0 = rope_cos[lane * D + d], c1 = rope_cos[lane * D + d + 1];
            const float s0 = rope_sin[lane * D + d], s1 = rope_sin[lane * D + d + 1];
            const float o0 = __fadd_rn(__fmul_rn(y0, c0), __fmul_rn(-y1, s0));
            const float o1 = __fadd_rn(__fmul_rn(y1, c1), __fmul_rn(y0, s1));
            y0 = rbf(o0);
            y1 = rbf(o1);
          }
          if (which == 0) {  // q = bf16(q * scale)  (attentions.py:113)
            y0 = rbf(y0 * scale);
            y1 = rbf(y1 * scale);
          }
          r[d] = __float2bfloat16_rn(y0);
          r[d + 1] = __float2bfloat16_rn(y1);
        }
      }
    }
    __syncwarp();

    // ---- lane i: scores over keys, softmax, PV ----
    if (lane < n) {
      float qreg[D];
#pragma unroll
      for (int d = 0; d < D; ++d) qreg[d] = __bfloat162float(sq[lane * D + d]);
      float s[kMaxN];
      float mx = -INFINITY;
#pragma unroll 1
      for (int j = 0; j < n; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) acc = fmaf(qreg[d], __bfloat162float(sk[j * D + d]), acc);
        acc = rbf(acc);  // bf16 matmul output
        s[j] = acc;
        mx = fmaxf(mx, acc);
      }
      float den = 0.f;
#pragma unroll 1
      for (int j = 0; j < n; ++j) {
        s[j] = expf(s[j] - mx);
        den += s[j];
      }
      float o[D];
#pragma unroll
      for (int d = 0; d < D; ++d) o[d] = 0.f;
#pragma unroll 1
      for (int j = 0; j < n; ++j) {
        const float p = rbf(s[j] / den);  // softmax in fp32, cast to bf16
#pragma unroll
        for (int d = 0; d < D; ++d) o[d] = fmaf(p, __bfloat162float(sv[j * D + d]), o[d]);
      }
      bf16* dst = out + (size_t)(row0 + (long long)lane * tok_stride) * C + (size_t)h * D;
#pragma unroll
      for (int d = 0; d < D; d += 8) {
        uint4 u;
        u.x = pack_bf16x2(o[d], o[d + 1]);
        u.y = pack_bf16x2(o[d + 2], o[d + 3]);
        u.z = pack_bf16x2(o[d + 4], o[d + 5]);
        u.w = pack_bf16x2(o[d + 6], o[d + 7]);
        *reinterpret_cast<uint4*>(dst + d) = u;
      }
    }
    __syncwarp();
  }
}

}  // namespace vsb

using namespace vsb;

extern "C" int vsb_attn_short(const vsb_bf16* qkv, vsb_bf16* out, const vsb_bf16* wq, const vsb_bf16* wk,
                              const float* rope_cos, const float* rope_sin, int n_outer, int n_inner,
                              long long outer_stride, long long inner_stride, long long tok_stride, int n, int H,
                              int D, float eps, float scale, void* stream) {
  if (!qkv || !out || !wq || !wk || n_outer <= 0 || n_inner <= 0 || n <= 0 || H <= 0)
    return fail(VSB_ERR_INVALID, "attn_short: bad args");
  if (n > kMaxN) return fail(VSB_ERR_UNSUPPORTED, "attn_short: n=%d > %d (use vsb_attn_flash)", n, kMaxN);
  if ((rope_cos == nullptr) != (rope_sin == nullptr)) return fail(VSB_ERR_INVALID, "attn_short: rope tables");
  if (!aligned16(qkv) || !aligned16(out)) return fail(VSB_ERR_UNSUPPORTED, "attn_short: alignment");
  const long long total = (long long)n_outer * n_inner * H;
  long long blocks = (total + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (D == 72) {
    size_t smem = (size_t)kWarpsPerBlock * 3 * kMaxN * 72 * sizeof(bf16);
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(attn_short_kernel<72>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      attr = true;
    }
    attn_short_kernel<72><<<(int)blocks, kWarpsPerBlock * 32, smem, st>>>(
        (const bf16*)qkv, (bf16*)out, (const bf16*)wq, (const bf16*)wk, rope_cos, rope_sin, n_outer, n_inner,
        outer_stride, inner_stride, tok_stride, n, H, eps, scale);
  } else if (D == 64) {
    size_t smem = (size_t)kWarpsPerBlock * 3 * kMaxN * 64 * sizeof(bf16);
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(attn_short_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      attr = true;
    }
    attn_short_kernel<64><<<(int)blocks, kWarpsPerBlock * 32, smem, st>>>(
        (const bf16*)qkv, (bf16*)out, (const bf16*)wq, (const bf16*)wk, rope_cos, rope_sin, n_outer, n_inner,
        outer_stride, inner_stride, tok_stride, n, H, eps, scale);
  } else {
    return fail(VSB_ERR_UNSUPPORTED, "attn_short: head_dim %d (72 or 64 only)", D);
  }
  return check_launch("attn_short");
}
