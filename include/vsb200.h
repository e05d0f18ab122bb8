/* vsb200.h -- C-ABI of libvsb200.so: the B200-native (sm_100a) kernels behind the VideoSys DiT denoising hot path.
 *
 * The reference (NUS-HPC-AI-Lab/VideoSys @ 4cce4778) is pure Python and exposes no FFI/plugin boundary
 * (SURVEY.md section 8b); this header is the boundary a maintainer binds with ctypes (INTEGRATION.md shows the
 * stub).  Each entry cites the reference code it replaces (paths relative to /root/reference/videosys/).
 *
 * Conventions
 *   - plain pointers and sizes; every pointer is DEVICE memory unless named host_*; bf16 = uint16 storage.
 *   - all kernels are asynchronous on `stream` (a cudaStream_t passed as void*); no allocation, no host sync.
 *   - return 0 on success, a negative vsb_status otherwise; vsb_last_error() gives the text (per host thread).
 *   - no CPU fallback and no multi-backend dispatch: unsupported shape/alignment = VSB_ERR_UNSUPPORTED.
 *   - one process per GPU; entries are not re-entrant per device (same as one torch stream).
 */
#ifndef VSB200_H_
#define VSB200_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  VSB_OK = 0,
  VSB_ERR_INVALID = -1,     /* bad argument (null pointer, non-positive size) */
  VSB_ERR_UNSUPPORTED = -2, /* shape / alignment outside what the sm_100a kernels handle */
  VSB_ERR_CUDA = -3,        /* CUDA runtime / driver error (text in vsb_last_error) */
  VSB_ERR_NO_DEVICE = -4    /* no sm_100 device: the library never falls back to the CPU */
} vsb_status;

typedef uint16_t vsb_bf16;

int vsb_version(void);
const char* vsb_last_error(void);
/* Checks that `device` is compute capability 10.x and loads cuTensorMapEncodeTiled. */
int vsb_init(int device);
/* Number of kernels this library has launched since load (all entries); bench.py reports the delta. */
unsigned long long vsb_launch_count(void);
/* Encoded TMA tensor maps are cached by (base, shape, strides, box, swizzle): which = 0 -> hits, 1 -> misses (host
 * cuTensorMapEncodeTiled calls actually made).  Option "tmap_cache" (default 1) turns the cache off. */
unsigned long long vsb_tmap_cache_stats(int which);
/* Run-time kernel selection knobs (not a backend switch: every choice is an sm_100a kernel of this library).
 *   "gemm_2sm" (default 1): CTA-pair (tcgen05 cta_group::2, M = 256) GEMM for M >= 1024 when N % 192 == 0 or
 *   N % 256 == 0; 0 forces the single-CTA (M = 128) kernel everywhere. */
int vsb_set_option(const char* name, int value);
/*   "attn_variant" (default -1 = auto): schedule of vsb_attn_flash.  2 = 64-key tiles with a double-buffered S in TMEM
 *   and four MMA issuer warps, one CTA per pair of query tiles; 3 = the same tiles under persistent CTAs (one per SM,
 *   running ahead across query pairs; auto picks it for nk <= 1024, where per-CTA fixed costs dominate); 0 = 128-key
 *   tiles with the two softmax warpgroups ping-ponging (first version, kept for comparison).
 *   "attn_poly_exp" (default 0; variants 2, 3): 1 / 2 / 3 = 25 / 37.5 / 50 % of the exp2 as a polynomial on the FMA
 *   pipe (no gain measured at sustained clocks; kept as a knob).
 *   "attn_pingpong" (default 1, variant 0): the warpgroups alternate on the MUFU phase.
 *   Same results up to fp32 rounding.
 * vsb_debug_attn_trace: device buffer of 9*16*4 int64 that CTA (0,0,0) of vsb_attn_flash fills with clock64()
 * timestamps (profiling aid, NULL disables). */
int vsb_debug_attn_trace(void* device_buffer);
/*   "attn_variant" 4 = variant 2 with the query rows resident in TMEM (S = Q K^T issued as TS MMAs: the A operand no
 *   longer re-read from shared memory on every K step) and 8 instead of 6 K/V stages; 5 = 4 with the softmax row sum
 *   accumulated by the tensor core (ones written into the zero padding column d = 72 of every V tile; head_dim 72).
 *   6 = up to 320 keys (text cross-attention): K/V of a (batch, head) resident in shared memory, 160-key score tiles
 *   (auto for nk <= 320).
 *   "dsp_rowwise" (default 1): vsb_dsp_scatter decodes indices once per token row; 0 = the first version (per vector).
 *   "ln_occupancy" (default 3): 4 = vsb_ln_modulate compiled for 4 resident blocks per SM (64 registers).
 *   "tmap_cache" (default 1): cache of encoded TMA tensor maps. */

/* ---- AdaLN: LayerNorm(eps, no affine) -> x*(1+scale)+shift with per-frame t / t0 select --------------------
 * replaces norm1/norm2 + t2i_modulate + t_mask_select: models/transformers/open_sora_transformer_3d.py:47-48,
 * :152-160, :196-200, :260-264.  Rounds to bf16 at the same points as the eager chain (after LN, after 1+scale,
 * after the multiply, after the add).
 *   x, out   [B, T, S, C] bf16 (out may alias x); C % 8 == 0, C <= 3072
 *   mod      [2, B, 6, C] bf16: mod[0] = scale_shift_table + t, mod[1] = table + t0 (see vsb_modulation_table)
 *   x_mask   [B, T] uint8 (nonzero -> use mod[0]) or NULL (always mod[0])
 *   shift_row/scale_row: which of the 6 rows (0,1 for attention; 3,4 for the MLP) */
int vsb_ln_modulate(const vsb_bf16* x, vsb_bf16* out, const vsb_bf16* mod, const uint8_t* x_mask, int shift_row,
                    int scale_row, int B, int T, int S, int C, float eps, void* stream);

/* Same with nn.LayerNorm(C, eps, elementwise_affine=True) in front (gamma, beta [C]): the video / text streams of
 * CogVideoXLayerNormZero (models/modules/normalization.py:51-57): mod = the 6-way chunk of linear(silu(temb)) laid
 * out as [1 or 2, B, 6, C]; rows (0,1) video, (3,4) text. */
int vsb_ln_modulate_affine(const vsb_bf16* x, vsb_bf16* out, const vsb_bf16* mod, const uint8_t* x_mask,
                           const vsb_bf16* gamma, const vsb_bf16* beta, int shift_row, int scale_row, int B, int T, int S,
                           int C, float eps, void* stream);

/* mod[0,b,r,:] = bf16(table[r,:] + t[b, r*C:(r+1)*C]); mod[1] likewise from t0 (t0 may be NULL -> mod[1]=mod[0]).
 * replaces open_sora_transformer_3d.py:177-184.  table [6,C], t/t0 [B,6C], mod [2,B,6,C]. */
int vsb_modulation_table(const vsb_bf16* table, const vsb_bf16* t, const vsb_bf16* t0, vsb_bf16* mod, int B, int C,
                         int rows, void* stream);

/* ---- patch embedding of the latent + position embedding + sequence-parallel split ------------------------------
 * replaces open_sora_transformer_3d.py:568-572 (x_embedder: Conv3d kernel = stride = (1, ph, pw); rearrange; + pos_emb)
 * and :577 (split_sequence along the patch axis); the same for Latte's 2-D PatchEmbed (latte_transformer_3d.py:1245).
 *   z    latent, element (b, c, t, y, x) at b*batch_stride + c*chan_stride + t*frame_stride + y*W + x (elements);
 *        rows / columns beyond H / W read as zero (the reference pads to the patch grid)
 *   w    [C, Cin*ph*pw] (= conv.weight storage, taps ordered (c, ky, kx)); bias [C] or NULL; pos [S_total, C] or NULL
 *   out  [B, T, S_local, C]: patch columns s0 .. s0+S_local-1 of every (batch, frame); columns >= S_total are zero
 * Rounds like the eager chain: conv (fp32 accumulate) -> bf16, + bias -> bf16, + pos -> bf16.
 * Returns 1 without launching unless Cin*ph*pw == 16, C % 8 == 0 and C <= 2048. */
int vsb_patch_embed(const vsb_bf16* z, const vsb_bf16* w, const vsb_bf16* bias, const vsb_bf16* pos, vsb_bf16* out, int B,
                    int Cin, int T, int H, int W, long long batch_stride, long long chan_stride, long long frame_stride,
                    int ph, int pw, int C, int s0, int S_local, void* stream);

/* ---- gate * y (+ per-frame select) + residual, optional PAB cache write ---------------------------------------
 * replaces open_sora_transformer_3d.py:219-228 and :270-284.  gated = bf16(gate*y); out = bf16(x + gated).
 *   cache_out (nullable): receives `gated` (what the reference keeps as last_attn, :224-225). */
int vsb_gate_residual(const vsb_bf16* x, const vsb_bf16* y, vsb_bf16* out, vsb_bf16* cache_out, const vsb_bf16* mod,
                      const uint8_t* x_mask, int gate_row, int B, int T, int S, int C, void* stream);

/* out = bf16(x + y) over n elements: cross-attention residual (:240) and PAB replay of a cached tensor
 * (:192-193 -> :228, :234-235).  One coalesced pass: 2 reads + 1 write. */
int vsb_residual_add(const vsb_bf16* x, const vsb_bf16* y, vsb_bf16* out, size_t n, void* stream);

/* ---- per-head RMSNorm of q and k inside a packed qkv buffer (spatial blocks, no RoPE) -------------------------
 * replaces LlamaRMSNorm on q,k: models/modules/normalization.py:28-33 via attentions.py:75.
 *   qkv [rows, 3, H, D] bf16, normalised in place for the q and k thirds; wq, wk [D] bf16. */
int vsb_qk_rmsnorm(vsb_bf16* qkv, const vsb_bf16* wq, const vsb_bf16* wk, size_t rows, int H, int D, float eps,
                   void* stream);
/* Same, followed by RoPE on q and k (attentions.py:76-78 -> rotary_embedding_torch rotate_queries_or_keys: interleaved
 * pairs, fp32 math, cast back): the pre-pass of TEMPORAL attention over >= 30 frames, where the reference leaves
 * native_attention for F.scaled_dot_product_attention (attentions.py:95-100).  rope_cos / rope_sin [pos_mod, D] fp32;
 * token row r sits at position (r / pos_div) % pos_mod
 * (token-major [B, T, S] activation: pos_div = S, pos_mod = T).
 * wq == wk == NULL: RoPE only (q / k stay un-normalised): the temporal attention of Vchitect beyond 64 frames
 * (attentions.py:688-701 apply_rotary_emb, complex multiply on the same interleaved pairs; no q/k norm). */
int vsb_qk_rmsnorm_rope(vsb_bf16* qkv, const vsb_bf16* wq, const vsb_bf16* wk, size_t rows, int H, int D, float eps,
                        const float* rope_cos, const float* rope_sin, int pos_div, int pos_mod, void* stream);

/* Half-rotation RoPE on q and k in place (Open-Sora-Plan RoPE1D / RoPE2D / RoPE3D:
 * models/transformers/open_sora_plan_v110_transformer_3d.py:136-252 applied :1217-1232,
 * open_sora_plan_v120_transformer_3d.py:63-118 applied :921-925): every head is cut into D / (2*half) blocks (one per
 * position axis), inside a block out[d] = x[d]*cos + rotate_half(x)[d]*sin with rotate_half = (-second half, first half),
 * each product and the sum rounded to the 16-bit dtype as the reference's eager ops do.
 *   rope_cos, rope_sin_signed [pos_mod, D] fp32: row p holds, for every channel of a head, the reference's 16-bit cos / sin
 *   of that channel's axis position (sin negated on the first half of each block); token row r uses table row
 *   (r / pos_div) % pos_mod.  half even, D a multiple of 2*half. */
int vsb_qk_rope_halves(vsb_bf16* qkv, size_t rows, int H, int D, int half, const float* rope_cos,
                       const float* rope_sin_signed, int pos_div, int pos_mod, void* stream);
/* Per-head LayerNorm(D, eps, affine) of q and k in place (CogVideoX: diffusers Attention(qk_norm="layer_norm"),
 * models/transformers/cogvideox_transformer_3d.py:241-242 -> processor :130-133).  wq,bq,wk,bk [D] bf16. */
int vsb_qk_layernorm(vsb_bf16* qkv, const vsb_bf16* wq, const vsb_bf16* bq, const vsb_bf16* wk, const vsb_bf16* bk,
                     size_t rows, int H, int D, float eps, void* stream);

/* ---- short-sequence attention (n < 30): RMSNorm(q,k) -> [RoPE] -> native_attention, one warp per (seq, head) --
 * replaces OpenSoraAttention.forward's N<30 path: attentions.py:59-78,95-97,111-120 with the reference's op order
 * (bf16(q*scale), bf16 scores, fp32 softmax, bf16 probs).  Reads the packed qkv of the token-major activation
 * without any rearrange: sequence (o,i) token j lives at row o*outer_stride + i*inner_stride + j*tok_stride.
 *   qkv [rows,3,H,D] bf16; out [rows,H*D] bf16; rope_cos/rope_sin [n, D] fp32 or NULL; n == 1 copies v (:65-66).
 *   n <= 64: two instantiations, up to 32 tokens (OpenSora's 15 .. 20 frames) and 33 .. 64 (Vchitect's 40-frame temporal
 *   attention, models/modules/attentions.py:707-768, with flags = 3 and RoPE tables from freqs_cis).
 *   flags: bit 0 = no q/k RMSNorm (wq, wk may be NULL): diffusers Attention as used by Latte
 *          (models/transformers/latte_transformer_3d.py:259-268,611-619); bit 1 = SDPA rounding (fp32 scores, scale
 *          inside the softmax) instead of native_attention's bf16 intermediate steps. */
int vsb_attn_short(const vsb_bf16* qkv, vsb_bf16* out, const vsb_bf16* wq, const vsb_bf16* wk, const float* rope_cos,
                   const float* rope_sin, int n_outer, int n_inner, long long outer_stride, long long inner_stride,
                   long long tok_stride, int n, int H, int D, float eps, float scale, int flags, void* stream);

/* ---- GEMM on tcgen05: out[M,N] = act(A[M,K] @ W[N,K]^T + bias[N]) ---------------------------------------------
 * replaces every nn.Linear on the path (attentions.py:59,107,156-157; timm Mlp fc1/fc2) and the tanh-GELU between
 * fc1 and fc2 (models/modules/activations.py:3).  bf16 in, fp32 accumulate in TMEM, bf16 out.
 *   act: 0 = none, 1 = gelu_tanh applied to bf16(acc+bias) (the eager rounding point).
 *   Requirements: K % 8 == 0, N % 8 == 0, 16-byte aligned pointers, row-major contiguous. */
int vsb_gemm_bias_act(const vsb_bf16* A, const vsb_bf16* W, const vsb_bf16* bias, vsb_bf16* out, int M, int N, int K,
                      int act, void* stream);

/* Same GEMM with the residual branch of the block fused into the epilogue:
 *   y = bf16(A @ W^T + bias);  g = gate_row >= 0 ? bf16(gate * y) : y;  out = bf16(resid + g)
 * gate = mod[sel(b,t), b, gate_row, :] with the per-frame t/t0 select of x_mask (as in vsb_gate_residual); M == B*T*S.
 * replaces proj/fc2 Linear + open_sora_transformer_3d.py:219-228,:270-284 (gate_row 2 / 5) and cross proj + :240
 * (gate_row -1).  out may alias resid.  Returns 1 WITHOUT launching when the CTA-pair kernel does not take the shape
 * (M < 1024 or N not a multiple of 192/256): the caller then uses vsb_gemm_bias_act + vsb_gate_residual. */
int vsb_gemm_bias_residual(const vsb_bf16* A, const vsb_bf16* W, const vsb_bf16* bias, const vsb_bf16* resid,
                           vsb_bf16* out, const vsb_bf16* mod, const uint8_t* x_mask, int gate_row, int M, int N, int K,
                           int B, int T, int S, void* stream);

/* ---- flash attention on tcgen05 (spatial self-attention and text cross-attention) -------------------------------
 * replaces F.scaled_dot_product_attention at attentions.py:100 and :268 (bool key mask = per-batch key count).
 * q/k/v are strided views: element (b, n, h, d) at base + b*batch_stride + n*row_stride + h*D + d (strides in
 * elements, multiples of 8).  out [nb, nq, H*D] contiguous.  kv_lens (host int array, nullable) = valid keys per
 * batch (<= nk, nb <= 64 when given).  head_dim 72 / 64: the tcgen05 kernels; 96 (Open-Sora-Plan v1.2.0,
 * open_sora_plan_v120_transformer_3d.py:929-931), 128, 80, 48, 32: the same contract on the warp-level tensor path
 * (mma.sync, FlashAttention-2 schedule, csrc/attn_mma.cu). */
int vsb_attn_flash(const vsb_bf16* q, const vsb_bf16* k, const vsb_bf16* v, vsb_bf16* out, int nb, int nq, int nk,
                   int H, int D, long long q_row_stride, long long q_batch_stride, long long kv_row_stride,
                   long long kv_batch_stride, const int* host_kv_lens, float scale, void* stream);
/* Same with a strided OUTPUT view: element (b, n, h, d) at out + b*out_batch_stride + n*out_row_stride + h*D + d.
 * Temporal attention over >= 30 frames runs on the token-major activation with batch = patch s, row = frame t:
 * q/k/v row stride S*3C, batch stride 3C; out row stride S*C, batch stride C (one call per CFG sample). */
int vsb_attn_flash_strided(const vsb_bf16* q, const vsb_bf16* k, const vsb_bf16* v, vsb_bf16* out, int nb, int nq, int nk,
                           int H, int D, long long q_row_stride, long long q_batch_stride, long long kv_row_stride,
                           long long kv_batch_stride, long long out_row_stride, long long out_batch_stride,
                           const int* host_kv_lens, float scale, void* stream);

/* ---- Pyramid Attention Broadcast gate (host integer logic, bit-exact) -------------------------------------------
 * replaces PABManager.if_broadcast_{spatial,temporal,cross}: core/pab/pab_mgr.py:54-91.
 * Returns 1 (reuse the cached tensor) or 0; *count advances on every call and wraps modulo steps.
 * has_timestep = 0 models `timestep is None`. */
int vsb_pab_gate(int broadcast_on, int has_timestep, int timestep, int* count, int range, int lo, int hi, int steps);

/* ---- DSP reshard (dimension switch) over NVLink peer memory ----------------------------------------------------
 * replaces STDiT3Block.dynamic_switch -> all_to_all_with_pad -> _all_to_all_func:
 * open_sora_transformer_3d.py:288-315, core/distributed/comm.py:282-304,104-108.
 * One kernel packs each destination rank's slice and stores it straight into that rank's receive window with
 * 128-bit peer stores (no staging copies, zero padding synthesised on the fly), then signals a per-peer flag.
 *   peer_recv[r]   device pointer (peer-mapped) to rank r's receive window for this direction
 *   peer_flags[r]  device pointer (peer-mapped) to rank r's flag array [world] (uint32 epoch counters)
 *   to_spatial_shard = 0: local [B, T, Sl, C] (S-sharded) -> recv [B, Tp/sp, Sp, C] (T-sharded)
 *   to_spatial_shard = 1: local [B, Tl, S(+pad), C] -> recv [B, Tp, Sl, C]
 * vsb_dsp_wait blocks the stream until all `world` peers have delivered epoch `epoch`. */
int vsb_dsp_scatter(const vsb_bf16* local, void* const* host_peer_recv, void* const* host_peer_flags, int rank,
                    int world, int to_spatial_shard, int B, int T, int S, int C, unsigned epoch, void* stream);
int vsb_dsp_wait(void* my_flags, int world, unsigned epoch, void* stream);
/* epoch == 0 (both calls above and below): take the next value of a DEVICE-side counter kept in the flag array
 * (slots 32 = sent, 33 = waited), so a captured CUDA graph of a whole denoising step advances the epochs on every
 * replay without host involvement.  Do not mix explicit and device-side epochs on one flag array.
 *
 * Producer- / consumer-fused forms of the same switch (north_star: "in-kernel sequence-dim all-to-all issuing P2P stores
 * fused with the preceding [producer]"; in STDiT3 the producer of the S-shard -> T-shard switch is the AdaLN modulate,
 * open_sora_transformer_3d.py:196-216, and the consumer of the switch back is the gate + residual, :213-228):
 *
 * vsb_ln_modulate_dsp = vsb_ln_modulate whose store path IS the reshard: row (b, t, sl) goes to peer (b*T+t) / Tl,
 *   window viewed as [Tl, Sg, C], row ((b*T+t) % Tl, rank*Sl + sl), Tl = ceil(B*T / world) ((batch, frame) sequences
 *   are scattered, not frames: 2*20 = 40 split 8 ways exactly); zero rows for padded sequences, padded columns never
 *   sent; publishes the epoch like vsb_dsp_scatter.  Follow with vsb_dsp_wait on the same flag array.
 * vsb_dsp_signal publishes the next epoch without moving data (the proj GEMM wrote into this rank's OWN window).
 * vsb_gate_residual_dsp = vsb_gate_residual whose y operand is PULLED with 128-bit peer loads from the producers'
 *   windows host_peer_y[r] (each viewed as [Tl, Sg, C]); call after vsb_dsp_wait on the signalled flag array. */
int vsb_ln_modulate_dsp(const vsb_bf16* x, const vsb_bf16* mod, const uint8_t* x_mask, int shift_row, int scale_row,
                        int B, int T, int Sl, int C, float eps, void* const* host_peer_recv, void* const* host_peer_flags,
                        int rank, int world, int Sg, unsigned epoch, void* stream);
int vsb_dsp_signal(void* const* host_peer_flags, int rank, int world, unsigned epoch, void* stream);
int vsb_gate_residual_dsp(const vsb_bf16* x, void* const* host_peer_y, vsb_bf16* out, vsb_bf16* cache_out,
                          const vsb_bf16* mod, const uint8_t* x_mask, int gate_row, int B, int T, int Sl, int C, int rank,
                          int world, int Sg, void* stream);

/* Symmetric receive windows for the reshard: cudaMalloc'ed (zeroed) here so that a CUDA IPC handle maps the exact
 * base address; handles (64 bytes) are exchanged once at initialize() over the process group's store.
 * replaces nothing in the reference (NCCL owns its buffers); this is the B200-native plumbing for P2P stores. */
int vsb_dsp_alloc(void** out, size_t bytes);
int vsb_dsp_free(void* p);
int vsb_ipc_get_handle(void* devptr, void* out_handle64);
int vsb_ipc_open_handle(const void* handle64, void** out_devptr);
int vsb_ipc_close_handle(void* devptr);

/* ---- IEEE fp16 twins ----------------------------------------------------------------------------------------------
 * The reference runs CogVideoX-2b and Latte in torch.float16 (pipelines/cogvideox/pipeline_cogvideox.py:138-139,
 * pipelines/latte/pipeline_latte.py:201).  Every kernel entry above that touches activations exists a second time with
 * the suffix _f16: same arguments, same tile schedules, fp32 accumulation and rounding points, the 16-bit storage type
 * being IEEE half instead of bfloat16 (each kernel source is compiled twice: csrc/vsb_common.cuh, -DVSB_HALF).  The
 * DSP entries are bf16 only (sequence parallelism is an OpenSora path). */
typedef uint16_t vsb_f16;
int vsb_ln_modulate_f16(const vsb_f16* x, vsb_f16* out, const vsb_f16* mod, const uint8_t* x_mask, int shift_row,
                    int scale_row, int B, int T, int S, int C, float eps, void* stream);
int vsb_ln_modulate_affine_f16(const vsb_f16* x, vsb_f16* out, const vsb_f16* mod, const uint8_t* x_mask,
                           const vsb_f16* gamma, const vsb_f16* beta, int shift_row, int scale_row, int B, int T, int S,
                           int C, float eps, void* stream);
int vsb_modulation_table_f16(const vsb_f16* table, const vsb_f16* t, const vsb_f16* t0, vsb_f16* mod, int B, int C,
                         int rows, void* stream);
int vsb_patch_embed_f16(const vsb_f16* z, const vsb_f16* w, const vsb_f16* bias, const vsb_f16* pos, vsb_f16* out, int B,
                    int Cin, int T, int H, int W, long long batch_stride, long long chan_stride, long long frame_stride,
                    int ph, int pw, int C, int s0, int S_local, void* stream);
int vsb_gate_residual_f16(const vsb_f16* x, const vsb_f16* y, vsb_f16* out, vsb_f16* cache_out, const vsb_f16* mod,
                      const uint8_t* x_mask, int gate_row, int B, int T, int S, int C, void* stream);
int vsb_residual_add_f16(const vsb_f16* x, const vsb_f16* y, vsb_f16* out, size_t n, void* stream);
int vsb_qk_rmsnorm_f16(vsb_f16* qkv, const vsb_f16* wq, const vsb_f16* wk, size_t rows, int H, int D, float eps,
                   void* stream);
int vsb_qk_rmsnorm_rope_f16(vsb_f16* qkv, const vsb_f16* wq, const vsb_f16* wk, size_t rows, int H, int D, float eps,
                        const float* rope_cos, const float* rope_sin, int pos_div, int pos_mod, void* stream);
int vsb_qk_rope_halves_f16(vsb_f16* qkv, size_t rows, int H, int D, int half, const float* rope_cos,
                           const float* rope_sin_signed, int pos_div, int pos_mod, void* stream);
int vsb_qk_layernorm_f16(vsb_f16* qkv, const vsb_f16* wq, const vsb_f16* bq, const vsb_f16* wk, const vsb_f16* bk,
                     size_t rows, int H, int D, float eps, void* stream);
int vsb_attn_short_f16(const vsb_f16* qkv, vsb_f16* out, const vsb_f16* wq, const vsb_f16* wk, const float* rope_cos,
                   const float* rope_sin, int n_outer, int n_inner, long long outer_stride, long long inner_stride,
                   long long tok_stride, int n, int H, int D, float eps, float scale, int flags, void* stream);
int vsb_gemm_bias_act_f16(const vsb_f16* A, const vsb_f16* W, const vsb_f16* bias, vsb_f16* out, int M, int N, int K,
                      int act, void* stream);
int vsb_gemm_bias_residual_f16(const vsb_f16* A, const vsb_f16* W, const vsb_f16* bias, const vsb_f16* resid,
                           vsb_f16* out, const vsb_f16* mod, const uint8_t* x_mask, int gate_row, int M, int N, int K,
                           int B, int T, int S, void* stream);
int vsb_attn_flash_f16(const vsb_f16* q, const vsb_f16* k, const vsb_f16* v, vsb_f16* out, int nb, int nq, int nk,
                   int H, int D, long long q_row_stride, long long q_batch_stride, long long kv_row_stride,
                   long long kv_batch_stride, const int* host_kv_lens, float scale, void* stream);
int vsb_attn_flash_strided_f16(const vsb_f16* q, const vsb_f16* k, const vsb_f16* v, vsb_f16* out, int nb, int nq, int nk,
                           int H, int D, long long q_row_stride, long long q_batch_stride, long long kv_row_stride,
                           long long kv_batch_stride, long long out_row_stride, long long out_batch_stride,
                           const int* host_kv_lens, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VSB200_H_ */
