/*
 * univst_b200 -- C ABI of the B200 (sm_100a) kernels behind UniVST's three-branch DDIM denoising hot path.
 *
 * The reference (QuanjianSong/UniVST) is pure Python on top of torch/diffusers: it has no FFI of its own.
 * Every entry point below therefore names the reference *Python call site* it replaces (file:line relative to
 * the reference checkout); INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns 0 (UNIVST_OK) or a negative error code and never throws;
 *     univst_last_error() returns a human-readable message for the last failure on the calling thread;
 *   - all tensor pointers are DEVICE pointers owned by the caller, fp16 unless the name says otherwise,
 *     16-byte aligned; activations are frames-major / channels-last: [images, H, W, C] == [tokens, C];
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - thread-compatible: the only process-wide state is a cache of device properties, the opt-in split-K switch and
 *     workspace registry (univst_gemm_tune / univst_gemm_set_workspace, keyed by stream) and a ring of arrival counters of
 *     the GroupNorm statistics kernel in device memory; nothing else outlives a call.
 */
#ifndef UNIVST_B200_H_
#define UNIVST_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UNIVST_OK 0
#define UNIVST_ERR_INVALID (-1)  /* bad argument / unsupported shape */
#define UNIVST_ERR_CUDA (-2)     /* a CUDA runtime / driver call failed */
#define UNIVST_ERR_NO_DEVICE (-3)/* not running on an sm_100 device */

#define UNIVST_ABI_VERSION 1

int univst_abi_version(void);
const char* univst_last_error(void);
/* 0 when the current device is compute capability 10.x, UNIVST_ERR_NO_DEVICE otherwise. */
int univst_device_check(void);

/* ------------------------------------------------------------------------------------------------------------
 * Fused epilogue of the tensor-core GEMM / implicit-GEMM convolution:
 *   y = (acc + bias[col] + rowvec[row / rows_per_group][col] + residual[row][col]) * out_scale
 *   geglu != 0 : the weight rows were interleaved per tile so that each BN-wide tile holds BN/2 value columns
 *                followed by BN/2 gate columns; y = (val + bias) * gelu_erf(gate + bias), N_out = N / 2
 *   bias2      : y = fp16(y) + bias2[col]   (the algebraically dead temporal attention of the SD backbone,
 *                models/attention.py:331-346 -- to_out weight is zero so it reduces to its bias)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct univst_epilogue {
  const void* bias;     /* fp16 [N] or NULL */
  const void* rowvec;   /* fp16 [M / rows_per_group, rowvec_ld] or NULL (time-embedding projection per branch) */
  int32_t rows_per_group;
  int32_t rowvec_ld;    /* row stride of rowvec in halves */
  int32_t act;          /* 1: y = silu(fp16(y)) (time-embedding MLP, unet_3d_condition.py:359-365, resnet.py:355);
                           2: y = gelu_tanh(fp16(y)) (feed-forward of the SD3 MMDiT blocks, third-party diffusers) */
  const void* residual; /* fp16 [M, ldr] or NULL */
  int32_t ldr;
  const void* bias2;    /* fp16 [N_out] or NULL */
  int32_t geglu;
  float out_scale;      /* 1 / output_scale_factor (resnet.py:392) */
} univst_epilogue_t;

/* D[M, N_out] = epilogue( [A | A2][M, K] * W[N, K]^T ).  A covers K columns [0, K1), A2 (optional) [K1, K):
 * the skip-connection concat of the up blocks (unet_3d_blocks.py:523) is folded into the K loop.
 * Replaces nn.Linear / 1x1 PseudoConv3d call sites: attention.py:123,127,141,143 (proj_in/out),
 * pnp_utils.py:39-43,97 and attention.py:375-377,425 (to_q/k/v/out), diffusers FeedForward (attention.py:329),
 * resnet.py:390 (conv_shortcut). */
int univst_gemm_f16(const void* A, int32_t lda, const void* A2, int32_t lda2, int32_t K1, const void* W, int32_t M,
                    int32_t N, int32_t K, void* D, int32_t ldd, const univst_epilogue_t* ep, void* stream);
/* Split-K for launches with few output tiles and a long reduction (the deep UNet levels when a frame shard holds only a
 * few images: 5 tiles x 180 k-blocks would keep 5 of 148 SMs busy).  Opt-in: univst_gemm_tune(max_tiles) splits the
 * K loop of GEMM / conv launches with at most max_tiles tiles (0 = never, the default) over the idle SMs; the slices park
 * their fp32 accumulators in a caller-owned workspace registered for the launching stream (first 256 KiB = arrival
 * counters, zeroed by the caller once); then every slice adds up ITS share of the tile over all slices in slice order and
 * runs the epilogue on it (a reduce-scatter through L2) -- deterministic, but the fp32 summation order differs from the
 * unsplit kernel's.  No workspace registered for a stream: no split on it. */
int univst_gemm_set_workspace(void* ws, int64_t bytes, void* stream);
int univst_gemm_tune(int32_t splitk_max_tiles);

/* Y[NB*H*W, Cout] = epilogue( conv3x3(X, Wt[Cout, 3, 3, C1 + C2], padding 1) ), H x W = OUTPUT size (power-of-two sizes up to W = 128 tile exactly; any other size takes row-block / row-segment tiles).
 * Implicit GEMM: the 9 taps are 4-D TMA boxes over the NHWC activation with out-of-bounds zero fill (no im2col).
 *   stride 1: X is [NB, H, W, C1]; X2 (optional) is [NB, H, W, C2] -- the skip-connection concat of the up blocks
 *             (unet_3d_blocks.py:523,618) is folded into the K loop, the concatenated tensor never exists.
 *   stride 2: X holds the four input parity planes [4 = (row parity, col parity)][NB, H, W, C1] produced by
 *             univst_space_to_depth2_f16 (downsamplers, resnet.py:216, padding 1).
 * Replaces PseudoConv3d.forward (resnet.py:57-80; the temporal Conv1d is a Dirac identity and is skipped). */
int univst_conv3x3_f16(const void* X, const void* X2, int32_t NB, int32_t H, int32_t W, int32_t C1, int32_t C2,
                       const void* Wt, int32_t Cout, int32_t stride, void* Y, int32_t ldy,
                       const univst_epilogue_t* ep, void* stream);
/* The same implicit GEMM with other tap tables (the VAE around the loop, SURVEY.md 8f row 2 -- third-party diffusers
 * AutoencoderKLTemporalDecoder, called at stable_diffusion.py:385,810,830 and ddim_inversion.py:29):
 *   univst_conv3x3_s2_pad_after_f16: 3x3, stride 2, one zero row / column AFTER the image (Downsample2D(padding=0) pads
 *     (0, 1, 0, 1) before its stride-2 conv); X = the four parity planes, H x W = output size;
 *   univst_conv_temporal3_f16: the (3, 1, 1) temporal convolutions of the temporal decoder.  X: [NB clips, F frames, HW
 *     pixels, C] channels-last, Wt: [Cout, 3, C] (tap t = frame offset t - 1), zero frames beyond the clip's ends. */
int univst_conv3x3_s2_pad_after_f16(const void* X, int32_t NB, int32_t H, int32_t W, int32_t C, const void* Wt, int32_t Cout,
                                    void* Y, int32_t ldy, const univst_epilogue_t* ep, void* stream);
int univst_conv_temporal3_f16(const void* X, int32_t NB, int32_t F, int32_t HW, int32_t C, const void* Wt, int32_t Cout,
                              void* Y, int32_t ldy, const univst_epilogue_t* ep, void* stream);
/* Small kernels of the VAE legs: in-place softmax(scale * x) over the columns of every row (the single-head mid-block
 * attention: head dim 512 runs as QK^T GEMM -> softmax -> PV GEMM); decoder rows -> uint8 pixels with the reference's
 * rounding (stable_diffusion.py:812-814); uint8 pixels -> zero-padded encoder input rows (:826-827); KL posterior sample
 * (or mode when noise == NULL) of [mean | logvar] moment rows, scaled, in the (C, F, hw) latent layout. */
int univst_softmax_rows_f16(void* X, int32_t ld, int32_t rows, int32_t cols, float scale, void* stream);
int univst_frames_to_u8(const void* X, int32_t ld, int64_t pixels, void* out, void* stream);
int univst_u8_to_frames_f16(const void* in, int64_t pixels, int32_t cpad, void* out, void* stream);
int univst_vae_sample_f16(const void* moments, int32_t ld, const void* noise, int32_t C, int32_t F, int32_t HW, float scaling,
                          void* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused sparse-causal attention: O[img] = softmax(Q[img] K^T / sqrt(d)) V, where the K/V sequence of image `img` is
 * the concatenation of the K/V of the images kv_src[img * nsrc + 0 .. nsrc) (never materialised: every KV tile is a
 * TMA box over the source image).  Q: [NI * N, ldq] with head h at columns [h d, (h+1) d); K, V: [NIkv * Nkv, ldkv].
 * kv_src is a DEVICE int32 array.  Replaces pnp_utils.py:59-92 (patched attn1, sources [max(f-1,0), 0]),
 * models/attention.py:384-420 (stock SparseCausalAttention, sources [max(f-1,0), f, 0]) and the cross-attention
 * SDPA of attention.py:316-323 (one shared 77-token source).
 * ---------------------------------------------------------------------------------------------------------- */
int univst_sc_attention_f16(const void* Q, int32_t ldq, const void* K, const void* V, int32_t ldkv, int32_t NI,
                            int32_t NIkv, int32_t H, int32_t d, int32_t N, int32_t Nkv, const int32_t* kv_src,
                            int32_t nsrc, void* O, int32_t ldo, void* stream);
/* Frame-sharded form (one clip over several GPUs): the K/V sources "previous frame" of a shard's first
 * frame and "first frame of the clip" live on other ranks.  Instead of exchanging them (NCCL send/recv + broadcast) the
 * kernel's TMA producer reads those tiles straight from PEER memory over NVLink (the buffers must be P2P-mapped, e.g.
 * torch symmetric memory) while the tensor pipe works on the previous tile -- the halo transfer is the attention
 * kernel's own loads.  Source index NI + b names image b*Fl + Fl-1 of (K_prev, V_prev) = the previous rank's buffer,
 * NI + B + b names image b*Fl of (K_first, V_first) = rank 0's buffer; all buffers share ldkv and the [NI*N, .] row
 * layout.  The caller orders the launch after the peers' projections (a cross-rank barrier on the stream). */
int univst_sc_attention_sharded_f16(const void* Q, int32_t ldq, const void* K, const void* V, int32_t ldkv, int32_t NI,
                                    int32_t H, int32_t d, int32_t N, const int32_t* kv_src, int32_t nsrc, void* O,
                                    int32_t ldo, const void* K_prev, const void* V_prev, const void* K_first,
                                    const void* V_first, int32_t B, int32_t Fl, void* stream);
/* Joint attention of the SD3 / SD3.5 MMDiT processors (backbones/video_diffusion_sd3/pnp_utils.py:9-132, :135-271): the
 * key / value sequence of an image is [K/V of source frames in the first tensor | the text tokens of a second tensor with
 * its own token count] -- source indices >= NIkv name image (src - NIkv) of (K2, V2).  Called once with the image tokens
 * as queries and once with the text tokens (the reference concatenates them into one query sequence, :100). */
int univst_joint_attention_f16(const void* Q, int32_t ldq, const void* K, const void* V, int32_t ldkv, int32_t NI,
                               int32_t NIkv, int32_t H, int32_t d, int32_t N, int32_t Nkv, const void* K2, const void* V2,
                               int32_t ldkv2, int32_t NIkv2, int32_t Nkv2, const int32_t* kv_src, int32_t nsrc, void* O,
                               int32_t ldo, void* stream);
/* Per-head RMS norm of the Q and K column blocks of a fused [rows, ld] = [Q | K | V] buffer, in place (attn.norm_q /
 * norm_k / norm_added_q / norm_added_k, sd3/pnp_utils.py:46-49, :92-95; diffusers RMSNorm semantics).  wq / wk: fp16 [d]
 * or NULL to skip that block. */
int univst_rmsnorm_heads_f16(void* QKV, int32_t ld, int32_t rows, int32_t H, int32_t d, const void* wq, const void* wk,
                             float eps, void* stream);
/* AdaIN-guided shift of the edit branch in the SD3 layout (sd3/pnp_utils.py:180-193 + attention_adain :287-300), in
 * place on the fused [3 F N, ld] buffer: style statistics per (frame, head, channel) over the tokens, content instance
 * norm over (tokens, head_dim) per (frame, head). */
int64_t univst_sd3_shift_workspace_bytes(int32_t F, int32_t C, int32_t d);
int univst_sd3_attn_shift_f16(void* QKV, int32_t ld, int32_t F, int32_t N, int32_t H, int32_t d, float alpha, float beta,
                              float gamma, void* workspace, void* stream);

/* Tuning hook (no reference counterpart): tile / exp2 variant of the head-dim <= 64 kernel, collapsing of repeated
 * K/V sources (exact: a source that occurs c times is streamed once with log2 c added to its scores) and the start
 * stagger of the softmax groups.  Negative values restore the defaults (environment UNIVST_ATTN_*). */
int univst_attention_tune(int32_t variant, int32_t dedupe, int32_t stagger);

/* Cross-attention over a short context (<= 80 tokens: the 77 CLIP tokens), one K/V "image" per branch named by
 * kv_src[NI].  Same math as univst_sc_attention_f16 with one source; a dedicated kernel because the context fits in a
 * few KB of shared memory (diffusers Attention / AttnProcessor2_0 at models/attention.py:316-323).
 * univst_cross_attention_supported(d, Nkv) tells whether the shape is covered (otherwise use univst_sc_attention_f16). */
int univst_cross_attention_supported(int32_t d, int32_t Nkv);
int univst_cross_attention_f16(const void* Q, int32_t ldq, const void* K, const void* V, int32_t ldkv, int32_t NI,
                               int32_t NIkv, int32_t H, int32_t d, int32_t N, int32_t Nkv, const int32_t* kv_src, void* O,
                               int32_t ldo, void* stream);

/* Temporal self-attention of the AnimateDiff motion modules: for every (branch, pixel, head) softmax over the F
 * frames.  QKV is the fused projection output [B F N, ld] (rows ordered branch, frame, pixel; Q at column 0, K at H d,
 * V at 2 H d), O [B F N, ldo].  Replaces VersatileAttention.forward, backbones/animatediff/models/motion_module.py:
 * 276-336 (including both "(b f) d c <-> (b d) f c" rearranges); F <= 32, d % 8 == 0. */
int univst_temporal_attention_f16(const void* QKV, int32_t ld, int32_t B, int32_t F, int32_t N, int32_t H, int32_t d,
                                  void* O, int32_t ldo, void* stream);

/* AdaIN-guided attention shift of the edit branch, in place on the fused [3 F N, ld] = [Q | K | V] buffer
 * (branch-major: 0 content, 1 style, 2 edit).  Replaces pnp_utils.py:47-57 + attention_adain :114-125. */
int64_t univst_attn_shift_workspace_bytes(int32_t F, int32_t C);
int univst_attn_shift_f16(void* QKV, int32_t ld, int32_t F, int32_t N, int32_t C, float alpha, float beta, float gamma,
                          void* workspace, void* stream);
/* The same with (alpha, beta, gamma) read from three floats in DEVICE memory: the launch no longer depends on the DDIM
 * step, so a captured UNet forward can be replayed for every step of the shift window (beta changes per step,
 * pnp_utils.py:50).  univst_set_floats writes up to 64 host floats (passed by value at launch) to device memory. */
int univst_attn_shift_dev_f16(void* QKV, int32_t ld, int32_t F, int32_t N, int32_t C, const float* abg, void* workspace,
                              void* stream);
int univst_set_floats(float* dst, const float* vals, int32_t n, void* stream);

/* GroupNorm(+SiLU) over [NB, rows, C1 + C2] channels-last (second source optional = skip-connection concat):
 * statistics per (batch, group) over rows x channels-per-group.  resnet.py:338,369 / unet_3d_condition.py:439
 * (NB = branches, rows = F H W: statistics span the frames) and attention.py:121 (NB = images, rows = H W). */
int64_t univst_groupnorm_workspace_bytes(int32_t NB, int32_t groups);
int univst_groupnorm_f16(const void* X1, const void* X2, int32_t C1, int32_t C2, int32_t NB, int32_t rows,
                         int32_t groups, const void* gamma, const void* beta, float eps, int32_t silu, void* Y,
                         void* workspace, void* stream);
/* Split form for frame-sharded execution (the reference's GroupNorm statistics span all frames, resnet.py:338): local
 * (sum, sum of squares) per (batch, group) into sums[NB, groups, 2]; the caller all-reduces them over the ranks and
 * applies with stat_rows = the global row count. */
int univst_groupnorm_stats_f16(const void* X1, const void* X2, int32_t C1, int32_t C2, int32_t NB, int32_t rows,
                               int32_t groups, float* sums, void* workspace, void* stream);
int univst_groupnorm_apply_f16(const void* X1, const void* X2, int32_t C1, int32_t C2, int32_t NB, int32_t rows,
                               int32_t groups, const float* sums, int64_t stat_rows, const void* gamma, const void* beta,
                               float eps, int32_t silu, void* Y, void* stream);
/* LayerNorm over the last axis of [rows, C] (attention.py:290,312,329). */
int univst_layernorm_f16(const void* X, int32_t rows, int32_t C, const void* gamma, const void* beta, float eps, void* Y,
                         void* stream);
/* The two elementwise pieces of the SD3 / SD3.5 MMDiT blocks around the (built) joint-attention processors -- third-party
 * diffusers JointTransformerBlock, which backbones/video_diffusion_sd3/models/transformer_3D_model.py:12-113 drives:
 * adaLN modulation  y = LayerNorm(x, no affine) * (1 + scale[s]) + shift[s]  with per-sample rows scale / shift [samples, ld]
 * (sample s = row / rows_per_sample), and the gated residual  out = x + gate[s] * y. */
int univst_layernorm_modulate_f16(const void* X, int32_t rows, int32_t C, const void* scale, const void* shift, int32_t ld,
                                  int32_t rows_per_sample, float eps, void* Y, void* stream);
int univst_gated_add_f16(const void* x, const void* y, const void* gate, int32_t ld_gate, int32_t rows_per_sample,
                         int64_t rows, int32_t C, void* out, void* stream);

/* Nearest x2 upsample of [NB, H, W, C] (resnet.py:145) and the parity-plane rearrangement feeding the stride-2 conv. */
int univst_upsample2x_f16(const void* X, int32_t NB, int32_t H, int32_t W, int32_t C, void* Y, void* stream);
int univst_space_to_depth2_f16(const void* X, int32_t NB, int32_t Ho, int32_t Wo, int32_t C, void* Y, void* stream);

/* (B, C, F, hw) latents <-> channels-last [(b f) hw, Cpad] (Z: HOST array of B device pointers). */
int univst_pack_latents_f16(const void* const* Z, int32_t B, int32_t C, int32_t F, int32_t HW, int32_t Cpad, void* out,
                            void* stream);
int univst_unpack_latents_f16(const void* X, int32_t ld, int32_t B, int32_t C, int32_t F, int32_t HW, void* Y,
                              void* stream);
/* diffusers Timesteps(dim, flip_sin_to_cos=True, freq_shift=0) for B fp32 timesteps (unet_3d_condition.py:359). */
int univst_timestep_embedding_f16(const float* t, int32_t B, int32_t dim, void* out, void* stream);

/* Per-step latent arithmetic of video_style_transfer (stable_diffusion.py:687-704, :761) on (C, F, hw) fp16 latents. */
int univst_mask_resize_u8(const uint8_t* mask, int32_t F, int32_t Hin, int32_t Win, int32_t Hout, int32_t Wout,
                          void* out, void* stream);
int univst_latent_blend_f16(const void* a, const void* b, const void* mask, int32_t C, int32_t F, int32_t HW, void* out,
                            void* stream);
int univst_latent_adain_f16(const void* cnt, const void* sty, int32_t C, int32_t F, int32_t HW, void* out, void* stream);
/* The blend on frame-major (F, C, hw) latents, the layout of the SD3 loop (video_diffusion_sd3/pipelines/
 * custom_pipeline.py:297-304, :316): out = (1 - m[f, hw]) a + m[f, hw] b.  That loop's latent_adain (sd3 pnp_utils.py:
 * 304-316, statistics per (frame, channel) plane) is univst_latent_adain_f16 with C = F * C planes and F = 1. */
int univst_latent_blend_fc_f16(const void* a, const void* b, const void* mask, int32_t F, int32_t C, int32_t HW, void* out,
                               void* stream);
/* DDIM step, eta = 0 (diffusers DDIMScheduler.step) and, with the alphas swapped, next_step of
 * inversion_tools/ddim_inversion.py:190-204.  eps is read from the channels-last conv_out buffer of `branch`. */
int univst_ddim_step_f16(const void* z, const void* eps_nhwc, int32_t ld, int32_t branch, int32_t C, int32_t F,
                         int32_t HW, float alpha_t, float alpha_prev, void* z_out, void* x0_out, void* stream);
/* Frames <-> pixels exchange of the frame-sharded AnimateDiff motion modules (the temporal attention of
 * backbones/animatediff/models/motion_module.py:279 runs over ALL frames of a pixel): the local [rows, C] activations
 * are stored straight into the ranks' symmetric-memory buffers (dst: HOST array of P device pointers, peers mapped over
 * NVLink) at their place in the owner's layout.  dir 0: rows (b, local frame, pixel) -> rank pixel / (N / P), row
 * (b, global frame, local pixel); dir 1: the inverse.  The caller places a cross-rank barrier after it. */
int univst_exchange_push_f16(int32_t dir, const void* src, int32_t ld, void* const* dst, int32_t rank, int32_t P, int32_t B,
                             int32_t Fl, int32_t N, int32_t C, void* stream);
/* The same with the cross-rank synchronisation (univst_xrank_*, below) as the tail of the kernel: no barrier launch. */
int univst_exchange_push_xrank_f16(int32_t dir, const void* src, int32_t ld, void* const* dst, void* const* ctl,
                                   int32_t rank, int32_t P, int32_t B, int32_t Fl, int32_t N, int32_t C, void* stream);
/* K/V halo of the frame-sharded sparse-causal attention (attention.py:395-410 needs frame f-1 and frame 0 of every
 * branch): nblk blocks of [rows, cols] halves (block b at src + b * src_blk_rows * ld_src) are stored into every non-null
 * dst[r] (HOST array of P device pointers into the ranks' symmetric-memory projection buffers; block b at
 * dst[r] + b * dst_blk_rows * ld_dst).  One read, up to P peer writes over NVLink; the caller places the barrier. */
int univst_halo_push_f16(const void* src, int32_t ld_src, int64_t src_blk_rows, void* const* dst, int32_t P, int32_t ld_dst,
                         int64_t dst_blk_rows, int32_t nblk, int32_t rows, int32_t cols, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Cross-rank synchronisation of the frame-sharded UNet over peer-mapped memory (no reference counterpart: the
 * reference is single-GPU; SURVEY.md 8e lists the exchanges frame sharding needs).  Every rank owns one control block
 * of univst_xrank_ctl_bytes() bytes, zero-initialised, in memory that ALL ranks have mapped (e.g. torch symmetric
 * memory); `ctl` is a HOST array of `world` device pointers, ctl[r] = rank r's block as mapped into this process.
 * A synchronisation stores this rank's next epoch into every peer's block and spins until every peer's epoch has
 * arrived in the local block; the epoch counter lives in the block and is advanced on the device, so launches take no
 * per-call host state and can be captured in a CUDA graph.  All ranks must issue the same sequence of synchronising
 * calls.  A wait that exceeds 10 s sets the sticky error word (uint32 at byte 72 of the local block) instead of hanging.
 * ---------------------------------------------------------------------------------------------------------- */
int64_t univst_xrank_ctl_bytes(void);
int32_t univst_xrank_slot_floats(void);
int univst_xrank_barrier(void* const* ctl, int32_t rank, int32_t world, void* stream);
/* Up to two block copies into peer memory (same block layout as univst_halo_push_f16; dst[r] = NULL skips rank r)
 * followed by the synchronisation as the tail of the same kernel: when it retires, the peers' copies of this step have
 * landed locally.  K/V halo of attn1 (attention.py:395-410): copy 0 = my last frame -> rank + 1's "previous frame" bank,
 * copy 1 (rank 0) = the clip's first frame -> every rank's "first frame" bank.  Also the gather of the noise prediction. */
typedef struct univst_push {
  const void* src;
  int32_t ld_src;
  int64_t src_blk_rows;
  void* dst[16];
  int32_t ld_dst;
  int64_t dst_blk_rows;
  int32_t nblk, rows, cols;
  void* mc_dst; /* optional: MULTICAST address of the destination (same offset in every rank's buffer, e.g. torch symmetric
                   memory's multicast_ptr): one multimem.st per 16 bytes, replicated by the NVSwitch, instead of one
                   store per peer -- the frame-0 K/V "broadcast" then costs rank 0 one copy of egress, not world - 1 */
} univst_push_t;
int univst_xrank_push_f16(const univst_push_t* pushes, int32_t npush, void* const* ctl, int32_t rank, int32_t world,
                          void* stream);
/* univst_groupnorm_f16 whose statistics span the rows of ALL ranks (resnet.py:338,369; unet_3d_condition.py:439 with the
 * frames sharded): chunk partials -> fold + store into every rank's slot + synchronise + add the slots in rank order
 * (bit-identical statistics on every rank) -> apply with rows * world rows.  Three launches, no collective call. */
int univst_groupnorm_xrank_f16(const void* X1, const void* X2, int32_t C1, int32_t C2, int32_t NB, int32_t rows,
                               int32_t groups, const void* gamma, const void* beta, float eps, int32_t silu, void* Y,
                               void* workspace, void* const* ctl, int32_t rank, int32_t world, void* stream);
int univst_axpby_f16(const void* a, const void* b, float wa, float wb, int64_t n, void* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Point-matching mask propagation (src/mask_propagation.py:72-83): aff = exp(<tar_n, src_m> / T) on L2-normalised
 * features; per target point keep the entries >= its topk-th largest (ties kept), normalise, transport the labels.
 * feat_tar [N, C], feat_src [C, M] (the reference's layouts), segs [Ccls, M] -> segs_tar [Ccls, N]; all fp32 device
 * pointers.  Optional outputs for index-set parity tests: thresholds [N] (the per-target kept threshold) and
 * kept_idx [N, kept_cap] (the kept source indices in ascending order, -1 padded).
 * ---------------------------------------------------------------------------------------------------------- */
int64_t univst_maskprop_workspace_bytes(int32_t N, int32_t C, int32_t M);
int univst_maskprop_f32(const float* feat_tar, const float* feat_src, const float* segs, int32_t N, int32_t C, int32_t M,
                        int32_t Ccls, float temperature, int32_t topk, float* segs_tar, float* thresholds,
                        int32_t* kept_idx, int32_t kept_cap, void* workspace, void* stream);

/* Sliding-window flow-warp smoothing, one key frame (src/cal_optica_flow.py:20-46; window loop of
 * pipelines/stable_diffusion.py:725-751).  frames: [F, H, W, 3] uint8, updated IN PLACE at `key`; for each of the
 * n <= 4 neighbours: fwd = flow(key -> neighbour), bwd = flow(neighbour -> key), [H, W, 2] fp32 device pointers
 * (neighbour_idx, fwd_flows, bwd_flows are HOST arrays).  Bit-exact with NumPy + cv2.remap(INTER_LINEAR, constant 0).
 * univst_mask_select_u8: out = keep_mask ? orig : est per pixel (stable_diffusion.py:751). */
int univst_flow_warp_key_u8(void* frames, int32_t F, int32_t H, int32_t W, int32_t key, int32_t n_neighbours,
                            const int32_t* neighbour_idx, const void* const* fwd_flows, const void* const* bwd_flows,
                            float threshold, void* stream);
int univst_mask_select_u8(const uint8_t* keep_mask, const void* orig, const void* est, int64_t npix, void* out,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNIVST_B200_H_ */
