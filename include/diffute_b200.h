/* diffute_b200.h — C-ABI of the B200-native DiffUTE sampling kernels (libdiffute_b200.so).
 *
 * Drop-in boundary.  chenhaoxing/DiffUTE has no native code and no FFI: its sampling path calls the
 * Python objects of the third-party `diffusers` package (app.ipynb:545-553 loads them, app.ipynb:772-819
 * runs them).  The Python classes in diffute_b200/ mirror those call signatures; every device operation
 * they perform goes through the entry points below.  Each entry point names the reference call site(s)
 * whose arithmetic it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises the host,
 *     never allocates: scratch memory is passed by the caller (`workspace`);
 *   - return 0 on success, <0 on error (DFU_ERR_*); dfu_last_error() gives the text.  Nothing throws;
 *   - activations are NHWC ("channels last"): an image tensor [B,H,W,C] is also the token tensor [B,H*W,C];
 *   - "f16 operand" tensors hold `planes` consecutive copies: plane 0 = RN fp16 of the value, plane 1
 *     (only in DFU_PREC_FP16X2 mode) = RN fp16 of the rounding residual.  Contractions run on tcgen05
 *     tensor cores with fp32 accumulation in TMEM, 1 pass (FP16) or 3 passes (FP16X2: hi*hi + lo*hi + hi*lo).
 */
#ifndef DIFFUTE_B200_H_
#define DIFFUTE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFU_OK 0
#define DFU_ERR_INVALID (-1)
#define DFU_ERR_CUDA (-2)
#define DFU_ERR_DRIVER (-3)
#define DFU_ERR_WORKSPACE (-4)

/* ---- library ------------------------------------------------------------------------------- */
int dfu_version(void);
const char* dfu_last_error(void);
/* Number of SMs of the current device (<0 on error). */
int dfu_num_sms(void);

/* ---- tensor-core contraction core -------------------------------------------------------------
 * D[M,N] = sum over operand groups g, passes p, taps t, channel chunks c of  A_g[.,t,c] * B_g[.,t,c]^T
 * Replaces every F.conv2d (3x3, 1x1, stride 1/2, fused nearest-2x upsample via a pre-expanded operand,
 * fused skip-concat via a concatenated operand, fused 1x1 shortcut as a second operand group) and
 * F.linear inside diffusers' UNet2DConditionModel / AutoencoderKL:
 *   reference call sites app.ipynb:814 (unet), :793 (vae.encode), :819 (vae.decode);
 *   module math SURVEY.md A.1/A.2 (ResnetBlock2D conv1/conv2/conv_shortcut, Attention to_q/k/v/out,
 *   Transformer2DModel proj_in/out, FeedForward GEGLU/out, Down/Upsample2D conv).
 */
typedef struct DfuGemmOperand {
  const void* a;        /* f16 activations */
  int32_t a_mode;       /* 0: matrix [rows, a_ld]; 1: NHWC image [imgs, a_h, a_w, a_c] walked by taps */
  int32_t a_rows;       /* mode 0: total rows incl. planes; mode 1: total images incl. planes */
  int32_t a_ld;         /* mode 0: row stride in elements (>= K, multiple of 8) */
  int32_t a_h, a_w, a_c;/* mode 1: stored image extent and channels (a_c multiple of 64) */
  int32_t a_plane;      /* rows (mode 0) or images (mode 1) between the hi and lo planes */
  const void* b;        /* f16 weights, K-major: [b_rows, b_ld] */
  int32_t b_rows;       /* total rows incl. planes */
  int32_t b_ld;         /* row stride in elements (>= ntaps * k_per_tap, multiple of 8) */
  int32_t b_plane;      /* rows between the hi and lo planes */
  int32_t ntaps;        /* 1 (linear / 1x1) or 9 */
  int32_t k_per_tap;    /* contraction length per tap (multiple of 64) */
  int8_t tap_dn[9];     /* per tap: image offset (parity plane for stride-2), row offset, column offset */
  int8_t tap_dy[9];
  int8_t tap_dx[9];
  int8_t b_static;      /* 1: B is a constant (weights) — its tiles may be fetched before the kernel producing A ends */
  int8_t _pad[4];
} DfuGemmOperand;

#define DFU_EPI_F32 0     /* out_f32[m, n] = v */
#define DFU_EPI_F16 1     /* out_f16 planes [m, n] = split(v) */
#define DFU_EPI_GEGLU 2   /* weights packed in 32-row blocks (16 value rows, 16 gate rows):
                             out_f16[m, n/2] = split((a+bias_a) * gelu_erf(g+bias_g)) */

typedef struct DfuGemm {
  int32_t m, n;            /* output rows (B*H*W for conv) and columns */
  int32_t ngroups;         /* 1 or 2 */
  int32_t npass;           /* 1 (DFU_PREC_FP16) or 3 (DFU_PREC_FP16X2) */
  DfuGemmOperand g[2];
  int32_t batch;           /* 0/1: one problem.  > 1: `batch` independent products of m / batch rows each (matrix operands,
                              one group): batch b uses A rows [b*a_batch_rows, ...), B rows [b*b_batch_rows, ...) and
                              writes output / residual rows [b*m/batch, (b+1)*m/batch) — the per-sample Q K^T and P V of the
                              VAE's single-head attention in ONE launch */
  int32_t a_batch_rows, b_batch_rows;
  int32_t conv;            /* 1: rows are output pixels of a [B,H,W] grid (mode-1 operands) */
  int32_t B, H, W;         /* output grid when conv=1 */
  /* epilogue: v = alpha*acc + bias[n] + rowvec[(m / rows_per_sample), n] + residual[m, n] */
  int32_t epi;             /* DFU_EPI_* */
  int32_t act;             /* 0: none; 1: exact-erf GELU applied to v before it is stored (DFU_EPI_F32 / DFU_EPI_F16):
                              the non-gated MLP of the TrOCR ViT glyph encoder (app.ipynb:546-548, :773-776) */
  float alpha;
  const float* bias;       /* [n] or NULL (packed like the weights for GEGLU) */
  const float* rowvec;     /* [samples, rowvec_ld] or NULL (time-embedding add) */
  int32_t rowvec_ld;
  int32_t rows_per_sample;
  const float* residual;   /* [m, ldr] fp32 or NULL */
  int32_t ldr;
  float* out_f32;          /* DFU_EPI_F32 */
  int32_t ldo;
  void* out_f16;           /* DFU_EPI_F16 / GEGLU: [planes][m][ldh] */
  int32_t ldh;
  int32_t out_planes;      /* 1 or 2 */
  int64_t out_plane_stride;/* elements between planes */
  /* tiling (0 = choose automatically) */
  int32_t block_n;         /* multiple of 16 (32 for GEGLU), <= 256, divides n */
  int32_t splits;          /* split-K factor */
  int32_t stages;
  int32_t kernel;          /* 0: choose; 1: one 128 x block_n tile per CTA (split-K slices reduce inside a thread-block
                              cluster); 2: persistent CTA pairs — tcgen05 cta_group::2, 256 x block_n tiles, two TMEM
                              accumulators so a tile's epilogue overlaps the next tile's main loop (splits must be 1) */
  void* workspace;         /* fp32 [splits, m, n] when splits > 1 */
  size_t workspace_bytes;
  void* sync_words;        /* reserved (accepted and ignored): backed a grid-barrier second stage that was measured
                              slower than the PDL-overlapped reduce launch and removed */
} DfuGemm;

int dfu_gemm(const DfuGemm* desc, void* stream);
/* The tiling dfu_gemm would choose: out[8] = {block_n, splits, stages, tiles_m, tiles_n, k_blocks, kernel (1|2), 0}. No GPU needed. */
int dfu_gemm_plan(const DfuGemm* desc, int32_t* out);
/* Process-wide counters: out[6] = {gemm launches, split-K launches, fused second stages, separate reduce launches,
 * persistent pair-kernel launches, last grid}. */
void dfu_gemm_stats(int64_t* out);
/* Workspace bytes dfu_gemm needs for this descriptor with automatic tiling (0 if none). */
size_t dfu_gemm_workspace(const DfuGemm* desc);

/* ---- normalisation / operand casts (memory-bound, fp32 NHWC in, fp16 operand planes out) ------
 * dfu_groupnorm: diffusers ResnetBlock2D.norm1/norm2 + SiLU, Transformer2DModel.norm, conv_norm_out, VAE
 * group_norm (reached from app.ipynb:814/:793/:819).  The input may be the channel concat
 * of two tensors (`torch.cat([h, skip], 1)` of the up blocks) which is never materialised in fp32.
 * Writes any of: normalised(+SiLU) fp16 operand `out16`, the same in fp32 `out32` (feeds the few-channel
 * fp32 output convs), and `raw16`, the un-normalised cast of the input (operand of the 1x1 conv_shortcut).
 * workspace: dfu_groupnorm_workspace() bytes of per-chunk partial sums (deterministic two-stage reduction).
 * Small and medium maps run ONE launch (thread-block clusters, statistics exchanged through distributed shared memory,
 * the input read once); larger maps run a statistics launch + an apply launch.  sync_words: reserved (accepted and
 * ignored; it backed grid-barrier / arrival-counter variants that were measured slower and removed).
 */
size_t dfu_groupnorm_workspace(int B, int HW, int C, int groups);
int dfu_groupnorm(const float* src0, int C0, const float* src1, int C1, int B, int HW, int groups,
                  const float* gamma, const float* beta, float eps, int silu, void* out16, int planes,
                  int64_t plane_stride, float* out32, void* raw16, void* workspace, size_t workspace_bytes,
                  void* sync_words, void* stream);
/* BasicTransformerBlock.norm1/2/3 (LayerNorm, eps 1e-5) over [M, C] tokens -> fp16 operand planes (out16) and / or an
 * fp32 copy (out32: the ViT glyph encoder's final layernorm returns fp32 last_hidden_state). */
int dfu_layernorm(const float* x, int M, int C, const float* gamma, const float* beta, float eps, void* out16,
                  int planes, int64_t plane_stride, float* out32, void* stream);
/* ViT patch embedding operand (TrOCR encoder, app.ipynb:773-776): NCHW fp32 pixel_values [B, C, H, W] ->
 * fp16 planes [B * (1 + (H/P)*(W/P)), C*P*P], row 0 of every sample zero (the CLS slot: its embedding comes in through
 * the GEMM's residual operand), row 1 + py*(W/P) + px = the patch flattened in (c, ky, kx) order = Conv2d's weight order. */
int dfu_patchify_f16(const float* x, int B, int C, int H, int W, int P, void* out16, int planes, int64_t plane_stride,
                     void* stream);
/* fp32 NHWC -> fp16 operand. mode 0: as is; 1: nearest 2x upsample (Upsample2D's F.interpolate folded into the
 * operand of its conv); 2: space-to-depth parity planes [py*2+px][B][H/2][W/2][C] (Downsample2D stride-2 conv). */
int dfu_cast_f16(const float* x, int B, int H, int W, int C, int mode, void* out16, int planes,
                 int64_t plane_stride, void* stream);

/* ---- small fp32 kernels ---------------------------------------------------------------------- */
/* diffusers `Timesteps` (flip_sin_to_cos, freq_shift): t[B] fp32 -> [B, dim]. */
int dfu_timestep_embedding(const float* t, int B, int dim, int flip_sin_to_cos, float freq_shift, float* out,
                           void* stream);
/* out[b,n] = act_out(bias[n] + sum_k W[n,k] * act_in(x[b,k])), B <= 16: TimestepEmbedding linear_1/linear_2 and all
 * ResnetBlock2D.time_emb_proj layers (stacked into one W) in one launch each. */
int dfu_gemv(const float* x, int B, int K, int ldx, const float* W, const float* bias, int N, int silu_in,
             int silu_out, float* out, int ldo, void* stream);
/* Few-input-channel conv (UNet conv_in over the never-materialised cat([latents, mask, masked_latents]) of
 * app.ipynb:811; VAE encoder conv_in; VAE post_quant_conv + decoder conv_in): NCHW fp32 sources -> NHWC fp32.
 * bstrideN: elements between samples of source N (0 broadcasts). nhwc=1: sources are NHWC. pre_scale multiplies
 * the inputs (1/scaling_factor).  w is TRANSPOSED: [Cin*k*k][Cout] (k = c*k*k + ky*k + kx). */
int dfu_conv_small_in(const float* src0, int c0, int64_t bstride0, const float* src1, int c1, int64_t bstride1,
                      const float* src2, int c2, int64_t bstride2, int nhwc, int B, int H, int W, int ksz,
                      const float* w, const float* bias, int Cout, float pre_scale, float* out, void* stream);
/* Few-output-channel conv (UNet conv_out, VAE conv_out + quant_conv): NHWC fp32 -> NCHW fp32, weights
 * [Cout][k*k][Cin]; optional trailing 1x1 (w2,b2); optional fused scheduler update
 * prev = coef[0]*sample + coef[1]*result  (DDIMScheduler.step collapsed, app.ipynb:816 / SURVEY a12). */
int dfu_conv_small_out(const float* x, int B, int H, int W, int Cin, int ksz, const float* w, const float* bias,
                       int Cout, const float* w2, const float* b2, int Cout2, float* out, const float* sample,
                       float* prev, const float* coef, const uint32_t* seed, void* stream);
/* Gaussian noise as a pure function of (seed, step, element index): Philox4x32-10 keyed by `seed`, counter (index lo,
 * index hi, step, 0), the first two output words -> 24-bit uniforms (k + 0.5) / 2^24 -> sqrt(-2 ln u1) cos(2 pi u2).
 * This is the stream dfu_conv_small_out's fused ancestral step adds (`seed` != NULL: coef = {cx, ce, sigma, step};
 * prev = cx * sample + ce * eps + sigma * z[element], DDPMScheduler.step at app.ipynb:816).  out [n] and / or the raw
 * words bits [n][2]. */
int dfu_philox_normal(uint64_t seed, uint32_t step, int64_t n, float* out, uint32_t* bits, void* stream);
/* y = a*x + b*e (+ c*n): scheduler.step / add_noise / get_velocity in collapsed-coefficient form. */
int dfu_axpbypcz(const float* x, const float* e, const float* n, float a, float b, float c, float* y,
                 int64_t total, void* stream);
/* DDIMScheduler.step / DDPMScheduler.step (app.ipynb:816) in coefficient form, any prediction_type, optional
 * clip_sample:  y = p0 * clamp?(a0*x + a1*m, -1, 1) + d0*x + d1*m + sn*n;  x0_out (optional) = the clamped x0. */
int dfu_scheduler_step(const float* x, const float* m, const float* n, float a0, float a1, float p0, float d0,
                       float d1, float sn, int clip, float* y, float* x0_out, int64_t total, void* stream);
/* add_noise / get_velocity (train_diffute_v1.py:897, :907) with per-sample coefficients: y[b] = ca[b]*x[b] + cb[b]*e[b]. */
int dfu_axpby_rows(const float* x, const float* e, const float* ca, const float* cb, float* y, int B,
                   int64_t per_row, void* stream);
/* DiagonalGaussianDistribution.sample()/mode() * scale from NCHW moments [B, 2*Cz, h, w] (eps NULL = mode). */
int dfu_gaussian_sample(const float* moments, const float* eps, int B, int Cz, int HW, float scale, float* z,
                        void* stream);
/* Row softmax of fp32 scores*scale -> fp16 operand planes (single-head d=512 VAE attention). */
int dfu_softmax_rows(const float* s, int rows, int n, int lds, float scale, void* p16, int ldp, int planes,
                     int64_t plane_stride, void* stream);
/* fp16 [planes][rows][cols (row stride ld_in)] -> [planes][cols][rows]. */
int dfu_transpose_f16(const void* in, int planes, int rows, int cols, int ld_in, int64_t in_plane, void* out,
                      int64_t out_plane, void* stream);

/* ---- the reference's pre-/post-processing around the loop, on the GPU ---------------
 * text_editing, /root/reference/app.ipynb:705-748: window [y_s, y_s+ch) x [x_s, x_s+cw) of the uint8 HWC photograph
 * (already clipped to the image) -> alb.Resize(out_size, out_size) = cv2.resize INTER_LINEAR on uint8 (OpenCV's 11-bit
 * fixed point, bit-exact) -> alb.Normalize(0.5, 0.5) -> ToTensorV2.  image_out / masked_out [3][S][S] fp32 in [-1, 1]
 * (masked: pixels under the text box (bx0, by0)-(bx1, by1), both corners INCLUSIVE as PIL's rectangle, app.ipynb:370-383,
 * are zeroed before the resize), mask_out [S][S] the resized 0/1 mask, mask_lat [S/f][S/f] its nearest-neighbour
 * reduction to the latent grid (app.ipynb:787-790).  Any output may be NULL. */
int dfu_glue_preprocess(const uint8_t* image, int h, int w, int x_s, int y_s, int cw, int ch, int bx0, int by0,
                        int bx1, int by1, int out_size, int lat_factor, float* image_out, float* masked_out,
                        float* mask_out, float* mask_lat, void* stream);
/* app.ipynb:821-841: decoded [3][S][S] fp32 in [-1, 1] -> (x / 2 + 0.5) * 255 -> cv2.resize on float32 to (r_w, r_h)
 * (double-precision fraction, fused multiply-add lerp, horizontal pass first: what the IPP build of the wheel
 * computes) -> pasted at (x_s, y_s) -> of that only the numpy slice [by0:by1, bx0:bx1] (end EXCLUSIVE) replaces the
 * photograph -> round half to even -> uint8 (wrap != 0: modulo 256 like numpy's astype; 0: clamped).  out [h][w][3]. */
int dfu_glue_composite(const float* decoded, int S, const uint8_t* image, int h, int w, int x_s, int y_s, int r_w,
                       int r_h, int bx0, int by0, int bx1, int by1, int wrap, uint8_t* out, void* stream);

/* TrOCRProcessor's image side (app.ipynb:773-774) = ViTImageProcessor: uint8 HWC glyph image [h][w][3] -> PIL
 * Image.resize((S, S), BILINEAR) (Pillow's antialiased 22-bit fixed-point resample, bit-exact) -> * 1/255 ->
 * (x - 0.5) / 0.5 -> out [3][S][S] fp32 (one sample of `pixel_values`).  workspace: coefficient tables. */
size_t dfu_glyph_preprocess_workspace(int h, int w, int out_size);
int dfu_glyph_preprocess(const uint8_t* image, int h, int w, int out_size, void* workspace, size_t workspace_bytes,
                         float* out, void* stream);

/* ---- fused attention core (head dim 64) --------------------------------------------------------
 * out[b, q, h*64:(h+1)*64] = softmax(Q_h K_h^T * scale) V_h for every sample b and head h: diffusers `Attention`
 * core of BasicTransformerBlock.attn1 (self, Nk = Nq = H*W) and attn2 (cross, Nk = 577 glyph tokens), SURVEY A.1,
 * reached from app.ipynb:814.  Q/K/V are fp16 operand matrices [planes][B*N][ld] with head h at columns
 * col0 + 64*h (so a fused QKV projection output can be passed three times with different col0).  The result is
 * written as an fp16 operand [planes][B*Nq][ldo] for the to_out projection.
 * Work distribution: the (sample, head, 128-query tile) items x 64-key blocks sequence is cut into equal contiguous
 * ranges, one per CTA ("stream-K"); kv_splits = 0 chooses the number of ranges automatically (2 x SMs for long key
 * sequences, one item per CTA for short ones), kv_splits = k > 0 forces items x k ranges (1 = never cut an item).  Items
 * cut across CTAs leave partial (O, max, sum) in `workspace` (dfu_attention_workspace bytes) and a merge kernel combines
 * the pieces in key order (deterministic, bit-identical run to run).
 */
size_t dfu_attention_workspace(int B, int heads, int Nq, int Nk, int kv_splits);
/* The distribution dfu_attention would use (no GPU needed): out[5] = {CTAs, work items, 64-key blocks per item, largest
 * number of pieces an item is cut into (<= 8), 1 if the range <-> CTA maps are consistent}. */
int dfu_attention_plan(int B, int heads, int Nq, int Nk, int kv_splits, int32_t* out);
int dfu_attention(const void* q, int ldq, int q_col0, int64_t q_plane_stride, const void* k, int ldk, int k_col0,
                  const void* v, int ldv, int v_col0, int64_t kv_plane_stride, int B, int heads, int Nq, int Nk,
                  int planes, float scale, void* out, int ldo, int64_t out_plane_stride, int kv_splits,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ---- load-time weight packing --------------------------------------------------------------------
 * One launch repacks a whole list of parameters from the diffusers state-dict layout (what
 * UNet2DConditionModel.from_pretrained / AutoencoderKL.from_pretrained read, app.ipynb:550-553) into the kernels'
 * layouts.  Job j covers rows*taps*cin elements; prefix[j] = elements of jobs 0..j-1 (device, int64).
 *   src  fp32 [rows][cin][taps]  (torch [Cout, Cin, kh, kw] flattened; taps = 1 for linear / vectors)
 *   mode 0: dst[(dst_row0 + r') * dst_ld + tap*cin + ci]  — K-major, tap-major; r' = r, or the GEGLU interleave
 *           (blocks of 16 value rows / 16 gate rows) when geglu = 1; planes 0 -> fp32, 1 -> fp16, 2 -> fp16 hi + lo
 *           (lo at + plane_stride elements);
 *   mode 1: fp32 transposed dst[(ci*taps + tap) * dst_ld + dst_row0 + r]  (dfu_conv_small_in weights).
 *   src2 (optional) is added elementwise before packing (conv2 bias + folded conv_shortcut bias). */
typedef struct DfuPackJob {
  const float* src;
  const float* src2;
  void* dst;
  int64_t plane_stride;
  int32_t rows, cin, taps;
  int32_t mode;
  int32_t dst_row0, dst_ld;
  int32_t geglu;
  int32_t planes;
} DfuPackJob;
int dfu_pack_weights(const DfuPackJob* jobs_dev, const int64_t* prefix_dev, int njobs, int64_t total_elements,
                     void* stream);

/* ---- diagnostics --------------------------------------------------------------------------------
 * In-kernel timeline records (one per CTA: grid id, kernel tag, SM, globaltimer / clock64 phase stamps) for
 * scripts/trace_step.py.  `buf` = device u64 array: [0] next record (zeroed by the caller), [1] capacity in records,
 * records of 12 u64 from [8]; NULL disables.  Only libdiffute_b200_trace.so (built with -DDFU_TRACE) records anything;
 * in the product library these return -1 and no kernel contains tracing code.  One setter per translation unit.
 */
int dfu_trace_set_gemm(void* buf);
int dfu_trace_set_gemm2(void* buf);
int dfu_trace_set_attn(void* buf);
int dfu_trace_set_norm(void* buf);
int dfu_trace_set_misc(void* buf);

#ifdef __cplusplus
}
#endif
#endif /* DIFFUTE_B200_H_ */
