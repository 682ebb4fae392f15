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
  int32_t b_ld;         /* row stride in elements = total K of this group (ntaps * k_per_tap) */
  int32_t b_plane;      /* rows between the hi and lo planes */
  int32_t ntaps;        /* 1 (linear / 1x1) or 9 */
  int32_t k_per_tap;    /* contraction length per tap (multiple of 64) */
  int8_t tap_dn[9];     /* per tap: image offset (parity plane for stride-2), row offset, column offset */
  int8_t tap_dy[9];
  int8_t tap_dx[9];
  int8_t _pad[5];
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
  int32_t conv;            /* 1: rows are output pixels of a [B,H,W] grid (mode-1 operands) */
  int32_t B, H, W;         /* output grid when conv=1 */
  /* epilogue: v = alpha*acc + bias[n] + rowvec[(m / rows_per_sample), n] + residual[m, n] */
  int32_t epi;             /* DFU_EPI_* */
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
  int32_t block_n;         /* multiple of 32, <= 256, divides n */
  int32_t splits;          /* split-K factor */
  int32_t stages;
  void* workspace;         /* fp32 [splits, m, n] when splits > 1 */
  size_t workspace_bytes;
} DfuGemm;

int dfu_gemm(const DfuGemm* desc, void* stream);
/* Workspace bytes dfu_gemm needs for this descriptor with automatic tiling (0 if none). */
size_t dfu_gemm_workspace(const DfuGemm* desc);

#ifdef __cplusplus
}
#endif
#endif /* DIFFUTE_B200_H_ */
