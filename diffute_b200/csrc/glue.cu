// diffute_b200 — the reference's pre-/post-processing around the sampling loop, on the GPU.
//
// text_editing (app.ipynb:663-771, :821-841) crops a window of the photograph around the text box, resizes it to
// 512 x 512 with albumentations' Resize (= cv2.resize INTER_LINEAR on uint8), normalises to [-1, 1], does the same with
// the masked photograph and the rectangular mask, and after vae.decode resizes the result back with cv2.resize on
// float32 and pastes the text box into the photograph.  Both kernels reproduce the third-party arithmetic bit for
// bit (oracle/glue.py, pinned against cv2 4.13 / PIL in tests/test_glue_oracle.py):
//   uint8 path   OpenCV's fixed-point bilinear: 11-bit weights cvRound(w * 2048) from the float32 fraction, horizontal
//                pass into int, vertical pass (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2, rows
//                clamped vertically / weight moved onto the clamped pixel horizontally, exact 2x decimation routed to
//                the area average;
//   float path   (IPP build of the wheel) fraction taken in double, fused multiply-add lerp, horizontal pass first.
// Byte / integer work, HBM-bound and tiny (3 x 512 x 512 outputs): one thread per destination pixel, no staging.
#include "common.cuh"
#include "kernels.h"

namespace dfu {

struct U8Coef {
  int s, w0, w1;
};

// cv::resize INTER_LINEAR coefficient of destination index d (resize.cpp: fx = (float)((dx + 0.5) * scale_x - 0.5))
__device__ __forceinline__ U8Coef coef_u8(int d, int ssize, double scale, bool horizontal) {
  // (explicit round-to-nearest steps: no contraction into a double-precision fma, which the host code does not do)
  float f = static_cast<float>(__dsub_rn(__dmul_rn(__dadd_rn(static_cast<double>(d), 0.5), scale), 0.5));
  int s = static_cast<int>(floorf(f));
  f = __fsub_rn(f, static_cast<float>(s));
  if (horizontal) {
    if (s < 0) {
      f = 0.f;
      s = 0;
    }
    if (s >= ssize - 1) {
      f = 0.f;
      s = ssize - 1;
    }
  }
  U8Coef c;
  c.s = s;
  c.w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));  // saturate_cast<short>(cvRound(w * INTER_RESIZE_COEF_SCALE))
  c.w1 = __float2int_rn(__fmul_rn(f, 2048.f));
  return c;
}

struct PreParams {
  const uint8_t* image;  // [h][w][3]
  int h, w;
  int x_s, y_s, cw, ch;        // crop window inside the image
  int bx0, by0, bx1, by1;      // text box, both corners inclusive (PIL rectangle)
  int S, lat_factor;
  double scale_x, scale_y;
  float* image_out;   // [3][S][S]
  float* masked_out;  // [3][S][S]
  float* mask_out;    // [S][S]
  float* mask_lat;    // [S / f][S / f]
};

__device__ __forceinline__ int vresize_u8(int r0, int r1, int b0, int b1) {
  const int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__global__ void __launch_bounds__(256) glue_preprocess_kernel(const PreParams p) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x;
  const int dy = blockIdx.y;
  if (dx >= p.S) return;
  const bool area = p.cw == 2 * p.S && p.ch == 2 * p.S;  // INTER_LINEAR at exactly 2x decimation = INTER_AREA fast path
  int xs[2], ys[2], a0, a1, b0, b1;
  if (area) {
    xs[0] = 2 * dx;
    xs[1] = 2 * dx + 1;
    ys[0] = 2 * dy;
    ys[1] = 2 * dy + 1;
    a0 = a1 = b0 = b1 = 0;
  } else {
    const U8Coef cx = coef_u8(dx, p.cw, p.scale_x, true);
    const U8Coef cy = coef_u8(dy, p.ch, p.scale_y, false);
    xs[0] = cx.s;
    xs[1] = min(cx.s + 1, p.cw - 1);
    ys[0] = min(max(cy.s, 0), p.ch - 1);
    ys[1] = min(max(cy.s + 1, 0), p.ch - 1);
    a0 = cx.w0;
    a1 = cx.w1;
    b0 = cy.w0;
    b1 = cy.w1;
  }
  // the four source pixels: 3 channels each, and whether they lie under the text box
  int px[2][2][3];
  int mk[2][2];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int x = p.x_s + xs[i], y = p.y_s + ys[j];
      const uint8_t* s = p.image + (static_cast<size_t>(y) * p.w + x) * 3;
      px[j][i][0] = s[0];
      px[j][i][1] = s[1];
      px[j][i][2] = s[2];
      mk[j][i] = (x >= p.bx0 && x <= p.bx1 && y >= p.by0 && y <= p.by1) ? 1 : 0;
    }
  auto interp = [&](int v00, int v01, int v10, int v11) -> int {
    if (area) return (v00 + v01 + v10 + v11 + 2) >> 2;
    return vresize_u8(v00 * a0 + v01 * a1, v10 * a0 + v11 * a1, b0, b1);
  };
  const size_t plane = static_cast<size_t>(p.S) * p.S;
  const size_t o = static_cast<size_t>(dy) * p.S + dx;
  const float k = static_cast<float>(1.0 / 127.5);  // alb.Normalize: (x - mean * 255) * (1 / (std * 255)) in float32
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (p.image_out) {
      const int v = interp(px[0][0][c], px[0][1][c], px[1][0][c], px[1][1][c]);
      p.image_out[c * plane + o] = __fmul_rn(__fsub_rn(static_cast<float>(v), 127.5f), k);
    }
    if (p.masked_out) {  // prepare_mask_and_masked_image: pixels under the mask are zero BEFORE the resize
      const int v = interp(mk[0][0] ? 0 : px[0][0][c], mk[0][1] ? 0 : px[0][1][c], mk[1][0] ? 0 : px[1][0][c],
                           mk[1][1] ? 0 : px[1][1][c]);
      p.masked_out[c * plane + o] = __fmul_rn(__fsub_rn(static_cast<float>(v), 127.5f), k);
    }
  }
  if (p.mask_out || p.mask_lat) {
    const float m = static_cast<float>(interp(mk[0][0], mk[0][1], mk[1][0], mk[1][1]));
    if (p.mask_out) p.mask_out[o] = m;
    const int f = p.lat_factor;
    if (p.mask_lat && dx % f == 0 && dy % f == 0)  // F.interpolate(nearest) to the latent grid = every f-th pixel
      p.mask_lat[static_cast<size_t>(dy / f) * (p.S / f) + dx / f] = m;
  }
}

struct CompParams {
  const float* decoded;  // [3][S][S] in [-1, 1]
  int S;
  const uint8_t* image;  // [h][w][3]
  int h, w;
  int x_s, y_s, r_w, r_h;
  int bx0, by0, bx1, by1;  // numpy slice [by0:by1, bx0:bx1] (end EXCLUSIVE, app.ipynb:840)
  int wrap;
  double scale_x, scale_y;
  uint8_t* out;  // [h][w][3]
};

// value of output byte (y, x, c): the photograph's, or inside the pasted text box the resized decoded image
__device__ __forceinline__ uint8_t composite_byte(const CompParams& p, int y, int x, int c, size_t o) {
  const int dx = x - p.x_s, dy = y - p.y_s;
  const bool inside = x >= p.bx0 && x < p.bx1 && y >= p.by0 && y < p.by1 && dx >= 0 && dx < p.r_w && dy >= 0 && dy < p.r_h;
  if (!inside) return p.image[o];
  // cv2.resize float32: the fraction is taken in double precision
  const double fxd = __dsub_rn(__dmul_rn(__dadd_rn(static_cast<double>(dx), 0.5), p.scale_x), 0.5);
  const double fyd = __dsub_rn(__dmul_rn(__dadd_rn(static_cast<double>(dy), 0.5), p.scale_y), 0.5);
  int sx = static_cast<int>(floor(fxd)), sy = static_cast<int>(floor(fyd));
  float fx = static_cast<float>(fxd - static_cast<double>(sx));
  const float fy = static_cast<float>(fyd - static_cast<double>(sy));
  if (sx < 0) {
    fx = 0.f;
    sx = 0;
  }
  if (sx >= p.S - 1) {
    fx = 0.f;
    sx = p.S - 1;
  }
  const int x0 = sx, x1 = min(sx + 1, p.S - 1);
  const int y0 = min(max(sy, 0), p.S - 1), y1 = min(max(sy + 1, 0), p.S - 1);
  const float* d = p.decoded + static_cast<size_t>(c) * p.S * p.S;
  auto px = [&](int yy, int xx) {  // image = (image_vae / 2 + 0.5) * 255.0 in float32 (app.ipynb:822)
    return __fmul_rn(__fadd_rn(__fmul_rn(d[static_cast<size_t>(yy) * p.S + xx], 0.5f), 0.5f), 255.0f);
  };
  const float v00 = px(y0, x0), v01 = px(y0, x1), v10 = px(y1, x0), v11 = px(y1, x1);
  const float r0 = __fmaf_rn(__fsub_rn(v01, v00), fx, v00);
  const float r1 = __fmaf_rn(__fsub_rn(v11, v10), fx, v10);
  const float v = __fmaf_rn(__fsub_rn(r1, r0), fy, r0);
  int q = __float2int_rn(v);  // np.round: half to even
  q = p.wrap ? (q & 255) : (q < 0 ? 0 : (q > 255 ? 255 : q));
  return static_cast<uint8_t>(q);
}

// One thread per 4 output bytes of the flattened [h][w][3] image: rows that the text box does not touch (nearly all of
// the photograph) are copied as 32-bit words, coalesced; only words with a byte inside the box take the per-byte path.
__global__ void __launch_bounds__(256) glue_composite_kernel(const CompParams p) {
  const size_t total = static_cast<size_t>(p.h) * p.w * 3;
  const size_t o = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (o >= total) return;
  const size_t row_bytes = static_cast<size_t>(p.w) * 3;
  const int y_first = static_cast<int>(o / row_bytes);
  const int y_last = static_cast<int>(min(o + 3, total - 1) / row_bytes);
  const bool rows_hit = y_last >= p.by0 && y_first < p.by1;  // (conservative) some byte may lie in the box's rows
  if (!rows_hit && o + 4 <= total) {
    *reinterpret_cast<uint32_t*>(p.out + o) = *reinterpret_cast<const uint32_t*>(p.image + o);
    return;
  }
  for (int k = 0; k < 4 && o + k < total; ++k) {
    const size_t b = o + k;
    const int y = static_cast<int>(b / row_bytes);
    const int rem = static_cast<int>(b - static_cast<size_t>(y) * row_bytes);
    p.out[b] = composite_byte(p, y, rem / 3, rem % 3, b);
  }
}

}  // namespace dfu

extern "C" int dfu_glue_preprocess(const uint8_t* image, int h, int w, int x_s, int y_s, int cw, int ch, int bx0, int by0,
                                   int bx1, int by1, int out_size, int lat_factor, float* image_out, float* masked_out,
                                   float* mask_out, float* mask_lat, void* stream) {
  using namespace dfu;
  DFU_REQUIRE(image && h > 0 && w > 0, "glue_preprocess: bad image");
  DFU_REQUIRE(cw > 0 && ch > 0 && x_s >= 0 && y_s >= 0 && x_s + cw <= w && y_s + ch <= h,
              "glue_preprocess: window (%d, %d, %d x %d) leaves the %d x %d image", x_s, y_s, cw, ch, w, h);
  DFU_REQUIRE(out_size > 0 && (mask_lat == nullptr || (lat_factor > 0 && out_size % lat_factor == 0)),
              "glue_preprocess: bad output size %d / latent factor %d", out_size, lat_factor);
  PreParams p;
  p.image = image;
  p.h = h;
  p.w = w;
  p.x_s = x_s;
  p.y_s = y_s;
  p.cw = cw;
  p.ch = ch;
  p.bx0 = bx0;
  p.by0 = by0;
  p.bx1 = bx1;
  p.by1 = by1;
  p.S = out_size;
  p.lat_factor = lat_factor > 0 ? lat_factor : 1;
  p.scale_x = 1.0 / (static_cast<double>(out_size) / static_cast<double>(cw));  // cv::resize: 1. / inv_scale_x
  p.scale_y = 1.0 / (static_cast<double>(out_size) / static_cast<double>(ch));
  p.image_out = image_out;
  p.masked_out = masked_out;
  p.mask_out = mask_out;
  p.mask_lat = mask_lat;
  dim3 grid((out_size + 255) / 256, out_size);
  glue_preprocess_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DFU_CHECK_CUDA(cudaGetLastError());
  return DFU_OK;
}

extern "C" int dfu_glue_composite(const float* decoded, int S, const uint8_t* image, int h, int w, int x_s, int y_s,
                                  int r_w, int r_h, int bx0, int by0, int bx1, int by1, int wrap, uint8_t* out,
                                  void* stream) {
  using namespace dfu;
  DFU_REQUIRE(decoded && image && out && S > 0 && h > 0 && w > 0, "glue_composite: bad arguments");
  DFU_REQUIRE(r_w > 0 && r_h > 0 && x_s >= 0 && y_s >= 0 && x_s + r_w <= w && y_s + r_h <= h,
              "glue_composite: pasted region (%d, %d, %d x %d) leaves the %d x %d image", x_s, y_s, r_w, r_h, w, h);
  CompParams p;
  p.decoded = decoded;
  p.S = S;
  p.image = image;
  p.h = h;
  p.w = w;
  p.x_s = x_s;
  p.y_s = y_s;
  p.r_w = r_w;
  p.r_h = r_h;
  p.bx0 = bx0;
  p.by0 = by0;
  p.bx1 = bx1;
  p.by1 = by1;
  p.wrap = wrap;
  p.scale_x = 1.0 / (static_cast<double>(r_w) / static_cast<double>(S));
  p.scale_y = 1.0 / (static_cast<double>(r_h) / static_cast<double>(S));
  p.out = out;
  const size_t words = (static_cast<size_t>(h) * w * 3 + 3) / 4;
  glue_composite_kernel<<<static_cast<unsigned>((words + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  DFU_CHECK_CUDA(cudaGetLastError());
  return DFU_OK;
}

// ---------------------------------------------------------------------------------------------
// TrOCRProcessor's image side (app.ipynb:773-774: processor(images=[draw_ttf]).pixel_values) = ViTImageProcessor:
// PIL Image.resize((S, S), BILINEAR) -> * 1/255 -> (x - 0.5) / 0.5 -> CHW float32.  Pillow's resample (Resample.c) is
// an antialiased separable filter: triangle of support max(scale, 1), weights computed in double, normalised, turned
// into 22-bit fixed point (int)(0.5 + w * 2^22), horizontal pass first with a uint8 intermediate (clip8), then the
// vertical pass.  Reproduced bit for bit (oracle/glue.py::pil_resize_bilinear, pinned against Pillow and against
// transformers' ViTImageProcessor); the double-precision steps use explicit round-to-nearest intrinsics so that no
// multiply-add is contracted.
// ---------------------------------------------------------------------------------------------
namespace dfu {

constexpr int kPilBits = 22;

__device__ __forceinline__ double pil_triangle(double x) {
  if (x < 0.0) x = -x;
  return x < 1.0 ? __dsub_rn(1.0, x) : 0.0;
}

// one thread per destination index: coefficient row K[xx][ksize] and bounds B[xx] = (first source index, count)
__global__ void pil_coeffs_kernel(int in_size, int out_size, int ksize, int* __restrict__ K, int* __restrict__ B) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  if (xx >= out_size) return;
  const double scale = __ddiv_rn(static_cast<double>(in_size), static_cast<double>(out_size));
  const double fs = scale < 1.0 ? 1.0 : scale;
  const double support = fs;  // bilinear: filter support 1.0 x filterscale
  const double ss = __ddiv_rn(1.0, fs);
  const double center = __dmul_rn(__dadd_rn(static_cast<double>(xx), 0.5), scale);
  int xmin = static_cast<int>(__dadd_rn(__dsub_rn(center, support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = static_cast<int>(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  auto weight = [&](int x) {
    return pil_triangle(__dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss));
  };
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) ww = __dadd_rn(ww, weight(x));
  int* k = K + static_cast<size_t>(xx) * ksize;
  for (int x = 0; x < ksize; ++x) {
    double w = x < xmax ? weight(x) : 0.0;
    if (x < xmax && ww != 0.0) w = __ddiv_rn(w, ww);
    k[x] = static_cast<int>(__dadd_rn(0.5, __dmul_rn(w, static_cast<double>(1 << kPilBits))));
  }
  B[2 * xx] = xmin;
  B[2 * xx + 1] = xmax;
}

__device__ __forceinline__ int clip8(int v) {
  v >>= kPilBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__global__ void __launch_bounds__(256)
pil_resize_normalize_kernel(const uint8_t* __restrict__ img, int h, int w, int S, const int* __restrict__ Kx,
                            const int* __restrict__ Bx, int ksx, const int* __restrict__ Ky, const int* __restrict__ By,
                            int ksy, float* __restrict__ out) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  const int yy = blockIdx.y;
  if (xx >= S) return;
  const bool need_h = w != S, need_v = h != S;  // Pillow skips a pass whose size does not change
  const int xmin = need_h ? Bx[2 * xx] : xx, xcnt = need_h ? Bx[2 * xx + 1] : 1;
  const int ymin = need_v ? By[2 * yy] : yy, ycnt = need_v ? By[2 * yy + 1] : 1;
  const int* kx = Kx + static_cast<size_t>(xx) * ksx;
  const int* ky = Ky + static_cast<size_t>(yy) * ksy;
  int acc[3] = {1 << (kPilBits - 1), 1 << (kPilBits - 1), 1 << (kPilBits - 1)};
  int last[3] = {0, 0, 0};
  for (int ty = 0; ty < ycnt; ++ty) {
    const uint8_t* row = img + (static_cast<size_t>(ymin + ty) * w + xmin) * 3;
    int hs[3];
    if (need_h) {
      hs[0] = hs[1] = hs[2] = 1 << (kPilBits - 1);
      for (int tx = 0; tx < xcnt; ++tx) {
        const int k = kx[tx];
        hs[0] += row[tx * 3] * k;
        hs[1] += row[tx * 3 + 1] * k;
        hs[2] += row[tx * 3 + 2] * k;
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) hs[c] = clip8(hs[c]);  // the horizontal pass writes a uint8 image
    } else {
      hs[0] = row[0];
      hs[1] = row[1];
      hs[2] = row[2];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      last[c] = hs[c];
      if (need_v) acc[c] += hs[c] * ky[ty];
    }
  }
  const size_t plane = static_cast<size_t>(S) * S;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int v = need_v ? clip8(acc[c]) : last[c];
    // transformers' PIL / numpy backend: rescale = float32(float64(x) * (1 / 255)), normalize = (x - 0.5) / 0.5 in float32
    const float x = static_cast<float>(__dmul_rn(static_cast<double>(v), 1.0 / 255.0));
    out[c * plane + static_cast<size_t>(yy) * S + xx] = __fdiv_rn(__fsub_rn(x, 0.5f), 0.5f);
  }
}

__host__ inline int pil_ksize(int in_size, int out_size) {
  const double scale = static_cast<double>(in_size) / static_cast<double>(out_size);
  const double support = scale < 1.0 ? 1.0 : scale;
  return static_cast<int>(ceil(support)) * 2 + 1;
}

}  // namespace dfu

extern "C" size_t dfu_glyph_preprocess_workspace(int h, int w, int out_size) {
  using namespace dfu;
  if (h <= 0 || w <= 0 || out_size <= 0) return 0;
  return static_cast<size_t>(out_size) * (pil_ksize(w, out_size) + pil_ksize(h, out_size) + 4) * sizeof(int);
}

extern "C" int dfu_glyph_preprocess(const uint8_t* image, int h, int w, int out_size, void* workspace,
                                    size_t workspace_bytes, float* out, void* stream) {
  using namespace dfu;
  DFU_REQUIRE(image && out && h > 0 && w > 0 && out_size > 0, "glyph_preprocess: bad arguments");
  const size_t need = dfu_glyph_preprocess_workspace(h, w, out_size);
  DFU_REQUIRE(workspace && workspace_bytes >= need, "glyph_preprocess: workspace %zu < %zu bytes", workspace_bytes, need);
  const int S = out_size, ksx = pil_ksize(w, S), ksy = pil_ksize(h, S);
  int* Kx = static_cast<int*>(workspace);
  int* Bx = Kx + static_cast<size_t>(S) * ksx;
  int* Ky = Bx + 2 * S;
  int* By = Ky + static_cast<size_t>(S) * ksy;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  pil_coeffs_kernel<<<(S + 127) / 128, 128, 0, st>>>(w, S, ksx, Kx, Bx);
  pil_coeffs_kernel<<<(S + 127) / 128, 128, 0, st>>>(h, S, ksy, Ky, By);
  dim3 grid((S + 255) / 256, S);
  pil_resize_normalize_kernel<<<grid, 256, 0, st>>>(image, h, w, S, Kx, Bx, ksx, Ky, By, ksy, out);
  DFU_CHECK_CUDA(cudaGetLastError());
  return DFU_OK;
}
