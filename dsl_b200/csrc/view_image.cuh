// Per-pixel arithmetic of the PIXEL side of the reference's view pipelines: Resize(keep_ratio, cv2
// INTER_LINEAR on uint8) -> PatchShuffle -> RandomFlip(horizontal) -> Normalize(to_rgb) -> Pad, i.e.
// mmdet/datasets/pipelines/transforms.py:218-247 (_resize_img -> mmcv.imrescale -> cv2.resize), :2180-2199 (PatchShuffle
// pixel moves), :374-383 (imflip), :668-683 (imnormalize), :729-741 (impad_to_multiple), evaluated backwards from one
// OUTPUT pixel. Every operation is spelled with a fixed rounding (no fused multiply-adds, double where OpenCV uses
// double), so the result is the same bits as the CPU pipeline's. The functions are __host__ __device__: geometry.cu
// calls them from view_images_kernel; tests/view_image_host.cpp compiles the same header with g++ so the arithmetic is
// pinned against the oracle on the GPU-less build box (test infrastructure: nothing in libdslb.so runs on the host).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define DSLB_HD __host__ __device__ __forceinline__
#else
#define DSLB_HD static inline
#endif

namespace dslb {

struct ImageViewDev {   // == dslb_image_view_t
  int src_h, src_w;     // source image: uint8 HWC, 3 channels
  int img_h, img_w;     // size after Resize (img_shape), from mmcv.rescale_size on the host
  int ps_mode;          // PatchShuffle: 0 off, 1 'flip' (columns), 2 'flop' (rows)
  int ps_crop;          // min(int(round(extent * PS_place)), extent); 0 or extent = no-op
  int flip;             // RandomFlip horizontal
  int reserved;
};

// products / sums with one rounding each, also where the compiler would otherwise contract them into an FMA
DSLB_HD double vi_dmul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  volatile double r = a * b;
  return r;
#endif
}
DSLB_HD double vi_dadd(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  volatile double r = a + b;
  return r;
#endif
}
DSLB_HD float vi_fsub(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  volatile float r = a - b;
  return r;
#endif
}

// OpenCV's INTER_LINEAR source index and 11-bit weights of destination index d along one axis (modules/imgproc/src/
// resize.cpp, resizeGeneric_ set-up): scale = 1 / (dn / sn) in double, f = (float)((d + 0.5) * scale - 0.5),
// i = floor(f), weights cvRound((1 - f) * 2048), cvRound(f * 2048) (round half to even). Horizontally the weight is
// reset where the tap leaves the image; vertically only the row indices are clamped.
struct LinTap {
  int i0, i1;   // clamped source indices
  int a0, a1;   // weights, a0 + a1 == 2048 up to the two roundings
};

// scale of one axis as cv::resize computes it: inv_scale = (double)dsize / ssize, scale = 1. / inv_scale
DSLB_HD double vi_axis_scale(int dn, int sn) { return 1.0 / ((double)dn / (double)sn); }

DSLB_HD LinTap vi_linear_tap(int d, double scale, int sn, bool vertical) {
  float f = (float)vi_dadd(vi_dmul((double)d + 0.5, scale), -0.5);
  int i = (int)floorf(f);
  f = vi_fsub(f, (float)i);
  if (!vertical) {
    if (i < 0) { f = 0.f; i = 0; }
    if (i >= sn - 1) { f = 0.f; i = sn - 1; }
  }
  LinTap t;
  t.a1 = (int)rintf(f * 2048.f);                 // exact scaling by a power of two, then round half to even
  t.a0 = (int)rintf(vi_fsub(1.f, f) * 2048.f);
  t.i0 = i < 0 ? 0 : (i > sn - 1 ? sn - 1 : i);
  t.i1 = i + 1 < 0 ? 0 : (i + 1 > sn - 1 ? sn - 1 : i + 1);
  return t;
}

// cv2.resize(INTER_LINEAR) of one uint8 sample: horizontal pass in int32 (pixel * 11-bit weight), vertical pass
// (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2 (VResizeLinear<uchar, int, short, FixedPtCast<..., 22>>).
DSLB_HD int vi_bilinear_u8(int p00, int p01, int p10, int p11, const LinTap& tx, const LinTap& ty) {
  const int s0 = p00 * tx.a0 + p01 * tx.a1;
  const int s1 = p10 * tx.a0 + p11 * tx.a1;
  return (((ty.a0 * (s0 >> 4)) >> 16) + ((ty.a1 * (s1 >> 4)) >> 16) + 2) >> 2;
}

// position in the RESIZED image that output pixel (y, x) of the view shows: undo RandomFlip, then PatchShuffle
// ('flip': shuffled[:, j] = resized[:, (j + crop) mod w]; 'flop': the same on rows).
DSLB_HD void vi_source_pos(const ImageViewDev& v, int y, int x, int& ry, int& rx) {
  if (v.flip) x = v.img_w - 1 - x;
  if (v.ps_mode == 1 && v.ps_crop > 0 && v.ps_crop < v.img_w) {
    x += v.ps_crop;
    if (x >= v.img_w) x -= v.img_w;
  } else if (v.ps_mode == 2 && v.ps_crop > 0 && v.ps_crop < v.img_h) {
    y += v.ps_crop;
    if (y >= v.img_h) y -= v.img_h;
  }
  ry = y;
  rx = x;
}

// mmcv.imnormalize on one resized uint8 sample: float32 subtraction of the float32 mean, then the float32 value times
// the DOUBLE reciprocal of the float32 std, rounded to float32 once (what cv2.subtract / cv2.multiply do with float64
// scalar operands on a float32 image; transforms.py:665-666 stores mean / std as float32).
DSLB_HD float vi_normalize(int pix, float mean, double inv_std) {
  return (float)vi_dmul((double)vi_fsub((float)pix, mean), inv_std);
}

// The three channels of one output pixel from its taps. to_rgb: output channel c shows source channel 2 - c.
DSLB_HD void vi_pixel_taps(const uint8_t* src, int src_w, const LinTap& tx, const LinTap& ty, const float* mean,
                           const double* inv_std, int to_rgb, float* out3) {
  const uint8_t* r0 = src + (long long)ty.i0 * src_w * 3;
  const uint8_t* r1 = src + (long long)ty.i1 * src_w * 3;
  for (int c = 0; c < 3; ++c) {
    const int sc = to_rgb ? 2 - c : c;
    const int p = vi_bilinear_u8(r0[tx.i0 * 3 + sc], r0[tx.i1 * 3 + sc], r1[tx.i0 * 3 + sc], r1[tx.i1 * 3 + sc], tx, ty);
    out3[c] = vi_normalize(p, mean[c], inv_std[c]);
  }
}

// Output pixel (y, x), y < img_h, x < img_w, of the view.
DSLB_HD void vi_pixel(const uint8_t* src, const ImageViewDev& v, int y, int x, const float* mean, const double* inv_std,
                      int to_rgb, float* out3) {
  int ry, rx;
  vi_source_pos(v, y, x, ry, rx);
  const LinTap tx = vi_linear_tap(rx, vi_axis_scale(v.img_w, v.src_w), v.src_w, false);
  const LinTap ty = vi_linear_tap(ry, vi_axis_scale(v.img_h, v.src_h), v.src_h, true);
  vi_pixel_taps(src, v.src_w, tx, ty, mean, inv_std, to_rgb, out3);
}

// ---- one thread of view_images_kernel ------------------------------------------------------------------------------
struct ViewImageParams {
  float mean[3];
  double inv_std[3];
  int to_rgb;
};

constexpr int VI_TX = 32, VI_TY = 8, VI_ROWS = 4;   // 32 x 8 threads per 32 x 32 output tile, four rows per thread

// Thread (tx, ty) of block (bx, by, bz): column x = bx * 32 + tx of image bz, rows by * 32 + j * 8 + ty. Keeps the column
// tap and the row scale (the double divisions happen once), walks its four rows and writes the three channel planes of
// out [B][3][H][W]; everything outside img_h x img_w is written as zero. The __global__ wrapper in geometry.cu passes
// its blockIdx / threadIdx; tests/view_image_host.cpp walks the same grid on the host.
DSLB_HD void vi_thread(const uint8_t* const* srcs, const ImageViewDev* views, const ViewImageParams& prm, float* out, int H,
                       int W, int bx, int by, int bz, int tx_, int ty_) {
  const ImageViewDev v = views[bz];
  const uint8_t* src = srcs[bz];
  const int x = bx * VI_TX + tx_;
  if (x >= W) return;
  const size_t plane = (size_t)H * W;
  float* ob = out + (size_t)bz * 3 * plane;
  const bool in_x = x < v.img_w;
  LinTap tx = {0, 0, 0, 0};
  if (in_x) {
    int ry, rx;
    vi_source_pos(v, 0, x, ry, rx);
    tx = vi_linear_tap(rx, vi_axis_scale(v.img_w, v.src_w), v.src_w, false);
  }
  const double scale_y = vi_axis_scale(v.img_h, v.src_h);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int j = 0; j < VI_ROWS; ++j) {
    const int y = by * (VI_TY * VI_ROWS) + j * VI_TY + ty_;
    if (y >= H) break;
    float o[3] = {0.f, 0.f, 0.f};
    if (in_x && y < v.img_h) {
      int ry, rx;
      vi_source_pos(v, y, x, ry, rx);
      const LinTap ty = vi_linear_tap(ry, scale_y, v.src_h, true);
      vi_pixel_taps(src, v.src_w, tx, ty, prm.mean, prm.inv_std, prm.to_rgb, o);
    }
    const size_t at = (size_t)y * W + x;
    ob[at] = o[0];
    ob[plane + at] = o[1];
    ob[2 * plane + at] = o[2];
  }
}

}  // namespace dslb
