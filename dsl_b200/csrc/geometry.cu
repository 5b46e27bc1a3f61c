// Device-side view geometry of the data pipeline: the box part of Resize -> PatchShuffle -> RandomFlip
// of the reference's train pipelines (configs/fcos_semi/*.py:70-92; mmdet/datasets/pipelines/transforms.py:249-257,
// 2168-2248, 397-429) and the zero-padded batch assembly (Pad size_divisor=32 + collate). With it the boxes the EMA
// teacher found on the weak view (original-image coordinates after `rescale`) can be carried to the student's view
// without leaving the device. Arithmetic is fp32, left to right, no fused multiply-adds, exactly like the NumPy
// float32 expressions of the reference.
#include "common.h"
#include "view_image.cuh"

namespace dslb {

struct ViewDev {   // == dslb_view_t
  float sx, sy;     // Resize scale_factor (w_scale, h_scale)
  int img_w, img_h; // image size AFTER the resize (img_shape): clip range, PatchShuffle / flip extent
  int clip;         // Resize.bbox_clip_border
  int ps_mode;      // PatchShuffle: 0 off, 1 'flip' (cut at column ps_crop), 2 'flop' (cut at row ps_crop)
  int ps_crop;      // crop_w / crop_h = min(int(round(extent * place)), extent); 0 or extent -> no-op
  int flip;         // RandomFlip horizontal
};

constexpr int VIEW_MAX_IMGS = 32;

// One warp per image: lane 0 walks the image's boxes in order (a PatchShuffle cut can split a box in two, so the output
// position depends on all earlier boxes), the block then packs the per-image runs.
__global__ void __launch_bounds__(32 * VIEW_MAX_IMGS) view_boxes_kernel(
    const float* __restrict__ boxes, const long long* __restrict__ labels, const int* __restrict__ off,
    const ViewDev* __restrict__ views, int B, int max_out, float* __restrict__ stage, long long* __restrict__ stage_lab,
    float* __restrict__ out_boxes, long long* __restrict__ out_labels, int* __restrict__ out_off) {
  __shared__ int cnt[VIEW_MAX_IMGS];
  __shared__ int base[VIEW_MAX_IMGS + 1];
  const int img = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (img < B && lane == 0) {
    const ViewDev v = views[img];
    const int b0 = off[img], b1 = off[img + 1];
    float* st = stage + (long long)b0 * 2 * 4;          // at most two output boxes per input box
    long long* sl = stage_lab ? stage_lab + (long long)b0 * 2 : nullptr;
    int n = 0;
    const float fw = (float)v.img_w, fh = (float)v.img_h;
    const bool ps = (v.ps_mode == 1 && v.ps_crop != 0 && v.ps_crop != v.img_w) ||
                    (v.ps_mode == 2 && v.ps_crop != 0 && v.ps_crop != v.img_h);
    const float cw = v.ps_mode == 1 ? (float)v.ps_crop : fw;   // 'flip': crop_h = h; 'flop': crop_w = w
    const float chh = v.ps_mode == 2 ? (float)v.ps_crop : fh;
    for (int i = b0; i < b1; ++i) {
      float x1 = __fmul_rn(boxes[4 * i], v.sx), y1 = __fmul_rn(boxes[4 * i + 1], v.sy);
      float x2 = __fmul_rn(boxes[4 * i + 2], v.sx), y2 = __fmul_rn(boxes[4 * i + 3], v.sy);
      if (v.clip) {
        x1 = fminf(fmaxf(x1, 0.f), fw); x2 = fminf(fmaxf(x2, 0.f), fw);
        y1 = fminf(fmaxf(y1, 0.f), fh); y2 = fminf(fmaxf(y2, 0.f), fh);
      }
      float o[2][4];
      int k = 1;
      o[0][0] = x1; o[0][1] = y1; o[0][2] = x2; o[0][3] = y2;
      if (ps) {
        const float ax = __fadd_rn(__fsub_rn(x1, cw), 1.f), bx = __fadd_rn(__fsub_rn(x2, cw), 1.f);
        const float ay = __fadd_rn(__fsub_rn(y1, chh), 1.f), by = __fadd_rn(__fsub_rn(y2, chh), 1.f);
        if (__fmul_rn(ax, bx) >= 0.f && __fmul_rn(ay, by) >= 0.f) {   // the box lies on one side of the cut
          if (v.ps_mode == 1) {
            if (ax < 0.f) { x1 = __fsub_rn(__fadd_rn(x1, fw), cw); x2 = __fsub_rn(__fadd_rn(x2, fw), cw); }
            if (__fadd_rn(__fsub_rn(x2, cw), 1.f) > 0.f) { x1 = __fsub_rn(x1, cw); x2 = __fsub_rn(x2, cw); }
          } else {
            if (ay < 0.f) { y1 = __fsub_rn(__fadd_rn(y1, fh), chh); y2 = __fsub_rn(__fadd_rn(y2, fh), chh); }
            if (__fadd_rn(__fsub_rn(y2, chh), 1.f) > 0.f) { y1 = __fsub_rn(y1, chh); y2 = __fsub_rn(y2, chh); }
          }
          o[0][0] = x1; o[0][1] = y1; o[0][2] = x2; o[0][3] = y2;
        } else {                                                      // it straddles the cut: two boxes
          k = 2;
          if (v.ps_mode == 1) {
            o[0][0] = __fsub_rn(__fadd_rn(x1, fw), cw); o[0][1] = y1; o[0][2] = __fsub_rn(fw, 1.f); o[0][3] = y2;
            o[1][0] = 0.f; o[1][1] = y1; o[1][2] = __fsub_rn(x2, cw); o[1][3] = y2;
          } else {
            o[0][0] = x1; o[0][1] = __fsub_rn(__fadd_rn(y1, fh), chh); o[0][2] = x2; o[0][3] = __fsub_rn(fh, 1.f);
            o[1][0] = x1; o[1][1] = 0.f; o[1][2] = x2; o[1][3] = __fsub_rn(y2, chh);
          }
        }
      }
      for (int j = 0; j < k; ++j) {
        float a = o[j][0], c = o[j][2];
        if (v.flip) {   // flipped[0] = w - x2, flipped[2] = w - x1
          const float t = __fsub_rn(fw, c);
          c = __fsub_rn(fw, a);
          a = t;
        }
        st[4 * n] = a; st[4 * n + 1] = o[j][1]; st[4 * n + 2] = c; st[4 * n + 3] = o[j][3];
        if (sl) sl[n] = labels[i];
        ++n;
      }
    }
    cnt[img] = n;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int b = 0; b < B; ++b) {
      base[b] = t;
      t += cnt[b];
      if (t > max_out) t = max_out;
    }
    base[B] = t;
    for (int b = 0; b <= B; ++b) out_off[b] = base[b];
  }
  __syncthreads();
  if (img < B) {
    const int n = base[img + 1] - base[img];
    const float* st = stage + (long long)off[img] * 2 * 4;
    for (int e = lane; e < n * 4; e += 32) out_boxes[(long long)base[img] * 4 + e] = st[e];
    if (out_labels && stage_lab)
      for (int e = lane; e < n; e += 32) out_labels[base[img] + e] = stage_lab[(long long)off[img] * 2 + e];
  }
}

// out[b] (C, H, W) = zeros with img_b (C, h_b, w_b) in the top-left corner: Pad + collate of the reference's loader.
__global__ void pad_batch_kernel(const float* const* __restrict__ imgs, const int* __restrict__ hw, float* __restrict__ out,
                                 int B, int C, int H, int W) {
  const long long total = (long long)B * C * H * W;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(t % W);
    const int y = (int)((t / W) % H);
    const int c = (int)((t / ((long long)W * H)) % C);
    const int b = (int)(t / ((long long)W * H * C));
    const int h = hw[2 * b], w = hw[2 * b + 1];
    out[t] = (y < h && x < w) ? imgs[b][((long long)c * h + y) * w + x] : 0.f;
  }
}

// Pixel side of the same pipelines (view_image.cuh): one launch turns B uint8 HWC source images into the zero-padded
// fp32 NCHW batch the network reads. HBM-bound: 12 B written per output pixel, <= 12 source bytes gathered (adjacent
// threads read adjacent source pixels, so the taps come from L1 / L2). A 32 x 8 thread block covers a 32 x 32 output
// tile; a warp writes 128 contiguous bytes of one channel plane. The per-thread body (vi_thread) lives in the header so
// that the host test harness can walk the same grid.
__global__ void __launch_bounds__(VI_TX * VI_TY) view_images_kernel(
    const uint8_t* const* __restrict__ srcs, const ImageViewDev* __restrict__ views, const ViewImageParams prm,
    float* __restrict__ out, int H, int W) {
  vi_thread(srcs, views, prm, out, H, W, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, threadIdx.y);
}

}  // namespace dslb

using namespace dslb;

extern "C" size_t dslb_view_boxes_workspace_bytes(int max_in) { return (size_t)max_in * 2 * (4 * 4 + 8); }

extern "C" int dslb_view_boxes(const float* boxes, const int64_t* labels, const int32_t* off, const dslb_view_t* views,
                               int B, int max_in, int max_out, void* workspace, size_t ws_bytes, float* out_boxes,
                               int64_t* out_labels, int32_t* out_off, void* stream) {
  DSLB_CHECK_ARG(boxes && off && views && workspace && out_boxes && out_off, "dslb_view_boxes: null argument");
  DSLB_CHECK_ARG(B >= 1 && B <= VIEW_MAX_IMGS && B * 32 <= 1024, "dslb_view_boxes: at most 32 images per call");
  DSLB_CHECK_ARG(ws_bytes >= dslb_view_boxes_workspace_bytes(max_in), "dslb_view_boxes: workspace too small");
  DSLB_CHECK_ARG((labels == nullptr) == (out_labels == nullptr), "dslb_view_boxes: labels and out_labels go together");
  static_assert(sizeof(dslb_view_t) == sizeof(ViewDev), "dslb_view_t layout");
  float* stage = (float*)workspace;
  long long* stage_lab = labels ? (long long*)((char*)workspace + (size_t)max_in * 2 * 16) : nullptr;
  view_boxes_kernel<<<1, 32 * B, 0, (cudaStream_t)stream>>>(boxes, (const long long*)labels, off, (const ViewDev*)views, B,
                                                          max_out, stage, stage_lab, out_boxes, (long long*)out_labels,
                                                          out_off);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

namespace dslb {
// Scale-invariant extra sample (semi_epoch_based_runner.py:186-204): the half-resolution copy of the LAST image of the
// batch takes that image's boxes divided by two. Packed lists: image B-1 = [off[B-1], off[B]) is appended as image B.
__global__ void append_scaled_boxes_kernel(float4* __restrict__ boxes, long long* __restrict__ labels,
                                           int* __restrict__ off, int B, float scale, int max_boxes) {
  const int a = off[B - 1], b = off[B];
  const int n = min(b - a, max_boxes - b);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float4 v = boxes[a + i];
    boxes[b + i] = make_float4(__fmul_rn(v.x, scale), __fmul_rn(v.y, scale), __fmul_rn(v.z, scale), __fmul_rn(v.w, scale));
    if (labels) labels[b + i] = labels[a + i];
  }
  if (threadIdx.x == 0) off[B + 1] = b + max(n, 0);
}
}  // namespace dslb

extern "C" int dslb_append_scaled_boxes(float* boxes, int64_t* labels, int32_t* off, int B, float scale, int max_boxes,
                                        void* stream) {
  DSLB_CHECK_ARG(boxes && off && B >= 1 && max_boxes >= 0, "dslb_append_scaled_boxes: bad arguments");
  dslb::append_scaled_boxes_kernel<<<1, 128, 0, (cudaStream_t)stream>>>((float4*)boxes, (long long*)labels, off, B, scale,
                                                                        max_boxes);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" int dslb_pad_batch(const float* const* imgs_dev, const int32_t* hw_dev, float* out, int B, int C, int H, int W,
                              void* stream) {
  DSLB_CHECK_ARG(imgs_dev && hw_dev && out && B >= 1 && C >= 1 && H >= 1 && W >= 1, "dslb_pad_batch: bad arguments");
  const long long total = (long long)B * C * H * W;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
  pad_batch_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(imgs_dev, hw_dev, out, B, C, H, W);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" int dslb_view_images(const uint8_t* const* srcs_dev, const dslb_image_view_t* views_dev, int B,
                                const float* mean, const float* std, int to_rgb, float* out, int H, int W, void* stream) {
  DSLB_CHECK_ARG(srcs_dev && views_dev && mean && std && out, "dslb_view_images: null argument");
  DSLB_CHECK_ARG(B >= 1 && B <= 65535 && H >= 1 && W >= 1, "dslb_view_images: bad batch or output size");
  static_assert(sizeof(dslb_image_view_t) == sizeof(ImageViewDev), "dslb_image_view_t layout");
  ViewImageParams prm;
  for (int c = 0; c < 3; ++c) {
    DSLB_CHECK_ARG(std[c] != 0.f, "dslb_view_images: std must be non-zero");
    prm.mean[c] = mean[c];
    prm.inv_std[c] = 1.0 / (double)std[c];
  }
  prm.to_rgb = to_rgb ? 1 : 0;
  const dim3 grid(cdiv(W, VI_TX), cdiv(H, VI_TY * VI_ROWS), B);
  DSLB_CHECK_ARG(grid.y <= 65535, "dslb_view_images: output too tall");
  view_images_kernel<<<grid, dim3(VI_TX, VI_TY), 0, (cudaStream_t)stream>>>(srcs_dev, (const ImageViewDev*)views_dev, prm,
                                                                            out, H, W);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}
