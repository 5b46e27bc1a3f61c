// Teacher-side decode + score gating (reference: FCOSHead._get_bboxes, mmdet/models/dense_heads/fcos_head.py:406-527;
// the `scores > score_thr` gate of multiclass_nms, mmdet/core/post_processing/bbox_nms.py:34-67).
//   kernel A  per FPN point: max_c(sigmoid(cls_c)) * sigmoid(centerness)  -> the key of the per-level top-nms_pre
//   kernel B  for the selected points: distance2bbox + clip to img_shape + /scale_factor, then every class whose raw
//             sigmoid score passes score_thr is emitted as a candidate (box, score*centerness, label, point)
// Both stream the fp32 logits once (HBM-bound).
#include "common.h"

namespace dslb {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// 4 threads per point; thread `part` reads the float4 groups part, part + 4, ... of the point's C logits (any C % 4 == 0:
// COCO's 80 and VOC's 20 classes, configs/fcos_semi/voc/*.py), one shuffle-max at the end. max is order independent, so
// the result does not depend on how the classes are split.
__global__ void point_scores_kernel(const float* __restrict__ cls, const float* __restrict__ regctr,
                                    float* __restrict__ out, long long npts, int C, int ld_cls) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pt = t >> 2;
  const int part = (int)(t & 3);
  float m = -INFINITY;
  if (pt < npts) {
    const float4* p = reinterpret_cast<const float4*>(cls + pt * ld_cls);
    for (int i = part; i < C / 4; i += 4) {
      const float4 v = __ldg(p + i);
      m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
  }
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
  if (pt < npts && part == 0) out[pt] = sigmoidf_(m) * sigmoidf_(__ldg(regctr + pt * 8 + 4));
}

// Per-level top-nms_pre of the point scores (fcos_head.py:452-460: `max_scores.topk(nms_pre)` per image and level) as a
// radix select: one block per (level, image); four 8-bit passes over the order-preserving integer image of the fp32
// scores find the K-th largest key T, then every point with key > T and the lowest-index points with key == T are
// compacted into sel[n][0..K). The SET equals torch.topk's wherever the K-th score is unique; ties at the cut are
// resolved towards the lower point index (torch leaves them unspecified). The order inside sel is unspecified: the
// consumer (decode_gate) claims candidate slots atomically anyway and the NMS sorts by score.
struct TopkLevel {
  const float* scores;  // [B][n]
  long long* sel;       // [B][K]
  int n, K;
};
struct TopkParams {
  TopkLevel lv[8];
  int nlv, B;
};

__device__ __forceinline__ uint32_t order_key(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(1024) topk_points_kernel(const __grid_constant__ TopkParams P) {
  __shared__ int hist[256];
  __shared__ uint32_t s_prefix;
  __shared__ int s_remaining, s_gt, s_warp[32], s_eq_base;
  const TopkLevel L = P.lv[blockIdx.x];
  const int img = blockIdx.y;
  const float* __restrict__ sc = L.scores + (long long)img * L.n;
  long long* __restrict__ out = L.sel + (long long)img * L.K;
  const int n = L.n, K = L.K, t = threadIdx.x;
  if (t == 0) {
    s_prefix = 0u;
    s_remaining = K;
    s_gt = 0;
    s_eq_base = 0;
  }
  uint32_t mask = 0u;
  for (int pass = 3; pass >= 0; --pass) {
    for (int i = t; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    for (int i = t; i < n; i += blockDim.x) {
      const uint32_t k = order_key(__ldg(sc + i));
      if ((k & mask) == prefix) atomicAdd(&hist[(k >> (8 * pass)) & 255], 1);
    }
    __syncthreads();
    if (t == 0) {
      int cum = 0, rem = s_remaining;
      for (int b = 255; b >= 0; --b) {
        if (cum + hist[b] >= rem) {
          s_prefix = prefix | ((uint32_t)b << (8 * pass));
          s_remaining = rem - cum;
          break;
        }
        cum += hist[b];
      }
    }
    mask |= 0xffu << (8 * pass);
    __syncthreads();
  }
  const uint32_t T = s_prefix;
  const int need_eq = s_remaining;   // points with key == T still to take (>= 1)
  const int n_gt = K - need_eq;      // points with key > T: all taken
  const int lane = t & 31, wid = t >> 5, nw = blockDim.x >> 5;
  for (int i0 = 0; i0 < n; i0 += blockDim.x) {
    const int i = i0 + t;
    const uint32_t k = i < n ? order_key(__ldg(sc + i)) : 0u;
    const bool gt = i < n && k > T, eq = i < n && k == T;
    if (gt) out[atomicAdd(&s_gt, 1)] = i;
    // ordered compaction of the ties: block-wide exclusive scan of `eq` in index order
    const unsigned bal = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) s_warp[wid] = __popc(bal);
    __syncthreads();
    int before = s_eq_base;
    for (int w = 0; w < wid; ++w) before += s_warp[w];
    const int slot = before + __popc(bal & ((1u << lane) - 1u));
    if (eq && slot < need_eq) out[n_gt + slot] = i;
    __syncthreads();
    if (t == 0) {
      int tot = 0;
      for (int w = 0; w < nw; ++w) tot += s_warp[w];
      s_eq_base += tot;
    }
    __syncthreads();
  }
}

struct DecodeParams {
  const float* cls;      // [B*hw][ld_cls]
  const float* regctr;   // [B*hw][8]  (eval mode: bbox already multiplied by the stride)
  const long long* sel;  // [B][K] selected point indices inside the level (row-major y*w+x), or NULL = all points
  const float* img_hw;   // [B][2] (H, W) of img_shape
  const float* scale_factor;  // [B][4] or NULL
  float* out_boxes;      // [B][cap][4]
  float* out_scores;     // [B][cap]
  int* out_labels;       // [B][cap]
  int* out_points;       // [B][cap] global point id (level offset + y*w+x) for tests
  int* counts;           // [B]
  int* overflow;         // sticky flag: set to 1 when an image exceeds cap (or NULL)
  int B, K, C, hw, w, stride, ld_cls, cap, point_offset;
  float score_thr;
};

__global__ void decode_gate_kernel(const __grid_constant__ DecodeParams P) {
  const int cv = P.C / 4;
  const long long total = (long long)P.B * P.K * cv;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(e % cv);
    const long long sk = e / cv;
    const int k = (int)(sk % P.K);
    const int n = (int)(sk / P.K);
    const long long pl = P.sel ? P.sel[(long long)n * P.K + k] : k;
    const long long lp = (long long)n * P.hw + pl;
    const float4 xv = __ldg(reinterpret_cast<const float4*>(P.cls + lp * P.ld_cls + c4 * 4));
    const float x[4] = {xv.x, xv.y, xv.z, xv.w};
    float s[4];
    bool any = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s[j] = sigmoidf_(x[j]);
      any |= s[j] > P.score_thr;
    }
    if (!any) continue;
    const int y = (int)(pl / P.w), xq = (int)(pl - (long long)y * P.w);
    const float px = (float)(xq * P.stride + P.stride / 2), py = (float)(y * P.stride + P.stride / 2);
    const float4 d = __ldg(reinterpret_cast<const float4*>(P.regctr + lp * 8));
    const float ctr = sigmoidf_(__ldg(P.regctr + lp * 8 + 4));
    const float H = P.img_hw[n * 2], W = P.img_hw[n * 2 + 1];
    float b[4] = {px - d.x, py - d.y, px + d.z, py + d.w};  // distance2bbox, core/bbox/transforms.py:119-162
    b[0] = fminf(fmaxf(b[0], 0.f), W);
    b[1] = fminf(fmaxf(b[1], 0.f), H);
    b[2] = fminf(fmaxf(b[2], 0.f), W);
    b[3] = fminf(fmaxf(b[3], 0.f), H);
    if (P.scale_factor) {
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = __fdiv_rn(b[j], P.scale_factor[n * 4 + j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (s[j] > P.score_thr) {  // gate on the RAW class score; centerness is applied after (bbox_nms.py:51-62)
        const int slot = atomicAdd(P.counts + n, 1);
        if (slot < P.cap) {
          const long long o = (long long)n * P.cap + slot;
          reinterpret_cast<float4*>(P.out_boxes)[o] = make_float4(b[0], b[1], b[2], b[3]);
          P.out_scores[o] = __fmul_rn(s[j], ctr);
          P.out_labels[o] = c4 * 4 + j;
          P.out_points[o] = P.point_offset + (int)pl;
        } else if (P.overflow) {
          *P.overflow = 1;
        }
      }
    }
  }
}

}  // namespace dslb

using namespace dslb;

extern "C" int dslb_fcos_point_scores(const float* cls, const float* regctr, float* out, long long npts, int C,
                                      int ld_cls, void* stream) {
  DSLB_CHECK_ARG(cls && regctr && out, "dslb_fcos_point_scores: null argument");
  DSLB_CHECK_ARG(C > 0 && C % 4 == 0 && ld_cls % 4 == 0, "dslb_fcos_point_scores: C and ld_cls must be multiples of 4");
  if (npts == 0) return DSLB_OK;
  const long long threads = npts * 4;
  point_scores_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(cls, regctr, out, npts, C,
                                                                                           ld_cls);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" int dslb_fcos_topk_points(const float* const* scores, int64_t* const* sel, const int32_t* n, const int32_t* K,
                                     int nlevels, int B, void* stream) {
  DSLB_CHECK_ARG(scores && sel && n && K, "dslb_fcos_topk_points: null argument");
  DSLB_CHECK_ARG(nlevels >= 0 && nlevels <= 8 && B >= 0, "dslb_fcos_topk_points: nlevels %d not in [0,8]", nlevels);
  if (nlevels == 0 || B == 0) return DSLB_OK;
  TopkParams P;
  P.nlv = nlevels;
  P.B = B;
  for (int i = 0; i < nlevels; ++i) {
    DSLB_CHECK_ARG(scores[i] && sel[i] && K[i] > 0 && K[i] <= n[i], "dslb_fcos_topk_points: level %d needs 0 < K <= n", i);
    P.lv[i].scores = scores[i];
    P.lv[i].sel = (long long*)sel[i];
    P.lv[i].n = n[i];
    P.lv[i].K = K[i];
  }
  topk_points_kernel<<<dim3(nlevels, B), 1024, 0, (cudaStream_t)stream>>>(P);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" int dslb_fcos_decode_gate(const float* cls, const float* regctr, const int64_t* sel, int B, int K, int C,
                                     int h, int w, int stride, int ld_cls, const float* img_hw,
                                     const float* scale_factor, float score_thr, int point_offset, float* out_boxes,
                                     float* out_scores, int32_t* out_labels, int32_t* out_points, int32_t* counts,
                                     int cap, int32_t* overflow, void* stream) {
  DSLB_CHECK_ARG(cls && regctr && img_hw && out_boxes && out_scores && out_labels && out_points && counts,
                 "dslb_fcos_decode_gate: null argument");
  DSLB_CHECK_ARG(C % 4 == 0 && K >= 0 && cap > 0, "dslb_fcos_decode_gate: bad sizes");
  if (K == 0 || B == 0) return DSLB_OK;
  DecodeParams P;
  P.cls = cls;
  P.regctr = regctr;
  P.sel = (const long long*)sel;
  P.img_hw = img_hw;
  P.scale_factor = scale_factor;
  P.out_boxes = out_boxes;
  P.out_scores = out_scores;
  P.out_labels = out_labels;
  P.out_points = out_points;
  P.counts = counts;
  P.overflow = overflow;
  P.B = B;
  P.K = K;
  P.C = C;
  P.hw = h * w;
  P.w = w;
  P.stride = stride;
  P.ld_cls = ld_cls;
  P.cap = cap;
  P.point_offset = point_offset;
  P.score_thr = score_thr;
  const long long total = (long long)B * K * (C / 4);
  long long blocks = (total + 255) / 256;
  const long long capb = (long long)num_sms() * 8;
  if (blocks > capb) blocks = capb;
  decode_gate_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(P);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}
