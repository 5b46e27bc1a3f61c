// Teacher-side decode + score gating (reference: FCOSHead._get_bboxes, mmdet/models/dense_heads/fcos_head.py:406-527;
// the `scores > score_thr` gate of multiclass_nms, mmdet/core/post_processing/bbox_nms.py:34-67).
//   kernel A  per FPN point: max_c(sigmoid(cls_c)) * sigmoid(centerness)  -> the key of the per-level top-nms_pre
//   kernel B  for the selected points: distance2bbox + clip to img_shape + /scale_factor, then every class whose raw
//             sigmoid score passes score_thr is emitted as a candidate (box, score*centerness, label, point)
// Both stream the fp32 logits once (HBM-bound).
#include "common.h"

namespace dslb {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// 4 threads per point, each reading C/4 logits (C % 16 == 0): coalesced enough, one shuffle-max at the end
__global__ void point_scores_kernel(const float* __restrict__ cls, const float* __restrict__ regctr,
                                    float* __restrict__ out, long long npts, int C, int ld_cls) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pt = t >> 2;
  const int part = (int)(t & 3);
  float m = -INFINITY;
  if (pt < npts) {
    const int per = C / 4;
    const float4* p = reinterpret_cast<const float4*>(cls + pt * ld_cls + part * per);
    for (int i = 0; i < per / 4; ++i) {
      const float4 v = __ldg(p + i);
      m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
  }
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
  if (pt < npts && part == 0) out[pt] = sigmoidf_(m) * sigmoidf_(__ldg(regctr + pt * 8 + 4));
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
  DSLB_CHECK_ARG(C % 16 == 0 && ld_cls % 4 == 0, "dslb_fcos_point_scores: C must be a multiple of 16");
  if (npts == 0) return DSLB_OK;
  const long long threads = npts * 4;
  point_scores_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(cls, regctr, out, npts, C,
                                                                                           ld_cls);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" int dslb_fcos_decode_gate(const float* cls, const float* regctr, const int64_t* sel, int B, int K, int C,
                                     int h, int w, int stride, int ld_cls, const float* img_hw,
                                     const float* scale_factor, float score_thr, int point_offset, float* out_boxes,
                                     float* out_scores, int32_t* out_labels, int32_t* out_points, int32_t* counts,
                                     int cap, void* stream) {
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
