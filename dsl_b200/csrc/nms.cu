// Teacher post-processing that the reference runs on the host (D2H + numpy + JSON): class-aware NMS of the gated
// candidates (multiclass_nms -> mmcv batched_nms, mmdet/core/post_processing/bbox_nms.py:78-94) and the pseudo-label
// rule chain of UnlabelPredHook + SemiCOCODataset (mmdet/runner/hooks/unlabel_pred_hook.py:20-38,142-165;
// mmdet/datasets/semicoco.py:220-269). Everything stays on the device and feeds the student's target kernel directly.
//   K1 sort      per image: 64-bit keys (score desc, candidate id asc) bitonic-sorted in shared memory; also max coord
//   K2 mask      64x64 IoU tiles on the class-offset boxes -> suppression bit matrix
//   K3 reduce    one warp per image walks the sorted list greedily, emits the first max_det survivors
//   K4 labels    per image (<= 128 detections): hook gate + int() truncation + per-class NMS + dataset filter rule
#include "common.h"

namespace dslb {

constexpr int NMS_TILE = 64;

__device__ __forceinline__ bool iou_gt(const float4& a, const float4& b, float thr) {
  // mmcv nms_cuda_kernel.cuh devIoU, offset 0; every operation rounded separately (no FMA contraction)
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float width = fmaxf(__fsub_rn(right, left), 0.f), height = fmaxf(__fsub_rn(bottom, top), 0.f);
  const float inter = __fmul_rn(width, height);
  const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const float sb = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter)) > thr;
}

// ---- K1: sort the candidates of one image. key = (~score_bits << 32) | id  (scores are >= 0, so the bit pattern is
// monotonic): ascending key order = descending score, ties by ascending candidate id = (point * C + label).
__global__ void __launch_bounds__(1024) nms_sort_kernel(const float* __restrict__ scores, const int* __restrict__ labels,
                                                        const int* __restrict__ points, const float* __restrict__ boxes,
                                                        const int* __restrict__ counts, int cap, int C,
                                                        int* __restrict__ order, float* __restrict__ maxc) {
  extern __shared__ unsigned long long keys[];  // npow2 entries
  __shared__ float smax[32];
  const int n_img = blockIdx.x;
  const int n = min(counts[n_img], cap);
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  float m = 0.f;  // boxes are clipped to >= 0, so 0 is a valid identity for the max
  for (int i = threadIdx.x; i < np2; i += blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < n) {
      const long long o = (long long)n_img * cap + i;
      const unsigned sb = __float_as_uint(scores[o]);
      const unsigned id = (unsigned)points[o] * (unsigned)C + (unsigned)labels[o];
      k = ((unsigned long long)(~sb) << 32) | id;
      const float4 b = reinterpret_cast<const float4*>(boxes)[o];
      m = fmaxf(m, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
    }
    keys[i] = k;
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t = fmaxf(t, smax[i]);
    maxc[n_img] = t;
  }
  for (int k = 2; k <= np2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = keys[i], b = keys[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            keys[i] = b;
            keys[l] = a;
          }
        }
      }
      __syncthreads();
    }
  // the id alone does not locate the slot (slots were claimed atomically): second pass maps ids back to slots through
  // a rank search. Cheaper: re-read every slot, binary-search its key in the sorted array, store slot at that rank.
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const long long o = (long long)n_img * cap + i;
    const unsigned sb = __float_as_uint(scores[o]);
    const unsigned id = (unsigned)points[o] * (unsigned)C + (unsigned)labels[o];
    const unsigned long long k = ((unsigned long long)(~sb) << 32) | id;
    int lo = 0, hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (keys[mid] < k) lo = mid + 1; else hi = mid;
    }
    order[(long long)n_img * cap + lo] = i;
  }
}

// ---- K2: suppression bits. Block (cb, rb, image): rows rb*64.., columns cb*64.. of the SORTED list.
__global__ void __launch_bounds__(NMS_TILE) nms_mask_kernel(const float* __restrict__ boxes, const int* __restrict__ labels,
                                                            const int* __restrict__ order, const int* __restrict__ counts,
                                                            const float* __restrict__ maxc, int cap, float thr,
                                                            unsigned long long* __restrict__ mask) {
  const int n_img = blockIdx.z;
  const int n = min(counts[n_img], cap);
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (rb * NMS_TILE >= n || cb * NMS_TILE >= n || cb < rb) return;
  __shared__ float4 cbox[NMS_TILE];
  const float off1 = __fadd_rn(maxc[n_img], 1.f);  // batched_nms: offsets = idxs * (max_coordinate + 1)
  const int words = cap / 64;
  auto load = [&](int sorted_idx) {
    const long long o = (long long)n_img * cap + order[(long long)n_img * cap + sorted_idx];
    float4 b = reinterpret_cast<const float4*>(boxes)[o];
    const float off = __fmul_rn((float)labels[o], off1);
    b.x = __fadd_rn(b.x, off);
    b.y = __fadd_rn(b.y, off);
    b.z = __fadd_rn(b.z, off);
    b.w = __fadd_rn(b.w, off);
    return b;
  };
  const int ccount = min(n - cb * NMS_TILE, NMS_TILE);
  if ((int)threadIdx.x < ccount) cbox[threadIdx.x] = load(cb * NMS_TILE + threadIdx.x);
  __syncthreads();
  const int r = rb * NMS_TILE + threadIdx.x;
  if (r < n) {
    const float4 a = load(r);
    unsigned long long bits = 0;
    const int start = (rb == cb) ? threadIdx.x + 1 : 0;
    for (int j = start; j < ccount; ++j)
      if (iou_gt(a, cbox[j], thr)) bits |= 1ull << j;
    mask[((long long)n_img * cap + r) * words + cb] = bits;
  }
}

// ---- K3: greedy walk, one warp per image. dets [B][max_det][5] (x1,y1,x2,y2,score), det_labels, det_count.
__global__ void __launch_bounds__(32) nms_reduce_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                                                        const int* __restrict__ labels, const int* __restrict__ order,
                                                        const int* __restrict__ counts,
                                                        const unsigned long long* __restrict__ mask, int cap, int max_det,
                                                        float* __restrict__ dets, int* __restrict__ det_labels,
                                                        int* __restrict__ det_count) {
  extern __shared__ unsigned long long remv[];  // cap / 64 words
  const int n_img = blockIdx.x;
  const int n = min(counts[n_img], cap);
  const int words = cap / 64;
  const int lane = threadIdx.x;
  for (int w = lane; w < words; w += 32) remv[w] = 0;
  __syncwarp();
  int kept = 0;
  const int nblk = (n + 63) / 64;
  for (int i = 0; i < n && kept < max_det; ++i) {
    const int wi = i >> 6;
    if (remv[wi] & (1ull << (i & 63))) continue;  // uniform across the warp (shared memory)
    const long long o = (long long)n_img * cap + order[(long long)n_img * cap + i];
    if (lane < 4) dets[((long long)n_img * max_det + kept) * 5 + lane] = boxes[o * 4 + lane];
    if (lane == 4) dets[((long long)n_img * max_det + kept) * 5 + 4] = scores[o];
    if (lane == 5) det_labels[(long long)n_img * max_det + kept] = labels[o];
    ++kept;
    const unsigned long long* row = mask + ((long long)n_img * cap + i) * words;
    for (int w = wi + lane; w < nblk; w += 32) remv[w] |= row[w];
    __syncwarp();
  }
  if (lane == 0) det_count[n_img] = kept;
}

// ---- K4: hook + dataset rule chain, one block of 128 threads per image, one detection per thread.
struct PseudoParams {
  const float* dets;       // [B][max_det][5], score-sorted survivors of multiclass_nms
  const int* det_labels;   // [B][max_det]
  const int* det_count;    // [B]
  const double* thr_class; // [C] per-class ignore threshold (adathres "thres", default 0.3), fp64 like the JSON value
  const float* img_wh;     // [B][2] (width, height) of the image the boxes live in
  float* gt_boxes;         // [max_boxes][4] packed over images
  long long* gt_labels;    // [max_boxes]
  int* gt_off;             // [B+1]
  float* ig_boxes;         // [max_boxes][4]
  int* ig_off;             // [B+1]
  int B, max_det, C, max_boxes;
  float nms_iou;
  double infer_score_thr, ignore_lo;
  // adaptive-threshold statistics (unlabel_pred_hook.py:295-343), optional: per-class count / score sum of the boxes
  // the hook would have written to its JSON file and adathres() would have counted
  long long* stat_cnt;      // [C] or null
  double* stat_cum;         // [C]
  const double* stat_prev;  // [C] last epoch's thresholds (-inf = class absent from the history) or null = first pass
  // optional export of what the hook writes to the image's JSON file (alive after its NMS, before the dataset rule)
  float* sv_boxes;          // [B][max_det][4] or null
  float* sv_scores;         // [B][max_det]
  int* sv_labels;           // [B][max_det]
  int* sv_count;            // [B]
};

__global__ void __launch_bounds__(128) pseudo_label_kernel(const __grid_constant__ PseudoParams P) {
  // single block; images are processed one after the other so the packed output offsets are a running sum
  __shared__ float4 sbox[128];
  __shared__ double sscore[128];
  __shared__ int slabel[128];
  __shared__ int alive[128];
  __shared__ int rank_of[128];
  __shared__ int gt_base, ig_base;
  const int t = threadIdx.x;
  if (t == 0) {
    gt_base = 0;
    ig_base = 0;
    P.gt_off[0] = 0;
    P.ig_off[0] = 0;
  }
  __syncthreads();
  for (int n_img = 0; n_img < P.B; ++n_img) {
    const int n = min(P.det_count[n_img], min(P.max_det, 128));
    // -- hook gate (parse_det_results): score >= infer_score_thre, int() truncation, round(score, 6)
    bool ok = false;
    float4 b = make_float4(0, 0, 0, 0);
    double s6 = 0.0;
    int lab = -1;
    if (t < n) {
      const float* d = P.dets + ((long long)n_img * P.max_det + t) * 5;
      const float sc = d[4];
      lab = P.det_labels[(long long)n_img * P.max_det + t];
      ok = !((double)sc < P.infer_score_thr);
      b = make_float4(truncf(d[0]), truncf(d[1]), truncf(d[2]), truncf(d[3]));
      s6 = rint((double)sc * 1e6) / 1e6;
      // per-class NMS loop of save_results2file runs over range(0, len(id2cat) - 1) (unlabel_pred_hook.py:156); the
      // reference's category file carries a trailing background entry (tools/coco_convert2_semicoco_json.py:47-48,
      // voc_convert2_semivoc_json.py:63), so len(id2cat) - 1 == num_classes: every real class survives
      ok = ok && lab >= 0 && lab < P.C;
      // mmcv nms(score_threshold=0.1): scores > 0.1 in fp32
      ok = ok && ((float)s6 > 0.1f);  // hard-coded score_threshold=0.1 of the hook's nms call (:163)
    }
    sbox[t] = b;
    sscore[t] = s6;
    slabel[t] = lab;
    alive[t] = ok ? 1 : 0;
    __syncthreads();
    // -- order inside a class: score desc (mmcv nms sorts), ties by input order; output order: class asc, score desc
    int rk = 0;
    if (t < n && ok) {
      for (int j = 0; j < n; ++j) {
        if (j == t || !alive[j]) continue;
        const bool before = slabel[j] < lab || (slabel[j] == lab && ((float)sscore[j] > (float)s6 ||
                                                                     ((float)sscore[j] == (float)s6 && j < t)));
        rk += before ? 1 : 0;
      }
    }
    __syncthreads();
    rank_of[t] = -1;
    __syncthreads();
    if (t < n && ok) rank_of[rk] = t;
    __syncthreads();
    int m = 0;  // number of gated detections
    for (int j = 0; j < n; ++j) m += alive[j];
    // -- greedy per-class NMS in rank order (IoU on the truncated boxes, > thr suppresses)
    for (int r = 0; r < m; ++r) {
      const int i = rank_of[r];
      if (alive[i] && t < n && ok && t != i && slabel[i] == lab) {
        // t is later in the order than i  <=>  its rank is larger
        if (rk > r && iou_gt(sbox[i], b, P.nms_iou)) alive[t] = 0;
      }
      __syncthreads();
    }
    // -- dataset rule (semicoco.py:220-269), in output order; thread 0 walks the (<= 128) survivors
    if (t == 0) {
      const float W = P.img_wh[n_img * 2], H = P.img_wh[n_img * 2 + 1];
      int g = gt_base, q = ig_base, sv = 0;
      for (int r = 0; r < m; ++r) {
        const int i = rank_of[r];
        if (!alive[i]) continue;
        const float4 bb = sbox[i];
        if (P.sv_boxes) {
          const long long at = (long long)n_img * P.max_det + sv;
          reinterpret_cast<float4*>(P.sv_boxes)[at] = bb;
          P.sv_scores[at] = (float)sscore[i];
          P.sv_labels[at] = slabel[i];
          ++sv;
        }
        if (P.stat_cnt) {
          // adathres() reads the hook's JSON, i.e. every box alive here (before the dataset's geometry filter): counted
          // when score >= 0.3 (no history file yet) or >= last epoch's threshold of its class
          const double sj = (double)(float)sscore[i];
          const double gate = P.stat_prev ? P.stat_prev[slabel[i]] : 0.3;
          if (sj >= gate) {
            P.stat_cnt[slabel[i]] += 1;
            P.stat_cum[slabel[i]] += sj;
          }
        }
        const float iw = fmaxf(0.f, fminf(bb.z, W) - fmaxf(bb.x, 0.f));
        const float ih = fmaxf(0.f, fminf(bb.w, H) - fmaxf(bb.y, 0.f));
        if (iw * ih == 0.f) continue;
        if (bb.z - bb.x < 1.f || bb.w - bb.y < 1.f) continue;
        const double sc = (double)(float)sscore[i];  // JSON carries the fp32 value of the nms output
        const bool ignore = sc < P.thr_class[slabel[i]] && sc >= P.ignore_lo;
        if (ignore) {
          if (q < P.max_boxes) reinterpret_cast<float4*>(P.ig_boxes)[q] = bb;
          ++q;
        } else {
          if (g < P.max_boxes) {
            reinterpret_cast<float4*>(P.gt_boxes)[g] = bb;
            P.gt_labels[g] = slabel[i];
          }
          ++g;
        }
      }
      gt_base = min(g, P.max_boxes);
      ig_base = min(q, P.max_boxes);
      P.gt_off[n_img + 1] = gt_base;
      P.ig_off[n_img + 1] = ig_base;
      if (P.sv_boxes) P.sv_count[n_img] = sv;
    }
    __syncthreads();
  }
}

// adathres() tail (unlabel_pred_hook.py:344-361), fp64 like the reference's Python floats: over the classes that were
// counted at all, mean = sum(count) / #classes; weight_c = (mean / cum_c)^gamma2;
// thr_c = clip((cum_c / mean)^gamma1 * base, lo, hi). Classes never counted keep `absent_thr` / weight 0 and get
// prev_out = -inf ("not in history": always counted next epoch).
__global__ void adathres_finalize_kernel(const long long* __restrict__ cnt, const double* __restrict__ cum, int C,
                                         double gamma1, double gamma2, double base, double lo, double hi,
                                         double absent_thr, double* __restrict__ thr_out,
                                         double* __restrict__ weight_out, double* __restrict__ prev_out) {
  __shared__ double s_mean;
  if (threadIdx.x == 0) {
    long long tot = 0;
    int present = 0;
    for (int c = 0; c < C; ++c) {
      tot += cnt[c];
      present += cnt[c] > 0 ? 1 : 0;
    }
    s_mean = present ? (double)tot / (double)present : 0.0;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (cnt[c] > 0) {
      const double t = pow(cum[c] / s_mean, gamma1) * base;
      thr_out[c] = fmax(fmin(t, hi), lo);
      weight_out[c] = pow(s_mean / cum[c], gamma2);
      if (prev_out) prev_out[c] = thr_out[c];
    } else {
      thr_out[c] = absent_thr;
      weight_out[c] = 0.0;
      if (prev_out) prev_out[c] = -INFINITY;
    }
  }
}

}  // namespace dslb

using namespace dslb;

extern "C" size_t dslb_nms_workspace_bytes(int B, int cap) {
  // order [B][cap] int32 + maxc [B] float (padded) + mask [B][cap][cap/64] u64
  return (size_t)B * cap * 4 + 256 + (size_t)B * cap * (cap / 64) * 8;
}

extern "C" int dslb_multiclass_nms(const float* boxes, const float* scores, const int32_t* labels, const int32_t* points,
                                   const int32_t* counts, int B, int cap, int num_classes, float iou_thr, int max_det,
                                   void* workspace, size_t ws_bytes, float* dets, int32_t* det_labels,
                                   int32_t* det_count, void* stream) {
  DSLB_CHECK_ARG(boxes && scores && labels && points && counts && workspace && dets && det_labels && det_count,
                 "dslb_multiclass_nms: null argument");
  DSLB_CHECK_ARG(B >= 1 && cap >= 64 && cap <= 8192 && (cap & (cap - 1)) == 0, "dslb_multiclass_nms: cap must be a power of two in [64, 8192]");
  DSLB_CHECK_ARG(ws_bytes >= dslb_nms_workspace_bytes(B, cap), "dslb_multiclass_nms: workspace too small");
  DSLB_CHECK_ARG(max_det >= 1, "dslb_multiclass_nms: max_det");
  cudaStream_t s = (cudaStream_t)stream;
  int* order = (int*)workspace;
  float* maxc = (float*)((char*)workspace + (size_t)B * cap * 4);
  unsigned long long* mask = (unsigned long long*)((char*)workspace + (size_t)B * cap * 4 + 256);
  static bool attr_set = false;
  if (!attr_set) {
    DSLB_CHECK_CUDA(cudaFuncSetAttribute(nms_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
    attr_set = true;
  }
  nms_sort_kernel<<<B, 1024, (size_t)cap * 8, s>>>(scores, labels, points, boxes, counts, cap, num_classes, order, maxc);
  DSLB_CHECK_CUDA(cudaGetLastError());
  const int nb = cap / NMS_TILE;
  nms_mask_kernel<<<dim3(nb, nb, B), NMS_TILE, 0, s>>>(boxes, labels, order, counts, maxc, cap, iou_thr, mask);
  DSLB_CHECK_CUDA(cudaGetLastError());
  nms_reduce_kernel<<<B, 32, (size_t)(cap / 64) * 8, s>>>(boxes, scores, labels, order, counts, mask, cap, max_det, dets,
                                                          det_labels, det_count);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

static int pseudo_labels_launch(const float* dets, const int32_t* det_labels, const int32_t* det_count,
                                const double* thr_class, const float* img_wh, int B, int max_det, int num_classes,
                                double infer_score_thr, float nms_iou, double ignore_lo, int max_boxes, float* gt_boxes,
                                int64_t* gt_labels, int32_t* gt_off, float* ig_boxes, int32_t* ig_off, int64_t* stat_cnt,
                                double* stat_cum, const double* stat_prev, float* sv_boxes, float* sv_scores,
                                int32_t* sv_labels, int32_t* sv_count, void* stream) {
  DSLB_CHECK_ARG(dets && det_labels && det_count && thr_class && img_wh && gt_boxes && gt_labels && gt_off && ig_boxes &&
                     ig_off,
                 "dslb_pseudo_labels: null argument");
  DSLB_CHECK_ARG(B >= 1 && max_det >= 1 && max_det <= 128, "dslb_pseudo_labels: max_det must be in [1, 128]");
  DSLB_CHECK_ARG((stat_cnt == nullptr) == (stat_cum == nullptr), "dslb_pseudo_labels: stat_cnt and stat_cum go together");
  PseudoParams P;
  P.dets = dets; P.det_labels = det_labels; P.det_count = det_count; P.thr_class = thr_class; P.img_wh = img_wh;
  P.gt_boxes = gt_boxes; P.gt_labels = (long long*)gt_labels; P.gt_off = gt_off; P.ig_boxes = ig_boxes; P.ig_off = ig_off;
  P.B = B; P.max_det = max_det; P.C = num_classes; P.max_boxes = max_boxes;
  P.infer_score_thr = infer_score_thr; P.nms_iou = nms_iou; P.ignore_lo = ignore_lo;
  P.stat_cnt = (long long*)stat_cnt; P.stat_cum = stat_cum; P.stat_prev = stat_prev;
  P.sv_boxes = sv_boxes; P.sv_scores = sv_scores; P.sv_labels = sv_labels; P.sv_count = sv_count;
  pseudo_label_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(P);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" int dslb_pseudo_labels_stats(const float* dets, const int32_t* det_labels, const int32_t* det_count,
                                        const double* thr_class, const float* img_wh, int B, int max_det,
                                        int num_classes, double infer_score_thr, float nms_iou, double ignore_lo,
                                        int max_boxes, float* gt_boxes, int64_t* gt_labels, int32_t* gt_off,
                                        float* ig_boxes, int32_t* ig_off, int64_t* stat_cnt, double* stat_cum,
                                        const double* stat_prev, void* stream) {
  return pseudo_labels_launch(dets, det_labels, det_count, thr_class, img_wh, B, max_det, num_classes, infer_score_thr,
                              nms_iou, ignore_lo, max_boxes, gt_boxes, gt_labels, gt_off, ig_boxes, ig_off, stat_cnt,
                              stat_cum, stat_prev, nullptr, nullptr, nullptr, nullptr, stream);
}

extern "C" int dslb_pseudo_labels_saved(const float* dets, const int32_t* det_labels, const int32_t* det_count,
                                        const double* thr_class, const float* img_wh, int B, int max_det,
                                        int num_classes, double infer_score_thr, float nms_iou, double ignore_lo,
                                        int max_boxes, float* gt_boxes, int64_t* gt_labels, int32_t* gt_off,
                                        float* ig_boxes, int32_t* ig_off, float* saved_boxes, float* saved_scores,
                                        int32_t* saved_labels, int32_t* saved_count, void* stream) {
  DSLB_CHECK_ARG(saved_boxes && saved_scores && saved_labels && saved_count, "dslb_pseudo_labels_saved: null argument");
  return pseudo_labels_launch(dets, det_labels, det_count, thr_class, img_wh, B, max_det, num_classes, infer_score_thr,
                              nms_iou, ignore_lo, max_boxes, gt_boxes, gt_labels, gt_off, ig_boxes, ig_off, nullptr,
                              nullptr, nullptr, saved_boxes, saved_scores, saved_labels, saved_count, stream);
}

extern "C" int dslb_pseudo_labels(const float* dets, const int32_t* det_labels, const int32_t* det_count,
                                  const double* thr_class, const float* img_wh, int B, int max_det, int num_classes,
                                  double infer_score_thr, float nms_iou, double ignore_lo, int max_boxes, float* gt_boxes,
                                  int64_t* gt_labels, int32_t* gt_off, float* ig_boxes, int32_t* ig_off, void* stream) {
  return dslb_pseudo_labels_stats(dets, det_labels, det_count, thr_class, img_wh, B, max_det, num_classes,
                                  infer_score_thr, nms_iou, ignore_lo, max_boxes, gt_boxes, gt_labels, gt_off, ig_boxes,
                                  ig_off, nullptr, nullptr, nullptr, stream);
}

extern "C" int dslb_adathres_finalize(const int64_t* stat_cnt, const double* stat_cum, int num_classes, double gamma1,
                                      double gamma2, double base, double lo, double hi, double absent_thr,
                                      double* thr_out, double* weight_out, double* prev_out, void* stream) {
  DSLB_CHECK_ARG(stat_cnt && stat_cum && thr_out && weight_out && num_classes >= 1, "dslb_adathres_finalize: bad arguments");
  adathres_finalize_kernel<<<1, 128, 0, (cudaStream_t)stream>>>((const long long*)stat_cnt, stat_cum, num_classes, gamma1,
                                                               gamma2, base, lo, hi, absent_thr, thr_out, weight_out,
                                                               prev_out);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}
