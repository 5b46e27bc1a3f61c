// Dense pseudo-label path of FCOSHead.loss (reference: mmdet/models/dense_heads/fcos_head.py:170-338):
//   kernel 1  per-point target assignment (get_points + get_targets for GT and for ignore boxes + ignore/unlabeled
//             weights + centerness targets), bit-exact with the reference's fp32 compare logic, and the two global
//             normalisers (num_pos, sum of centerness targets);
//   kernel 2  sigmoid focal loss over all points x classes, GIoU + centerness BCE on positives, the DSL
//             scale-invariant soft loss, the block-reduced loss sums, and d(loss)/d(head outputs) written directly in
//             the layouts the dgrad/wgrad convs consume.
// Both are HBM-bound (kernel 2 streams the fp32 logits once and writes the bf16 gradient once).
#include "common.h"
#include "loss_math.cuh"

#include <cuda_bf16.h>

namespace dslb {

constexpr int MAX_LEVELS = 8;
constexpr float FCOS_INF = 1e8f;  // fcos_head.py:11

struct LossLevel {
  const float* cls;
  const float* regctr;
  __nv_bfloat16* dcls_bf16;
  float* dcls_f32;
  __nv_bfloat16* dregctr_bf16;
  float* dregctr_f32;
  int h, w, stride;
  int ld_cls, ld_dcls, ld_dreg;
  float rr_lo, rr_hi, scale, cs_radius;  // cs_radius = float(stride * center_sample_radius)
  long long pt_begin;                    // first flat point index of this level (level-major, image-major inside)
};

struct LossParams {
  LossLevel lv[MAX_LEVELS];
  int nlevels, B, C;
  long long npoints;
  // targets
  const float* gt_boxes;
  const long long* gt_labels;
  const int* gt_off;
  const float* ig_boxes;
  const int* ig_off;
  int center_sampling, norm_on_bbox;
  float loss_weight;
  int n_labeled;  // images [0, n_labeled) are "labeled" (weight 1), the rest x loss_weight
  long long* labels;
  float* bbox_targets;
  float* weights;
  float* ctr_targets;
  double* counts;  // [0] num_pos, [1] sum of centerness targets (this rank)
  // loss
  const float* norm;  // [0] max(mean_ranks(num_pos),1)  [1] max(mean_ranks(ctr_sum),1e-6)
  float alpha, gamma;
  double* loss_sums;  // [0] cls [1] bbox [2] centerness [3] sisoft
  float* dscale;      // [nlevels] gradient of the per-level Scale parameter
  const float* level_scales;  // [nlevels] device copy of the Scale values (overrides lv[].scale when non-null)
  float si_weight;    // 0 = off; else soft_weight (or soft_weight/1000 while warming up); needs odd B
};

// ------------------------------------------------------------------------------------------------ kernel 1
// One pass of fcos_head.py:623-705 for one point against the boxes of its image. Returns the chosen GT index
// (first minimum of the INF-masked areas) and whether any box matched. No fused multiply-adds anywhere: every
// operation is a single IEEE fp32 op, as in the reference's elementwise torch code.
struct Assign {
  int idx;
  bool matched;
  float l, t, r, b;
};

__device__ __forceinline__ Assign assign_point(float px, float py, const float4* __restrict__ boxes, int nbox,
                                               bool center_sampling, float cs_radius, float rr_lo, float rr_hi) {
  Assign a;
  a.idx = 0;
  a.matched = false;
  a.l = a.t = a.r = a.b = 0.f;
  float best = 0.f;
  for (int g = 0; g < nbox; ++g) {
    const float4 bx = boxes[g];
    const float l = __fsub_rn(px, bx.x), r = __fsub_rn(bx.z, px);
    const float t = __fsub_rn(py, bx.y), b = __fsub_rn(bx.w, py);
    bool inside;
    if (center_sampling) {
      const float cx = __fmul_rn(__fadd_rn(bx.x, bx.z), 0.5f), cy = __fmul_rn(__fadd_rn(bx.y, bx.w), 0.5f);
      const float xmin = __fsub_rn(cx, cs_radius), ymin = __fsub_rn(cy, cs_radius);
      const float xmax = __fadd_rn(cx, cs_radius), ymax = __fadd_rn(cy, cs_radius);
      const float c0 = xmin > bx.x ? xmin : bx.x;
      const float c1 = ymin > bx.y ? ymin : bx.y;
      const float c2 = xmax > bx.z ? bx.z : xmax;
      const float c3 = ymax > bx.w ? bx.w : ymax;
      const float m = fminf(fminf(__fsub_rn(px, c0), __fsub_rn(py, c1)), fminf(__fsub_rn(c2, px), __fsub_rn(c3, py)));
      inside = m > 0.f;
    } else {
      inside = fminf(fminf(l, t), fminf(r, b)) > 0.f;
    }
    const float mx = fmaxf(fmaxf(l, t), fmaxf(r, b));
    const bool in_range = (mx >= rr_lo) && (mx <= rr_hi);
    float area = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
    if (!inside || !in_range) area = FCOS_INF;
    if (g == 0 || area < best) {  // strict '<': the first index wins ties (areas.min(dim=1), fcos_head.py:699)
      best = area;
      a.idx = g;
      a.l = l;
      a.t = t;
      a.r = r;
      a.b = b;
    }
  }
  a.matched = nbox > 0 && best != FCOS_INF;
  return a;
}

__global__ void __launch_bounds__(256) fcos_targets_kernel(const __grid_constant__ LossParams P) {
  __shared__ float4 s_gt[256];
  __shared__ float s_red[2][8];
  const int lvl = blockIdx.z, n = blockIdx.y;
  const LossLevel& L = P.lv[lvl];
  const int hw = L.h * L.w;
  const int p0 = blockIdx.x * 256;
  if (p0 >= hw) return;
  const int pt = p0 + threadIdx.x;
  const bool active = pt < hw;
  const int y = active ? pt / L.w : 0, x = active ? pt - (pt / L.w) * L.w : 0;
  // points: idx * stride + stride // 2 (integer-valued floats; anchor_free_head.py:287-321, fcos_head.py:550-560)
  const float px = (float)(x * L.stride + L.stride / 2), py = (float)(y * L.stride + L.stride / 2);

  // --- real / pseudo GT
  const int g0 = P.gt_off[n], g1 = P.gt_off[n + 1];
  Assign best;
  best.idx = 0;
  best.matched = false;
  best.l = best.t = best.r = best.b = 0.f;
  float best_area = 0.f;
  bool any = false;
  // boxes are staged through shared memory 256 at a time; the running (area, index) minimum carries across stages
  for (int base = g0; base < g1; base += 256) {
    const int cnt = min(256, g1 - base);
    __syncthreads();
    if (threadIdx.x < cnt) s_gt[threadIdx.x] = reinterpret_cast<const float4*>(P.gt_boxes)[base + threadIdx.x];
    __syncthreads();
    if (active) {
      for (int g = 0; g < cnt; ++g) {
        const float4 bx = s_gt[g];
        const float l = __fsub_rn(px, bx.x), r = __fsub_rn(bx.z, px);
        const float t = __fsub_rn(py, bx.y), b = __fsub_rn(bx.w, py);
        bool inside;
        if (P.center_sampling) {
          const float cx = __fmul_rn(__fadd_rn(bx.x, bx.z), 0.5f), cy = __fmul_rn(__fadd_rn(bx.y, bx.w), 0.5f);
          const float xmin = __fsub_rn(cx, L.cs_radius), ymin = __fsub_rn(cy, L.cs_radius);
          const float xmax = __fadd_rn(cx, L.cs_radius), ymax = __fadd_rn(cy, L.cs_radius);
          const float c0 = xmin > bx.x ? xmin : bx.x;
          const float c1 = ymin > bx.y ? ymin : bx.y;
          const float c2 = xmax > bx.z ? bx.z : xmax;
          const float c3 = ymax > bx.w ? bx.w : ymax;
          const float m =
              fminf(fminf(__fsub_rn(px, c0), __fsub_rn(py, c1)), fminf(__fsub_rn(c2, px), __fsub_rn(c3, py)));
          inside = m > 0.f;
        } else {
          inside = fminf(fminf(l, t), fminf(r, b)) > 0.f;
        }
        const float mx = fmaxf(fmaxf(l, t), fmaxf(r, b));
        const bool in_range = (mx >= L.rr_lo) && (mx <= L.rr_hi);
        float area = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
        if (!inside || !in_range) area = FCOS_INF;
        if (!any || area < best_area) {  // strict '<': first index wins ties (fcos_head.py:699)
          any = true;
          best_area = area;
          best.idx = base - g0 + g;
          best.l = l;
          best.t = t;
          best.r = r;
          best.b = b;
        }
      }
    }
  }
  const bool pos = active && any && best_area != FCOS_INF;

  // --- ignore boxes (labels all num_classes-1, fcos_head.py:208-215): only "matched or not" matters
  bool ig_hit = false;
  if (P.ig_off) {
    const int i0 = P.ig_off[n], i1 = P.ig_off[n + 1];
    for (int base = i0; base < i1; base += 256) {
      const int cnt = min(256, i1 - base);
      __syncthreads();
      if (threadIdx.x < cnt) s_gt[threadIdx.x] = reinterpret_cast<const float4*>(P.ig_boxes)[base + threadIdx.x];
      __syncthreads();
      if (active && !ig_hit) {
        const Assign a = assign_point(px, py, s_gt, cnt, P.center_sampling != 0, L.cs_radius, L.rr_lo, L.rr_hi);
        ig_hit = a.matched;
      }
    }
  }

  float ctr_t = 0.f;
  if (active) {
    const long long o = L.pt_begin + (long long)n * hw + pt;
    long long label = P.C;
    float tl = 0.f, tt = 0.f, tr = 0.f, tb = 0.f;
    if (g1 > g0) {
      // bbox_targets of BG points are those of the arg-min GT too (index 0 when nothing matched), as in the
      // reference's `bbox_targets[range(num_points), min_area_inds]`
      tl = best.l;
      tt = best.t;
      tr = best.r;
      tb = best.b;
      if (P.norm_on_bbox) {
        const float s = (float)L.stride;
        tl = __fdiv_rn(tl, s);
        tt = __fdiv_rn(tt, s);
        tr = __fdiv_rn(tr, s);
        tb = __fdiv_rn(tb, s);
      }
      if (pos) label = P.gt_labels[g0 + best.idx];
    }
    P.labels[o] = label;
    reinterpret_cast<float4*>(P.bbox_targets)[o] = make_float4(tl, tt, tr, tb);
    // ignore mask: weight 0 where an ignore box claims a BACKGROUND point (fcos_head.py:297-304)
    float wgt = (P.ig_off && ig_hit && !pos) ? 0.f : 1.f;
    if (P.loss_weight != 1.0f && n >= P.n_labeled) wgt = __fmul_rn(wgt, P.loss_weight);  // fcos_head.py:217-235
    P.weights[o] = wgt;
    if (pos) {  // fcos_head.py:707-726
      const float lr = __fdiv_rn(fminf(tl, tr), fmaxf(tl, tr));
      const float tbv = __fdiv_rn(fminf(tt, tb), fmaxf(tt, tb));
      ctr_t = __fsqrt_rn(__fmul_rn(lr, tbv));
    }
    P.ctr_targets[o] = ctr_t;
  }
  // block reduction of (num_pos, sum ctr_t)
  float c = pos ? 1.f : 0.f, s = ctr_t;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c += __shfl_xor_sync(0xffffffffu, c, o);
    s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_red[0][threadIdx.x >> 5] = c;
    s_red[1][threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double cc = 0, ss = 0;
    for (int k = 0; k < 8; ++k) {
      cc += s_red[0][k];
      ss += s_red[1][k];
    }
    if (cc != 0) atomicAdd(P.counts, cc);
    if (ss != 0) atomicAdd(P.counts + 1, ss);
  }
}

// norm[0] = max(mean over ranks of num_pos, 1); norm[1] = max(mean over ranks of ctr_sum, 1e-6)
// (reduce_mean, fcos_head.py:266,273-274; `counts` already holds the all-reduced SUM over `world` ranks)
__global__ void fcos_norm_kernel(const double* __restrict__ counts, float world, float* __restrict__ norm) {
  norm[0] = fmaxf((float)(counts[0] / world), 1.0f);
  norm[1] = fmaxf((float)(counts[1] / world), 1e-6f);
}

// ------------------------------------------------------------------------------------------------ kernel 2
__device__ __forceinline__ int loss_find_level(const LossParams& P, long long pt) {
  int l = 0;
  while (l + 1 < P.nlevels && pt >= P.lv[l + 1].pt_begin) ++l;
  return l;
}

__global__ void __launch_bounds__(256) fcos_loss_kernel(const __grid_constant__ LossParams P) {
  __shared__ double s_red[4][8];
  const int cv = P.C / 4;  // float4 chunks per point
  const long long total = P.npoints * cv;
  const float inv_npos = 1.f / P.norm[0], inv_den = 1.f / P.norm[1];
  float acc_cls = 0.f, acc_box = 0.f, acc_ctr = 0.f, acc_si = 0.f;

  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long pt = e / cv;
    const int c4 = (int)(e - pt * cv);
    const int lvl = loss_find_level(P, pt);
    const LossLevel& L = P.lv[lvl];
    const long long lp = pt - L.pt_begin;  // point index inside the level: n*hw + y*w + x
    const int hw = L.h * L.w;
    const long long label = P.labels[pt];
    const float wgt = P.weights[pt];
    const float4 xv = __ldg(reinterpret_cast<const float4*>(L.cls + lp * L.ld_cls + c4 * 4));
    const float x[4] = {xv.x, xv.y, xv.z, xv.w};
    float g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float ls, gr;
      focal_elem(x[k], label == (long long)(c4 * 4 + k), P.alpha, P.gamma, ls, gr);
      acc_cls += ls * wgt;
      g[k] = gr * wgt * inv_npos;
    }
    // DSL scale-invariant soft loss (fcos_head.py:312-333): level i of image B-2 vs level i-1 of image B-1 (cropped)
    if (P.si_weight != 0.f) {
      const int n = (int)(lp / hw);
      const int rem = (int)(lp - (long long)n * hw);
      const int y = rem / L.w, xq = rem - y * L.w;
      if (n == P.B - 2 && lvl >= 1) {
        const LossLevel& Q = P.lv[lvl - 1];
        const long long qp = ((long long)(P.B - 1) * Q.h + y) * Q.w + xq;
        const float4 qv = __ldg(reinterpret_cast<const float4*>(Q.cls + qp * Q.ld_cls + c4 * 4));
        const float q[4] = {qv.x, qv.y, qv.z, qv.w};
        const float inv_cnt = P.si_weight / ((float)P.C * (float)hw);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float d = x[k] - q[k];
          acc_si += d * d * inv_cnt;
          g[k] += 2.f * d * inv_cnt;
        }
      }
      if (n == P.B - 1 && lvl + 1 < P.nlevels) {
        const LossLevel& U = P.lv[lvl + 1];
        if (y < U.h && xq < U.w) {
          const long long up = ((long long)(P.B - 2) * U.h + y) * U.w + xq;
          const float4 uv = __ldg(reinterpret_cast<const float4*>(U.cls + up * U.ld_cls + c4 * 4));
          const float u[4] = {uv.x, uv.y, uv.z, uv.w};
          const float inv_cnt = P.si_weight / ((float)P.C * (float)(U.h * U.w));
#pragma unroll
          for (int k = 0; k < 4; ++k) g[k] -= 2.f * (u[k] - x[k]) * inv_cnt;
        }
      }
    }
    if (L.dcls_f32) *reinterpret_cast<float4*>(L.dcls_f32 + lp * P.C + c4 * 4) = make_float4(g[0], g[1], g[2], g[3]);
    if (L.dcls_bf16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(g[0], g[1]), hi = __floats2bfloat162_rn(g[2], g[3]);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&lo);
      u.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(L.dcls_bf16 + lp * L.ld_dcls + c4 * 4) = u;
    }

    // --- regression + centerness branch: one thread per point (the one holding class chunk 0)
    if (c4 == 0) {
      float dr[5] = {0.f, 0.f, 0.f, 0.f, 0.f};   // d loss / d (bbox_pred l,t,r,b ; centerness logit)
      float dconv[4] = {0.f, 0.f, 0.f, 0.f};     // d loss / d conv_reg output (through relu(scale * x))
      if (label >= 0 && label < P.C) {           // positive (fcos_head.py:262-263)
        const int rem = (int)(lp % hw);
        const int y = rem / L.w, xq = rem - y * L.w;
        const float px = (float)(xq * L.stride + L.stride / 2), py = (float)(y * L.stride + L.stride / 2);
        const float4 bp = __ldg(reinterpret_cast<const float4*>(L.regctr + lp * 8));
        const float cl = __ldg(L.regctr + lp * 8 + 4);
        const float4 bt = reinterpret_cast<const float4*>(P.bbox_targets)[pt];
        const float ct = P.ctr_targets[pt];
        const float pb[4] = {px - bp.x, py - bp.y, px + bp.z, py + bp.w};  // distance2bbox
        const float tb[4] = {px - bt.x, py - bt.y, px + bt.z, py + bt.w};
        float gb[4];
        const float lb = giou_loss_grad(pb, tb, 1e-6f, gb);
        // flatten_weights = unlabeled weight only (the ignore mask is NOT applied to positives, fcos_head.py:281-292)
        const int n = (int)(lp / hw);
        const float fw = (P.loss_weight != 1.0f && n >= P.n_labeled) ? P.loss_weight : 1.f;
        const float wb = ct * fw;
        acc_box += lb * wb;
        const float s = wb * inv_den;
        dr[0] = -gb[0] * s;  // x1 = px - l
        dr[1] = -gb[1] * s;
        dr[2] = gb[2] * s;
        dr[3] = gb[3] * s;
        // centerness: BCE with logits (losses/cross_entropy_loss.py:73-112)
        acc_ctr += (fmaxf(cl, 0.f) - cl * ct + log1pf(expf(-fabsf(cl)))) * fw;
        dr[4] = (1.f / (1.f + expf(-cl)) - ct) * fw * inv_npos;
        // chain through bbox_pred = relu(scale_l * conv_reg): conv output = bbox_pred / scale_l where positive
        const float bpv[4] = {bp.x, bp.y, bp.z, bp.w};
        const float lscale = P.level_scales ? __ldg(P.level_scales + lvl) : L.scale;
        float ds = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (bpv[k] > 0.f) {
            dconv[k] = dr[k] * lscale;
            ds += dr[k] * (bpv[k] / lscale);
          }
        }
        if (P.dscale && ds != 0.f) atomicAdd(P.dscale + lvl, ds);
      }
      if (L.dregctr_f32) {
        float* o = L.dregctr_f32 + lp * 8;
        *reinterpret_cast<float4*>(o) = make_float4(dr[0], dr[1], dr[2], dr[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(dr[4], 0.f, 0.f, 0.f);
      }
      if (L.dregctr_bf16) {
        __nv_bfloat162 a = __floats2bfloat162_rn(dconv[0], dconv[1]), b = __floats2bfloat162_rn(dconv[2], dconv[3]);
        __nv_bfloat162 c = __floats2bfloat162_rn(dr[4], 0.f);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&a);
        u.y = *reinterpret_cast<uint32_t*>(&b);
        u.z = *reinterpret_cast<uint32_t*>(&c);
        u.w = 0u;
        *reinterpret_cast<uint4*>(L.dregctr_bf16 + lp * L.ld_dreg) = u;
      }
    }
  }

  // block reduction -> 4 fp64 atomics per block
  float v[4] = {acc_cls * inv_npos, acc_box * inv_den, acc_ctr * inv_npos, acc_si};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if ((threadIdx.x & 31) == 0) s_red[k][threadIdx.x >> 5] = (double)v[k];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0;
    for (int k = 0; k < 8; ++k) s += s_red[threadIdx.x][k];
    if (s != 0) atomicAdd(P.loss_sums + threadIdx.x, s);
  }
}

}  // namespace dslb

using namespace dslb;

static int fill_loss_params(LossParams& P, const dslb_fcos_level_t* levels, int nlevels, int B, int C) {
  DSLB_CHECK_ARG(levels && nlevels >= 1 && nlevels <= MAX_LEVELS, "fcos: nlevels %d out of range", nlevels);
  DSLB_CHECK_ARG(B >= 1 && C >= 4 && C % 4 == 0, "fcos: B=%d C=%d unsupported (C must be a multiple of 4)", B, C);
  memset(&P, 0, sizeof(P));
  P.nlevels = nlevels;
  P.B = B;
  P.C = C;
  long long pt = 0;
  for (int i = 0; i < nlevels; ++i) {
    LossLevel& d = P.lv[i];
    const dslb_fcos_level_t& s = levels[i];
    DSLB_CHECK_ARG(s.h > 0 && s.w > 0 && s.stride > 0, "fcos level %d: bad geometry", i);
    d.cls = s.cls;
    d.regctr = s.regctr;
    d.dcls_bf16 = (__nv_bfloat16*)s.dcls_bf16;
    d.dcls_f32 = s.dcls_f32;
    d.dregctr_bf16 = (__nv_bfloat16*)s.dregctr_bf16;
    d.dregctr_f32 = s.dregctr_f32;
    d.h = s.h;
    d.w = s.w;
    d.stride = s.stride;
    d.ld_cls = s.ld_cls > 0 ? s.ld_cls : C;
    d.ld_dcls = s.ld_dcls;
    d.ld_dreg = s.ld_dreg;
    d.rr_lo = s.rr_lo;
    d.rr_hi = s.rr_hi;
    d.scale = s.scale;
    d.cs_radius = s.cs_radius;
    d.pt_begin = pt;
    pt += (long long)B * s.h * s.w;
  }
  P.npoints = pt;
  return DSLB_OK;
}

extern "C" int dslb_fcos_targets(const dslb_fcos_level_t* levels, int nlevels, int B, int num_classes,
                                 const float* gt_boxes, const int64_t* gt_labels, const int32_t* gt_off,
                                 const float* ig_boxes, const int32_t* ig_off, int center_sampling, int norm_on_bbox,
                                 float loss_weight, int n_labeled, int64_t* labels, float* bbox_targets, float* weights,
                                 float* ctr_targets, double* counts, void* stream) {
  LossParams P;
  int rc = fill_loss_params(P, levels, nlevels, B, num_classes);
  if (rc != DSLB_OK) return rc;
  DSLB_CHECK_ARG(gt_off && labels && bbox_targets && weights && ctr_targets && counts, "dslb_fcos_targets: null output");
  P.gt_boxes = gt_boxes;
  P.gt_labels = (const long long*)gt_labels;
  P.gt_off = gt_off;
  P.ig_boxes = ig_boxes;
  P.ig_off = ig_off;
  P.center_sampling = center_sampling;
  P.norm_on_bbox = norm_on_bbox;
  P.loss_weight = loss_weight;
  P.n_labeled = n_labeled;
  P.labels = (long long*)labels;
  P.bbox_targets = bbox_targets;
  P.weights = weights;
  P.ctr_targets = ctr_targets;
  P.counts = counts;
  int maxhw = 0;
  for (int i = 0; i < nlevels; ++i) maxhw = levels[i].h * levels[i].w > maxhw ? levels[i].h * levels[i].w : maxhw;
  dim3 grid((maxhw + 255) / 256, B, nlevels);
  fcos_targets_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" int dslb_fcos_norm(const double* counts, float world_size, float* norm, void* stream) {
  DSLB_CHECK_ARG(counts && norm && world_size >= 1.f, "dslb_fcos_norm: bad arguments");
  fcos_norm_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counts, world_size, norm);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" int dslb_fcos_loss(const dslb_fcos_level_t* levels, int nlevels, int B, int num_classes,
                              const int64_t* labels, const float* bbox_targets, const float* weights,
                              const float* ctr_targets, const float* norm, float alpha, float gamma, float loss_weight,
                              int n_labeled, float si_weight, const float* level_scales, double* loss_sums,
                              float* dscale, void* stream) {
  LossParams P;
  int rc = fill_loss_params(P, levels, nlevels, B, num_classes);
  if (rc != DSLB_OK) return rc;
  DSLB_CHECK_ARG(labels && bbox_targets && weights && ctr_targets && norm && loss_sums, "dslb_fcos_loss: null argument");
  DSLB_CHECK_ARG(si_weight == 0.f || (B % 2 == 1 && B >= 3), "dslb_fcos_loss: the scale-invariant loss needs an odd batch");
  for (int i = 0; i < nlevels; ++i) {
    DSLB_CHECK_ARG(levels[i].cls && levels[i].regctr, "dslb_fcos_loss: level %d has null inputs", i);
    DSLB_CHECK_ARG(!levels[i].dcls_bf16 || (levels[i].ld_dcls >= num_classes && levels[i].ld_dcls % 4 == 0),
                   "dslb_fcos_loss: level %d ld_dcls", i);
    DSLB_CHECK_ARG(!levels[i].dregctr_bf16 || (levels[i].ld_dreg >= 8 && levels[i].ld_dreg % 8 == 0),
                   "dslb_fcos_loss: level %d ld_dreg", i);
  }
  P.labels = (long long*)labels;
  P.bbox_targets = (float*)bbox_targets;
  P.weights = (float*)weights;
  P.ctr_targets = (float*)ctr_targets;
  P.norm = norm;
  P.alpha = alpha;
  P.gamma = gamma;
  P.loss_weight = loss_weight;
  P.n_labeled = n_labeled;
  P.si_weight = si_weight;
  P.loss_sums = loss_sums;
  P.dscale = dscale;
  P.level_scales = level_scales;
  const long long total = P.npoints * (num_classes / 4);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  fcos_loss_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(P);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}
