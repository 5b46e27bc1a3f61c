// Per-element loss arithmetic shared by the fused FCOSHead loss kernel (fcos_loss.cu) and the standalone LOSSES
// kernels (losses.cu): sigmoid focal loss, 1 - GIoU with its gradient, numerically stable softplus.
#pragma once
#include <cuda_runtime.h>

namespace dslb {

__device__ __forceinline__ float softplus(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

// focal loss of one (logit, one-hot target) and its derivative w.r.t. the logit (losses/focal_loss.py:11-56).
__device__ __forceinline__ void focal_elem(float x, bool is_target, float alpha, float gamma, float& loss, float& grad) {
  const float p = 1.f / (1.f + expf(-x));
  if (is_target) {
    const float q = 1.f - p;               // pt
    const float bce = softplus(-x);        // -log p
    const float qg = (gamma == 2.f) ? q * q : powf(q, gamma);
    loss = alpha * qg * bce;
    // d/dx [ a * q^g * bce ] = a * ( -g q^(g-1) * p q * bce + q^g * (-(1-p)) ),  dq/dx = -p q, dbce/dx = -(1-p)
    const float qgm1 = (gamma == 2.f) ? q : powf(q, gamma - 1.f);
    grad = alpha * (-gamma * qgm1 * p * q * bce - qg * q);
  } else {
    const float bce = softplus(x);         // -log(1-p)
    const float pg = (gamma == 2.f) ? p * p : powf(p, gamma);
    loss = (1.f - alpha) * pg * bce;
    const float pgm1 = (gamma == 2.f) ? p : powf(p, gamma - 1.f);
    grad = (1.f - alpha) * (gamma * pgm1 * p * (1.f - p) * bce + pg * p);
  }
}

// 1 - GIoU of (pred, target) boxes and its gradient w.r.t. the pred box corners
// (core/bbox/iou_calculators/iou2d_calculator.py:214-260 with eps = 1e-6, losses/iou_loss.py:85-102).
__device__ __forceinline__ float giou_loss_grad(const float (&p)[4], const float (&t)[4], float eps, float (&g)[4]) {
  const float pw = p[2] - p[0], ph = p[3] - p[1];
  const float a1 = pw * ph, a2 = (t[2] - t[0]) * (t[3] - t[1]);
  const float ltx = fmaxf(p[0], t[0]), lty = fmaxf(p[1], t[1]);
  const float rbx = fminf(p[2], t[2]), rby = fminf(p[3], t[3]);
  const float iw_raw = rbx - ltx, ih_raw = rby - lty;
  const float iw = fmaxf(iw_raw, 0.f), ih = fmaxf(ih_raw, 0.f);
  const float overlap = iw * ih;
  const float union_raw = a1 + a2 - overlap;
  const float uni = fmaxf(union_raw, eps);
  const float iou = overlap / uni;
  const float elx = fminf(p[0], t[0]), ely = fminf(p[1], t[1]);
  const float erx = fmaxf(p[2], t[2]), ery = fmaxf(p[3], t[3]);
  const float ew_raw = erx - elx, eh_raw = ery - ely;
  const float ew = fmaxf(ew_raw, 0.f), eh = fmaxf(eh_raw, 0.f);
  const float earea_raw = ew * eh;
  const float earea = fmaxf(earea_raw, eps);
  const float giou = iou - (earea - uni) / earea;
  // ---- backward. Selection weights of max/min: 1 to the larger (smaller), 0.5 each on ties (torch.max/min).
  auto wmax = [](float a, float b) { return a > b ? 1.f : (a == b ? 0.5f : 0.f); };  // d max(a,b) / d a
  auto wmin = [](float a, float b) { return a < b ? 1.f : (a == b ? 0.5f : 0.f); };  // d min(a,b) / d a
  const float ciw = iw_raw >= 0.f ? 1.f : 0.f, cih = ih_raw >= 0.f ? 1.f : 0.f;     // clamp(min=0) passes x >= 0
  const float cew = ew_raw >= 0.f ? 1.f : 0.f, ceh = eh_raw >= 0.f ? 1.f : 0.f;
  // d iw / d(p0,p2), d ih / d(p1,p3)
  const float diw[4] = {-ciw * wmax(p[0], t[0]), 0.f, ciw * wmin(p[2], t[2]), 0.f};
  const float dih[4] = {0.f, -cih * wmax(p[1], t[1]), 0.f, cih * wmin(p[3], t[3])};
  const float dew[4] = {-cew * wmin(p[0], t[0]), 0.f, cew * wmax(p[2], t[2]), 0.f};
  const float deh[4] = {0.f, -ceh * wmin(p[1], t[1]), 0.f, ceh * wmax(p[3], t[3])};
  const float da1[4] = {-ph, -pw, ph, pw};
  const float cu = union_raw > eps ? 1.f : (union_raw == eps ? 0.5f : 0.f);
  const float ce = earea_raw > eps ? 1.f : (earea_raw == eps ? 0.5f : 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float dov = diw[k] * ih + iw * dih[k];
    const float dun = cu * (da1[k] - dov);
    const float dea = ce * (dew[k] * eh + ew * deh[k]);
    const float diou = (dov * uni - overlap * dun) / (uni * uni);
    // giou = iou - 1 + uni/earea
    const float dg = diou + (dun * earea - uni * dea) / (earea * earea);
    g[k] = -dg;  // loss = 1 - giou
  }
  return 1.f - giou;
}

}  // namespace dslb
