// Standalone loss kernels behind the reference's LOSSES registry keys (the fused FCOSHead.loss kernel in fcos_loss.cu
// is what the training step uses; these serve a config that builds the loss modules on their own):
//   FocalLoss         mmdet/models/losses/focal_loss.py:11-56,59-102   (sigmoid focal loss, gamma / alpha)
//   GIoULoss          mmdet/models/losses/iou_loss.py:85-102,329-366   (1 - GIoU, eps)
//   CrossEntropyLoss  mmdet/models/losses/cross_entropy_loss.py:73-112 (use_sigmoid=True: BCE with logits)
// Each kernel makes ONE pass: element-wise (weighted) loss (optional), its fp64 sum (optional) and the gradient of
// the weighted loss w.r.t. the prediction (optional) — the reduction / avg_factor / loss_weight scaling of
// weight_reduce_loss (losses/utils.py:27-54) is a scalar the host wrapper applies.
#include "common.h"
#include "loss_math.cuh"

namespace dslb {

__device__ __forceinline__ void block_sum_to(double v, double* out) {
  __shared__ double red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && out) atomicAdd(out, v);
  }
}

__global__ void __launch_bounds__(256) focal_loss_kernel(const float* __restrict__ x, const long long* __restrict__ labels,
                                                         const float* __restrict__ weight, long long N, int C,
                                                         float alpha, float gamma, float* __restrict__ loss_elem,
                                                         double* __restrict__ loss_sum, float* __restrict__ dx) {
  const long long total = N * C;
  double acc = 0.0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long n = t / C;
    const int c = (int)(t - n * C);
    const float w = weight ? weight[n] : 1.f;
    float l, g;
    focal_elem(x[t], labels[n] == c, alpha, gamma, l, g);   // label == C (background) matches no column
    l *= w;
    if (loss_elem) loss_elem[t] = l;
    if (dx) dx[t] = g * w;
    acc += (double)l;
  }
  block_sum_to(acc, loss_sum);
}

__global__ void __launch_bounds__(256) giou_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                        const float* __restrict__ weight, long long n, float eps,
                                                        float* __restrict__ loss_elem, double* __restrict__ loss_sum,
                                                        float* __restrict__ dpred) {
  double acc = 0.0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const float4 p4 = reinterpret_cast<const float4*>(pred)[t];
    const float4 t4 = reinterpret_cast<const float4*>(target)[t];
    const float p[4] = {p4.x, p4.y, p4.z, p4.w}, q[4] = {t4.x, t4.y, t4.z, t4.w};
    float g[4];
    const float w = weight ? weight[t] : 1.f;
    const float l = giou_loss_grad(p, q, eps, g) * w;
    if (loss_elem) loss_elem[t] = l;
    if (dpred) reinterpret_cast<float4*>(dpred)[t] = make_float4(g[0] * w, g[1] * w, g[2] * w, g[3] * w);
    acc += (double)l;
  }
  block_sum_to(acc, loss_sum);
}

__global__ void __launch_bounds__(256) bce_logits_kernel(const float* __restrict__ x, const float* __restrict__ tgt,
                                                         const float* __restrict__ weight, long long n,
                                                         float* __restrict__ loss_elem, double* __restrict__ loss_sum,
                                                         float* __restrict__ dx) {
  double acc = 0.0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const float v = x[t], y = tgt[t];
    const float w = weight ? weight[t] : 1.f;
    // F.binary_cross_entropy_with_logits: (1 - y) * x + softplus(-x)
    const float l = ((1.f - y) * v + softplus(-v)) * w;
    if (loss_elem) loss_elem[t] = l;
    if (dx) dx[t] = (1.f / (1.f + expf(-v)) - y) * w;
    acc += (double)l;
  }
  block_sum_to(acc, loss_sum);
}

static int grid_for(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (b > cap) b = cap;
  return b < 1 ? 1 : (int)b;
}

}  // namespace dslb

using namespace dslb;

extern "C" int dslb_sigmoid_focal_loss(const float* logits, const int64_t* labels, const float* weight, long long N,
                                       int C, float alpha, float gamma, float* loss_elem, double* loss_sum,
                                       float* dlogits, void* stream) {
  DSLB_CHECK_ARG(logits && labels && N >= 0 && C >= 1, "dslb_sigmoid_focal_loss: bad arguments");
  if (N == 0) return DSLB_OK;
  focal_loss_kernel<<<grid_for(N * C), 256, 0, (cudaStream_t)stream>>>(logits, (const long long*)labels, weight, N, C,
                                                                        alpha, gamma, loss_elem, loss_sum, dlogits);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" int dslb_giou_loss(const float* pred, const float* target, const float* weight, long long n, float eps,
                              float* loss_elem, double* loss_sum, float* dpred, void* stream) {
  DSLB_CHECK_ARG(pred && target && n >= 0, "dslb_giou_loss: bad arguments");
  DSLB_CHECK_ARG(((uintptr_t)pred % 16) == 0 && ((uintptr_t)target % 16) == 0 && (!dpred || ((uintptr_t)dpred % 16) == 0),
                 "dslb_giou_loss: boxes must be 16-byte aligned (n, 4) fp32");
  if (n == 0) return DSLB_OK;
  giou_loss_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(pred, target, weight, n, eps, loss_elem, loss_sum, dpred);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" int dslb_bce_with_logits(const float* x, const float* target, const float* weight, long long n,
                                    float* loss_elem, double* loss_sum, float* dx, void* stream) {
  DSLB_CHECK_ARG(x && target && n >= 0, "dslb_bce_with_logits: bad arguments");
  if (n == 0) return DSLB_OK;
  bce_logits_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, target, weight, n, loss_elem, loss_sum, dx);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}
