// Flat-buffer parameter kernels (HBM-bound, float4-vectorised): EMA teacher update, gradient norm, clip + momentum
// SGD. Parameters of the detector live in ONE flat fp32 buffer (nn.Parameters are views), so each of these is a
// single launch over ~32 M floats instead of the reference's per-tensor Python loops.
#include "common.h"

namespace dslb {

// T <- S * c_s + T * c_t, each product and the sum rounded separately (matches the reference's fp32 expression
// `student * (1 - keep_rate) + value * keep_rate`, runner/hooks/semi_epoch_based_runner.py:398-404)
__global__ void ema_kernel(float* __restrict__ t, const float* __restrict__ s, long long n4, long long n, float c_s,
                           float c_t) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 a = reinterpret_cast<float4*>(t)[i];
    const float4 b = __ldg(reinterpret_cast<const float4*>(s) + i);
    a.x = __fadd_rn(__fmul_rn(b.x, c_s), __fmul_rn(a.x, c_t));
    a.y = __fadd_rn(__fmul_rn(b.y, c_s), __fmul_rn(a.y, c_t));
    a.z = __fadd_rn(__fmul_rn(b.z, c_s), __fmul_rn(a.z, c_t));
    a.w = __fadd_rn(__fmul_rn(b.w, c_s), __fmul_rn(a.w, c_t));
    reinterpret_cast<float4*>(t)[i] = a;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    t[i] = __fadd_rn(__fmul_rn(s[i], c_s), __fmul_rn(t[i], c_t));
}

__global__ void sqnorm_kernel(const float* __restrict__ g, long long n4, long long n, double* __restrict__ out) {
  __shared__ double sm[8];
  const long long stride = (long long)gridDim.x * blockDim.x;
  float acc = 0.f;
  double dacc = 0;
  int cnt = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(g) + i);
    acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    if (++cnt == 64) {  // flush the fp32 partial into fp64 regularly
      dacc += acc;
      acc = 0.f;
      cnt = 0;
    }
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += g[i] * g[i];
  dacc += acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = dacc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += sm[k];
    atomicAdd(out, s);
  }
}

// coef = min(max_norm / (sqrt(sqnorm) + 1e-6), 1)   (torch.nn.utils.clip_grad_norm_, norm_type 2); max_norm <= 0: 1
__global__ void clip_coef_kernel(const double* __restrict__ sqnorm, float max_norm, float* __restrict__ coef) {
  float c = 1.f;
  if (max_norm > 0.f) {
    const float total = (float)sqrt(*sqnorm);
    c = fminf(max_norm / (total + 1e-6f), 1.f);
  }
  coef[0] = c;
  coef[1] = (float)sqrt(*sqnorm);
}

// torch.optim.SGD (momentum, dampening 0, no nesterov) with the clip coefficient applied to the gradient first:
//   d = g*coef + wd*p;  buf = first ? d : mom*buf + d;  p -= lr*buf
// EMA: the teacher copy of the same parameters follows in the same pass, T <- P_new * c_s + T * c_t with the rounding of
// ema_kernel (one read of the student weights less than SGD followed by the flat EMA).
template <bool EMA>
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, long long n4,
                           long long n, const float* __restrict__ coef, const float* __restrict__ lr_scale, float lr,
                           float mom, float wd, int first, float* __restrict__ t, float c_s, float c_t) {
  const float c = coef ? coef[0] : 1.f;
  if (lr_scale) lr *= lr_scale[0];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 bv = first ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4*>(buf)[i];
    float d;
    d = gv.x * c + wd * pv.x; bv.x = first ? d : mom * bv.x + d; pv.x -= lr * bv.x;
    d = gv.y * c + wd * pv.y; bv.y = first ? d : mom * bv.y + d; pv.y -= lr * bv.y;
    d = gv.z * c + wd * pv.z; bv.z = first ? d : mom * bv.z + d; pv.z -= lr * bv.z;
    d = gv.w * c + wd * pv.w; bv.w = first ? d : mom * bv.w + d; pv.w -= lr * bv.w;
    reinterpret_cast<float4*>(buf)[i] = bv;
    reinterpret_cast<float4*>(p)[i] = pv;
    if (EMA) {
      float4 a = reinterpret_cast<float4*>(t)[i];
      a.x = __fadd_rn(__fmul_rn(pv.x, c_s), __fmul_rn(a.x, c_t));
      a.y = __fadd_rn(__fmul_rn(pv.y, c_s), __fmul_rn(a.y, c_t));
      a.z = __fadd_rn(__fmul_rn(pv.z, c_s), __fmul_rn(a.z, c_t));
      a.w = __fadd_rn(__fmul_rn(pv.w, c_s), __fmul_rn(a.w, c_t));
      reinterpret_cast<float4*>(t)[i] = a;
    }
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = g[i] * c + wd * p[i];
    const float b = first ? d : mom * buf[i] + d;
    buf[i] = b;
    const float pn = p[i] - lr * b;
    p[i] = pn;
    if (EMA) t[i] = __fadd_rn(__fmul_rn(pn, c_s), __fmul_rn(t[i], c_t));
  }
}

static inline int flat_grid(long long n4) {
  long long blocks = (n4 + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace dslb

using namespace dslb;

extern "C" int dslb_ema_update(float* teacher, const float* student, long long n, float c_student, float c_teacher,
                               void* stream) {
  DSLB_CHECK_ARG(teacher && student && n >= 0, "dslb_ema_update: bad arguments");
  DSLB_CHECK_ARG(((uintptr_t)teacher % 16) == 0 && ((uintptr_t)student % 16) == 0, "dslb_ema_update: 16-byte alignment");
  if (n == 0) return DSLB_OK;
  ema_kernel<<<flat_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(teacher, student, n / 4, n, c_student, c_teacher);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" int dslb_sq_norm(const float* g, long long n, double* out, void* stream) {
  DSLB_CHECK_ARG(g && out && ((uintptr_t)g % 16) == 0, "dslb_sq_norm: bad arguments");
  if (n == 0) return DSLB_OK;
  sqnorm_kernel<<<flat_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(g, n / 4, n, out);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" int dslb_clip_coef(const double* sqnorm, float max_norm, float* coef, void* stream) {
  DSLB_CHECK_ARG(sqnorm && coef, "dslb_clip_coef: null argument");
  clip_coef_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sqnorm, max_norm, coef);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" int dslb_sgd_step(float* p, const float* g, float* buf, long long n, const float* coef,
                             const float* lr_scale, float lr, float momentum, float weight_decay, int first_step,
                             void* stream) {
  DSLB_CHECK_ARG(p && g && buf, "dslb_sgd_step: null argument");
  DSLB_CHECK_ARG(((uintptr_t)p % 16) == 0 && ((uintptr_t)g % 16) == 0 && ((uintptr_t)buf % 16) == 0,
                 "dslb_sgd_step: 16-byte alignment");
  if (n == 0) return DSLB_OK;
  sgd_kernel<false><<<flat_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(p, g, buf, n / 4, n, coef, lr_scale, lr, momentum,
                                                                         weight_decay, first_step, nullptr, 0.f, 0.f);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" int dslb_sgd_ema_step(float* p, const float* g, float* buf, long long n, const float* coef,
                                 const float* lr_scale, float lr, float momentum, float weight_decay, int first_step,
                                 float* teacher, float c_student, float c_teacher, void* stream) {
  DSLB_CHECK_ARG(p && g && buf && teacher, "dslb_sgd_ema_step: null argument");
  DSLB_CHECK_ARG(((uintptr_t)p % 16) == 0 && ((uintptr_t)g % 16) == 0 && ((uintptr_t)buf % 16) == 0 &&
                     ((uintptr_t)teacher % 16) == 0,
                 "dslb_sgd_ema_step: 16-byte alignment");
  if (n == 0) return DSLB_OK;
  sgd_kernel<true><<<flat_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(p, g, buf, n / 4, n, coef, lr_scale, lr, momentum,
                                                                        weight_decay, first_step, teacher, c_student,
                                                                        c_teacher);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}
