// RLA_ResNet-specific glue (reference: mmdet/models/backbones/resnet_rla.py): the recurrent-state update between
// bottlenecks (tanh(BatchNorm(h + conv_out(x))), forward and backward, with the 2x2 average pooling of the state at
// stage boundaries) and the gradients of TRAINABLE BatchNorm affine parameters evaluated with frozen running statistics
// (norm_eval=True but requires_grad=True, resnet_rla.py:361-375,389-399), derived from the packed weight gradients of
// the convs the BatchNorm is folded into. All HBM-bound / latency-bound; no tensor-core work here.
#include <new>

#include "common.h"

#include <cuda_bf16.h>

namespace dslb {

static inline int rla_grid(long long items, int block, int waves) {
  long long blocks = (items + block - 1) / block;
  const long long cap = (long long)num_sms() * waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    f[2 * e] = __uint_as_float(w[e] << 16);
    f[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    w[e] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

constexpr int RLA_LD = 64;  // every recurrent-state tensor is stored with 64-channel rows (32 real + 32 zero)
constexpr int RLA_C = 32;

// pre[n][p][q][c] = (pool ? mean of the 2x2 block of h_old : h_old) + y_out, for the 8 channels of octet `oc`
__device__ __forceinline__ void rla_pre(const __nv_bfloat16* __restrict__ h_old, const __nv_bfloat16* __restrict__ y_out,
                                        long long pix, int oc, int Ho, int Wo, int pool, float* pre) {
  float y[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(y_out + pix * RLA_LD + oc * 8)), y);
  if (!pool) {
    float h[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(h_old + pix * RLA_LD + oc * 8)), h);
#pragma unroll
    for (int e = 0; e < 8; ++e) pre[e] = h[e] + y[e];
  } else {
    const int q = (int)(pix % Wo);
    const long long t = pix / Wo;
    const int p = (int)(t % Ho);
    const long long n = t / Ho;
    const int Wi = 2 * Wo, Hi = 2 * Ho;
    const __nv_bfloat16* base = h_old + ((n * Hi + 2 * p) * Wi + 2 * q) * RLA_LD + oc * 8;
    float a[8], b[8], c[8], d[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(base)), a);
    unpack8(__ldg(reinterpret_cast<const uint4*>(base + RLA_LD)), b);
    unpack8(__ldg(reinterpret_cast<const uint4*>(base + (long long)Wi * RLA_LD)), c);
    unpack8(__ldg(reinterpret_cast<const uint4*>(base + (long long)Wi * RLA_LD + RLA_LD)), d);
#pragma unroll
    for (int e = 0; e < 8; ++e) pre[e] = (a[e] + b[e] + c[e] + d[e]) * 0.25f + y[e];
  }
}

__global__ void __launch_bounds__(256) rla_state_fwd_kernel(
    const __nv_bfloat16* __restrict__ h_old, const __nv_bfloat16* __restrict__ y_out, const float* __restrict__ gamma,
    const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ var, float eps,
    __nv_bfloat16* __restrict__ hb, long long npix, int Ho, int Wo, int pool) {
  const int oc = threadIdx.x & 3;
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = oc * 8 + e;
    sc[e] = gamma[c] * rsqrtf(var[c] + eps);
    sh[e] = beta[c] - mean[c] * sc[e];
  }
  for (long long pix = (long long)blockIdx.x * 64 + (threadIdx.x >> 2); pix < npix; pix += (long long)gridDim.x * 64) {
    float pre[8], o[8];
    rla_pre(h_old, y_out, pix, oc, Ho, Wo, pool, pre);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = tanhf(pre[e] * sc[e] + sh[e]);
    *reinterpret_cast<uint4*>(hb + pix * RLA_LD + oc * 8) = pack8(o);
  }
}

// g = d_hb * (1 - hb^2) (tanh backward, from the stored bf16 output); dbeta += sum g; dgamma += sum g * xhat with
// xhat = (pre - mean) * rstd; d_pre = g * gamma * rstd. With pooling the state gradient 0.25 * d_pre is spread over the
// 2x2 source block of dh_old.
__global__ void __launch_bounds__(256) rla_state_bwd_kernel(
    const __nv_bfloat16* __restrict__ d_hb, const __nv_bfloat16* __restrict__ hb, const __nv_bfloat16* __restrict__ h_old,
    const __nv_bfloat16* __restrict__ y_out, const float* __restrict__ gamma, const float* __restrict__ mean,
    const float* __restrict__ var, float eps, __nv_bfloat16* __restrict__ d_pre, __nv_bfloat16* __restrict__ dh_old,
    float* __restrict__ dgamma, float* __restrict__ dbeta, long long npix, int Ho, int Wo, int pool) {
  __shared__ float red[64][4][17];
  const int oc = threadIdx.x & 3, pl = threadIdx.x >> 2;
  float rstd[8], mu[8], gs[8], sg[8], sb[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = oc * 8 + e;
    rstd[e] = rsqrtf(var[c] + eps);
    mu[e] = mean[c];
    gs[e] = gamma[c] * rstd[e];
    sg[e] = 0.f;
    sb[e] = 0.f;
  }
  for (long long pix = (long long)blockIdx.x * 64 + pl; pix < npix; pix += (long long)gridDim.x * 64) {
    float pre[8], t[8], dh[8], o[8];
    rla_pre(h_old, y_out, pix, oc, Ho, Wo, pool, pre);
    unpack8(__ldg(reinterpret_cast<const uint4*>(hb + pix * RLA_LD + oc * 8)), t);
    unpack8(__ldg(reinterpret_cast<const uint4*>(d_hb + pix * RLA_LD + oc * 8)), dh);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float g = dh[e] * (1.f - t[e] * t[e]);
      sb[e] += g;
      sg[e] += g * (pre[e] - mu[e]) * rstd[e];
      o[e] = g * gs[e];
    }
    *reinterpret_cast<uint4*>(d_pre + pix * RLA_LD + oc * 8) = pack8(o);
    if (pool) {
      const int q = (int)(pix % Wo);
      const long long tt = pix / Wo;
      const int p = (int)(tt % Ho);
      const long long n = tt / Ho;
      const int Wi = 2 * Wo, Hi = 2 * Ho;
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] *= 0.25f;
      const uint4 v = pack8(o);
      __nv_bfloat16* base = dh_old + ((n * Hi + 2 * p) * Wi + 2 * q) * RLA_LD + oc * 8;
      *reinterpret_cast<uint4*>(base) = v;
      *reinterpret_cast<uint4*>(base + RLA_LD) = v;
      *reinterpret_cast<uint4*>(base + (long long)Wi * RLA_LD) = v;
      *reinterpret_cast<uint4*>(base + (long long)Wi * RLA_LD + RLA_LD) = v;
    }
  }
  if (dgamma == nullptr) return;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    red[pl][oc][e] = sg[e];
    red[pl][oc][8 + e] = sb[e];
  }
  __syncthreads();
  if (threadIdx.x < 64) {   // thread = (channel 0..31, which sum)
    const int c = threadIdx.x & 31, which = threadIdx.x >> 5;
    float a = 0.f;
    for (int k = 0; k < 64; ++k) a += red[k][c >> 3][which * 8 + (c & 7)];
    atomicAdd((which ? dbeta : dgamma) + c, a);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Trainable BatchNorm affine parameters under frozen statistics. The conv runs with W' = W * gamma / sigma and the
// epilogue adds beta - mu * gamma / sigma, so with dW' (the packed weight gradient of the folded conv) and
// dbeta[c] = sum_pix dy[c] (column sums, already accumulated by dslb_colsum):
//     sum_pix dy[c] * conv(x, W)[c] = <dW'[c], W[c]>     =>     dgamma[c] = (<dW'[c], W[c]> - mu[c] * dbeta[c]) / sigma[c]
// A conv whose input is a concatenation contributes one (dW', W) piece per part.
struct BnGradDev {
  const float* dw[2];
  const float* w[2];
  int I[2], dw_ld[2], w_ld[2], rows[2];
  const float* mean;
  const float* var;
  const float* dbeta;
  float* dgamma;
  int O, RS;
  float eps;
  int work_begin;  // prefix sum of O
};

__global__ void __launch_bounds__(256) bn_affine_grads_kernel(const BnGradDev* __restrict__ D, int n, int total) {
  const int lane = threadIdx.x & 31;
  for (int item = blockIdx.x * 8 + (threadIdx.x >> 5); item < total; item += gridDim.x * 8) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (D[mid].work_begin <= item) lo = mid; else hi = mid - 1;
    }
    const BnGradDev& d = D[lo];
    const int o = item - d.work_begin;
    float acc = 0.f;
#pragma unroll
    for (int pc = 0; pc < 2; ++pc) {
      if (d.dw[pc] == nullptr) continue;
      const float* __restrict__ dw = d.dw[pc] + (long long)o * d.dw_ld[pc];
      const float* __restrict__ w = d.w[pc] + (long long)o * d.w_ld[pc] * d.RS;
      const long long plane = (long long)d.rows[pc] * d.dw_ld[pc];
      for (int tap = 0; tap < d.RS; ++tap)
        for (int i = lane; i < d.I[pc]; i += 32) acc += __ldg(dw + tap * plane + i) * __ldg(w + (long long)i * d.RS + tap);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) d.dgamma[o] = (acc - d.mean[o] * d.dbeta[o]) * rsqrtf(d.var[o] + d.eps);
  }
}

}  // namespace dslb

using namespace dslb;

struct dslb_bn_grad_plan {
  void* dev = nullptr;
  int n = 0;
  int total = 0;
};

extern "C" int dslb_rla_state_fwd(const void* h_old, const void* y_out, const float* bn_gamma, const float* bn_beta,
                                  const float* bn_mean, const float* bn_var, float eps, void* hb, int N, int Ho, int Wo,
                                  int pool, void* stream) {
  DSLB_CHECK_ARG(h_old && y_out && bn_gamma && bn_beta && bn_mean && bn_var && hb && N > 0 && Ho > 0 && Wo > 0,
                 "dslb_rla_state_fwd: bad arguments");
  const long long npix = (long long)N * Ho * Wo;
  rla_state_fwd_kernel<<<rla_grid(npix, 64, 8), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)h_old, (const __nv_bfloat16*)y_out, bn_gamma, bn_beta, bn_mean, bn_var, eps,
      (__nv_bfloat16*)hb, npix, Ho, Wo, pool ? 1 : 0);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" int dslb_rla_state_bwd(const void* d_hb, const void* hb, const void* h_old, const void* y_out,
                                  const float* bn_gamma, const float* bn_mean, const float* bn_var, float eps,
                                  void* d_pre, void* dh_old, float* dgamma, float* dbeta, int N, int Ho, int Wo, int pool,
                                  void* stream) {
  DSLB_CHECK_ARG(d_hb && hb && h_old && y_out && bn_gamma && bn_mean && bn_var && d_pre && N > 0 && Ho > 0 && Wo > 0,
                 "dslb_rla_state_bwd: bad arguments");
  DSLB_CHECK_ARG(!pool || dh_old, "dslb_rla_state_bwd: pooling needs dh_old");
  DSLB_CHECK_ARG((dgamma == nullptr) == (dbeta == nullptr), "dslb_rla_state_bwd: dgamma and dbeta go together");
  const long long npix = (long long)N * Ho * Wo;
  rla_state_bwd_kernel<<<rla_grid(npix, 64, 4), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)d_hb, (const __nv_bfloat16*)hb, (const __nv_bfloat16*)h_old, (const __nv_bfloat16*)y_out,
      bn_gamma, bn_mean, bn_var, eps, (__nv_bfloat16*)d_pre, (__nv_bfloat16*)dh_old, dgamma, dbeta, npix, Ho, Wo,
      pool ? 1 : 0);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" int dslb_bn_grad_plan_create(const dslb_bn_grad_desc_t* descs, int n, dslb_bn_grad_plan_t** out) {
  DSLB_CHECK_ARG(descs && out && n >= 1, "dslb_bn_grad_plan_create: bad arguments");
  BnGradDev* h = new (std::nothrow) BnGradDev[n];
  if (!h) {
    set_error("out of host memory");
    return DSLB_ENOMEM;
  }
  long long total = 0;
  for (int k = 0; k < n; ++k) {
    const dslb_bn_grad_desc_t& s = descs[k];
    BnGradDev& d = h[k];
    bool ok = s.dw0 && s.w0 && s.I0 > 0 && s.rows0 >= s.O && s.mean && s.var && s.dbeta && s.dgamma && s.O > 0 &&
              s.R > 0 && s.S > 0 && (s.dw1 == nullptr || (s.w1 && s.I1 > 0 && s.rows1 >= s.O));
    if (!ok) {
      delete[] h;
      set_error("dslb_bn_grad_plan_create: descriptor %d is invalid", k);
      return DSLB_EINVAL;
    }
    d.dw[0] = s.dw0; d.w[0] = s.w0; d.I[0] = s.I0; d.rows[0] = s.rows0;
    d.dw_ld[0] = s.dw_ld0 > 0 ? s.dw_ld0 : s.I0;
    d.w_ld[0] = s.w_ld0 > 0 ? s.w_ld0 : s.I0;
    d.dw[1] = s.dw1; d.w[1] = s.w1; d.I[1] = s.I1; d.rows[1] = s.rows1;
    d.dw_ld[1] = s.dw_ld1 > 0 ? s.dw_ld1 : s.I1;
    d.w_ld[1] = s.w_ld1 > 0 ? s.w_ld1 : s.I1;
    d.mean = s.mean; d.var = s.var; d.dbeta = s.dbeta; d.dgamma = s.dgamma;
    d.O = s.O; d.RS = s.R * s.S; d.eps = s.bn_eps;
    d.work_begin = (int)total;
    total += s.O;
  }
  dslb_bn_grad_plan* p = new (std::nothrow) dslb_bn_grad_plan();
  cudaError_t e = p ? cudaSuccess : cudaErrorMemoryAllocation;
  if (e == cudaSuccess) e = cudaMalloc(&p->dev, sizeof(BnGradDev) * n);
  if (e == cudaSuccess) e = cudaMemcpy(p->dev, h, sizeof(BnGradDev) * n, cudaMemcpyHostToDevice);
  delete[] h;
  if (e != cudaSuccess || total > 0x7fffffffLL) {
    set_error("dslb_bn_grad_plan_create: %s", e != cudaSuccess ? cudaGetErrorString(e) : "too many channels");
    if (p && p->dev) cudaFree(p->dev);
    delete p;
    return DSLB_ECUDA;
  }
  p->n = n;
  p->total = (int)total;
  *out = p;
  return DSLB_OK;
}

extern "C" int dslb_bn_grad_plan_run(const dslb_bn_grad_plan_t* p, void* stream) {
  DSLB_CHECK_ARG(p && p->dev, "dslb_bn_grad_plan_run: null plan");
  bn_affine_grads_kernel<<<rla_grid(p->total, 8, 8), 256, 0, (cudaStream_t)stream>>>((const BnGradDev*)p->dev, p->n,
                                                                                     p->total);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" void dslb_bn_grad_plan_destroy(dslb_bn_grad_plan_t* p) {
  if (!p) return;
  if (p->dev) cudaFree(p->dev);
  delete p;
}
