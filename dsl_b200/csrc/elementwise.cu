// HBM-bound SIMT kernels around the tensor-core convs: layout conversion, stem im2col, max-pool, FPN top-down
// add, ReLU masks, GroupNorm apply / backward, weight packing, frozen-BN folding. All are 16-byte vectorised over
// the NHWC channel dimension (8 bf16 per thread per access) and grid-stride over a multiple of the SM count.
#include "common.h"

#include <cuda_bf16.h>

namespace dslb {

static inline int grid_for(long long work_items, int block, int max_waves = 8) {
  long long blocks = (work_items + block - 1) / block;
  long long cap = (long long)num_sms() * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

__device__ __forceinline__ float bflo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bfhi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t packbf(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bflo(u.x); f[1] = bfhi(u.x); f[2] = bflo(u.y); f[3] = bfhi(u.y);
  f[4] = bflo(u.z); f[5] = bfhi(u.z); f[6] = bflo(u.w); f[7] = bfhi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = packbf(f[0], f[1]); u.y = packbf(f[2], f[3]); u.z = packbf(f[4], f[5]); u.w = packbf(f[6], f[7]);
  return u;
}

// ------------------------------------------------------------------------------------------------ layout
// NCHW fp32 -> NHWC bf16 (channels padded with zeros to Cpad, Cpad % 8 == 0). Tiled through shared memory so both
// sides are coalesced: block = 32 pixels x all channels of one (n, row segment).
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int C, int HW,
                                    int Cpad) {
  // one thread per (pixel, channel) pair, pixel fastest for reads; transposed via smem tile [32 ch][33 px]
  __shared__ float tile[32][33];
  const int npt = (HW + 31) / 32, nct = (Cpad + 31) / 32;
  const long long total = (long long)N * npt * nct;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const int ct = t % nct;
    const int pt = (t / nct) % npt;
    const int n = t / ((long long)nct * npt);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int k = ty; k < 32; k += 8) {
      const int c = ct * 32 + k, p = pt * 32 + tx;
      tile[k][tx] = (c < C && p < HW) ? x[((long long)n * C + c) * HW + p] : 0.f;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
      const int p = pt * 32 + k, c = ct * 32 + tx;
      if (p < HW && c < Cpad) y[((long long)n * HW + p) * Cpad + c] = __float2bfloat16_rn(tile[tx][k]);
    }
    __syncthreads();
  }
}

// pixel-major (NHWC) rows of `ld` elements (bf16 or fp32) -> NCHW fp32, first C channels
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, float* __restrict__ y, int N, int C, int HW, int ld) {
  __shared__ float tile[32][33];
  const int npt = (HW + 31) / 32, nct = (C + 31) / 32;
  const long long total = (long long)N * npt * nct;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const int ct = t % nct;
    const int pt = (t / nct) % npt;
    const int n = t / ((long long)nct * npt);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = ty; k < 32; k += 8) {
      const int p = pt * 32 + k, c = ct * 32 + tx;
      tile[k][tx] = (p < HW && c < C) ? (float)x[((long long)n * HW + p) * ld + c] : 0.f;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
      const int c = ct * 32 + k, p = pt * 32 + tx;
      if (c < C && p < HW) y[((long long)n * C + c) * HW + p] = tile[tx][k];
    }
    __syncthreads();
  }
}

// 3x3 stride-2 pad-1 max-pool, NHWC bf16. One thread = 8 channels of TWO horizontally adjacent output pixels: their
// windows share a column, so 15 loads (3 rows x 5 columns) make two outputs instead of 18, and the row-wise maxima of the
// five columns are formed once. Padding taps read -inf; max is exact in bf16, so it runs on packed bf16x2 values.
__device__ __forceinline__ uint4 max8(const uint4& a, const uint4& b) {
  uint4 m = a;
  __nv_bfloat162* x = reinterpret_cast<__nv_bfloat162*>(&m);
  const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
  for (int e = 0; e < 4; ++e) x[e] = __hmax2(x[e], y[e]);
  return m;
}

__global__ void maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int H,
                                    int W, int C, int Ho, int Wo) {
  const int cv = C / 8;
  const int Wp = (Wo + 1) >> 1;
  const long long total = (long long)N * Ho * Wp * cv;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c8 = t % cv;
    const long long pp = t / cv;
    const int qp = pp % Wp;
    const int p = (pp / Wp) % Ho;
    const int n = pp / ((long long)Wp * Ho);
    const int q = qp * 2;
    const uint4 ninf = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);
    uint4 u[15];   // all taps are loaded first: independent requests in flight
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = p * 2 - 1 + r;
#pragma unroll
      for (int s = 0; s < 5; ++s) {
        const int w = q * 2 - 1 + s;
        const bool ok = h >= 0 && h < H && w >= 0 && w < W;
        u[r * 5 + s] = ok ? __ldg(reinterpret_cast<const uint4*>(x + (((long long)n * H + h) * W + w) * C + c8 * 8)) : ninf;
      }
    }
    uint4 cm[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) cm[s] = max8(max8(u[s], u[5 + s]), u[10 + s]);
    __nv_bfloat16* dst = y + (((long long)n * Ho + p) * Wo + q) * C + c8 * 8;
    *reinterpret_cast<uint4*>(dst) = max8(max8(cm[0], cm[1]), cm[2]);
    if (q + 1 < Wo) *reinterpret_cast<uint4*>(dst + C) = max8(max8(cm[2], cm[3]), cm[4]);
  }
}

// ------------------------------------------------------------------------------------------------ FPN / ReLU
// dst[n,y,x,:] += src[n, (y*h)/H, (x*w)/W, :]   (F.interpolate nearest to dst's size, necks/fpn.py:163-172)
__global__ void upsample_add_kernel(__nv_bfloat16* __restrict__ dst, const __nv_bfloat16* __restrict__ src, int N, int H,
                                    int W, int h, int w, int C) {
  const int cv = C / 8;
  const long long total = (long long)N * H * W * cv;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c8 = t % cv;
    const long long pix = t / cv;
    const int x = pix % W;
    const int y = (pix / W) % H;
    const int n = pix / ((long long)W * H);
    const int ys = min((y * h) / H, h - 1), xs = min((x * w) / W, w - 1);
    float a[8], b[8];
    unpack8(*reinterpret_cast<const uint4*>(dst + pix * C + c8 * 8), a);
    unpack8(*reinterpret_cast<const uint4*>(src + (((long long)n * h + ys) * w + xs) * C + c8 * 8), b);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] += b[e];
    *reinterpret_cast<uint4*>(dst + pix * C + c8 * 8) = pack8(a);
  }
}

// backward of the above w.r.t. src: dsrc[n,ys,xs,:] += sum over the dst pixels that read it
__global__ void upsample_add_bwd_kernel(__nv_bfloat16* __restrict__ dsrc, const __nv_bfloat16* __restrict__ ddst, int N,
                                        int H, int W, int h, int w, int C) {
  const int cv = C / 8;
  const long long total = (long long)N * h * w * cv;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c8 = t % cv;
    const long long pix = t / cv;
    const int xs = pix % w;
    const int ys = (pix / w) % h;
    const int n = pix / ((long long)w * h);
    // dst rows y with (y*h)/H == ys  <=>  y in [ceil(ys*H/h), ceil((ys+1)*H/h) - 1]
    const int y0 = (ys * H + h - 1) / h, y1 = min(((ys + 1) * H + h - 1) / h, H);
    const int x0 = (xs * W + w - 1) / w, x1 = min(((xs + 1) * W + w - 1) / w, W);
    float a[8];
    unpack8(*reinterpret_cast<const uint4*>(dsrc + pix * C + c8 * 8), a);
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) {
        float b[8];
        unpack8(*reinterpret_cast<const uint4*>(ddst + (((long long)n * H + y) * W + x) * C + c8 * 8), b);
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] += b[e];
      }
    *reinterpret_cast<uint4*>(dsrc + pix * C + c8 * 8) = pack8(a);
  }
}

// mode 0: y = relu(x);  mode 1: y = (m > 0) ? x : 0;  mode 2: y = x + m
__global__ void relu_family_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ m,
                                   __nv_bfloat16* __restrict__ y, long long n8, int mode) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n8; t += (long long)gridDim.x * blockDim.x) {
    float a[8], b[8];
    unpack8(reinterpret_cast<const uint4*>(x)[t], a);
    if (mode != 0) unpack8(reinterpret_cast<const uint4*>(m)[t], b);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (mode == 0) a[e] = fmaxf(a[e], 0.f);
      else if (mode == 1) a[e] = (b[e] > 0.f) ? a[e] : 0.f;
      else a[e] += b[e];
    }
    reinterpret_cast<uint4*>(y)[t] = pack8(a);
  }
}

// ------------------------------------------------------------------------------------------------ GroupNorm
struct GnSeg {
  const __nv_bfloat16* x;   // pre-norm conv output [N*HW][C]
  __nv_bfloat16* y;         // fwd: normalised+ReLU output; bwd: gradient w.r.t. x
  const __nv_bfloat16* dz;  // bwd: gradient w.r.t. the post-ReLU output
  const double* stats;      // [N][G][DSLB_GN_STAT_STRIDE]
  const float* gamma;
  const float* beta;
  double* red;              // bwd: [N][C][2] (sum dy, sum dy*xhat), fp64
  float* dbias;             // bwd: [C] += sum dx
  float4* mr;               // [N][G] (mean, rstd, k1, k2) in fp32: .xy written by the forward apply, .zw by the backward
  const double* gsums;      // bwd: [N][G][DSLB_GN_STAT_STRIDE] group sums from the producing conv's epilogue, or null
  int HW, npix;             // pixels per image, N*HW
  long long work_begin;     // prefix of (npix * C/8) work items
};
struct GnParams {
  GnSeg seg[DSLB_MAX_SEGS];
  int nseg, C, cpg;
  float eps;
  long long total;
};

__device__ __forceinline__ int gn_find(const GnParams& P, long long t) {
  int si = 0;
  while (si + 1 < P.nseg && t >= P.seg[si + 1].work_begin) ++si;
  return si;
}

// (sum, sumsq) in fp64 -> (mean, rstd) in fp32, once per (map, image, group): keeps the slow fp64 pipe out of the
// per-element kernels.
__global__ void gn_finalize_kernel(const __grid_constant__ GnParams P, int maxN) {
  const int G = P.C / P.cpg;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.nseg * maxN * G) return;
  const int g = t % G;
  const int n = (t / G) % maxN;
  const GnSeg& s = P.seg[t / (G * maxN)];
  if (n >= s.npix / s.HW) return;
  const double* st = s.stats + ((long long)n * G + g) * DSLB_GN_STAT_STRIDE;
  const double m = (double)P.cpg * s.HW;
  const double mean = st[0] / m;
  double var = st[1] / m - mean * mean;
  if (var < 0) var = 0;
  float4* o = s.mr + n * G + g;
  o->x = (float)mean;
  o->y = (float)(1.0 / sqrt(var + (double)P.eps));
}

// y = relu((x - mean) * rstd * gamma + beta)   (mmcv ConvModule order conv -> GN -> ReLU, eps 1e-5)
__global__ void gn_apply_relu_kernel(const __grid_constant__ GnParams P) {
  const int cv = P.C / 8;
  const int G = P.C / P.cpg;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < P.total; t += (long long)gridDim.x * blockDim.x) {
    const GnSeg& s = P.seg[gn_find(P, t)];
    const long long tl = t - s.work_begin;
    const int c8 = tl % cv;
    const long long pix = tl / cv;
    const int n = pix / s.HW;
    const int g = (c8 * 8) / P.cpg;
    const float4 mr = __ldg(s.mr + n * G + g);
    const float rstd = mr.y, fmean = mr.x;
    float a[8];
    unpack8(*reinterpret_cast<const uint4*>(s.x + pix * P.C + c8 * 8), a);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(s.gamma + c8 * 8));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(s.gamma + c8 * 8) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(s.beta + c8 * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(s.beta + c8 * 8) + 1);
    const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = fmaxf((a[e] - fmean) * rstd * ga[e] + be[e], 0.f);
    *reinterpret_cast<uint4*>(s.y + pix * P.C + c8 * 8) = pack8(a);
  }
}

// (sum, sumsq) of one (image, group) in fp64 -> (mean, rstd) in fp32, the arithmetic of gn_finalize_kernel
__device__ __forceinline__ float2 gn_mean_rstd(const double* __restrict__ st, double m, float eps) {
  const double mean = st[0] / m;
  double var = st[1] / m - mean * mean;
  if (var < 0) var = 0;
  return make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
}

// Same, driven by the block table of the backward (block = GN_PIXB pixels of ONE image of ONE map; thread = one
// 8-channel group x one of 8 pixel lanes): per-thread constants (mean, rstd, gamma, beta) are computed once and the loop
// body is two 16-byte accesses and 16 FMAs — no per-element search or integer division. The (mean, rstd) of the
// thread's group come straight from the fp64 sums of the conv epilogue (no separate finalize launch); the first block of
// every image leaves them in `mr` for the backward.
__global__ void __launch_bounds__(256) gn_apply_relu_tab_kernel(const __grid_constant__ GnParams P,
                                                                const int* __restrict__ blk_seg,
                                                                const int* __restrict__ blk_pix0) {
  const GnSeg& s = P.seg[blk_seg[blockIdx.x]];
  const int pix0 = blk_pix0[blockIdx.x];
  const int n = pix0 / s.HW;
  const int pend = min(pix0 + 256, (n + 1) * s.HW);
  const int c8 = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int G = P.C / P.cpg;
  const int g = (c8 * 8) / P.cpg;
  const float2 mr = gn_mean_rstd(s.stats + ((long long)n * G + g) * DSLB_GN_STAT_STRIDE, (double)P.cpg * s.HW, P.eps);
  if (pix0 == n * s.HW && pl == 0 && (c8 * 8) % P.cpg == 0) {
    float4* o = s.mr + n * G + g;
    o->x = mr.x;
    o->y = mr.y;
  }
  float a_[8], b_[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float ga = __ldg(s.gamma + c8 * 8 + e) * mr.y;
    a_[e] = ga;
    b_[e] = __ldg(s.beta + c8 * 8 + e) - mr.x * ga;
  }
  const uint4* x = reinterpret_cast<const uint4*>(s.x) + c8;
  uint4* y = reinterpret_cast<uint4*>(s.y) + c8;
  for (int p = pix0 + pl; p < pend; p += 8) {
    float v[8];
    unpack8(x[(long long)p * 32], v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = fmaxf(fmaf(v[e], a_[e], b_[e]), 0.f);
    y[(long long)p * 32] = pack8(v);
  }
}

// Backward pass 1: per (image, channel) sums of dy and dy*xhat, dy = dz * [relu input > 0].
// Block = 256 threads = 32 channel-octets x 8 pixel lanes (C == 256), PIXB pixels of one image per block.
constexpr int GN_PIXB = 256;
__global__ void gn_bwd_reduce_kernel(const __grid_constant__ GnParams P, const int* __restrict__ blk_seg,
                                     const int* __restrict__ blk_pix0) {
  __shared__ float sm[8][256][2];
  const GnSeg& s = P.seg[blk_seg[blockIdx.x]];
  const int pix0 = blk_pix0[blockIdx.x];
  const int n = pix0 / s.HW;
  const int pend = min(pix0 + GN_PIXB, (n + 1) * s.HW);
  const int c8 = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int G = P.C / P.cpg;
  const int g = (c8 * 8) / P.cpg;
  const float4 mr = __ldg(s.mr + n * G + g);
  const float rstd = mr.y, fmean = mr.x;
  float ga[8], be[8], A[8], B[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    ga[e] = __ldg(s.gamma + c8 * 8 + e);
    be[e] = __ldg(s.beta + c8 * 8 + e);
    A[e] = 0.f;
    B[e] = 0.f;
  }
#pragma unroll 4
  for (int p = pix0 + pl; p < pend; p += 8) {   // unrolled: 8 independent 16-byte loads in flight per thread
    float x[8], d[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(s.x + (long long)p * P.C + c8 * 8)), x);
    unpack8(__ldg(reinterpret_cast<const uint4*>(s.dz + (long long)p * P.C + c8 * 8)), d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float xh = (x[e] - fmean) * rstd;
      const float dy = (fmaf(xh, ga[e], be[e]) > 0.f) ? d[e] : 0.f;
      A[e] += dy;
      B[e] += dy * xh;
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sm[pl][c8 * 8 + e][0] = A[e];
    sm[pl][c8 * 8 + e][1] = B[e];
  }
  __syncthreads();
  const int c = threadIdx.x;  // one channel per thread
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    a += sm[k][c][0];
    b += sm[k][c][1];
  }
  double* dst = s.red + ((long long)n * P.C + c) * 2;
  atomicAdd(dst, (double)a);
  atomicAdd(dst + 1, (double)b);
}

// Between the two passes: k1 = S1/m, k2 = S2/m with S1 = sum_{c in g} gamma_c A_c, S2 likewise with B (fp64 -> fp32).
__global__ void gn_bwd_finalize_kernel(const __grid_constant__ GnParams P, int maxN) {
  const int G = P.C / P.cpg;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.nseg * maxN * G) return;
  const int g = t % G;
  const int n = (t / G) % maxN;
  const GnSeg& s = P.seg[t / (G * maxN)];
  if (n >= s.npix / s.HW) return;
  const double m = (double)P.cpg * s.HW;
  double S1 = 0, S2 = 0;
  for (int c = g * P.cpg; c < (g + 1) * P.cpg; ++c) {
    const double gm = (double)__ldg(s.gamma + c);
    S1 += gm * s.red[((long long)n * P.C + c) * 2];
    S2 += gm * s.red[((long long)n * P.C + c) * 2 + 1];
  }
  float4* o = s.mr + n * G + g;
  o->z = (float)(S1 / m);
  o->w = (float)(S2 / m);
}

// Backward pass 2: dx = rstd * (gamma*dy - S1/m - xhat*S2/m), S1 = sum_{c in g} gamma_c A_c, S2 likewise with B;
// also accumulates dbias[c] += sum dx (the conv bias in front of the GroupNorm).
// FUSED: the producing conv's epilogue already left S1, S2 per (image, group) in `gsums` (dslb_conv_seg_t::gnb_sums), so
// there was no reduce pass: the two constants are formed here, and the per-(image, channel) sums A, B that dgamma /
// dbeta need (`red`) are accumulated in this pass too.
template <bool FUSED>
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const __grid_constant__ GnParams P,
                                                           const int* __restrict__ blk_seg,
                                                           const int* __restrict__ blk_pix0) {
  __shared__ float sm[FUSED ? 3 : 1][8][256];
  const GnSeg& s = P.seg[blk_seg[blockIdx.x]];
  const int pix0 = blk_pix0[blockIdx.x];
  const int n = pix0 / s.HW;
  const int pend = min(pix0 + GN_PIXB, (n + 1) * s.HW);
  const int c8 = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int G = P.C / P.cpg;
  const int g = (c8 * 8) / P.cpg;
  const float4 mr = __ldg(s.mr + n * G + g);
  const float rstd = mr.y, fmean = mr.x;
  float k1 = mr.z, k2 = mr.w;
  if (FUSED) {
    const double* gs = s.gsums + ((long long)n * G + g) * DSLB_GN_STAT_STRIDE;
    const double m = (double)P.cpg * s.HW;
    k1 = (float)(gs[0] / m);
    k2 = (float)(gs[1] / m);
  }
  float ga[8], be[8], D[8], A[8], B[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    ga[e] = __ldg(s.gamma + c8 * 8 + e);
    be[e] = __ldg(s.beta + c8 * 8 + e);
    D[e] = 0.f;
    A[e] = 0.f;
    B[e] = 0.f;
  }
#pragma unroll 4
  for (int p = pix0 + pl; p < pend; p += 8) {   // unrolled: 8 independent 16-byte loads in flight per thread
    float x[8], d[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(s.x + (long long)p * P.C + c8 * 8)), x);
    unpack8(__ldg(reinterpret_cast<const uint4*>(s.dz + (long long)p * P.C + c8 * 8)), d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float xh = (x[e] - fmean) * rstd;
      const float dy = (fmaf(xh, ga[e], be[e]) > 0.f) ? d[e] : 0.f;
      const float dx = rstd * (ga[e] * dy - k1 - xh * k2);
      if (FUSED) {
        A[e] += dy;
        B[e] = fmaf(dy, xh, B[e]);
      }
      d[e] = dx;
      D[e] += dx;
    }
    *reinterpret_cast<uint4*>(s.y + (long long)p * P.C + c8 * 8) = pack8(d);
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sm[0][pl][c8 * 8 + e] = D[e];
    if (FUSED) {
      sm[1][pl][c8 * 8 + e] = A[e];
      sm[2][pl][c8 * 8 + e] = B[e];
    }
  }
  __syncthreads();
  const int c = threadIdx.x;
  float a = 0.f, sa = 0.f, sb = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    a += sm[0][k][c];
    if (FUSED) {
      sa += sm[1][k][c];
      sb += sm[2][k][c];
    }
  }
  atomicAdd(s.dbias + c, a);
  if (FUSED) {
    double* dst = s.red + ((long long)n * P.C + c) * 2;
    atomicAdd(dst, (double)sa);
    atomicAdd(dst + 1, (double)sb);
  }
}

// dgamma[c] += sum_n B[n][c], dbeta[c] += sum_n A[n][c]
__global__ void gn_bwd_params_kernel(const double* __restrict__ red, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                     int N, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double a = 0, b = 0;
  for (int n = 0; n < N; ++n) {
    a += red[((long long)n * C + c) * 2];
    b += red[((long long)n * C + c) * 2 + 1];
  }
  dbeta[c] += (float)a;
  dgamma[c] += (float)b;
}

// ------------------------------------------------------------------------------------------------ weights
// OIHW fp32 -> packed bf16 [taps][rows_pad][cols_pad].
//   transpose == 0 (fprop):  out[r*S+s][o][i] = w[o][i][r][s] * oscale[o]
//   transpose == 1 (dgrad):  out[r*S+s][i][o] = w[o][i][R-1-r][S-1-s] * oscale[o]    (180-degree rotated, in/out swapped)
__global__ void pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int O, int I, int R,
                                   int S, int rows_pad, int cols_pad, const float* __restrict__ oscale, int transpose) {
  const long long total = (long long)R * S * rows_pad * cols_pad;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int col = t % cols_pad;
    const int row = (t / cols_pad) % rows_pad;
    const int tap = t / ((long long)cols_pad * rows_pad);
    const int r = tap / S, s = tap - r * S;
    float v = 0.f;
    if (!transpose) {
      if (row < O && col < I) v = w[(((long long)row * I + col) * R + r) * S + s] * (oscale ? oscale[row] : 1.f);
    } else {
      if (row < I && col < O)
        v = w[(((long long)col * I + row) * R + (R - 1 - r)) * S + (S - 1 - s)] * (oscale ? oscale[col] : 1.f);
    }
    out[t] = __float2bfloat16_rn(v);
  }
}

// packed fp32 wgrad [taps][rows][I] -> OIHW fp32 gradient:  g[o][i][r][s] (+)= dw[r*S+s][o][i] * oscale[o]
__global__ void unpack_wgrad_kernel(const float* __restrict__ dw, float* __restrict__ g, int O, int I, int R, int S,
                                    int rows, const float* __restrict__ oscale, int accumulate) {
  const long long total = (long long)O * I * R * S;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int s = t % S;
    const int r = (t / S) % R;
    const int i = (t / ((long long)S * R)) % I;
    const int o = t / ((long long)S * R * I);
    float v = dw[((long long)(r * S + s) * rows + o) * I + i] * (oscale ? oscale[o] : 1.f);
    g[t] = accumulate ? g[t] + v : v;
  }
}

// frozen BatchNorm (eval): scale = gamma / sqrt(var + eps), shift = beta - mean * scale   (resnet.py:647-656)
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, float eps,
                               float* __restrict__ scale, float* __restrict__ shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] / sqrtf(var[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] - mean[c] * sc;
}

// column sums of a pixel-major bf16 matrix: out[c] += sum_p x[p][c]   (conv bias gradients)
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, long long npix, int ld,
                              int C) {
  // block = 256 threads: 32 channel lanes x 8 pixel lanes, loops over channel tiles of 32
  __shared__ float sm[8][33];
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int c = c0 + cl;
    float a = 0.f;
    if (c < C)
      for (long long p = (long long)blockIdx.x * 8 + pl; p < npix; p += (long long)gridDim.x * 8)
        a += __bfloat162float(x[p * ld + c]);
    sm[pl][cl] = a;
    __syncthreads();
    if (pl == 0 && c < C) {
      float tsum = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) tsum += sm[k][cl];
      atomicAdd(out + c, tsum);
    }
    __syncthreads();
  }
}


// Vector variant (C % 8 == 0, ld % 8 == 0, 16-byte aligned rows): one thread = 8 consecutive channels (one 16-byte load
// per pixel), T = C / 8 threads per pixel row, 256 / T rows per block iteration, 4 independent loads in flight per
// thread; per-block shared-memory reduction, then one reduction per channel and block (V4: one 16-byte
// red.global.add.v4.f32 per 4 channels — the few output lines are shared by every block, so the op count matters).
template <bool V4>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out,
                                                         long long npix, int ld, int T, int R) {
  __shared__ float sm[256][9];
  const int t = threadIdx.x % T, r = threadIdx.x / T;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (r < R) {
    const long long step = (long long)gridDim.x * R;
    const __nv_bfloat16* base = x + t * 8;
    long long p = (long long)blockIdx.x * R + r;
    for (; p + 3 * step < npix; p += 4 * step) {
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = __ldg(reinterpret_cast<const uint4*>(base + (p + k * step) * ld));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[2 * e] += bflo(w[e]);
          acc[2 * e + 1] += bfhi(w[e]);
        }
      }
    }
    for (; p < npix; p += step) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + p * ld));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[2 * e] += bflo(w[e]);
        acc[2 * e + 1] += bfhi(w[e]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) sm[threadIdx.x][e] = acc[e];
  __syncthreads();
  if (V4) {
    // thread = 4 consecutive channels of octet tt: sum over the R row lanes, one vector reduction
    for (int idx = threadIdx.x; idx < T * 2; idx += 256) {
      const int tt = idx >> 1, e0 = (idx & 1) * 4;
      float a[4] = {0.f, 0.f, 0.f, 0.f};
      for (int rr = 0; rr < R; ++rr) {
#pragma unroll
        for (int e = 0; e < 4; ++e) a[e] += sm[rr * T + tt][e0 + e];
      }
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + tt * 8 + e0), "f"(a[0]), "f"(a[1]),
                   "f"(a[2]), "f"(a[3])
                   : "memory");
    }
  } else {
    // thread (t, e-th channel) for t < T: sum over the R row lanes
    for (int idx = threadIdx.x; idx < T * 8; idx += 256) {
      const int tt = idx >> 3, e = idx & 7;
      float a = 0.f;
      for (int rr = 0; rr < R; ++rr) a += sm[rr * T + tt][e];
      atomicAdd(out + idx, a);
    }
  }
}

}  // namespace dslb

using namespace dslb;

#define LAUNCH_CHECK()                       \
  do {                                       \
    DSLB_CHECK_CUDA(cudaGetLastError());     \
    return DSLB_OK;                          \
  } while (0)

extern "C" int dslb_nchw_to_nhwc_bf16(const float* x, void* y, int N, int C, int H, int W, int Cpad, void* stream) {
  DSLB_CHECK_ARG(x && y && N > 0 && C > 0 && Cpad >= C, "dslb_nchw_to_nhwc_bf16: bad arguments");
  const long long tiles = (long long)N * ((H * W + 31) / 32) * ((Cpad + 31) / 32);
  nchw_to_nhwc_kernel<<<grid_for(tiles, 1, 16), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)y, N, C, H * W, Cpad);
  LAUNCH_CHECK();
}

extern "C" int dslb_nhwc_to_nchw_f32(const void* x, float* y, int N, int C, int H, int W, int ld, int x_is_fp32,
                                     void* stream) {
  DSLB_CHECK_ARG(x && y && N > 0 && C > 0 && ld >= C, "dslb_nhwc_to_nchw_f32: bad arguments");
  const long long tiles = (long long)N * ((H * W + 31) / 32) * ((C + 31) / 32);
  if (x_is_fp32)
    nhwc_to_nchw_kernel<float><<<grid_for(tiles, 1, 16), 256, 0, (cudaStream_t)stream>>>((const float*)x, y, N, C, H * W, ld);
  else
    nhwc_to_nchw_kernel<__nv_bfloat16>
        <<<grid_for(tiles, 1, 16), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, y, N, C, H * W, ld);
  LAUNCH_CHECK();
}

namespace dslb {
// Scale-invariant extra input (SemiEpochBasedRunner.train, mmdet/runner/hooks/semi_epoch_based_runner.py:186-204):
// out[c] = zeros(H, W) with the bilinear (align_corners=False) resize of img[c] to (H/2, W/2) in its top-left corner.
// Index arithmetic follows torch's upsample_bilinear2d: src = scale * (dst + 0.5) - 0.5 clamped at 0, scale = in / out.
__global__ void si_half_image_kernel(const float* __restrict__ img, float* __restrict__ out, int C, int H, int W, int oh,
                                     int ow) {
  const long long total = (long long)C * H * W;
  const float sh = (float)H / (float)oh, sw = (float)W / (float)ow;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(t % W);
    const int y = (int)((t / W) % H);
    const int c = (int)(t / ((long long)W * H));
    float v = 0.f;
    if (y < oh && x < ow) {
      float fy = sh * ((float)y + 0.5f) - 0.5f;
      float fx = sw * ((float)x + 0.5f) - 0.5f;
      fy = fy < 0.f ? 0.f : fy;
      fx = fx < 0.f ? 0.f : fx;
      const int y0 = (int)fy, x0 = (int)fx;
      const int yp = y0 < H - 1 ? 1 : 0, xp = x0 < W - 1 ? 1 : 0;
      const float ly = fy - (float)y0, lx = fx - (float)x0;
      const float hy = 1.f - ly, hx = 1.f - lx;
      const float* p = img + ((long long)c * H + y0) * W + x0;
      v = hy * (hx * p[0] + lx * p[xp]) + ly * (hx * p[(long long)yp * W] + lx * p[(long long)yp * W + xp]);
    }
    out[t] = v;
  }
}
}  // namespace dslb

extern "C" int dslb_si_half_image(const float* img, float* out, int C, int H, int W, void* stream) {
  DSLB_CHECK_ARG(img && out && C > 0 && H >= 2 && W >= 2, "dslb_si_half_image: bad arguments");
  const long long total = (long long)C * H * W;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)dslb::num_sms() * 16) blocks = (long long)dslb::num_sms() * 16;
  dslb::si_half_image_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(img, out, C, H, W, H / 2, W / 2);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" int dslb_maxpool3x3s2(const void* x, void* y, int N, int H, int W, int C, void* stream) {
  DSLB_CHECK_ARG(x && y && C % 8 == 0, "dslb_maxpool3x3s2: C must be a multiple of 8");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = (long long)N * Ho * ((Wo + 1) / 2) * (C / 8);
  maxpool3x3s2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y,
                                                                                N, H, W, C, Ho, Wo);
  LAUNCH_CHECK();
}

extern "C" int dslb_upsample_add(void* dst, const void* src, int N, int H, int W, int h, int w, int C, void* stream) {
  DSLB_CHECK_ARG(dst && src && C % 8 == 0, "dslb_upsample_add: C must be a multiple of 8");
  const long long total = (long long)N * H * W * (C / 8);
  upsample_add_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)dst,
                                                                                (const __nv_bfloat16*)src, N, H, W, h, w, C);
  LAUNCH_CHECK();
}

extern "C" int dslb_upsample_add_bwd(void* dsrc, const void* ddst, int N, int H, int W, int h, int w, int C,
                                     void* stream) {
  DSLB_CHECK_ARG(dsrc && ddst && C % 8 == 0, "dslb_upsample_add_bwd: C must be a multiple of 8");
  const long long total = (long long)N * h * w * (C / 8);
  upsample_add_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (__nv_bfloat16*)dsrc, (const __nv_bfloat16*)ddst, N, H, W, h, w, C);
  LAUNCH_CHECK();
}

extern "C" int dslb_relu_family(const void* x, const void* m, void* y, long long n, int mode, void* stream) {
  DSLB_CHECK_ARG(x && y && n % 8 == 0 && mode >= 0 && mode <= 2 && (mode == 0 || m), "dslb_relu_family: bad arguments");
  relu_family_kernel<<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)m, (__nv_bfloat16*)y, n / 8, mode);
  LAUNCH_CHECK();
}

static int fill_gn_params(GnParams& P, const dslb_gn_seg_t* segs, int nseg, int C, int groups, float eps) {
  DSLB_CHECK_ARG(segs && nseg >= 1 && nseg <= DSLB_MAX_SEGS, "gn: nseg out of range");
  DSLB_CHECK_ARG(C % 8 == 0 && groups > 0 && C % groups == 0 && (C / groups) % 8 == 0,
                 "gn: C=%d groups=%d unsupported (channels per group must be a multiple of 8)", C, groups);
  P.nseg = nseg;
  P.C = C;
  P.cpg = C / groups;
  P.eps = eps;
  long long w = 0;
  for (int i = 0; i < nseg; ++i) {
    GnSeg& d = P.seg[i];
    d.x = (const __nv_bfloat16*)segs[i].x;
    d.y = (__nv_bfloat16*)segs[i].y;
    d.dz = (const __nv_bfloat16*)segs[i].dz;
    d.stats = segs[i].stats;
    d.gamma = segs[i].gamma;
    d.beta = segs[i].beta;
    d.red = segs[i].red;
    d.dbias = segs[i].dbias;
    d.mr = reinterpret_cast<float4*>(segs[i].mr);
    d.gsums = segs[i].gsums;
    DSLB_CHECK_ARG(d.x && d.y && d.stats && d.gamma && d.beta && d.mr, "gn: seg %d has a null pointer", i);
    d.HW = segs[i].HW;
    d.npix = segs[i].N * segs[i].HW;
    d.work_begin = w;
    w += (long long)d.npix * (C / 8);
  }
  P.total = w;
  return DSLB_OK;
}

extern "C" int dslb_gn_apply_relu_tab(const dslb_gn_seg_t* segs, int nseg, int C, int groups, float eps,
                                      const int* blk_tab_dev, int nblocks, void* stream) {
  DSLB_CHECK_ARG(C == 256 && blk_tab_dev && nblocks > 0, "dslb_gn_apply_relu_tab: C must be 256 and a block table given");
  GnParams P;
  int rc = fill_gn_params(P, segs, nseg, C, groups, eps);
  if (rc != DSLB_OK) return rc;
  gn_apply_relu_tab_kernel<<<nblocks, 256, 0, (cudaStream_t)stream>>>(P, blk_tab_dev, blk_tab_dev + nblocks);
  LAUNCH_CHECK();
}

namespace dslb {
// Split-bf16 ("bf16x3") operands of the accurate FCOSHead mode: a value v travels as hi = bf16(v), lo = bf16(v - hi) and a
// conv over the 3C-channel row [hi | lo | hi] against the weights [w_hi | w_hi | w_lo] accumulates
// hi*w_hi + lo*w_hi + hi*w_lo in fp32 on the tensor core: operand precision ~2^-17 instead of bf16's 2^-9.
__device__ __forceinline__ void split_store8(const float (&v)[8], __nv_bfloat16* row, int C, int c0) {
  float hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    hi[e] = __bfloat162float(__float2bfloat16_rn(v[e]));
    lo[e] = v[e] - hi[e];
  }
  const uint4 h = pack8(hi), l = pack8(lo);
  *reinterpret_cast<uint4*>(row + c0) = h;
  *reinterpret_cast<uint4*>(row + C + c0) = l;
  *reinterpret_cast<uint4*>(row + 2 * C + c0) = h;
}

// y[3C] = split(relu((x - mean) * rstd * gamma + beta)) with x in FP32 (the accurate mode keeps the tower maps in fp32)
__global__ void gn_apply_relu_split_kernel(const __grid_constant__ GnParams P) {
  const int cv = P.C / 8;
  const int G = P.C / P.cpg;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < P.total; t += (long long)gridDim.x * blockDim.x) {
    const GnSeg& s = P.seg[gn_find(P, t)];
    const long long tl = t - s.work_begin;
    const int c8 = (int)(tl % cv);
    const long long pix = tl / cv;
    const int n = (int)(pix / s.HW);
    const float4 mr = __ldg(s.mr + n * G + (c8 * 8) / P.cpg);
    const float* xp = reinterpret_cast<const float*>(s.x) + pix * P.C + c8 * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(xp)), b = __ldg(reinterpret_cast<const float4*>(xp) + 1);
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float ga = __ldg(s.gamma + c8 * 8 + e) * mr.y;
      v[e] = fmaxf(fmaf(x[e], ga, __ldg(s.beta + c8 * 8 + e) - mr.x * ga), 0.f);
    }
    split_store8(v, s.y + pix * 3 * P.C, P.C, c8 * 8);
  }
}

// bf16 map [npix][C] -> [npix][3C] = [x | 0 | x]   (a bf16 value is its own hi part)
__global__ void bf16_to_split_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long npix,
                                     int C) {
  const int cv = C / 8;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < npix * cv; t += (long long)gridDim.x * blockDim.x) {
    const long long pix = t / cv;
    const int c0 = (int)(t - pix * cv) * 8;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + pix * C + c0));
    __nv_bfloat16* row = y + pix * 3 * C;
    *reinterpret_cast<uint4*>(row + c0) = v;
    *reinterpret_cast<uint4*>(row + C + c0) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(row + 2 * C + c0) = v;
  }
}
}  // namespace dslb

extern "C" int dslb_gn_apply_relu_split(const dslb_gn_seg_t* segs, int nseg, int C, int groups, float eps, void* stream) {
  GnParams P;
  int rc = fill_gn_params(P, segs, nseg, C, groups, eps);
  if (rc != DSLB_OK) return rc;
  int maxN = 1;
  for (int i = 0; i < nseg; ++i) maxN = segs[i].N > maxN ? segs[i].N : maxN;
  const int nfin = nseg * maxN * groups;
  gn_finalize_kernel<<<(nfin + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P, maxN);
  DSLB_CHECK_CUDA(cudaGetLastError());
  gn_apply_relu_split_kernel<<<grid_for(P.total, 256), 256, 0, (cudaStream_t)stream>>>(P);
  LAUNCH_CHECK();
}

namespace dslb {
__global__ void scatter_f32_kernel(float* __restrict__ dst, const long long* __restrict__ idx, const float* __restrict__ src,
                                   int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = src[i];
}
__global__ void f64_to_f32_kernel(const double* __restrict__ src, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (float)src[i];
}
}  // namespace dslb

extern "C" int dslb_zero(void* p, size_t bytes, void* stream) {
  DSLB_CHECK_ARG(p || bytes == 0, "dslb_zero: null pointer");
  if (bytes) DSLB_CHECK_CUDA(cudaMemsetAsync(p, 0, bytes, (cudaStream_t)stream));
  return DSLB_OK;
}

extern "C" int dslb_scatter_f32(float* dst, const int64_t* idx, const float* src, int n, void* stream) {
  DSLB_CHECK_ARG(dst && idx && src && n >= 0, "dslb_scatter_f32: bad arguments");
  if (n == 0) return DSLB_OK;
  scatter_f32_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dst, (const long long*)idx, src, n);
  LAUNCH_CHECK();
}

extern "C" int dslb_f64_to_f32(const double* src, float* dst, int n, void* stream) {
  DSLB_CHECK_ARG(src && dst && n >= 0, "dslb_f64_to_f32: bad arguments");
  if (n == 0) return DSLB_OK;
  f64_to_f32_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(src, dst, n);
  LAUNCH_CHECK();
}

extern "C" int dslb_bf16_to_split(const void* x, void* y, long long npix, int C, void* stream) {
  DSLB_CHECK_ARG(x && y && npix >= 0 && C % 8 == 0, "dslb_bf16_to_split: bad arguments");
  if (npix == 0) return DSLB_OK;
  bf16_to_split_kernel<<<grid_for(npix * (C / 8), 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x,
                                                                                      (__nv_bfloat16*)y, npix, C);
  LAUNCH_CHECK();
}

extern "C" int dslb_gn_apply_relu(const dslb_gn_seg_t* segs, int nseg, int C, int groups, float eps, void* stream) {
  GnParams P;
  int rc = fill_gn_params(P, segs, nseg, C, groups, eps);
  if (rc != DSLB_OK) return rc;
  int maxN = 1;
  for (int i = 0; i < nseg; ++i) maxN = segs[i].N > maxN ? segs[i].N : maxN;
  const int nfin = nseg * maxN * groups;
  gn_finalize_kernel<<<(nfin + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P, maxN);
  DSLB_CHECK_CUDA(cudaGetLastError());
  gn_apply_relu_kernel<<<grid_for(P.total, 256), 256, 0, (cudaStream_t)stream>>>(P);
  LAUNCH_CHECK();
}

// Backward of conv-bias -> GroupNorm -> ReLU for up to DSLB_MAX_SEGS maps. `blk_tab` is caller-owned device memory of
// 2 * dslb_gn_bwd_blocks(...) ints, filled by dslb_gn_bwd_plan (host pointers in, one cudaMemcpyAsync).
extern "C" int dslb_gn_bwd_blocks(const dslb_gn_seg_t* segs, int nseg) {
  long long b = 0;
  for (int i = 0; i < nseg; ++i) b += (long long)segs[i].N * ((segs[i].HW + GN_PIXB - 1) / GN_PIXB);
  return (int)b;
}

extern "C" int dslb_gn_bwd_plan(const dslb_gn_seg_t* segs, int nseg, int* blk_tab_host) {
  DSLB_CHECK_ARG(segs && blk_tab_host, "dslb_gn_bwd_plan: null argument");
  const int nb = dslb_gn_bwd_blocks(segs, nseg);
  int k = 0;
  for (int i = 0; i < nseg; ++i)
    for (int n = 0; n < segs[i].N; ++n)
      for (int p = 0; p < segs[i].HW; p += GN_PIXB) {
        blk_tab_host[k] = i;
        blk_tab_host[nb + k] = n * segs[i].HW + p;
        ++k;
      }
  return DSLB_OK;
}

extern "C" int dslb_gn_bwd(const dslb_gn_seg_t* segs, int nseg, int C, int groups, float eps, const int* blk_tab_dev,
                           int nblocks, void* stream) {
  DSLB_CHECK_ARG(C == 256, "dslb_gn_bwd: only C == 256 is supported (got %d)", C);
  GnParams P;
  int rc = fill_gn_params(P, segs, nseg, C, groups, eps);
  if (rc != DSLB_OK) return rc;
  int with_sums = 0;
  for (int i = 0; i < nseg; ++i) {
    DSLB_CHECK_ARG(segs[i].dz && segs[i].red && segs[i].dbias, "dslb_gn_bwd: seg %d has a null dz / red / dbias", i);
    with_sums += segs[i].gsums != nullptr;
  }
  DSLB_CHECK_ARG(with_sums == 0 || with_sums == nseg, "dslb_gn_bwd: gsums must be given for every segment or for none");
  if (with_sums) {   // the group sums came out of the producing convs' epilogues: one pass
    gn_bwd_apply_kernel<true><<<nblocks, 256, 0, (cudaStream_t)stream>>>(P, blk_tab_dev, blk_tab_dev + nblocks);
    LAUNCH_CHECK();
  }
  gn_bwd_reduce_kernel<<<nblocks, 256, 0, (cudaStream_t)stream>>>(P, blk_tab_dev, blk_tab_dev + nblocks);
  DSLB_CHECK_CUDA(cudaGetLastError());
  int maxN = 1;
  for (int i = 0; i < nseg; ++i) maxN = segs[i].N > maxN ? segs[i].N : maxN;
  const int nfin = nseg * maxN * groups;
  gn_bwd_finalize_kernel<<<(nfin + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P, maxN);
  DSLB_CHECK_CUDA(cudaGetLastError());
  gn_bwd_apply_kernel<false><<<nblocks, 256, 0, (cudaStream_t)stream>>>(P, blk_tab_dev, blk_tab_dev + nblocks);
  LAUNCH_CHECK();
}

extern "C" int dslb_gn_bwd_params(const double* red, float* dgamma, float* dbeta, int N, int C, void* stream) {
  DSLB_CHECK_ARG(red && dgamma && dbeta, "dslb_gn_bwd_params: null argument");
  gn_bwd_params_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(red, dgamma, dbeta, N, C);
  LAUNCH_CHECK();
}

extern "C" int dslb_pack_weight(const float* w, void* out, int O, int I, int R, int S, int rows_pad, int cols_pad,
                                const float* oscale, int transpose, void* stream) {
  DSLB_CHECK_ARG(w && out && O > 0 && I > 0, "dslb_pack_weight: bad arguments");
  DSLB_CHECK_ARG(rows_pad >= (transpose ? I : O) && cols_pad >= (transpose ? O : I), "dslb_pack_weight: padding too small");
  const long long total = (long long)R * S * rows_pad * cols_pad;
  pack_weight_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)out, O, I, R, S, rows_pad,
                                                                               cols_pad, oscale, transpose);
  LAUNCH_CHECK();
}

extern "C" int dslb_unpack_wgrad(const float* dw, float* g, int O, int I, int R, int S, int rows, const float* oscale,
                                 int accumulate, void* stream) {
  DSLB_CHECK_ARG(dw && g && rows >= O, "dslb_unpack_wgrad: bad arguments");
  const long long total = (long long)O * I * R * S;
  unpack_wgrad_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dw, g, O, I, R, S, rows, oscale, accumulate);
  LAUNCH_CHECK();
}

extern "C" int dslb_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                            float* scale, float* shift, int C, void* stream) {
  DSLB_CHECK_ARG(gamma && beta && mean && var && scale && shift, "dslb_bn_fold: null argument");
  bn_fold_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gamma, beta, mean, var, eps, scale, shift, C);
  LAUNCH_CHECK();
}

extern "C" int dslb_colsum(const void* x, float* out, long long npix, int ld, int C, void* stream) {
  DSLB_CHECK_ARG(x && out && ld >= C, "dslb_colsum: bad arguments");
  if (C % 8 == 0 && C <= 2048 && ld % 8 == 0 && ((uintptr_t)x % 16) == 0 && getenv("DSLB_COLSUM_SCALAR") == nullptr) {
    const int T = C / 8, R = 256 / T;
    // ~16 rows per thread at least, capped at 3 blocks per SM: every block ends with C reductions onto the same few
    // cache lines, so fewer, longer blocks (4 x 16-byte loads in flight per thread) beat a wide grid
    const int grid = grid_for(npix / ((long long)R * 16) + 1, 1, 3);
    if (((uintptr_t)out % 16) == 0 && getenv("DSLB_COLSUM_NO_V4") == nullptr)
      colsum_vec_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, out, npix, ld, T, R);
    else
      colsum_vec_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, out, npix, ld, T, R);
    LAUNCH_CHECK();
  }
  colsum_kernel<<<grid_for(npix / 8 + 1, 1, 4), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, out, npix, ld, C);
  LAUNCH_CHECK();
}

