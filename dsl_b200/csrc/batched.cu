// Multi-tensor ("one launch for the whole network") parameter kernels: refresh of the derived bf16 conv operands
// from the fp32 master weights with the frozen BatchNorm folded in, and the reverse map from the packed fp32 weight
// gradients to the OIHW gradient views. A plan owns a device copy of its descriptor table; running it is ONE launch.
#include <new>

#include "common.h"

#include <cuda_bf16.h>

namespace dslb {

struct PackDescDev {
  const float* w;
  __nv_bfloat16* out;
  const float* bn_gamma;
  const float* bn_beta;
  const float* bn_mean;
  const float* bn_var;
  float* scale_out;
  float* shift_out;
  int O, I, R, S;
  int rows_pad, cols_pad, row_off, col_off;
  int rows, cols8;  // iteration space: rows x (cols8 * 8) per tap
  int real_cols, fill;
  int mode;
  float eps;
  int ldw;               // input channels per output-channel row of w in memory (>= I; > I for a slice of a wider tensor)
  int it;                // tiled kernel: input channels per tile
  __nv_bfloat16* out2;   // tiled kernel: the dgrad operand of a merged (mode 0 + mode 1) descriptor, else null
  long long work_begin;  // prefix sum of rows * cols8
};

// One thread = 8 consecutive packed columns of one packed row, ALL filter taps (so the R*S strided fp32 reads of one
// thread fall into the same few cache lines, and every tap plane gets one 16-byte store).
//   mode 0 (fprop)  out[tap][row_off + o][col_off + i] = w[o][i][r][s] * sc[o]
//   mode 1 (dgrad)  out[tap][row_off + i][col_off + o] = w[o][i][R-1-r][S-1-s] * sc[o]
//   mode 2 (im2col) out[0][row_off + o][(r*S+s)*I + i] = w[o][i][r][s] * sc[o]      (stem: K = R*S*I in one row)
__global__ void pack_batched_kernel(const PackDescDev* __restrict__ D, int n, long long total) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (D[mid].work_begin <= t) lo = mid; else hi = mid - 1;
    }
    const PackDescDev& d = D[lo];
    const unsigned tl = (unsigned)(t - d.work_begin);   // per-descriptor work fits 32 bits (checked at plan creation)
    const int c8 = (int)(tl % (unsigned)d.cols8);
    const int row = (int)(tl / (unsigned)d.cols8);
    const int RS = d.R * d.S;
    const int lim = d.real_cols - c8 * 8;
    const bool vec = (d.col_off & 7) == 0 && (d.fill || lim >= 8);
    float sc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) sc[e] = 1.f;
    bool row_real;
    if (d.mode == 1) {
      row_real = row < d.I;
      if (d.bn_gamma) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int o = c8 * 8 + e;
          if (o < d.O) sc[e] = d.bn_gamma[o] / sqrtf(d.bn_var[o] + d.eps);
        }
      }
    } else {
      row_real = row < d.O;
      if (row_real && d.bn_gamma) {
        const float s0 = d.bn_gamma[row] / sqrtf(d.bn_var[row] + d.eps);
#pragma unroll
        for (int e = 0; e < 8; ++e) sc[e] = s0;
        if (c8 == 0) {
          d.scale_out[row] = s0;
          d.shift_out[row] = d.bn_beta[row] - d.bn_mean[row] * s0;
        }
      }
    }
    const int ntap = d.mode == 2 ? 1 : RS;
    for (int tap = 0; tap < ntap; ++tap) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
      if (row_real) {
        if (d.mode == 0) {
          const float* src = d.w + ((long long)row * d.ldw) * RS + tap;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int i = c8 * 8 + e;
            if (i < d.I) v[e] = src[(long long)i * RS] * sc[e];
          }
        } else if (d.mode == 1) {
          const int rtap = RS - 1 - tap;  // 180-degree rotation
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int o = c8 * 8 + e;
            if (o < d.O) v[e] = d.w[((long long)o * d.ldw + row) * RS + rtap] * sc[e];
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int k = c8 * 8 + e;
            if (k < RS * d.I) {
              const int tp = k / d.I, i = k - tp * d.I;
              v[e] = d.w[((long long)row * d.ldw + i) * RS + tp] * sc[e];
            }
          }
        }
      }
      __nv_bfloat162 h[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
      __nv_bfloat16* dst = d.out + ((long long)tap * d.rows_pad + d.row_off + row) * d.cols_pad + d.col_off + c8 * 8;
      if (vec) {
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<uint4*>(h);
      } else {  // partial / unaligned sub-block (conv_reg + conv_centerness share one packed operand): real columns only
        const __nv_bfloat16* hv = reinterpret_cast<const __nv_bfloat16*>(h);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (e < lim) dst[e] = hv[e];
      }
    }
  }
}

// Tiled variant for the regular convs (O, I multiples of 32, no padding / offsets, <= 9 taps): one block = 32 output
// channels x `it` input channels x all taps, `it` = 32 * f with f the largest power of two such that f * R*S <= 9 and
// it | I (256 input channels for a 1x1 conv, 32 for a 3x3), so a tile is up to 32 x 288 floats whatever the filter.
// The OIHW rows are read fully coalesced into shared memory (scaled by the folded BatchNorm) ONCE and leave as the
// fprop operand [tap][o][i] and/or, transposed and rotated by 180 degrees, the dgrad operand [RS-1-tap][i][o]: a
// descriptor whose `out2` is set is the merge of a mode-0 and a mode-1 descriptor over the same weight (the plan
// pairs them at creation). `work_begin` counts tiles here.
constexpr int PT = 32;
constexpr int PT_MAX_TAPS = 9;
constexpr int PT_PITCH = PT * PT_MAX_TAPS + 1;

__device__ __forceinline__ uint32_t pack_bf162(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(256, 5) pack_tiled_kernel(const PackDescDev* __restrict__ D, int n, int total_tiles) {
  __shared__ float st[PT][PT_PITCH];
  __shared__ float ssc[PT];
  const int tid = threadIdx.x;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (D[mid].work_begin <= tile) lo = mid; else hi = mid - 1;
    }
    const PackDescDev& d = D[lo];
    const int tl = tile - (int)d.work_begin;
    const int IT = d.it;
    const int itiles = d.I / IT;
    const int ot = tl / itiles, it = tl - ot * itiles;
    const int o0 = ot * PT, i0 = it * IT;
    const int RS = d.R * d.S;
    const int pieces = (IT >> 5) * RS;   // 128-byte pieces per tile row, <= 9
    if (tid < PT) {
      float sc = 1.f;
      if (d.bn_gamma) {
        const int o = o0 + tid;
        sc = d.bn_gamma[o] / sqrtf(d.bn_var[o] + d.eps);
        if (d.scale_out && it == 0) {
          d.scale_out[o] = sc;
          d.shift_out[o] = d.bn_beta[o] - d.bn_mean[o] * sc;
        }
      }
      ssc[tid] = sc;
    }
    __syncthreads();
    {
      // warp wv reads rows wv, wv+8, ...: the (up to) 9 x 128-byte pieces of a row are independent loads issued back to
      // back, so every thread keeps ~9 requests in flight (the plain strided loop was latency bound)
      const int wv = tid >> 5, lane = tid & 31;
#pragma unroll
      for (int rr = 0; rr < PT / 8; ++rr) {
        const int oo = wv + 8 * rr;
        const float* __restrict__ src = d.w + ((long long)(o0 + oo) * d.ldw + i0) * RS + lane;
        float v[PT_MAX_TAPS];
#pragma unroll
        for (int j = 0; j < PT_MAX_TAPS; ++j)
          if (j < pieces) v[j] = __ldg(src + 32 * j);
        const float sc = ssc[oo];
#pragma unroll
        for (int j = 0; j < PT_MAX_TAPS; ++j)
          if (j < pieces) st[oo][lane + 32 * j] = v[j] * sc;
      }
    }
    __syncthreads();
    __nv_bfloat16* const out_f = d.mode == 0 ? d.out : nullptr;
    __nv_bfloat16* const out_t = d.mode == 1 ? d.out : d.out2;
    if (out_f) {
      if (RS == 1) {
        // 1x1: a tile row is `IT` consecutive packed columns; one lane = 2 of them (conflict-free shared reads,
        // 128-byte warp stores)
        const int half = IT >> 1;
        for (int idx = tid; idx < PT * half; idx += 256) {
          const int oo = idx / half, pr = idx - oo * half;
          const float* sp = &st[oo][2 * pr];
          *reinterpret_cast<uint32_t*>(out_f + (long long)(o0 + oo) * d.cols_pad + i0 + 2 * pr) = pack_bf162(sp[0], sp[1]);
        }
      } else {
        // 16-byte stores: one thread = 8 consecutive packed columns of one (tap, row)
        const int i8n = IT >> 3;
        const int octs = RS * PT * i8n;
        for (int idx = tid; idx < octs; idx += 256) {
          const int i8 = idx % i8n, r2 = idx / i8n;
          const int oo = r2 & (PT - 1), tap = r2 >> 5;
          const float* sp = &st[oo][(8 * i8) * RS + tap];
          uint4 o;
          o.x = pack_bf162(sp[0], sp[RS]);
          o.y = pack_bf162(sp[2 * RS], sp[3 * RS]);
          o.z = pack_bf162(sp[4 * RS], sp[5 * RS]);
          o.w = pack_bf162(sp[6 * RS], sp[7 * RS]);
          *reinterpret_cast<uint4*>(out_f + ((long long)tap * d.rows_pad + o0 + oo) * d.cols_pad + i0 + 8 * i8) = o;
        }
      }
    }
    if (out_t) {
      // transposed operand: rows are input channels, 32 output channels = one 64-byte run per (tap, input channel)
      const int rows_t = d.mode == 1 ? d.rows_pad : d.I, cols_t = d.mode == 1 ? d.cols_pad : d.O;
      const int octs = RS * IT * (PT / 8);
      for (int idx = tid; idx < octs; idx += 256) {
        const int o8 = idx & 3, r2 = idx >> 2;
        const int ii = r2 % IT, tap = r2 / IT;
        const float* sp = &st[8 * o8][ii * RS + tap];
        uint4 o;
        o.x = pack_bf162(sp[0], sp[PT_PITCH]);
        o.y = pack_bf162(sp[2 * PT_PITCH], sp[3 * PT_PITCH]);
        o.z = pack_bf162(sp[4 * PT_PITCH], sp[5 * PT_PITCH]);
        o.w = pack_bf162(sp[6 * PT_PITCH], sp[7 * PT_PITCH]);
        *reinterpret_cast<uint4*>(out_t + ((long long)(RS - 1 - tap) * rows_t + i0 + ii) * cols_t + o0 + 8 * o8) = o;
      }
    }
    __syncthreads();
  }
}

struct UnpackDescDev {
  const float* dw;
  float* g;
  const float* bn_gamma;
  const float* bn_var;
  int O, I, R, S;
  int rows, row_off;
  float eps;
  int ldd, ldg;          // columns per packed dw row (>= I), input channels per output-channel row of g (>= I)
  int vec4;              // 1x1 conv with I % 4 == 0 and 16-byte aligned buffers: work items are float4s
  long long work_begin;  // prefix of O*I (O*I/4 for vec4 descriptors)
};

// g[o][i][r][s] = dw[r*S+s][row_off + o][i] * sc[o]. One thread = one (o, i) pair, all taps: the reads of a warp are
// contiguous inside every tap plane and its R*S-float writes tile a contiguous span of g.
__global__ void unpack_batched_kernel(const UnpackDescDev* __restrict__ D, int n, long long total) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (D[mid].work_begin <= t) lo = mid; else hi = mid - 1;
    }
    const UnpackDescDev& d = D[lo];
    const int RS = d.R * d.S;
    if (d.vec4) {  // 1x1 conv, I % 4 == 0: work item = 4 consecutive input channels (a plain scaled float4 copy)
      const unsigned tl4 = (unsigned)(t - d.work_begin);
      const unsigned i4n = (unsigned)d.I >> 2;
      const int o = (int)(tl4 / i4n);
      const int i = (int)(tl4 - (unsigned)o * i4n) << 2;
      const float sc = d.bn_gamma ? d.bn_gamma[o] / sqrtf(d.bn_var[o] + d.eps) : 1.f;
      float4 v = __ldg(reinterpret_cast<const float4*>(d.dw + ((long long)(d.row_off + o)) * d.ldd + i));
      v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
      *reinterpret_cast<float4*>(d.g + (long long)o * d.ldg + i) = v;
      continue;
    }
    const unsigned tl = (unsigned)(t - d.work_begin);
    const int i = (int)(tl % (unsigned)d.I);
    const int o = (int)(tl / (unsigned)d.I);
    const float sc = d.bn_gamma ? d.bn_gamma[o] / sqrtf(d.bn_var[o] + d.eps) : 1.f;
    const float* src = d.dw + ((long long)(d.row_off + o)) * d.ldd + i;
    float* dst = d.g + ((long long)o * d.ldg + i) * RS;
    for (int tap = 0; tap < RS; ++tap) dst[tap] = src[(long long)tap * d.rows * d.ldd] * sc;
  }
}

// Coalesced variant: one work item = one output channel x 256 consecutive input channels. The 256 x RS values are
// gathered tap plane by tap plane (contiguous reads) into shared memory and leave as one contiguous run of g.
// `work_begin` counts (o, i-chunk) items here; serves every descriptor with 1 < RS <= 9.
__global__ void __launch_bounds__(256) unpack_tiled_kernel(const UnpackDescDev* __restrict__ D, int n, long long total) {
  __shared__ float st[256 * 9];
  const int tid = threadIdx.x;
  for (long long item = blockIdx.x; item < total; item += gridDim.x) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (D[mid].work_begin <= item) lo = mid; else hi = mid - 1;
    }
    const UnpackDescDev& d = D[lo];
    const int tl = (int)(item - d.work_begin);
    const int chunks = (d.I + 255) >> 8;
    const int o = tl / chunks, i0 = (tl - o * chunks) << 8;
    const int ni = min(256, d.I - i0);
    const int RS = d.R * d.S;
    const float sc = d.bn_gamma ? d.bn_gamma[o] / sqrtf(d.bn_var[o] + d.eps) : 1.f;
    if (tid < ni) {
      const float* __restrict__ src = d.dw + ((long long)(d.row_off + o)) * d.ldd + i0 + tid;
      const long long plane = (long long)d.rows * d.ldd;
      float v[9];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap)
        if (tap < RS) v[tap] = __ldg(src + tap * plane);   // independent loads, all in flight together
#pragma unroll
      for (int tap = 0; tap < 9; ++tap)
        if (tap < RS) st[tid * RS + tap] = v[tap] * sc;
    }
    __syncthreads();
    float* dst = d.g + ((long long)o * d.ldg + i0) * RS;
    for (int idx = tid; idx < ni * RS; idx += 256) dst[idx] = st[idx];
    __syncthreads();
  }
}

// y[n][2p][2q][:] = x[n][p][q][:], every other pixel of the [N][H][W][C] map = 0 (C % 8 == 0).
__global__ void zero_upsample2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int h,
                                      int w, int H, int W, int cv) {
  const long long total = (long long)N * H * W * cv;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c8 = t % cv;
    const long long pix = t / cv;
    const int X = pix % W;
    const int Y = (pix / W) % H;
    const int n = pix / ((long long)W * H);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (!(X & 1) && !(Y & 1) && (Y >> 1) < h && (X >> 1) < w)
      v = reinterpret_cast<const uint4*>(x)[(((long long)n * h + (Y >> 1)) * w + (X >> 1)) * cv + c8];
    reinterpret_cast<uint4*>(y)[t] = v;
  }
}

__global__ void regctr_affine_kernel(const float* __restrict__ scales, int stride, const float* __restrict__ reg_bias,
                                     const float* __restrict__ ctr_bias, const float* __restrict__ level_mult,
                                     float* __restrict__ rc_scale, float* __restrict__ rc_shift,
                                     float* __restrict__ scale_vals, int nl) {
  const int t = threadIdx.x;
  if (t >= nl * 8) return;
  const int l = t >> 3, j = t & 7;
  const float sv = scales[(long long)l * stride];
  const float sc = j < 4 ? sv * level_mult[l] : 1.f;
  const float b = j < 4 ? reg_bias[j] : (j == 4 ? ctr_bias[0] : 0.f);
  rc_scale[t] = sc;
  rc_shift[t] = b * sc;
  if (j == 0) scale_vals[l] = sv;
}

}  // namespace dslb

using namespace dslb;

struct dslb_table_plan {
  void* dev = nullptr;
  int n = 0;
  long long total = 0;
  int kind = 0;  // 0 pack, 1 unpack
  void* dev_tiled = nullptr;  // pack only: descriptors served by pack_tiled_kernel
  int n_tiled = 0;
  int tiles = 0;
};

extern "C" int dslb_pack_plan_create(const dslb_pack_desc_t* descs, int n, dslb_table_plan_t** out) {
  DSLB_CHECK_ARG(descs && out && n >= 1, "dslb_pack_plan_create: bad arguments");
  PackDescDev* h = new (std::nothrow) PackDescDev[n];
  if (!h) {
    set_error("out of host memory");
    return DSLB_ENOMEM;
  }
  PackDescDev* ht = new (std::nothrow) PackDescDev[n];
  if (!ht) {
    delete[] h;
    set_error("out of host memory");
    return DSLB_ENOMEM;
  }
  long long w = 0;
  int nslow = 0, ntiled = 0;
  long long tiles = 0;
  for (int k = 0; k < n; ++k) {
    const dslb_pack_desc_t& s = descs[k];
    PackDescDev d0;
    PackDescDev& d = d0;
    const bool ok = s.w && s.out && s.O > 0 && s.I > 0 && s.R > 0 && s.S > 0 && s.mode >= 0 && s.mode <= 2 &&
                    s.cols_pad % 8 == 0 && (!s.bn_gamma || (s.bn_beta && s.bn_mean && s.bn_var)) &&
                    (s.mode == 1 || !s.bn_gamma || (s.scale_out && s.shift_out));
    if (!ok) {
      delete[] h;
      delete[] ht;
      set_error("dslb_pack_plan_create: descriptor %d is invalid", k);
      return DSLB_EINVAL;
    }
    d.w = s.w;
    d.out = (__nv_bfloat16*)s.out;
    d.bn_gamma = s.bn_gamma;
    d.bn_beta = s.bn_beta;
    d.bn_mean = s.bn_mean;
    d.bn_var = s.bn_var;
    d.scale_out = s.scale_out;
    d.shift_out = s.shift_out;
    d.O = s.O; d.I = s.I; d.R = s.R; d.S = s.S;
    d.rows_pad = s.rows_pad; d.cols_pad = s.cols_pad; d.row_off = s.row_off; d.col_off = s.col_off;
    d.mode = s.mode;
    d.eps = s.bn_eps;
    d.ldw = s.w_ld > 0 ? s.w_ld : s.I;
    d.it = PT;
    d.out2 = nullptr;
    if (d.ldw < s.I) {
      delete[] h;
      delete[] ht;
      set_error("dslb_pack_plan_create: descriptor %d has w_ld < I", k);
      return DSLB_EINVAL;
    }
    const int real_rows = (s.mode == 1) ? s.I : s.O;
    const int real_cols = (s.mode == 1) ? s.O : (s.mode == 2 ? s.R * s.S * s.I : s.I);
    // full = rewrite the zero padding too; otherwise only the real sub-block (shared, offset outputs)
    d.rows = s.fill_padding ? s.rows_pad : real_rows;
    d.cols8 = s.fill_padding ? s.cols_pad / 8 : (real_cols + 7) / 8;
    d.real_cols = real_cols;
    d.fill = s.fill_padding ? 1 : 0;
    if (d.row_off + d.rows > s.rows_pad || d.col_off + (s.fill_padding ? s.cols_pad : real_cols) > s.cols_pad) {
      delete[] h;
      delete[] ht;
      set_error("dslb_pack_plan_create: descriptor %d does not fit its packed block", k);
      return DSLB_EINVAL;
    }
    // regular convs without padding go to the tiled (coalesced) kernel
    const bool tiled = s.mode <= 1 && s.O % PT == 0 && s.I % PT == 0 && s.R * s.S <= PT_MAX_TAPS && s.row_off == 0 &&
                       s.col_off == 0 && s.rows_pad == real_rows && s.cols_pad == real_cols && s.cols_pad % 8 == 0 &&
                       ((uintptr_t)s.out % 16) == 0;
    if (tiled) {
      int f = 1;
      while (2 * f * s.R * s.S <= PT_MAX_TAPS && (s.I / PT) % (2 * f) == 0) f *= 2;
      d.it = PT * f;
      ht[ntiled++] = d;
    } else {
      d.work_begin = w;
      w += (long long)d.rows * d.cols8;
      h[nslow++] = d;
    }
  }
  // a conv's fprop and dgrad operands come from the same weight: merge the two descriptors so the tile is read once
  for (int a = 0; a < ntiled; ++a) {
    if (ht[a].mode != 0) continue;
    for (int b = 0; b < ntiled; ++b) {
      const PackDescDev& t = ht[b];
      if (t.mode == 1 && t.w == ht[a].w && t.O == ht[a].O && t.I == ht[a].I && t.R == ht[a].R && t.S == ht[a].S &&
          t.ldw == ht[a].ldw && t.bn_gamma == ht[a].bn_gamma && t.bn_var == ht[a].bn_var && t.eps == ht[a].eps) {
        ht[a].out2 = t.out;
        ht[b].mode = -1;   // absorbed
        break;
      }
    }
  }
  {
    int m = 0;
    for (int k = 0; k < ntiled; ++k)
      if (ht[k].mode >= 0) ht[m++] = ht[k];
    ntiled = m;
  }
  for (int k = 0; k < ntiled; ++k) {
    ht[k].work_begin = tiles;
    tiles += (long long)(ht[k].O / PT) * (ht[k].I / ht[k].it);
  }
  dslb_table_plan* p = new (std::nothrow) dslb_table_plan();
  cudaError_t e = p ? cudaSuccess : cudaErrorMemoryAllocation;
  if (e == cudaSuccess && nslow) e = cudaMalloc(&p->dev, sizeof(PackDescDev) * nslow);
  if (e == cudaSuccess && nslow) e = cudaMemcpy(p->dev, h, sizeof(PackDescDev) * nslow, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && ntiled) e = cudaMalloc(&p->dev_tiled, sizeof(PackDescDev) * ntiled);
  if (e == cudaSuccess && ntiled) e = cudaMemcpy(p->dev_tiled, ht, sizeof(PackDescDev) * ntiled, cudaMemcpyHostToDevice);
  delete[] h;
  delete[] ht;
  if (e != cudaSuccess || tiles > 0x7fffffffLL) {
    set_error("dslb_pack_plan_create: %s", e != cudaSuccess ? cudaGetErrorString(e) : "too many tiles");
    if (p && p->dev) cudaFree(p->dev);
    if (p && p->dev_tiled) cudaFree(p->dev_tiled);
    delete p;
    return DSLB_ECUDA;
  }
  p->n = nslow;
  p->total = w;
  p->kind = 0;
  p->n_tiled = ntiled;
  p->tiles = (int)tiles;
  *out = p;
  return DSLB_OK;
}

extern "C" int dslb_unpack_plan_create(const dslb_unpack_desc_t* descs, int n, dslb_table_plan_t** out) {
  DSLB_CHECK_ARG(descs && out && n >= 1, "dslb_unpack_plan_create: bad arguments");
  UnpackDescDev* h = new (std::nothrow) UnpackDescDev[n];
  if (!h) {
    set_error("out of host memory");
    return DSLB_ENOMEM;
  }
  UnpackDescDev* ht = new (std::nothrow) UnpackDescDev[n];
  if (!ht) {
    delete[] h;
    set_error("out of host memory");
    return DSLB_ENOMEM;
  }
  long long w = 0, items = 0;
  int nslow = 0, ntiled = 0;
  for (int k = 0; k < n; ++k) {
    const dslb_unpack_desc_t& s = descs[k];
    UnpackDescDev d0;
    UnpackDescDev& d = d0;
    if (!(s.dw && s.g && s.O > 0 && s.I > 0 && s.R > 0 && s.S > 0 && s.rows >= s.row_off + s.O &&
          (!s.bn_gamma || s.bn_var))) {
      delete[] h;
      delete[] ht;
      set_error("dslb_unpack_plan_create: descriptor %d is invalid", k);
      return DSLB_EINVAL;
    }
    d.dw = s.dw; d.g = s.g; d.bn_gamma = s.bn_gamma; d.bn_var = s.bn_var;
    d.O = s.O; d.I = s.I; d.R = s.R; d.S = s.S; d.rows = s.rows; d.row_off = s.row_off; d.eps = s.bn_eps;
    d.ldd = s.dw_ld > 0 ? s.dw_ld : s.I;
    d.ldg = s.g_ld > 0 ? s.g_ld : s.I;
    if (d.ldd < s.I || d.ldg < s.I) {
      delete[] h;
      delete[] ht;
      set_error("dslb_unpack_plan_create: descriptor %d has dw_ld / g_ld < I", k);
      return DSLB_EINVAL;
    }
    d.vec4 = 0;
    if (s.R * s.S > 1 && s.R * s.S <= 9) {
      d.work_begin = items;
      items += (long long)s.O * ((s.I + 255) / 256);
      ht[ntiled++] = d;
    } else {
      d.vec4 = (s.R * s.S == 1 && s.I % 4 == 0 && ((uintptr_t)s.dw % 16) == 0 && ((uintptr_t)s.g % 16) == 0 &&
                ((long long)s.row_off * d.ldd) % 4 == 0 && d.ldd % 4 == 0 && d.ldg % 4 == 0) ? 1 : 0;
      d.work_begin = w;
      w += d.vec4 ? (long long)s.O * s.I / 4 : (long long)s.O * s.I;
      h[nslow++] = d;
    }
  }
  dslb_table_plan* p = new (std::nothrow) dslb_table_plan();
  cudaError_t e = p ? cudaSuccess : cudaErrorMemoryAllocation;
  if (e == cudaSuccess && nslow) e = cudaMalloc(&p->dev, sizeof(UnpackDescDev) * nslow);
  if (e == cudaSuccess && nslow) e = cudaMemcpy(p->dev, h, sizeof(UnpackDescDev) * nslow, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && ntiled) e = cudaMalloc(&p->dev_tiled, sizeof(UnpackDescDev) * ntiled);
  if (e == cudaSuccess && ntiled) e = cudaMemcpy(p->dev_tiled, ht, sizeof(UnpackDescDev) * ntiled, cudaMemcpyHostToDevice);
  delete[] h;
  delete[] ht;
  if (e != cudaSuccess || items > 0x7fffffffLL) {
    set_error("dslb_unpack_plan_create: %s", e != cudaSuccess ? cudaGetErrorString(e) : "too many work items");
    if (p && p->dev) cudaFree(p->dev);
    if (p && p->dev_tiled) cudaFree(p->dev_tiled);
    delete p;
    return DSLB_ECUDA;
  }
  p->n = nslow;
  p->total = w;
  p->kind = 1;
  p->n_tiled = ntiled;
  p->tiles = (int)items;
  *out = p;
  return DSLB_OK;
}

extern "C" int dslb_table_plan_run(const dslb_table_plan_t* p, void* stream) {
  DSLB_CHECK_ARG(p && (p->dev || p->dev_tiled), "dslb_table_plan_run: null plan");
  long long blocks = (p->total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (p->kind == 0) {
    if (p->n_tiled) {
      const int tb = p->tiles < num_sms() * 5 ? p->tiles : num_sms() * 5;
      pack_tiled_kernel<<<tb, 256, 0, (cudaStream_t)stream>>>((const PackDescDev*)p->dev_tiled, p->n_tiled, p->tiles);
    }
    if (p->n)
      pack_batched_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const PackDescDev*)p->dev, p->n, p->total);
  } else {
    if (p->n_tiled) {
      const int tb = p->tiles < num_sms() * 8 ? p->tiles : num_sms() * 8;
      unpack_tiled_kernel<<<tb, 256, 0, (cudaStream_t)stream>>>((const UnpackDescDev*)p->dev_tiled, p->n_tiled, p->tiles);
    }
    if (p->n)
      unpack_batched_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const UnpackDescDev*)p->dev, p->n, p->total);
  }
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" void dslb_table_plan_destroy(dslb_table_plan_t* p) {
  if (!p) return;
  if (p->dev) cudaFree(p->dev);
  if (p->dev_tiled) cudaFree(p->dev_tiled);
  delete p;
}

extern "C" int dslb_zero_upsample2(const void* x, void* y, int N, int h, int w, int H, int W, int C, void* stream) {
  DSLB_CHECK_ARG(x && y && C % 8 == 0 && H >= 2 * h - 1 && W >= 2 * w - 1, "dslb_zero_upsample2: bad arguments");
  const long long total = (long long)N * H * W * (C / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  zero_upsample2_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, N, h, w,
                                                                        H, W, C / 8);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" int dslb_fcos_regctr_affine(const float* scales, int scale_stride, const float* reg_bias, const float* ctr_bias,
                                       const float* level_mult, float* rc_scale, float* rc_shift, float* scale_vals,
                                       int nlevels, void* stream) {
  DSLB_CHECK_ARG(scales && reg_bias && ctr_bias && level_mult && rc_scale && rc_shift && scale_vals && nlevels >= 1 &&
                     nlevels <= 16,
                 "dslb_fcos_regctr_affine: bad arguments");
  regctr_affine_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(scales, scale_stride, reg_bias, ctr_bias, level_mult, rc_scale,
                                                            rc_shift, scale_vals, nlevels);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}
