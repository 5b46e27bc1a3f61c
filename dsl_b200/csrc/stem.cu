// ResNet stem: 7x7 stride-2 pad-3 conv (3 -> 64) + frozen BatchNorm + ReLU, straight from the NCHW fp32 image to the
// NHWC bf16 feature map in ONE kernel (reference: mmdet/models/backbones/resnet.py:597-610, 630-637).
//
// The im2col matrix is never materialised in HBM (it was 413 MB per pass at 4 x 800 x 1344): every CTA stages the 7
// input rows a 128-pixel output segment needs in shared memory (bf16, channel-interleaved), assembles the K-major
// 128-byte-swizzled A tile from it with 16-byte copies, and feeds tcgen05.mma (M128 x N64 x K16, fp32 accumulators in
// TMEM, double-buffered so the MMA of tile i overlaps the epilogue of tile i-1 and the gather of tile i+1).
//   pre-pass   NCHW fp32 -> NHWC bf16 with 4 channels per pixel (8 bytes), so a patch row is a plain byte range
//              that cp.async moves without touching registers (zero-fill outside the image = the conv padding).
//   K layout   k' = r*32 + s'*4 + c  (8 tap slots x 4 channels per filter row, K = 224 -> 256; slot s' = 0 is a pad tap,
//              slot s' = s + 1 carries filter tap s): the patch starts one pixel early (at an EVEN image column, so two
//              pixels move per 16-byte cp.async), and the 32 values of (pixel q, filter row r) are the 64 CONTIGUOUS
//              bytes patch[r][2q .. 2q+8): an A row is 28 aligned sixteen-byte chunks copied verbatim; the pad tap /
//              pad channel read real (finite) data and meet zero weights.
//   B operand  built in shared memory by the kernel itself from the fp32 OIHW master weight x BN scale.
#include <stdlib.h>
#include <string.h>

#include "common.h"
#include "ptx.cuh"

namespace dslb {

// ---- pre-pass: NCHW fp32 image -> NHWC bf16 with the channels padded 3 -> 4 (8 bytes per pixel)
__global__ void img_to_nhwc4_kernel(const float* __restrict__ img, uint2* __restrict__ out, int N, long long HW) {
  const long long total = (long long)N * HW;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long n = t / HW, p = t - n * HW;
    const float* src = img + n * 3 * HW + p;
    const __nv_bfloat162 a = __floats2bfloat162_rn(__ldg(src), __ldg(src + HW));
    const __nv_bfloat162 b = __floats2bfloat162_rn(__ldg(src + 2 * HW), 0.f);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&a);
    o.y = *reinterpret_cast<const uint32_t*>(&b);
    out[t] = o;
  }
}

constexpr int ST_NT = 512;                 // threads per CTA
constexpr int ST_PW = 264;                 // patch pixels per row: 2*127 + 7 taps + 1 pad tap, rounded up to even
constexpr int ST_ROWB = ST_PW * 8;         // bytes per patch row (4 bf16 per pixel)
constexpr int ST_KC = 4;                   // K = 7 rows x 32 (8 taps x 4 ch) = 224 -> 256 = 4 chunks of 64
constexpr int ST_A_BYTES = ST_KC * 128 * 128;
constexpr int ST_B_BYTES = ST_KC * 64 * 128;
constexpr int ST_PATCH_BYTES = 7 * ST_ROWB;
constexpr int ST_NPATCH = 3;               // patch ring: tiles i+1 and i+2 are in flight while tile i is consumed
constexpr int ST_SMEM = 1024 + 2 * ST_A_BYTES + ST_B_BYTES + ST_NPATCH * ST_PATCH_BYTES + 64 * 4 + 64;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t n = valid ? 16u : 0u;  // src-size 0 -> the 16 bytes (two pixels) are zero-filled (conv padding)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(n) : "memory");
}

// K layout k' = r*32 + s*4 + c (s = 0..7, c = 0..3; weights of s = 7 and c = 3 are zero): the 32 values of (pixel q,
// filter row r) are the 64 contiguous bytes patch[r][2q .. 2q+8), i.e. four aligned 16-byte chunks copied verbatim.
__global__ void __launch_bounds__(ST_NT, 1)
stem_conv_kernel(const uint2* __restrict__ x4, const float* __restrict__ w, const float* __restrict__ bn_gamma,
                 const float* __restrict__ bn_beta, const float* __restrict__ bn_mean,
                 const float* __restrict__ bn_var, float eps, __nv_bfloat16* __restrict__ out, int N, int H, int W,
                 int Ho, int Wo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                              // 2 buffers
  uint8_t* sB = smem + 2 * ST_A_BYTES;
  uint8_t* sPatch = sB + ST_B_BYTES;               // ST_NPATCH buffers (cp.async ring)
  float* sShift = reinterpret_cast<float*>(sPatch + ST_NPATCH * ST_PATCH_BYTES);
  uint64_t* mma_done = reinterpret_cast<uint64_t*>(sShift + 64);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_done + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qsegs = (Wo + 127) / 128;
  const int ntiles = N * Ho * qsegs;  // host checks that this fits 31 bits

  // ---- prologue: barriers, TMEM, B operand, shift, zeroed K padding (k' >= 224) of both A buffers
  if (tid == 0) {
    mbar_init(&mma_done[0], 1);
    mbar_init(&mma_done[1], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  for (int e = tid; e < 64 * 64 * ST_KC; e += ST_NT) {
    const int o = e / (64 * ST_KC), k = e - o * (64 * ST_KC);
    const int r = k >> 5, s = ((k & 31) >> 2) - 1, c = k & 3;   // tap slot 0 is the pad tap
    float v = 0.f;
    if (r < 7 && s >= 0 && c < 3) {
      const float sc = bn_gamma[o] / sqrtf(bn_var[o] + eps);
      v = w[((o * 3 + c) * 7 + r) * 7 + s] * sc;
    }
    const int kc = k >> 6, c16 = (k & 63) >> 3;
    *reinterpret_cast<__nv_bfloat16*>(sB + kc * 8192 + o * 128 + ((c16 ^ (o & 7)) << 4) + (k & 7) * 2) =
        __float2bfloat16_rn(v);
  }
  if (tid < 64) {
    const float sc = bn_gamma[tid] / sqrtf(bn_var[tid] + eps);
    sShift[tid] = bn_beta[tid] - bn_mean[tid] * sc;
  }
  for (int e = tid; e < 2 * 128 * 4; e += ST_NT) {  // k' in [224, 256): 16-byte chunks 4..7 of K-chunk 3
    const int b = e / 512, row = (e & 511) >> 2, c16 = 4 + (e & 3);
    *reinterpret_cast<uint4*>(sA + b * ST_A_BYTES + 3 * 16384 + row * 128 + ((c16 ^ (row & 7)) << 4)) =
        make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);

  // asynchronous patch load: 7 rows x 132 pixel PAIRS x 16 bytes, one cp.async per pair, zero-filled outside the image
  // (the patch starts at an even column and W is even, so a pair is never half inside).
  // Element e = tid + ST_NT * i of the patch <-> (row r, pair cp): tile independent, computed once.
  constexpr int ST_PAIRS = ST_PW / 2;
  constexpr int ST_LD = (7 * ST_PAIRS + ST_NT - 1) / ST_NT;
  int p_r[ST_LD], p_cx[ST_LD];
#pragma unroll
  for (int i = 0; i < ST_LD; ++i) {
    const int e = tid + i * ST_NT;
    p_r[i] = e < 7 * ST_PAIRS ? e / ST_PAIRS : -1;
    p_cx[i] = 2 * (e % ST_PAIRS);
  }
  auto load_patch = [&](int tile, int pb) {
    const int qs = tile % qsegs;
    const int np = tile / qsegs;       // n * Ho + p
    const int p = np % Ho, n = np / Ho;
    const int h0 = 2 * p - 3, w0 = 2 * (qs * 128) - 4;
    uint8_t* dst = sPatch + pb * ST_PATCH_BYTES;
    const uint2* img_n = x4 + (long long)n * H * W;
#pragma unroll
    for (int i = 0; i < ST_LD; ++i) {
      if (p_r[i] >= 0) {
        const int h = h0 + p_r[i], ww = w0 + p_cx[i];
        const bool ok = h >= 0 && h < H && ww >= 0 && ww + 1 < W;
        cp_async16(dst + p_r[i] * ST_ROWB + p_cx[i] * 8, img_n + (ok ? h * W + ww : 0), ok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // A-tile copy id = tid + ST_NT * j  <->  (filter row j, pixel ql, 16-byte piece i): only j changes inside the loop
  const int a_i = tid & 3, a_ql = (tid >> 2) & 127;
  const int a_src0 = a_ql * 16 + a_i * 16;
  auto epilogue = [&](int tile, int buf, uint32_t parity) {
    mbar_wait(&mma_done[buf], parity);
    tc_fence_after();
    const int qs = tile % qsegs;
    const long long rowbase = (long long)(tile / qsegs) * Wo;  // (n*Ho + p) * Wo
    const int q = qs * 128 + (warp & 3) * 32 + lane;
    const int ch0 = (warp >> 2) * 16;   // 16 warps: TMEM lane quadrant warp & 3, 16-channel column group warp >> 2
    uint32_t r0[16];
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + buf * 64 + ch0;
    tmem_ld16(taddr, r0);
    tmem_ld_wait();
    if (q < Wo) {
      uint4* dst = reinterpret_cast<uint4*>(out + (rowbase + q) * 64 + ch0);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const float4 s0 = *reinterpret_cast<const float4*>(sShift + ch0 + 8 * g);
        const float4 s1 = *reinterpret_cast<const float4*>(sShift + ch0 + 8 * g + 4);
        const float sh[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(__uint_as_float(r0[8 * g + j]) + sh[j], 0.f);
        uint4 o;
        __nv_bfloat162 h;
        h = __floats2bfloat162_rn(v[0], v[1]); o.x = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2bfloat162_rn(v[2], v[3]); o.y = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2bfloat162_rn(v[4], v[5]); o.z = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2bfloat162_rn(v[6], v[7]); o.w = *reinterpret_cast<uint32_t*>(&h);
        dst[g] = o;
      }
    }
    tc_fence_before();
  };

  int tile = blockIdx.x;
  // every thread commits exactly one group per tile slot (possibly empty), so wait_group counts stay uniform
  if (tile < ntiles) load_patch(tile, 0); else asm volatile("cp.async.commit_group;" ::: "memory");
  if (tile + (int)gridDim.x < ntiles) load_patch(tile + gridDim.x, 1); else asm volatile("cp.async.commit_group;" ::: "memory");
  int it = 0;
  int prev_tile = -1;
  for (; tile < ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    const int pbuf = it % ST_NPATCH;
    const int next2 = tile + 2 * gridDim.x;
    if (next2 < ntiles) load_patch(next2, (it + 2) % ST_NPATCH);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 2;" ::: "memory");  // this tile's patch has landed (two younger groups may be in flight)
    __syncthreads();  // ... for every thread's copies
    // ---- A tile: 128 rows x 28 sixteen-byte chunks copied from the patch (LDS.128 -> STS.128, conflict-free)
    uint8_t* a = sA + buf * ST_A_BYTES;
    const uint8_t* pt = sPatch + pbuf * ST_PATCH_BYTES;
    {
      // explicit shared-space accesses with 32-bit addresses (the generic LD.E / ST.E the compiler emitted for these
      // were 40 % of the kernel's stall samples); all seven loads are issued before the first store
      const uint32_t pt32 = smem_u32(pt) + a_src0, a32 = smem_u32(a) + a_ql * 128;
      uint4 v[7];
#pragma unroll
      for (int j = 0; j < 7; ++j)
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w)
                     : "r"(pt32 + j * ST_ROWB));
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int k = j * 32 + a_i * 8;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a32 + (k >> 6) * 16384 +
                                                                     ((((k & 63) >> 3) ^ (a_ql & 7)) << 4)),
                     "r"(v[j].x), "r"(v[j].y), "r"(v[j].z), "r"(v[j].w)
                     : "memory");
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_base = smem_u32(a), b_base = smem_u32(sB);
#pragma unroll
      for (int k = 0; k < 4 * ST_KC; ++k) {
        const uint64_t ad = make_sdesc(a_base + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
        const uint64_t bd = make_sdesc(b_base + (k >> 2) * 8192 + (k & 3) * 32, 16, 1024);
        umma_bf16(tmem_base + buf * 64, ad, bd, idesc, k != 0);
      }
      umma_commit(&mma_done[buf]);
    }
    if (prev_tile >= 0) epilogue(prev_tile, buf ^ 1, ((it - 1) >> 1) & 1);
    prev_tile = tile;
    // the next iteration's leading __syncthreads orders this tile's patch reads / TMEM reads before their buffers are reused
  }
  if (prev_tile >= 0) epilogue(prev_tile, (it - 1) & 1, ((it - 1) >> 1) & 1);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}


// ------------------------------------------------------------------------------------------------ direct-patch stem
// Same arithmetic, no A-tile assembly: with 8-byte pixels and stride 2, the K segment of (output pixel q, filter row r)
// is the 64 contiguous bytes patch[r][16 q .. 16 q + 64) — consecutive pixels start 16 bytes apart and OVERLAP. That is
// exactly the un-swizzled K-major canonical layout of tcgen05 (a core matrix = 8 rows x 16 bytes with rows 16 bytes
// apart) read with LBO = 16 (next 8-element K chunk) and SBO = 128 (next 8 rows): the MMA reads the patch row in
// place. Per tile: one elected thread TMA-loads the 7-row patch (one box of 7 rows x 18 blocks of 16 pixels, out-of-bounds
// rows / blocks zero-filled = the conv padding), another issues 14 UMMAs (7 filter rows x K 32), 8 epilogue warps turn
// the TMEM accumulator into the NHWC bf16 row segment (shift + ReLU -> swizzled slab -> TMA store, clipped at Wo).
// No CTA-wide barrier inside the tile loop (the old kernel spent 45 % of its stall samples at its two).
constexpr int S2_STAGES = 4;
// patch rows are loaded as 18 boxes of 16 pixels (128-byte TMA rows: the TMA unit's cost is per row, and 16-byte rows —
// one pixel pair — made the load the bottleneck), starting at the 16-pixel boundary 12 pixels before the patch
constexpr int S2_ROWB = 18 * 128;                                      // 2 304 bytes per patch row
constexpr int S2_LEAD = 12 * 8;                                        // bytes in front of the first patch pixel
constexpr int S2_PATCH = 7 * S2_ROWB;                                  // 16 128 bytes
constexpr int S2_PATCH_STRIDE = (S2_PATCH + 1023) / 1024 * 1024;       // 16 384
constexpr int S2_SMEM = 1024 + S2_STAGES * S2_PATCH_STRIDE + ST_B_BYTES + 2 * 16384 + 64 * 4 + 256;

struct alignas(64) Stem2Params {
  CUtensorMap tmX;   // NHWC4 image as [N][H][W/16][64 bf16 = 16 px]: box 64 x 18 x 7 x 1, no swizzle
  CUtensorMap tmY;   // output [N*Ho][Wo][64]: box 64 x 128 x 1, 128B swizzle
  const float* w;
  const float* bn_gamma;
  const float* bn_beta;
  const float* bn_mean;
  const float* bn_var;
  float eps;
  int Ho, qsegs, ntiles;
};

__device__ __forceinline__ uint64_t make_sdesc_noswz(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell); layout type 0 = no swizzle
  return d;
}
__device__ __forceinline__ void tma_load_4d_tile(const void* desc, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                                 int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* desc, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(desc)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__global__ void __launch_bounds__(384, 1) stem2_conv_kernel(const __grid_constant__ Stem2Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* const sPatch = smem;
  uint8_t* const sB = sPatch + S2_STAGES * S2_PATCH_STRIDE;
  uint8_t* const sOut = sB + ST_B_BYTES;              // 2 slabs (double-buffered TMA stores)
  float* const sShift = reinterpret_cast<float*>(sOut + 2 * 16384);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(sShift + 64);
  uint64_t* const full = bars;            // [S2_STAGES]
  uint64_t* const empty = bars + 4;       // [S2_STAGES]
  uint64_t* const tfull = bars + 8;       // [2]
  uint64_t* const tempty = bars + 10;     // [2]
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntiles = P.ntiles;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S2_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_mbar_init();
  } else if (warp == 2) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  // B operand [4 K-chunks][64 out][64 k] bf16, K-major, 128B swizzle, BN scale folded in; k' = r*32 + s'*4 + c
  for (int e = tid; e < 64 * 64 * ST_KC; e += blockDim.x) {
    const int o = e / (64 * ST_KC), k = e - o * (64 * ST_KC);
    const int r = k >> 5, s = ((k & 31) >> 2) - 1, c = k & 3;   // tap slot 0 is the pad tap
    float v = 0.f;
    if (r < 7 && s >= 0 && c < 3) {
      const float sc = P.bn_gamma[o] / sqrtf(P.bn_var[o] + P.eps);
      v = P.w[((o * 3 + c) * 7 + r) * 7 + s] * sc;
    }
    const int kc = k >> 6, c16 = (k & 63) >> 3;
    *reinterpret_cast<__nv_bfloat16*>(sB + kc * 8192 + o * 128 + ((c16 ^ (o & 7)) << 4) + (k & 7) * 2) =
        __float2bfloat16_rn(v);
  }
  if (tid < 64) {
    const float sc = P.bn_gamma[tid] / sqrtf(P.bn_var[tid] + P.eps);
    sShift[tid] = P.bn_beta[tid] - P.bn_mean[tid] * sc;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int qs = t % P.qsegs;
        const int np = t / P.qsegs;          // n * Ho + p
        const int p = np % P.Ho, n = np / P.Ho;
        const int h0 = 2 * p - 3, blk0 = qs * 16 - 1;   // first pixel 2*q0 - 4 lies 12 pixels into block 16*qs - 1
        mbar_wait(&empty[st], ph ^ 1);
        mbar_expect_tx(&full[st], S2_PATCH);
        tma_load_4d_tile(&P.tmX, &full[st], sPatch + st * S2_PATCH_STRIDE, 0, blk0, h0, n);
        if (++st == S2_STAGES) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      const uint32_t b_base = smem_u32(sB);
      int st = 0, it = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sPatch + st * S2_PATCH_STRIDE);
#pragma unroll
        for (int k = 0; k < 14; ++k) {      // k = 2 r + kk: filter row r, K elements [16 kk, 16 kk + 16) of its 32
          const int r = k >> 1, kk = k & 1;
          const uint64_t ad = make_sdesc_noswz(a_base + r * S2_ROWB + S2_LEAD + kk * 32, 16, 128);
          const uint64_t bd = make_sdesc(b_base + (k >> 2) * 8192 + (k & 3) * 32, 16, 1024);
          umma_bf16(tmem_base + acc * 64, ad, bd, idesc, k != 0);
        }
        umma_commit(&empty[st]);
        umma_commit(&tfull[acc]);
        if (++st == S2_STAGES) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // 8 epilogue warps: quadrant ew owns pixels [32 ew, 32 ew + 32); half eh owns channels [32 eh, 32 eh + 32)
    const int ew = warp & 3, eh = (warp - 4) >> 2;
    const int et = ew * 32 + lane;
    const bool leader = (warp == 4 && lane == 0);
    int it = 0;
    bool pending = false;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      uint8_t* const slab = sOut + (it & 1) * 16384;
      const uint32_t row = smem_u32(slab) + et * 128;
      // the store issued two tiles ago read this slab: drained before anyone overwrites it
      if (leader && pending) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * 64 + eh * 32;
      uint32_t ra[16], rb[16];
      tmem_ld16(taddr, ra);
      tmem_ld16(taddr + 16, rb);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
#pragma unroll
      for (int g = 0; g < 4; ++g) {        // four 16-byte slots of 8 channels
        const uint32_t* src = g < 2 ? ra + 8 * g : rb + 8 * (g - 2);
        const float* sh = sShift + eh * 32 + 8 * g;
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v0 = __uint_as_float(src[2 * j]) + sh[2 * j], v1 = __uint_as_float(src[2 * j + 1]) + sh[2 * j + 1];
          asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(o[j]) : "f"(v1), "f"(v0));
        }
        const uint32_t slot = (uint32_t)(eh * 4 + g);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((slot ^ (uint32_t)(et & 7)) << 4)), "r"(o[0]),
                     "r"(o[1]), "r"(o[2]), "r"(o[3])
                     : "memory");
      }
      fence_proxy_async();
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (leader) {
        const int qs = t % P.qsegs;
        tma_store_3d(&P.tmY, slab, 0, qs * 128, t / P.qsegs);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        pending = true;
      }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace dslb

using namespace dslb;

extern "C" size_t dslb_stem_workspace_bytes(int N, int H, int W) { return (size_t)N * H * W * 8; }

extern "C" int dslb_stem_conv(const float* img, const float* w, const float* bn_gamma, const float* bn_beta,
                              const float* bn_mean, const float* bn_var, float eps, void* workspace, void* out, int N,
                              int H, int W, void* stream) {
  DSLB_CHECK_ARG(img && w && bn_gamma && bn_beta && bn_mean && bn_var && workspace && out && N > 0 && H > 0 && W > 0,
                 "dslb_stem_conv: bad arguments");
  DSLB_CHECK_ARG(((uintptr_t)workspace % 16) == 0, "dslb_stem_conv: workspace must be 16-byte aligned");
  const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
  static bool attr_set = false;
  if (!attr_set) {
    DSLB_CHECK_CUDA(cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
    attr_set = true;
  }
  const long long HW = (long long)H * W;
  long long blocks = ((long long)N * HW + 255) / 256;
  if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
  img_to_nhwc4_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(img, (uint2*)workspace, N, HW);
  DSLB_CHECK_CUDA(cudaGetLastError());
  const long long ntiles = (long long)N * Ho * ((Wo + 127) / 128);
  DSLB_CHECK_ARG(ntiles < (1ll << 30) && (long long)H * W < (1ll << 31), "dslb_stem_conv: image too large");
  const int grid = (int)(ntiles < num_sms() ? ntiles : num_sms());
  static const bool old_stem = getenv("DSLB_OLD_STEM") != nullptr;
  if (!old_stem && W % 16 == 0 && ((uintptr_t)out % 16) == 0) {
    // direct-patch kernel: the image as [N][H][W/16][16 px x 4 ch] (128-byte innermost rows, no swizzle)
    Stem2Params P;
    memset(&P, 0, sizeof(P));
    const uint64_t xd[4] = {64, (uint64_t)(W / 16), (uint64_t)H, (uint64_t)N};
    const uint64_t xs[3] = {128, (uint64_t)W * 8, (uint64_t)H * W * 8};
    const uint32_t xb[4] = {64, 18, 7, 1};
    int rc = encode_tiled_bf16_swz(&P.tmX, workspace, 4, xd, xs, xb, 0);
    if (rc != DSLB_OK) return rc;
    const uint64_t yd[3] = {64, (uint64_t)Wo, (uint64_t)N * Ho};
    const uint64_t ys[2] = {128, (uint64_t)Wo * 128};
    const uint32_t yb[3] = {64, 128, 1};
    rc = encode_tiled_bf16(&P.tmY, out, 3, yd, ys, yb);
    if (rc != DSLB_OK) return rc;
    P.w = w;
    P.bn_gamma = bn_gamma;
    P.bn_beta = bn_beta;
    P.bn_mean = bn_mean;
    P.bn_var = bn_var;
    P.eps = eps;
    P.Ho = Ho;
    P.qsegs = (Wo + 127) / 128;
    P.ntiles = (int)ntiles;
    static bool attr2_set = false;
    if (!attr2_set) {
      DSLB_CHECK_CUDA(cudaFuncSetAttribute(stem2_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_SMEM));
      attr2_set = true;
    }
    stem2_conv_kernel<<<grid, 384, S2_SMEM, (cudaStream_t)stream>>>(P);
    DSLB_CHECK_CUDA(cudaGetLastError());
    return DSLB_OK;
  }
  stem_conv_kernel<<<grid, ST_NT, ST_SMEM, (cudaStream_t)stream>>>((const uint2*)workspace, w, bn_gamma, bn_beta, bn_mean,
                                                                  bn_var, eps, (__nv_bfloat16*)out, N, H, W, Ho, Wo);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}
