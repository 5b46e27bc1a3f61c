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
  stem_conv_kernel<<<grid, ST_NT, ST_SMEM, (cudaStream_t)stream>>>((const uint2*)workspace, w, bn_gamma, bn_beta, bn_mean,
                                                                  bn_var, eps, (__nv_bfloat16*)out, N, H, W, Ho, Wo);
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}
