// Weight-gradient implicit GEMM for sm_100a:  dW[tap][co][ci] += sum_pix dY[pix][co] * X[pix (+) tap][ci].
//
//   GEMM view     M = Cout (tiles of 128), N = Cin (tiles of <=256), K = N*Ho*Wo output pixels, one GEMM per
//                 filter tap; K is split over CTAs ("split-K") and partial tiles are reduced with fp32
//                 red.global.add into the packed gradient buffer.
//   A operand     dY tile [64 pixels][128 co] loaded by two tiled-TMA boxes of 64 channels: the pixel index is
//                 the MMA K dimension, so the operand is MN-major (128B swizzle, LBO = 8 KiB between the two
//                 64-channel blocks, SBO = 1 KiB between 8-pixel groups).
//   B operand     X tile [64 pixels (+) tap][bn ci] loaded by bn/64 im2col-TMA boxes, MN-major likewise.
//   Everything else (warp roles, mbarrier ring, TMEM double buffering) mirrors conv_igemm.cu.
#include <new>
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace dslb {

constexpr int WG_BK = 64;                     // pixels per pipeline stage
constexpr int WG_STAGES = 4;
constexpr int WG_BOX = 64 * 64 * 2;           // one [64 px][64 ch] box = 8 KiB
constexpr int WG_A_BYTES = 2 * WG_BOX;        // 128 co
constexpr int WG_B_BYTES_MAX = 4 * WG_BOX;    // up to 256 ci
constexpr int WG_STAGES2 = 6;                 // CTA pairs stage half of the X tile: 32 KiB stages
constexpr int WG_MAXST = 8;
constexpr int WG_SMEM = 1024 + WG_STAGES * (WG_A_BYTES + WG_B_BYTES_MAX) + 256;
static_assert(WG_STAGES2 * (WG_A_BYTES + WG_B_BYTES_MAX / 2) <= WG_STAGES * (WG_A_BYTES + WG_B_BYTES_MAX), "pair stages fit");
constexpr int WG_TMEM_COLS = 512;

struct alignas(128) WgSegDev {
  CUtensorMap tmDY;  // tiled 2-D [npix][ldy], box {64, 64}
  CUtensorMap tmX;   // im2col, 64 pixels x 64 channels
  float* dw;
  int npix, HoWo, Wo;
  int taps, S, stride, pad;
  int cout, ldy, dw_rows, cin;
  int m_tiles, n_tiles, bn;
  int ksplits, chunks, chunks_per_split;
  int job_begin;  // jobs of this seg: ((tap * m_tiles + mt) * n_tiles + nt) * ksplits + ks
};

struct alignas(128) WgParamsDev {
  WgSegDev seg[DSLB_MAX_SEGS];
  int nseg;
  int total_jobs;
  int cta2;   // 1: CTA pairs (cta_group::2): jobs 2j / 2j+1 are the two 128-channel Cout tiles of one (tap, Cin tile, K split);
              // they form one M = 256 MMA, each CTA staging its own dY tile and HALF of the X tile
};

struct WgJob {
  int si, tap, mt, nt, c_begin, c_end;
};

__device__ __forceinline__ WgJob wg_decode(const WgParamsDev* P, int job) {
  int si = 0;
  const int nseg = P->nseg;
  while (si + 1 < nseg && job >= P->seg[si + 1].job_begin) ++si;
  const WgSegDev& sg = P->seg[si];
  int j = job - sg.job_begin;
  WgJob o;
  o.si = si;
  int ks;
  if (P->cta2) {   // (tap, m pair, nt, ks, m parity): the parity is the CTA's rank in its pair
    const int lo = j & 1;
    j >>= 1;
    ks = j % sg.ksplits;
    j /= sg.ksplits;
    o.nt = j % sg.n_tiles;
    j /= sg.n_tiles;
    const int mp = j % (sg.m_tiles >> 1);
    o.tap = j / (sg.m_tiles >> 1);
    o.mt = 2 * mp + lo;
  } else {
    ks = j % sg.ksplits;
    j /= sg.ksplits;
    o.nt = j % sg.n_tiles;
    j /= sg.n_tiles;
    o.mt = j % sg.m_tiles;
    o.tap = j / sg.m_tiles;
  }
  o.c_begin = ks * sg.chunks_per_split;
  o.c_end = min(o.c_begin + sg.chunks_per_split, sg.chunks);
  return o;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// parameter block by value in the constant bank (see conv_igemm.cu)
template <bool CTA2>
__device__ __forceinline__ void conv_wgrad_body(const WgParamsDev* P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NST = CTA2 ? WG_STAGES2 : WG_STAGES;
  constexpr int B_STAGE = CTA2 ? WG_B_BYTES_MAX / 2 : WG_B_BYTES_MAX;
  uint8_t* sA = smem;
  uint8_t* sB = smem + NST * WG_A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + WG_STAGES * (WG_A_BYTES + WG_B_BYTES_MAX));
  uint64_t* empty = full + WG_MAXST;
  uint64_t* tfull = empty + WG_MAXST;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], CTA2 ? 8 : 4);
    }
    fence_mbar_init();
  } else if (warp == 2) {
    if (CTA2) {
      tmem_alloc2(tmem_slot, WG_TMEM_COLS);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, WG_TMEM_COLS);
      tmem_relinquish();
    }
  }
  if (CTA2) {
    __syncwarp();
    cluster_sync_all();
  }
  const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total = P->total_jobs;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int job = blockIdx.x; job < total; job += gridDim.x) {
        const WgJob jb = wg_decode(P, job);
        const WgSegDev& sg = P->seg[jb.si];
        const int r = jb.tap / sg.S;
        const int s = jb.tap - r * sg.S;
        const int co0 = jb.mt * 128;
        const int a_boxes = (co0 + 64 < sg.ldy) ? 2 : 1;  // second 64-channel block may not exist
        const int b_boxes = CTA2 ? sg.bn / 128 : sg.bn / 64;   // a pair's CTA stages half of the Cin tile
        const uint32_t tx = (CTA2 ? 2 : 1) * (a_boxes + b_boxes) * WG_BOX;
        const int ci0 = jb.nt * sg.bn + (CTA2 ? (int)cta_rank * (sg.bn / 2) : 0);
        for (int kc = jb.c_begin; kc < jb.c_end; ++kc) {
          const int pix0 = kc * WG_BK;
          const int n_img = pix0 / sg.HoWo;
          const int rem = pix0 - n_img * sg.HoWo;
          const int p = rem / sg.Wo;
          const int q = rem - p * sg.Wo;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* a = sA + stage * WG_A_BYTES;
          uint8_t* b = sB + stage * B_STAGE;
          if (CTA2) {
            if (cta_rank == 0) mbar_expect_tx(&full[stage], tx);
            const uint32_t bar = mapa_rank(smem_u32(&full[stage]), 0);
            for (int i = 0; i < a_boxes; ++i) tma_load_2d_cta2(&sg.tmDY, bar, a + i * WG_BOX, co0 + 64 * i, pix0);
            for (int i = 0; i < b_boxes; ++i)
              tma_load_im2col_4d_cta2(&sg.tmX, bar, b + i * WG_BOX, ci0 + 64 * i, q * sg.stride - sg.pad,
                                      p * sg.stride - sg.pad, n_img, (uint16_t)s, (uint16_t)r);
          } else {
            mbar_expect_tx(&full[stage], tx);
            for (int i = 0; i < a_boxes; ++i) tma_load_2d(&sg.tmDY, &full[stage], a + i * WG_BOX, co0 + 64 * i, pix0);
            for (int i = 0; i < b_boxes; ++i)
              tma_load_im2col_4d(&sg.tmX, &full[stage], b + i * WG_BOX, ci0 + 64 * i, q * sg.stride - sg.pad,
                                 p * sg.stride - sg.pad, n_img, (uint16_t)s, (uint16_t)r);
          }
          if (++stage == NST) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if ((!CTA2 || cta_rank == 0) && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int job = blockIdx.x; job < total; job += gridDim.x, ++it) {
        const WgJob jb = wg_decode(P, job);
        const WgSegDev& sg = P->seg[jb.si];
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        const uint32_t idesc = make_idesc_bf16(CTA2 ? 256 : 128, sg.bn, 1, 1);
        for (int kc = jb.c_begin; kc < jb.c_end; ++kc) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + stage * WG_A_BYTES);
          const uint32_t b_base = smem_u32(sB + stage * B_STAGE);
#pragma unroll
          for (int k = 0; k < WG_BK / 16; ++k) {
            const uint64_t ad = make_sdesc(a_base + k * 2048, WG_BOX, 1024);
            const uint64_t bd = make_sdesc(b_base + k * 2048, WG_BOX, 1024);
            if (CTA2) umma_bf16_cta2(d_tmem, ad, bd, idesc, (kc > jb.c_begin) || (k != 0));
            else umma_bf16(d_tmem, ad, bd, idesc, (kc > jb.c_begin) || (k != 0));
          }
          if (CTA2) umma_commit_cta2(&empty[stage]);
          else umma_commit(&empty[stage]);
          if (++stage == NST) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (CTA2) umma_commit_cta2(&tfull[acc]);
        else umma_commit(&tfull[acc]);
      }
    }
  } else if (warp >= 4) {
    const int ew = warp & 3;
    int it = 0;
    for (int job = blockIdx.x; job < total; job += gridDim.x, ++it) {
      const WgJob jb = wg_decode(P, job);
      const WgSegDev& sg = P->seg[jb.si];
      const int acc = it & 1;
      const int co = jb.mt * 128 + ew * 32 + lane;
      const bool valid = co < sg.cout && jb.c_end > jb.c_begin;
      float* __restrict__ dst =
          sg.dw + ((long long)jb.tap * sg.dw_rows + co) * sg.cin + jb.nt * sg.bn;
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * 256;
      for (int c0 = 0; c0 < sg.bn; c0 += 16) {
        uint32_t rr[16];
        tmem_ld16(taddr + c0, rr);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            red_add_v4(dst + c0 + j, __uint_as_float(rr[j]), __uint_as_float(rr[j + 1]),
                       __uint_as_float(rr[j + 2]), __uint_as_float(rr[j + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTA2) mbar_arrive_cluster(mapa_rank(smem_u32(&tempty[acc]), 0));
        else mbar_arrive(&tempty[acc]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CTA2) {
    __syncwarp();
    cluster_sync_all();
  }
  if (warp == 2) {
    tc_fence_after();
    if (CTA2) tmem_dealloc2(tmem_base, WG_TMEM_COLS);
    else tmem_dealloc(tmem_base, WG_TMEM_COLS);
  }
}

__global__ void __launch_bounds__(256, 1) conv_wgrad_kernel(const __grid_constant__ WgParamsDev PP) {
  conv_wgrad_body<false>(&PP);
}
__global__ void __launch_bounds__(256, 1) conv_wgrad_cta2_kernel(const __grid_constant__ WgParamsDev PP) {
  conv_wgrad_body<true>(&PP);
}

}  // namespace dslb

using namespace dslb;

struct dslb_wgrad_plan {
  WgParamsDev* dev = nullptr;  // HOST copy, passed by value at launch
  int total_jobs = 0;
  double flops = 0.0;
  int cta2 = 0;
};

extern "C" int dslb_wgrad_plan_create(const dslb_wgrad_seg_t* segs, int nseg, dslb_wgrad_plan_t** out) {
  DSLB_CHECK_ARG(segs && out, "dslb_wgrad_plan_create: null argument");
  DSLB_CHECK_ARG(nseg >= 1 && nseg <= DSLB_MAX_SEGS, "dslb_wgrad_plan_create: nseg %d not in [1,%d]", nseg,
                 DSLB_MAX_SEGS);
  WgParamsDev* h = new (std::nothrow) WgParamsDev();
  if (!h) {
    set_error("out of host memory");
    return DSLB_ENOMEM;
  }
  memset(h, 0, sizeof(*h));
  // First pass: geometry + the number of (tap, m, n) tile groups and K chunks, to size the split-K.
  long long base_jobs = 0, total_chunk_jobs = 0;
  for (int i = 0; i < nseg; ++i) {
    const dslb_wgrad_seg_t& s = segs[i];
    WgSegDev& d = h->seg[i];
#define SEG_CHECK(cond, ...)  \
  if (!(cond)) {              \
    set_error(__VA_ARGS__);   \
    delete h;                 \
    return DSLB_EINVAL;       \
  }
    SEG_CHECK(s.x && s.dy && s.dw, "wgrad seg %d: null x/dy/dw", i);
    SEG_CHECK(s.Cin > 0 && s.Cin % 64 == 0, "wgrad seg %d: Cin=%d must be a multiple of 64", i, s.Cin);
    SEG_CHECK(s.ldy >= s.Cout && s.ldy % 64 == 0, "wgrad seg %d: ldy=%d must be a multiple of 64 >= Cout", i,
              s.ldy);
    SEG_CHECK(s.dw_rows >= s.Cout, "wgrad seg %d: dw_rows < Cout", i);
    SEG_CHECK(s.R >= 1 && s.S >= 1 && s.R <= 7 && s.S <= 7 && s.stride >= 1 && s.stride <= 2 && s.pad >= 0,
              "wgrad seg %d: unsupported filter", i);
    SEG_CHECK(((uintptr_t)s.x % 16) == 0 && ((uintptr_t)s.dy % 16) == 0 && ((uintptr_t)s.dw % 16) == 0,
              "wgrad seg %d: pointers must be 16-byte aligned", i);
    const int Ho = (s.H + 2 * s.pad - s.R) / s.stride + 1;
    const int Wo = (s.W + 2 * s.pad - s.S) / s.stride + 1;
    SEG_CHECK(Ho > 0 && Wo > 0, "wgrad seg %d: empty output", i);
#undef SEG_CHECK
    d.dw = s.dw;
    d.npix = s.N * Ho * Wo;
    d.HoWo = Ho * Wo;
    d.Wo = Wo;
    d.taps = s.R * s.S;
    d.S = s.S;
    d.stride = s.stride;
    d.pad = s.pad;
    d.cout = s.Cout;
    d.ldy = s.ldy;
    d.dw_rows = s.dw_rows;
    d.cin = s.Cin;
    d.bn = s.Cin >= 256 ? 256 : s.Cin;  // Cin in {64,128,192,256,512,...}: tile of 256 or the whole thing
    if (s.Cin % d.bn != 0) d.bn = 64;
    d.m_tiles = cdiv(s.Cout, 128);
    d.n_tiles = s.Cin / d.bn;
    d.chunks = cdiv(d.npix, WG_BK);
    base_jobs += (long long)d.taps * d.m_tiles * d.n_tiles;
    total_chunk_jobs += (long long)d.taps * d.m_tiles * d.n_tiles * d.chunks;
  }
  // Aim for ~2 jobs per SM; never split below 8 chunks (512 pixels) per job.
  const int sms = num_sms();
  long long target = total_chunk_jobs / (2LL * sms);
  if (target < 8) target = 8;
  int jobs = 0;
  double flops = 0.0;
  for (int i = 0; i < nseg; ++i) {
    const dslb_wgrad_seg_t& s = segs[i];
    WgSegDev& d = h->seg[i];
    d.ksplits = (int)((d.chunks + target - 1) / target);
    if (d.ksplits < 1) d.ksplits = 1;
    d.chunks_per_split = cdiv(d.chunks, d.ksplits);
    d.ksplits = cdiv(d.chunks, d.chunks_per_split);
    d.job_begin = jobs;
    jobs += d.taps * d.m_tiles * d.n_tiles * d.ksplits;
    int rc = encode_im2col_bf16(&d.tmX, s.x, s.N, s.H, s.W, s.Cin, s.R, s.S, s.stride, s.pad, WG_BK);
    if (rc != DSLB_OK) {
      delete h;
      return rc;
    }
    const uint64_t yd[2] = {(uint64_t)s.ldy, (uint64_t)d.npix};
    const uint64_t ys[1] = {(uint64_t)s.ldy * 2};
    const uint32_t yb[2] = {64, 64};
    rc = encode_tiled_bf16(&d.tmDY, s.dy, 2, yd, ys, yb);
    if (rc != DSLB_OK) {
      delete h;
      return rc;
    }
    flops += 2.0 * (double)d.npix * s.Cout * s.Cin * s.R * s.S;
  }
  (void)base_jobs;
  h->nseg = nseg;
  h->total_jobs = jobs;
  // CTA pairs: every segment's Cout is a multiple of 256 (two 128-channel tiles = one M = 256 MMA) and its Cin tile splits
  // into two halves of whole 64-channel boxes. DSLB_CTA2=0 switches it off.
  bool cta2 = !(getenv("DSLB_CTA2") != nullptr && getenv("DSLB_CTA2")[0] == '0');
  for (int i = 0; i < nseg && cta2; ++i) {
    const WgSegDev& d = h->seg[i];
    if (d.cout % 256 != 0 || d.bn % 128 != 0 || d.ldy < d.cout) cta2 = false;
  }
  h->cta2 = cta2 ? 1 : 0;

  dslb_wgrad_plan* plan = new (std::nothrow) dslb_wgrad_plan();
  if (!plan) {
    delete h;
    set_error("out of host memory");
    return DSLB_ENOMEM;
  }
  plan->dev = h;
  plan->total_jobs = jobs;
  plan->flops = flops;
  plan->cta2 = cta2 ? 1 : 0;
  static bool attr_set = false;
  if (!attr_set) {
    DSLB_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_cta2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    DSLB_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    attr_set = true;
  }
  *out = plan;
  return DSLB_OK;
}

extern "C" int dslb_wgrad_plan_run(const dslb_wgrad_plan_t* plan, void* stream) {
  DSLB_CHECK_ARG(plan && plan->dev, "dslb_wgrad_plan_run: null plan");
  const int grid = plan->total_jobs < num_sms() ? plan->total_jobs : num_sms();
  if (plan->cta2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(grid & ~1));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = WG_SMEM;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    DSLB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_wgrad_cta2_kernel, *plan->dev));
    DSLB_CHECK_CUDA(cudaGetLastError());
    return DSLB_OK;
  }
  DSLB_CHECK_CUDA(launch_pdl(conv_wgrad_kernel, dim3(grid), dim3(256), WG_SMEM, (cudaStream_t)stream, *plan->dev));
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" void dslb_wgrad_plan_destroy(dslb_wgrad_plan_t* plan) {
  if (!plan) return;
  delete plan->dev;
  delete plan;
}

extern "C" double dslb_wgrad_plan_flops(const dslb_wgrad_plan_t* plan) { return plan ? plan->flops : 0.0; }
