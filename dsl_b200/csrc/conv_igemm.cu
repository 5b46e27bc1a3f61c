// Implicit-GEMM convolution for sm_100a: NHWC bf16 activations, fp32 accumulation in TMEM.
//
//   GEMM view     M = N*Ho*Wo output pixels (tiles of 128 consecutive pixels), N = Cout (tiles of <=256),
//                 K = R*S*Cin walked as (tap, 64-channel chunk).
//   A operand     im2col-mode TMA: one cp.async.bulk.tensor.4d...im2col per (tap, chunk) lands a 128-pixel x
//                 64-channel K-major tile (128B swizzle) in shared memory; the conv halo / stride / image
//                 wrap-around is done by the TMA unit (OOB -> 0), there is no materialised im2col.
//   B operand     tiled TMA on the packed weights [tap][Cout_pad][Cin].
//   MMA           tcgen05.mma.cta_group::1.kind::f16, M=128, N=bn, K=16, issued by one elected thread;
//                 accumulators double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps
//                 the MMAs of tile i+1.
//   Warp roles    warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-11 epilogue (two
//                 warpgroups, one per 64-channel slab): tcgen05.ld 32x32b -> scale/shift/residual/ReLU/mask/
//                 GroupNorm statistics -> 128B-swizzled smem staging tile -> TMA store (bf16) or direct stores.
//   Scheduling    persistent: grid = min(#tiles, #SMs), static round-robin over the tile list of up to
//                 DSLB_MAX_SEGS independent convs (e.g. 5 FPN levels x 2 FCOSHead towers in one launch).
#include <new>
#include <stdlib.h>

#include "common.h"
#include "conv_epilogue.cuh"
#include "conv_halo.h"
#include "ptx.cuh"

namespace dslb {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int STAGES = 4;       // operand stages of the widest (256-column) tiles
constexpr int MAX_STAGES = 8;   // narrower B tiles leave room for a deeper ring
constexpr int A_BYTES = BM * BK * 2;        // 16 KiB
constexpr int B_BYTES_MAX = 256 * BK * 2;   // 32 KiB
constexpr int BAR_BYTES = 256;
constexpr int STAT_BYTES = 4 * 32 * 2 * 4;  // GroupNorm partial sums [4 warps][32 halves][2] fp32
constexpr int OUT_BYTES = 2 * BM * 128;  // output staging: two 128B-swizzled [128 px][64 ch] bf16 slabs
constexpr int CONV_SMEM = 1024 + STAGES * (A_BYTES + B_BYTES_MAX) + OUT_BYTES + BAR_BYTES + STAT_BYTES;
constexpr int TMEM_COLS = 512;
constexpr int IDENT_BYTES = 64 * 128;
static_assert(1024 + 3 * (A_BYTES + B_BYTES_MAX) + 2 * OUT_BYTES + IDENT_BYTES + BAR_BYTES + STAT_BYTES <= CONV_SMEM,
              "3-stage layout (double slabs + identity tile) must fit the 4-stage budget");
static_assert(CONV_SMEM <= 232448, "more than 227 KiB of shared memory");

struct alignas(128) ConvSegDev {
  CUtensorMap tmA;
  CUtensorMap tmB;
  CUtensorMap tmY;  // output store map (staged epilogue only)
  CUtensorMap tmAux;  // residual (aux_kind 1) or ReLU mask (aux_kind 2) tile load map, same geometry as tmY
  CUtensorMap tmRes;  // residual tiles [128 px][64 ch] for the identity-MMA path (res_mma)
  void* y;
  const void* residual;
  const void* relu_mask;
  const float* scale;
  const float* shift;
  double* stats;             // GroupNorm sums of this output (forward), or of the GroupNorm backward below (gnb_x set)
  const void* gnb_x;         // GroupNorm BACKWARD sums: pre-norm map [npix][cout] bf16 of the norm this output flows into
  const float4* gnb_mr;      // [N][groups] (mean, rstd, -, -) of that norm
  const float* gnb_gamma;    // [cout]
  const float* gnb_beta;     // [cout]
  int npix, HoWo, Wo;
  int m_tiles, n_tiles, tile_begin;
  int cin_chunks, taps, S, stride, pad;
  int cout, bn, ldc;
  int out_fp32, relu_nch, cpg, groups;
  int scatter2, Hs, Ws;
  int staged;  // 1: bf16 output goes through the smem staging tile + TMA store
  int aux_kind;  // 0: none; 1: residual, 2: ReLU mask, 3: gnb_x tile arrive by TMA in the staging slab, consumed in place
  int res_mma;   // 1: the residual is added on the tensor core: after the K loop its [128 px][64 ch] tiles travel
                 // through the operand ring as A tiles and are multiplied by a shared-memory 64x64 identity into
                 // the matching 64 accumulator columns, so the epilogue never sees it (`residual` is null then)
};

struct alignas(128) ConvParamsDev {
  ConvSegDev seg[DSLB_MAX_SEGS];
  int nseg;
  int total_tiles;
  int nstages;  // operand pipeline depth: as many (A tile + widest B tile of the plan) stages as fit, <= MAX_STAGES
  int b_stride; // bytes between the B tiles of consecutive stages (= widest tile of the plan x 128 B)
  int nbuf;     // staging slabs per epilogue warpgroup: 1, or 2 with TMA-prefetched residual / mask tiles
  int pair;       // 1: every CTA works on PAIRS of consecutive 128-pixel tiles (same n-tile) that share the B operand:
                  // a stage holds two A tiles + one B tile, each k-iteration issues MMAs into both TMEM accumulators
                  // (no double buffering) -> 64 instead of 96 operand bytes per MMA clock for 256-wide tiles
  int has_ident;  // some segment uses res_mma: 8 KiB identity tile after the staging slabs (3-stage layout only)
  int any_aux;  // some segment brings its residual / mask tiles in by TMA (nbuf == 2); with nbuf == 2 and no aux the
                // second slab double-buffers the TMA stores instead
  int cta2;     // 1: launched as CTA PAIRS (cluster of 2, conv_igemm_cta2_kernel): two consecutive 128-pixel tiles of one
                // segment form ONE tcgen05.mma.cta_group::2 of M = 256; each CTA keeps its own A tile and HALF of the
                // 256-row weight tile (b_stride = 16 KiB), so a k-iteration feeds 32 KiB per SM instead of 48 KiB
};

__device__ __forceinline__ int find_seg(const ConvParamsDev* P, int tile) {
  int si = 0;
  const int nseg = P->nseg;
  while (si + 1 < nseg && tile >= P->seg[si + 1].tile_begin) ++si;
  return si;
}

// scale/shift -> residual -> ReLU(c < relu_nch) -> mask, on one 16-channel chunk of one output pixel.
// Fast path (all 16 channels real): vector loads, no per-element guards.
// aux_kind 1 / 2: the residual / mask values of this chunk were brought in by TMA and are passed in (a0, a1).
__device__ __forceinline__ void epilogue_math(const ConvSegDev& sg, const uint32_t (&rr)[16], float (&v)[16],
                                              int cb, bool fullchunk, bool valid, long long row, int aux_kind,
                                              const uint4& a0, const uint4& a1) {
  const float* __restrict__ scale = sg.scale;
  const float* __restrict__ shift = sg.shift;
  const __nv_bfloat16* __restrict__ resid = reinterpret_cast<const __nv_bfloat16*>(sg.residual);
  const __nv_bfloat16* __restrict__ rmask = reinterpret_cast<const __nv_bfloat16*>(sg.relu_mask);
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rr[j]);
  if (fullchunk) {
    if (scale) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(scale + cb) + j);
        v[4 * j] *= t.x;
        v[4 * j + 1] *= t.y;
        v[4 * j + 2] *= t.z;
        v[4 * j + 3] *= t.w;
      }
    }
    if (shift) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(shift + cb) + j);
        v[4 * j] += t.x;
        v[4 * j + 1] += t.y;
        v[4 * j + 2] += t.z;
        v[4 * j + 3] += t.w;
      }
    }
    if (aux_kind == 1) {
      const uint32_t w[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[2 * j] += bf16_lo(w[j]);
        v[2 * j + 1] += bf16_hi(w[j]);
      }
    } else if (valid && resid) {
      const uint4* rp = reinterpret_cast<const uint4*>(resid + row + cb);
      const uint4 r0 = rp[0], r1 = rp[1];
      const uint32_t w[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[2 * j] += bf16_lo(w[j]);
        v[2 * j + 1] += bf16_hi(w[j]);
      }
    }
    if (cb + 16 <= sg.relu_nch) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (cb < sg.relu_nch) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (cb + j < sg.relu_nch) v[j] = fmaxf(v[j], 0.f);
    }
    if (aux_kind == 2) {
      const uint32_t w[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (!(bf16_lo(w[j]) > 0.f)) v[2 * j] = 0.f;
        if (!(bf16_hi(w[j]) > 0.f)) v[2 * j + 1] = 0.f;
      }
    } else if (valid && rmask) {
      const uint4* mp = reinterpret_cast<const uint4*>(rmask + row + cb);
      const uint4 m0 = mp[0], m1 = mp[1];
      const uint32_t w[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (!(bf16_lo(w[j]) > 0.f)) v[2 * j] = 0.f;
        if (!(bf16_hi(w[j]) > 0.f)) v[2 * j + 1] = 0.f;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int c = min(cb + j, sg.cout - 1);
      if (scale) v[j] *= __ldg(scale + c);
      if (shift) v[j] += __ldg(shift + c);
      if (valid && resid && cb + j < sg.cout) v[j] += __bfloat162float(resid[row + cb + j]);
      if (cb + j < sg.relu_nch) v[j] = fmaxf(v[j], 0.f);
      if (valid && rmask && cb + j < sg.cout && !(__bfloat162float(rmask[row + cb + j]) > 0.f)) v[j] = 0.f;
    }
  }
}

// Shared-memory carve-up + pipeline barriers, common to both kernels below.
struct ConvSmem {
  uint8_t *sA, *sB, *sOut, *sIdent;
  uint64_t *full, *empty, *tfull, *tempty, *auxfull;
  uint32_t* tmem_slot;
  float* sStat;
  int nst, a_stride, b_stride;
};

__device__ __forceinline__ ConvSmem conv_carve(const ConvParamsDev* P, uint8_t* smem_raw) {
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // Layouts sharing one budget: as many (A + widest B) stages as fit next to one staging slab per epilogue warpgroup
  // (compute-bound plans) or next to four slabs (double-buffered slabs of the 2-group kernel / one slab for each of the
  // 4 groups of the fast kernel) plus, for res_mma plans, the identity tile.
  ConvSmem S;
  S.nst = P->nstages;
  S.b_stride = P->b_stride;
  S.a_stride = P->pair ? 2 * A_BYTES : A_BYTES;
  S.sA = smem;
  S.sB = smem + S.nst * S.a_stride;
  S.sOut = smem + S.nst * (S.a_stride + S.b_stride);
  S.sIdent = S.sOut + P->nbuf * OUT_BYTES;  // [64][64] bf16 identity, K-major, 128B swizzle (res_mma plans)
  uint8_t* sBar = S.sIdent + (P->has_ident ? IDENT_BYTES : 0);
  S.full = reinterpret_cast<uint64_t*>(sBar);
  S.empty = S.full + MAX_STAGES;
  S.tfull = S.empty + MAX_STAGES;
  S.tempty = S.tfull + 2;
  S.auxfull = S.tempty + 2;  // [warpgroup][slab buffer] (2-group kernel) / [warpgroup] (fast kernel)
  S.tmem_slot = reinterpret_cast<uint32_t*>(S.auxfull + 4);
  S.sStat = reinterpret_cast<float*>(sBar + BAR_BYTES);
  return S;
}

// barrier init, TMEM allocation, identity tile; returns the TMEM base address. Ends with a CTA-wide barrier.
template <bool CTA2 = false>
__device__ __forceinline__ uint32_t conv_prologue(const ConvParamsDev* P, const ConvSmem& S, int epi_warps) {
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();  // the next kernel of the stream may set itself up behind this one
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < MAX_STAGES; ++i) {
      mbar_init(&S.full[i], 1);
      mbar_init(&S.empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&S.tfull[i], 1);
      mbar_init(&S.tempty[i], CTA2 ? 2 * epi_warps : epi_warps);   // pair: the peer's epilogue warps arrive here too
    }
    for (int i = 0; i < 4; ++i) mbar_init(&S.auxfull[i], 1);
    fence_mbar_init();
  } else if (warp == 2) {
    if (CTA2) {
      tmem_alloc2(S.tmem_slot, TMEM_COLS);
      tmem_relinquish2();
    } else {
      tmem_alloc(S.tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  if (CTA2) {   // the peer's barriers are initialised before anything can arrive on them remotely
    __syncwarp();
    cluster_sync_all();
  }
  if (P->has_ident) {
    // row n holds 1.0 at k == n: 16-byte slot (n >> 3) ^ (n & 7) of the 128-byte row, element n & 7 inside it
    for (int i = threadIdx.x; i < IDENT_BYTES / 16; i += blockDim.x) {
      const int n = i >> 3, slot = i & 7;
      uint4 z = make_uint4(0u, 0u, 0u, 0u);
      if (slot == ((n >> 3) ^ (n & 7))) {
        const uint32_t one = 0x3f80u << ((n & 1) * 16);
        const int w = (n & 7) >> 1;
        if (w == 0) z.x = one;
        else if (w == 1) z.y = one;
        else if (w == 2) z.z = one;
        else z.w = one;
      }
      reinterpret_cast<uint4*>(S.sIdent)[i] = z;
    }
    fence_proxy_async();
  }
  pdl_wait();  // everything above overlapped the predecessor's tail; from here on its outputs are read
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *S.tmem_slot;
}

// ------------------------------------------------------------------ TMA producer (one elected thread of warp 0)
__device__ __forceinline__ void conv_producer(const ConvParamsDev* P, const ConvSmem& S) {
  const int total = P->total_tiles, nst = S.nst;
  const int nsub = P->pair ? 2 : 1;   // 128-pixel tiles per work item
  int stage = 0;
  uint32_t phase = 0;
  for (int tile = blockIdx.x * nsub; tile < total; tile += gridDim.x * nsub) {
    const ConvSegDev& sg = P->seg[find_seg(P, tile)];
    const int tl = tile - sg.tile_begin;
    const int nt = tl / sg.m_tiles;
    const int mt = tl - nt * sg.m_tiles;
    int n_img[2], cw[2], ch[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int pix0 = (mt + h) * BM;   // a padding tile (pix0 >= npix) reads out-of-bounds images: zero fill
      n_img[h] = pix0 / sg.HoWo;
      const int rem = pix0 - n_img[h] * sg.HoWo;
      const int p = rem / sg.Wo;
      const int q = rem - p * sg.Wo;
      cw[h] = q * sg.stride - sg.pad;
      ch[h] = p * sg.stride - sg.pad;
    }
    const uint32_t tx = nsub * A_BYTES + sg.bn * (BK * 2);
    for (int tap = 0; tap < sg.taps; ++tap) {
      const int r = tap / sg.S;
      const int s = tap - r * sg.S;
      for (int kc = 0; kc < sg.cin_chunks; ++kc) {
        mbar_wait(&S.empty[stage], phase ^ 1);
        mbar_expect_tx(&S.full[stage], tx);
        for (int h = 0; h < nsub; ++h)
          tma_load_im2col_4d(&sg.tmA, &S.full[stage], S.sA + stage * S.a_stride + h * A_BYTES, kc * BK, cw[h], ch[h],
                             n_img[h], (uint16_t)s, (uint16_t)r);
        tma_load_3d(&sg.tmB, &S.full[stage], S.sB + stage * S.b_stride, kc * BK, nt * sg.bn, tap);
        if (++stage == nst) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    if (sg.res_mma) {
      for (int j = 0; j < sg.bn / 64; ++j) {
        mbar_wait(&S.empty[stage], phase ^ 1);
        mbar_expect_tx(&S.full[stage], nsub * A_BYTES);
        for (int h = 0; h < nsub; ++h)
          tma_load_2d(&sg.tmRes, &S.full[stage], S.sA + stage * S.a_stride + h * A_BYTES, nt * sg.bn + 64 * j,
                      (mt + h) * BM);
        if (++stage == nst) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ MMA issuer (one elected thread of warp 1)
__device__ __forceinline__ void conv_mma(const ConvParamsDev* P, const ConvSmem& S, uint32_t tmem_base) {
  const int total = P->total_tiles, nst = S.nst;
  const int nsub = P->pair ? 2 : 1;
  int stage = 0;
  uint32_t phase = 0;
  int it = 0;   // work items done by this CTA
  for (int tile = blockIdx.x * nsub; tile < total; tile += gridDim.x * nsub, ++it) {
    const ConvSegDev& sg = P->seg[find_seg(P, tile)];
    const int kiters = sg.taps * sg.cin_chunks;
    // single mode: accumulators alternate per tile (the epilogue of tile i overlaps the MMAs of tile i+1);
    // pair mode: sub-tile h owns accumulator h, both must have been drained
    const int acc0 = nsub == 2 ? 0 : (it & 1);
    const uint32_t par = nsub == 2 ? ((it & 1) ^ 1) : (((it >> 1) & 1) ^ 1);
    mbar_wait(&S.tempty[acc0], par);
    if (nsub == 2) mbar_wait(&S.tempty[1], par);
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + acc0 * 256;
    const uint32_t idesc = make_idesc_bf16(BM, sg.bn, 0, 0);
    for (int ki = 0; ki < kiters; ++ki) {
      mbar_wait(&S.full[stage], phase);
      tc_fence_after();
      const uint32_t a_base = smem_u32(S.sA + stage * S.a_stride);
      const uint32_t b_base = smem_u32(S.sB + stage * S.b_stride);
      for (int h = 0; h < nsub; ++h) {
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t ad = make_sdesc(a_base + h * A_BYTES + k * 32, 16, 1024);
          const uint64_t bd = make_sdesc(b_base + k * 32, 16, 1024);
          umma_bf16(d_tmem + h * 256, ad, bd, idesc, (ki | k) != 0);
        }
      }
      umma_commit(&S.empty[stage]);
      if (++stage == nst) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (sg.res_mma) {
      const uint32_t idesc64 = make_idesc_bf16(BM, 64, 0, 0);
      const uint32_t i_base = smem_u32(S.sIdent);
      for (int j = 0; j < sg.bn / 64; ++j) {
        mbar_wait(&S.full[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(S.sA + stage * S.a_stride);
        for (int h = 0; h < nsub; ++h) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(d_tmem + h * 256 + 64 * j, make_sdesc(a_base + h * A_BYTES + k * 32, 16, 1024),
                      make_sdesc(i_base + k * 32, 16, 1024), idesc64, 1);
        }
        umma_commit(&S.empty[stage]);
        if (++stage == nst) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    umma_commit(&S.tfull[acc0]);
    if (nsub == 2) umma_commit(&S.tfull[1]);
  }
}

// ------------------------------------------------------------------ CTA-pair variants (cta_group::2)
// Both CTAs of the pair run the producer for THEIR tile (tile = blockIdx.x + k * gridDim.x; blockIdx.x = 2 * pair + rank, the
// pair's two tiles are consecutive 128-pixel tiles of one segment): own im2col A tile + rows [128 rank, 128 rank + 128) of
// the weight tile into their own shared memory, every byte accounted on the LEADER's full barrier.
__device__ __forceinline__ void conv_producer_cta2(const ConvParamsDev* P, const ConvSmem& S, uint32_t rank) {
  const int total = P->total_tiles, nst = S.nst;
  int stage = 0;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const ConvSegDev& sg = P->seg[find_seg(P, tile)];
    const int tl = tile - sg.tile_begin;
    const int nt = tl / sg.m_tiles;           // m_tiles is even: a pair shares its n-tile
    const int mt = tl - nt * sg.m_tiles;
    const int pix0 = mt * BM;
    const int n_img = pix0 / sg.HoWo;
    const int rem = pix0 - n_img * sg.HoWo;
    const int p = rem / sg.Wo;
    const int q = rem - p * sg.Wo;
    const int cw = q * sg.stride - sg.pad, ch = p * sg.stride - sg.pad;
    const int half = sg.bn / 2;
    for (int tap = 0; tap < sg.taps; ++tap) {
      const int r = tap / sg.S;
      const int s = tap - r * sg.S;
      for (int kc = 0; kc < sg.cin_chunks; ++kc) {
        mbar_wait(&S.empty[stage], phase ^ 1);     // local: the leader's commit multicasts to both CTAs
        if (rank == 0) mbar_expect_tx(&S.full[stage], 2 * (A_BYTES + half * (BK * 2)));
        const uint32_t bar = mapa_rank(smem_u32(&S.full[stage]), 0);
        tma_load_im2col_4d_cta2(&sg.tmA, bar, S.sA + stage * S.a_stride, kc * BK, cw, ch, n_img, (uint16_t)s, (uint16_t)r);
        tma_load_3d_cta2(&sg.tmB, bar, S.sB + stage * S.b_stride, kc * BK, nt * sg.bn + (int)rank * half, tap);
        if (++stage == nst) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  }
}

// Leader only: one tcgen05.mma.cta_group::2 of M = 256 x N = bn per 16-channel K step; commits multicast to both CTAs.
__device__ __forceinline__ void conv_mma_cta2(const ConvParamsDev* P, const ConvSmem& S, uint32_t tmem_base) {
  const int total = P->total_tiles, nst = S.nst;
  int stage = 0;
  uint32_t phase = 0;
  int it = 0;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
    const ConvSegDev& sg = P->seg[find_seg(P, tile)];
    const int kiters = sg.taps * sg.cin_chunks;
    const int acc = it & 1;
    mbar_wait(&S.tempty[acc], ((it >> 1) & 1) ^ 1);     // both CTAs' epilogues have drained this accumulator
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + acc * 256;
    const uint32_t idesc = make_idesc_bf16(2 * BM, sg.bn, 0, 0);
    for (int ki = 0; ki < kiters; ++ki) {
      mbar_wait(&S.full[stage], phase);
      tc_fence_after();
      const uint32_t a_base = smem_u32(S.sA + stage * S.a_stride);
      const uint32_t b_base = smem_u32(S.sB + stage * S.b_stride);
#pragma unroll
      for (int k = 0; k < BK / 16; ++k)
        umma_bf16_cta2(d_tmem, make_sdesc(a_base + k * 32, 16, 1024), make_sdesc(b_base + k * 32, 16, 1024), idesc,
                       (ki | k) != 0);
      umma_commit_cta2(&S.empty[stage]);
      if (++stage == nst) {
        stage = 0;
        phase ^= 1;
      }
    }
    umma_commit_cta2(&S.tfull[acc]);
  }
}

// The whole parameter block (tile table + TMA descriptors, <= 10 KiB) travels as a __grid_constant__ kernel parameter:
// it lives in the constant bank, so the per-tile / per-chunk reads of segment fields are constant-cache hits instead
// of dependent global loads that every "memory"-clobbering barrier asm would force again.
template <bool CTA2>
__device__ __forceinline__ void conv_igemm_body(const ConvParamsDev* P) {
  extern __shared__ uint8_t smem_raw[];
  const ConvSmem S = conv_carve(P, smem_raw);
  const int nbuf = P->nbuf;
  uint8_t* const sOut = S.sOut;
  uint64_t* const tfull = S.tfull;
  uint64_t* const tempty = S.tempty;
  uint64_t* const auxfull = S.auxfull;
  float* const sStat = S.sStat;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t tmem_base = conv_prologue<CTA2>(P, S, 8);
  const int total = P->total_tiles;
  const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;

  if (warp == 0) {
    if (elect_one()) {
      if (CTA2) conv_producer_cta2(P, S, cta_rank);
      else conv_producer(P, S);
    }
  } else if (warp == 1) {
    if (CTA2) {
      if (cta_rank == 0 && elect_one()) conv_mma_cta2(P, S, tmem_base);
    } else if (elect_one()) {
      conv_mma(P, S, tmem_base);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (2 warpgroups x 4 warps)
    const int ew = warp & 3;          // TMEM lane quadrant this warp may read
    const int eg = (warp - 4) >> 2;   // warpgroup: handles 64-channel slab `eg` of every 128-channel round
    const int et = ew * 32 + lane;    // 0..127: row of the tile owned by this thread
    const bool leader = (ew == 0 && lane == 0);  // one per warpgroup: owns that group's staging slab + TMA stores
    int it = 0;
    bool store_pending = false;       // (leader) a TMA store may still be reading the staging buffer
    int sbuf = 0;                     // staging slab of the current round (toggles every round when nbuf == 2)
    const bool dbl_store = (nbuf == 2 && !P->any_aux);
    uint32_t aux_par0 = 0, aux_par1 = 0;
    // (leader) TMA-load the residual / mask tile of work item (tile_n, round r0_n) into slab `buf` of this group
    auto aux_issue = [&](int tile_n, int r0_n, int buf) {
      if (tile_n >= total) return;
      const ConvSegDev& s2 = P->seg[find_seg(P, tile_n)];
      if (!s2.aux_kind) return;
      const int cbeg2 = r0_n + eg * 64;
      if (cbeg2 >= min(r0_n + 128, s2.bn)) return;
      const int tl2 = tile_n - s2.tile_begin;
      const int nt2 = tl2 / s2.m_tiles;
      const int mt2 = tl2 - nt2 * s2.m_tiles;
      uint64_t* bar = &auxfull[eg * 2 + buf];
      mbar_expect_tx(bar, BM * 128);
      tma_load_2d(&s2.tmAux, bar, sOut + (eg * nbuf + buf) * (BM * 128), nt2 * s2.bn + cbeg2, mt2 * BM);
    };
    if (leader && nbuf == 2) aux_issue(blockIdx.x, 0, 0);
    // pair mode: a work item = two consecutive 128-pixel tiles whose accumulators sit side by side in TMEM; they are
    // drained one after the other exactly like two tiles of the single mode (accumulator it & 1, same barrier phases)
    const int nsub = P->pair ? 2 : 1;
    for (int base = blockIdx.x * nsub; base < total; base += gridDim.x * nsub)
    for (int hsub = 0; hsub < nsub; ++hsub, ++it) {
      const int tile = base + hsub;
      const ConvSegDev& sg = P->seg[find_seg(P, tile)];
      const int tl = tile - sg.tile_begin;
      const int nt = tl / sg.m_tiles;
      const int mt = tl - nt * sg.m_tiles;
      const int pix = mt * BM + et;
      const bool valid = pix < sg.npix;
      const int n_img = pix / sg.HoWo;
      long long opix = pix;
      if (sg.scatter2) {
        const int rem = pix - n_img * sg.HoWo;
        const int p = rem / sg.Wo;
        const int q = rem - p * sg.Wo;
        opix = ((long long)n_img * sg.Hs + 2 * p) * sg.Ws + 2 * q;
      }
      const long long row = opix * sg.ldc;
      const int acc = it & 1;
      const int bn = sg.bn;
      const bool staged = sg.staged != 0;
      const bool do_stats = sg.stats != nullptr && mt * BM < sg.npix;   // (a pair-mode padding tile has no pixels)
      // all valid pixels of this tile in one image => GroupNorm partial sums reduce per tile
      const int pix_first = mt * BM;
      const int pix_last = min(pix_first + BM, sg.npix) - 1;
      const int n_tile = pix_first / sg.HoWo;
      const bool tile_uniform = (pix_last / sg.HoWo) == n_tile;

      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * 256;

      for (int r0 = 0; r0 < bn; r0 += 128) {  // rounds of <=128 channels (= the staging buffer)
        const int rend = min(r0 + 128, bn);
        // A round of <= 64 channels (N <= 64 layers: stem, layer1 conv1 / conv2) would leave warpgroup 1 idle: both
        // groups then split the chunks of warpgroup 0's slab and hand it over with 256-thread barriers.
        const bool shared = staged && (rend - r0 <= 64);
        int cbeg = r0 + eg * 64;
        int cend = min(cbeg + 64, rend);
        if (shared) {
          const int half = (((rend - r0 + 15) >> 4) + 1) >> 1;
          cbeg = eg ? r0 + half * 16 : r0;
          cend = eg ? rend : min(r0 + half * 16, rend);
        }
        const int aux_here = (nbuf == 2 && !shared && cbeg < rend) ? sg.aux_kind : 0;
        uint8_t* const slab_base = sOut + ((shared ? 0 : eg) * nbuf + sbuf) * (BM * 128);
        if (shared) {
          if (leader && eg == 0 && store_pending) {
            if (dbl_store) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else {
              asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              store_pending = false;
            }
          }
          asm volatile("bar.sync 7, 256;" ::: "memory");  // warpgroup 0's slab is free
        } else if (dbl_store) {
          // Two slabs, no TMA-loaded operands: the store issued a round ago (other slab) may still be in flight; only
          // the one before it, which read THIS round's slab, has to be drained.
          if (leader && store_pending) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
        } else if (nbuf == 2) {
          // Double-buffered slabs: the slab of this round was released a whole round ago (its store was drained
          // before the previous round's hand-over barrier). The leader drains the previous store and immediately
          // queues the NEXT work item's residual / mask tile into the other slab, one full round ahead of its use.
          if (leader) {
            if (store_pending) {
              asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              store_pending = false;
            }
            if (r0 + 128 < bn) aux_issue(tile, r0 + 128, sbuf ^ 1);
            else aux_issue(tile + gridDim.x, 0, sbuf ^ 1);
          }
          if (aux_here) {
            mbar_wait(&auxfull[eg * 2 + sbuf], sbuf ? aux_par1 : aux_par0);
            if (sbuf) aux_par1 ^= 1; else aux_par0 ^= 1;
          }
        } else if (staged) {
          if (leader && store_pending) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            store_pending = false;
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");  // this group's staging slab is free
        }
        // Lean path for the common tile: bf16 output through the staging slab, every 16-channel chunk inside Cout,
        // no per-channel scale, no GroupNorm statistics, ReLU on all or none of the tile's channels, residual / mask
        // (if any) already in the slab. Shared memory is addressed through 32-bit shared-space ld / st, ReLU rides on
        // the bf16x2 conversion and the mask is applied to the packed result, so a chunk costs ~50-80 instructions.
        const int relu_nch = sg.relu_nch;
        const int tile_c0 = nt * bn;
        const bool fast = staged && !do_stats && aux_here != 3 && sg.scale == nullptr && tile_c0 + bn <= sg.cout &&
                          (relu_nch >= tile_c0 + bn || relu_nch <= tile_c0) &&
                          (sg.residual == nullptr || aux_here == 1) && (sg.relu_mask == nullptr || aux_here == 2);
        if (fast) {
          const float* __restrict__ shp = sg.shift ? sg.shift + tile_c0 : nullptr;
          const bool relu_all = relu_nch >= tile_c0 + bn;
          const uint32_t slab_row = smem_u32(slab_base) + et * 128;
          const uint32_t sw = et & 7;
          fast_dispatch<true>((shp ? 6 : 0) + aux_here * 2 + (relu_all ? 1 : 0), taddr, cbeg, cend, r0, slab_row, sw, shp);
        } else
        for (int c0 = cbeg; c0 < cend; c0 += 16) {
          uint32_t rr[16];
          tmem_ld16(taddr + c0, rr);
          uint4 a0 = make_uint4(0u, 0u, 0u, 0u), a1 = a0;
          if (aux_here) {  // this thread's 32 bytes of the TMA-loaded tile (same swizzled slots it will overwrite)
            const int chx = ((c0 - r0) & 63) >> 3;
            a0 = *reinterpret_cast<const uint4*>(slab_base + et * 128 + ((chx ^ (et & 7)) << 4));
            a1 = *reinterpret_cast<const uint4*>(slab_base + et * 128 + (((chx + 1) ^ (et & 7)) << 4));
          }
          tmem_ld_wait();
          const int cb = nt * bn + c0;  // first global output channel of this chunk
          const bool fullchunk = (cb + 16 <= sg.cout);
          float v[16];
          epilogue_math(sg, rr, v, cb, fullchunk, valid, row, aux_here == 3 ? 0 : aux_here, a0, a1);
          if (do_stats) {
            // GroupNorm partial sums of the two 8-channel halves of this chunk, reduced over the warp's 32 pixels
            // with a 6-shuffle butterfly; lanes 0/8/16/24 end up with (s1,h0) (s2,h0) (s1,h1) (s2,h1).
            float a = 0.f, b = 0.f, c = 0.f, d = 0.f;
            if (valid && sg.gnb_x) {
              // backward of conv-bias -> GroupNorm -> ReLU in front of this output: v is dz (the gradient w.r.t. the
              // post-ReLU map), rounded to the bf16 the apply pass will read back. Per 8-channel half:
              //   (a | b) = sum gamma * dy,  (c | d) = sum gamma * dy * xhat,  dy = dz * [xhat * gamma + beta > 0]
              // (x tile: brought into the staging slab by TMA like a residual tile, aux_kind 3; else read from memory)
              uint4 x0 = a0, x1 = a1;
              if (aux_here != 3) {
                const uint4* xp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(sg.gnb_x) +
                                                                 (long long)pix * sg.cout + cb);
                x0 = __ldg(xp);
                x1 = __ldg(xp + 1);
              }
              const float4 m0 = __ldg(sg.gnb_mr + n_img * sg.groups + cb / sg.cpg);
              const float4 m1 = __ldg(sg.gnb_mr + n_img * sg.groups + (cb + 8) / sg.cpg);
              const uint32_t xw[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(sg.gnb_gamma + cb) + q);
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(sg.gnb_beta + cb) + q);
                const float gq[4] = {g4.x, g4.y, g4.z, g4.w}, bq[4] = {b4.x, b4.y, b4.z, b4.w};
                const float mean = q < 2 ? m0.x : m1.x, rstd = q < 2 ? m0.y : m1.y;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int j = 4 * q + e;
                  const uint32_t w2 = xw[j >> 1];
                  const float xv = __uint_as_float((j & 1) ? (w2 & 0xffff0000u) : (w2 << 16));
                  const float xh = (xv - mean) * rstd;
                  const float dzr = __bfloat162float(__float2bfloat16_rn(v[j]));
                  const float gd = fmaf(xh, gq[e], bq[e]) > 0.f ? gq[e] * dzr : 0.f;
                  if (q < 2) {
                    a += gd;
                    c = fmaf(gd, xh, c);
                  } else {
                    b += gd;
                    d = fmaf(gd, xh, d);
                  }
                }
              }
            } else if (valid) {
              if (fullchunk) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  a += v[j];
                  c = fmaf(v[j], v[j], c);
                  b += v[8 + j];
                  d = fmaf(v[8 + j], v[8 + j], d);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  if (cb + j < sg.cout) {
                    a += v[j];
                    c = fmaf(v[j], v[j], c);
                  }
                  if (cb + 8 + j < sg.cout) {
                    b += v[8 + j];
                    d = fmaf(v[8 + j], v[8 + j], d);
                  }
                }
              }
            }
            if (tile_uniform) {
              const bool hi16 = (lane & 16) != 0;
              float x = (hi16 ? b : a) + __shfl_xor_sync(0xffffffffu, hi16 ? a : b, 16);
              float y = (hi16 ? d : c) + __shfl_xor_sync(0xffffffffu, hi16 ? c : d, 16);
              const bool hi8 = (lane & 8) != 0;
              float z = (hi8 ? y : x) + __shfl_xor_sync(0xffffffffu, hi8 ? x : y, 8);
              z += __shfl_xor_sync(0xffffffffu, z, 4);
              z += __shfl_xor_sync(0xffffffffu, z, 2);
              z += __shfl_xor_sync(0xffffffffu, z, 1);
              if ((lane & 7) == 0) sStat[(ew * 32 + (c0 >> 3) + (lane >> 4)) * 2 + ((lane >> 3) & 1)] = z;
            } else if (valid) {  // rare: the tile straddles two images
              const int cpg = sg.cpg;
              if (cb < sg.cout) {
                double* dst = sg.stats + ((long long)n_img * sg.groups + cb / cpg) * DSLB_GN_STAT_STRIDE;
                atomicAdd(dst, (double)a);
                atomicAdd(dst + 1, (double)c);
              }
              if (cb + 8 < sg.cout) {
                double* dst = sg.stats + ((long long)n_img * sg.groups + (cb + 8) / cpg) * DSLB_GN_STAT_STRIDE;
                atomicAdd(dst, (double)b);
                atomicAdd(dst + 1, (double)d);
              }
            }
          }
          if (staged) {
            // 128B-swizzled [128 rows][64 ch] slabs, the layout the TMA store expects
            const int cl = c0 - r0;
            uint8_t* slab = slab_base + et * 128;
            const int ch = (cl & 63) >> 3;  // 16-byte chunk index inside the 128-byte row
            uint4 o0, o1;
            o0.x = pack_bf16(v[0], v[1]);
            o0.y = pack_bf16(v[2], v[3]);
            o0.z = pack_bf16(v[4], v[5]);
            o0.w = pack_bf16(v[6], v[7]);
            o1.x = pack_bf16(v[8], v[9]);
            o1.y = pack_bf16(v[10], v[11]);
            o1.z = pack_bf16(v[12], v[13]);
            o1.w = pack_bf16(v[14], v[15]);
            *reinterpret_cast<uint4*>(slab + ((ch ^ (et & 7)) << 4)) = o0;
            *reinterpret_cast<uint4*>(slab + (((ch + 1) ^ (et & 7)) << 4)) = o1;
          } else if (valid) {
            if (sg.out_fp32) {
              float* yp = reinterpret_cast<float*>(sg.y) + row + cb;
              if (fullchunk) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  reinterpret_cast<float4*>(yp)[j] =
                      make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (cb + j < sg.cout) yp[j] = v[j];
              }
            } else {
              __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(sg.y) + row + cb;
              if (fullchunk) {
                uint4 o0, o1;
                o0.x = pack_bf16(v[0], v[1]);
                o0.y = pack_bf16(v[2], v[3]);
                o0.z = pack_bf16(v[4], v[5]);
                o0.w = pack_bf16(v[6], v[7]);
                o1.x = pack_bf16(v[8], v[9]);
                o1.y = pack_bf16(v[10], v[11]);
                o1.z = pack_bf16(v[12], v[13]);
                o1.w = pack_bf16(v[14], v[15]);
                reinterpret_cast<uint4*>(yp)[0] = o0;
                reinterpret_cast<uint4*>(yp)[1] = o1;
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (cb + j < sg.cout) yp[j] = __float2bfloat16_rn(v[j]);
              }
            }
          }
        }
        if (rend == bn) {  // accumulator fully drained: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CTA2) mbar_arrive_cluster(mapa_rank(smem_u32(&tempty[acc]), 0));   // the pair's MMA warp lives in rank 0
            else mbar_arrive(&tempty[acc]);
          }
        }
        if (shared) {
          fence_proxy_async();
          asm volatile("bar.sync 8, 256;" ::: "memory");
          if (leader && eg == 0) {
            tma_store_2d(&sg.tmY, slab_base, nt * bn + r0, pix_first);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            store_pending = true;
          }
        } else if (staged) {
          fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA (async proxy)
          asm volatile("bar.sync %0, 128;" ::"r"(3 + eg) : "memory");
          if (leader && cbeg < rend) {
            tma_store_2d(&sg.tmY, slab_base, nt * bn + cbeg, pix_first);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            store_pending = true;
          }
        }
        if (nbuf == 2) sbuf ^= 1;
      }
      if (do_stats && tile_uniform) {
        asm volatile("bar.sync 5, 256;" ::: "memory");  // both warpgroups have written their sStat entries
        if (eg == 0 && et < 64) {
          const int hidx = et >> 1, k = et & 1;  // 8-channel half index inside the tile, (sum | sumsq)
          const int cfirst = nt * bn + hidx * 8;
          if (hidx * 8 < bn && cfirst < sg.cout) {
            const float tot = sStat[(0 * 32 + hidx) * 2 + k] + sStat[(1 * 32 + hidx) * 2 + k] +
                              sStat[(2 * 32 + hidx) * 2 + k] + sStat[(3 * 32 + hidx) * 2 + k];
            double* dst = sg.stats + ((long long)n_tile * sg.groups + cfirst / sg.cpg) * DSLB_GN_STAT_STRIDE;
            atomicAdd(dst + k, (double)tot);
          }
        }
        asm volatile("bar.sync 6, 256;" ::: "memory");  // sStat may be overwritten by the next tile
      }
    }
    if (leader && store_pending) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (CTA2) {   // the peer may still be reading this CTA's operands / signalling its barriers
    __syncwarp();
    cluster_sync_all();
  }
  if (warp == 2) {
    tc_fence_after();
    if (CTA2) tmem_dealloc2(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

__global__ void __launch_bounds__(384, 1) conv_igemm_kernel(const __grid_constant__ ConvParamsDev PP) {
  conv_igemm_body<false>(&PP);
}

// Same kernel launched as clusters of two CTAs (cudaLaunchAttributeClusterDimension = 2): see ConvParamsDev::cta2.
__global__ void __launch_bounds__(384, 1) conv_igemm_cta2_kernel(const __grid_constant__ ConvParamsDev PP) {
  conv_igemm_body<true>(&PP);
}

// Fast kernel for plans whose every tile takes the lean epilogue (bf16 staged output, full chunks, no scale / GroupNorm
// statistics, residual on the tensor core, mask by TMA) and whose tiles are short and wide (<= 16 k-iterations, up to 256
// columns): SIXTEEN epilogue warps = four warpgroups, each owning one 64-channel slab of the tile and one staging slab.
// The 8-warp epilogue of the general kernel is issue / latency bound on such tiles (2 warps per scheduler); with 4 per
// scheduler the per-tile epilogue time halves and the store / mask-load latency of one group hides behind the others.
__global__ void __launch_bounds__(640, 1) conv_igemm_fast4_kernel(const __grid_constant__ ConvParamsDev PP) {
  const ConvParamsDev* P = &PP;
  extern __shared__ uint8_t smem_raw[];
  const ConvSmem S = conv_carve(P, smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t tmem_base = conv_prologue(P, S, 16);
  const int total = P->total_tiles;

  if (warp == 0) {
    if (elect_one()) conv_producer(P, S);
  } else if (warp == 1) {
    if (elect_one()) conv_mma(P, S, tmem_base);
  } else if (warp >= 4) {
    const int ew = warp & 3;          // TMEM lane quadrant
    const int eg = (warp - 4) >> 2;   // warpgroup 0..3: channels [64 eg, 64 eg + 64) of the tile
    const int et = ew * 32 + lane;
    const bool leader = (ew == 0 && lane == 0);
    uint8_t* const slab = S.sOut + eg * (BM * 128);
    const uint32_t slab_row = smem_u32(slab) + et * 128;
    const uint32_t sw = et & 7;
    uint64_t* const auxbar = &S.auxfull[eg];
    uint32_t aux_par = 0;
    // (leader) queue the mask tile of work item `tile_n` into this group's slab; the slab must be free
    auto aux_issue = [&](int tile_n) {
      if (tile_n >= total) return;
      const ConvSegDev& s2 = P->seg[find_seg(P, tile_n)];
      if (!s2.aux_kind || eg * 64 >= s2.bn) return;
      const int tl2 = tile_n - s2.tile_begin;
      const int nt2 = tl2 / s2.m_tiles;
      const int mt2 = tl2 - nt2 * s2.m_tiles;
      mbar_expect_tx(auxbar, BM * 128);
      tma_load_2d(&s2.tmAux, auxbar, slab, nt2 * s2.bn + eg * 64, mt2 * BM);
    };
    if (leader) aux_issue(blockIdx.x);
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const ConvSegDev& sg = P->seg[find_seg(P, tile)];
      const int tl = tile - sg.tile_begin;
      const int nt = tl / sg.m_tiles;
      const int mt = tl - nt * sg.m_tiles;
      const int acc = it & 1;
      const int bn = sg.bn;
      const int cbeg = eg * 64, cend = min(cbeg + 64, bn);
      const bool active = cbeg < bn;
      const int aux_here = active ? sg.aux_kind : 0;
      // slab hand-over: the leader drained the previous store before it queued the mask tile / reached this barrier
      if (aux_here) {
        mbar_wait(auxbar, aux_par);
        aux_par ^= 1;
      } else {
        asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
      }
      mbar_wait(&S.tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      if (active) {
        const int tile_c0 = nt * bn;
        const float* __restrict__ shp = sg.shift ? sg.shift + tile_c0 : nullptr;
        const bool relu_all = sg.relu_nch >= tile_c0 + bn;
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * 256;
        fast_dispatch<false>((shp ? 6 : 0) + aux_here * 2 + (relu_all ? 1 : 0), taddr, cbeg, cend, 0, slab_row, sw, shp);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.tempty[acc]);
      if (active) fence_proxy_async();
      asm volatile("bar.sync %0, 128;" ::"r"(5 + eg) : "memory");
      if (leader) {
        if (active) {
          tma_store_2d(&sg.tmY, slab, nt * bn + cbeg, mt * BM);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        aux_issue(tile + gridDim.x);
      }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace dslb

// ============================================================================================ host side
using namespace dslb;

struct dslb_conv_plan {
  ConvParamsDev* dev = nullptr;  // HOST copy, passed by value at launch
  int total_tiles = 0;
  double flops = 0.0;
  int fast4 = 0;  // launch conv_igemm_fast4_kernel (16 epilogue warps)
  dslb_halo_plan* halo = nullptr;  // narrow 3x3 stride-1 conv: halo-tile kernel of conv_halo.cu instead
  int cta2 = 0;   // launch conv_igemm_cta2_kernel as clusters of two CTAs
};

static int pick_bn(int cout_pad) {
  // largest tile width <= 256 that divides cout_pad (cout_pad is a multiple of 16)
  for (int bn = 256; bn >= 16; bn -= 16)
    if (cout_pad % bn == 0) return bn;
  return 16;
}

extern "C" int dslb_conv_plan_create(const dslb_conv_seg_t* segs, int nseg, dslb_conv_plan_t** out) {
  DSLB_CHECK_ARG(segs && out, "dslb_conv_plan_create: null argument");
  DSLB_CHECK_ARG(nseg >= 1 && nseg <= DSLB_MAX_SEGS, "dslb_conv_plan_create: nseg %d not in [1,%d]", nseg,
                 DSLB_MAX_SEGS);
  if (nseg == 1 && segs[0].x && segs[0].w && segs[0].y && segs[0].N > 0 && segs[0].ldc >= segs[0].Cout &&
      ((uintptr_t)segs[0].x % 16) == 0 && ((uintptr_t)segs[0].w % 16) == 0 && ((uintptr_t)segs[0].y % 16) == 0 &&
      halo_eligible(segs[0])) {
    dslb_halo_plan* hp = nullptr;
    const int rc = halo_plan_create(segs[0], &hp);
    if (rc != DSLB_OK) return rc;
    dslb_conv_plan* plan = new (std::nothrow) dslb_conv_plan();
    if (!plan) {
      halo_plan_destroy(hp);
      set_error("out of host memory");
      return DSLB_ENOMEM;
    }
    plan->halo = hp;
    plan->flops = 2.0 * (double)segs[0].N * segs[0].H * segs[0].W * segs[0].Cout * segs[0].Cin * 9.0;
    *out = plan;
    return DSLB_OK;
  }
  ConvParamsDev* h = new (std::nothrow) ConvParamsDev();
  if (!h) {
    set_error("out of host memory");
    return DSLB_ENOMEM;
  }
  memset(h, 0, sizeof(*h));
  int tiles = 0;
  double flops = 0.0;
  bool any_aux = false, any_ident = false;
  // Tile width: the widest tile (<= 256) is the most MMA-efficient, but the deep layers have so few 128-pixel row
  // tiles (layer4: 33) that 256-wide tiles leave most SMs idle. Estimate waves x per-tile cycles for caps 256 / 128 /
  // 64 (k-iteration of a 256-wide tile ~ 542 clk, ~3000 clk of per-tile fill + epilogue) and take the cheapest.
  int bn_cap = 256;
  {
    const int sms = num_sms();
    double best = 0.0;
    for (int cap = 256; cap >= 64; cap >>= 1) {
      long long ntile = 0;
      double work = 0.0;
      bool ok = true;
      for (int i = 0; i < nseg && ok; ++i) {
        const dslb_conv_seg_t& s = segs[i];
        if (s.stride < 1 || s.R < 1 || s.S < 1 || s.cout_pad < 16) { ok = false; break; }
        const int Ho = (s.H + 2 * s.pad - s.R) / s.stride + 1, Wo = (s.W + 2 * s.pad - s.S) / s.stride + 1;
        if (Ho <= 0 || Wo <= 0) { ok = false; break; }
        int bn = pick_bn(s.cout_pad);
        if (bn > cap && s.cout_pad % cap == 0) bn = cap;
        // residual / mask tiles by TMA need > 64 channels; GroupNorm statistics were tuned for full-width tiles
        if (cap < 128 && (s.residual || s.relu_mask)) bn = pick_bn(s.cout_pad) > 128 && s.cout_pad % 128 == 0 ? 128 : pick_bn(s.cout_pad);
        if (s.gn_stats || s.gnb_x) bn = pick_bn(s.cout_pad);
        const long long t = (long long)cdiv(s.N * Ho * Wo, BM) * (s.cout_pad / bn);
        const double kit = (double)s.R * s.S * (s.Cin / 64);
        ntile += t;
        work += (double)t * (kit * 542.0 * bn / 256.0 + 3000.0);
      }
      if (!ok || ntile == 0) break;
      const double waves = (double)((ntile + sms - 1) / sms);
      const double cost = waves * (work / (double)ntile);
      if (cap == 256 || cost < 0.9 * best) {
        if (cap == 256 || cost < best) {
          best = cost;
          bn_cap = cap;
        }
      }
    }
  }
  for (int i = 0; i < nseg; ++i) {
    const dslb_conv_seg_t& s = segs[i];
    ConvSegDev& d = h->seg[i];
    int rc = DSLB_OK;
#define SEG_CHECK(cond, ...)       \
  if (!(cond)) {                   \
    set_error(__VA_ARGS__);        \
    delete h;                      \
    return DSLB_EINVAL;            \
  }
    SEG_CHECK(s.x && s.w && s.y, "conv seg %d: null x/w/y", i);
    SEG_CHECK(s.N > 0 && s.H > 0 && s.W > 0, "conv seg %d: bad N/H/W", i);
    SEG_CHECK(s.Cin > 0 && s.Cin % 64 == 0, "conv seg %d: Cin=%d must be a multiple of 64", i, s.Cin);
    SEG_CHECK(s.Cout > 0 && s.cout_pad >= s.Cout && s.cout_pad % 16 == 0,
              "conv seg %d: cout_pad=%d must be a multiple of 16 and >= Cout=%d", i, s.cout_pad, s.Cout);
    SEG_CHECK(s.R >= 1 && s.S >= 1 && s.R <= 7 && s.S <= 7 && s.stride >= 1 && s.stride <= 2 && s.pad >= 0,
              "conv seg %d: unsupported filter %dx%d stride %d pad %d", i, s.R, s.S, s.stride, s.pad);
    SEG_CHECK(s.ldc >= s.Cout && s.ldc % (s.out_fp32 ? 4 : 8) == 0, "conv seg %d: ldc=%d misaligned", i, s.ldc);
    SEG_CHECK(((uintptr_t)s.x % 16) == 0 && ((uintptr_t)s.w % 16) == 0 && ((uintptr_t)s.y % 16) == 0,
              "conv seg %d: x/w/y must be 16-byte aligned", i);
    const int Ho = (s.H + 2 * s.pad - s.R) / s.stride + 1;
    const int Wo = (s.W + 2 * s.pad - s.S) / s.stride + 1;
    SEG_CHECK(Ho > 0 && Wo > 0, "conv seg %d: empty output", i);
    if (s.gn_stats) {
      SEG_CHECK(s.gn_cpg == 8 || s.gn_cpg == 16, "conv seg %d: gn_cpg=%d must be 8 or 16", i, s.gn_cpg);
      SEG_CHECK(s.Cout % s.gn_cpg == 0, "conv seg %d: Cout %% gn_cpg != 0", i);
    }
    if (s.gnb_x) {
      SEG_CHECK(!s.gn_stats && s.gnb_mr && s.gnb_gamma && s.gnb_beta && s.gnb_sums,
                "conv seg %d: gnb_x needs gnb_mr / gnb_gamma / gnb_beta / gnb_sums and excludes gn_stats", i);
      SEG_CHECK((s.gn_cpg == 8 || s.gn_cpg == 16) && s.Cout % 16 == 0 && s.cout_pad == s.Cout && s.Cout % s.gn_cpg == 0 &&
                    !s.scatter2 && !s.out_fp32,
                "conv seg %d: gnb_x needs gn_cpg 8 or 16, Cout %% 16 == 0, a bf16 non-scattered output", i);
      SEG_CHECK(((uintptr_t)s.gnb_x % 16) == 0 && ((uintptr_t)s.gnb_mr % 16) == 0 && ((uintptr_t)s.gnb_gamma % 16) == 0 &&
                    ((uintptr_t)s.gnb_beta % 16) == 0,
                "conv seg %d: gnb pointers must be 16-byte aligned", i);
    }
    if (s.scatter2) SEG_CHECK(s.Hs >= 2 * Ho - 1 && s.Ws >= 2 * Wo - 1, "conv seg %d: scatter map too small", i);
#undef SEG_CHECK
    d.y = s.y;
    d.residual = s.residual;
    d.relu_mask = s.relu_mask;
    d.scale = s.scale;
    d.shift = s.shift;
    d.stats = s.gnb_x ? s.gnb_sums : s.gn_stats;
    d.gnb_x = s.gnb_x;
    d.gnb_mr = reinterpret_cast<const float4*>(s.gnb_mr);
    d.gnb_gamma = s.gnb_gamma;
    d.gnb_beta = s.gnb_beta;
    d.npix = s.N * Ho * Wo;
    d.HoWo = Ho * Wo;
    d.Wo = Wo;
    d.bn = pick_bn(s.cout_pad);
    if (!s.gn_stats && !s.gnb_x) {
      if (d.bn > bn_cap && s.cout_pad % bn_cap == 0) d.bn = bn_cap;
      if (bn_cap < 128 && (s.residual || s.relu_mask))
        d.bn = pick_bn(s.cout_pad) > 128 && s.cout_pad % 128 == 0 ? 128 : pick_bn(s.cout_pad);
    }
    d.m_tiles = cdiv(d.npix, BM);
    d.n_tiles = s.cout_pad / d.bn;
    d.tile_begin = tiles;
    tiles += d.m_tiles * d.n_tiles;
    d.cin_chunks = s.Cin / 64;
    d.taps = s.R * s.S;
    d.S = s.S;
    d.stride = s.stride;
    d.pad = s.pad;
    d.cout = s.Cout;
    d.ldc = s.ldc;
    d.out_fp32 = s.out_fp32;
    d.relu_nch = s.relu_nch;
    d.cpg = (s.gn_stats || s.gnb_x) ? s.gn_cpg : 16;
    d.groups = (s.gn_stats || s.gnb_x) ? s.Cout / s.gn_cpg : 1;
    d.scatter2 = s.scatter2;
    d.Hs = s.Hs;
    d.Ws = s.Ws;
    rc = encode_im2col_bf16(&d.tmA, s.x, s.N, s.H, s.W, s.Cin, s.R, s.S, s.stride, s.pad, BM);
    if (rc != DSLB_OK) {
      delete h;
      return rc;
    }
    const uint64_t wd[3] = {(uint64_t)s.Cin, (uint64_t)s.cout_pad, (uint64_t)(s.R * s.S)};
    const uint64_t ws[2] = {(uint64_t)s.Cin * 2, (uint64_t)s.cout_pad * s.Cin * 2};
    const uint32_t wb[3] = {64, (uint32_t)d.bn, 1};
    rc = encode_tiled_bf16(&d.tmB, s.w, 3, wd, ws, wb);
    if (rc != DSLB_OK) {
      delete h;
      return rc;
    }
    d.staged = (!s.out_fp32 && !s.scatter2) ? 1 : 0;
    if (d.staged) {
      const uint64_t yd[2] = {(uint64_t)s.Cout, (uint64_t)d.npix};
      const uint64_t ys[1] = {(uint64_t)s.ldc * 2};
      const uint32_t yb[2] = {64, (uint32_t)BM};
      rc = encode_tiled_bf16(&d.tmY, s.y, 2, yd, ys, yb);
      if (rc != DSLB_OK) {
        delete h;
        return rc;
      }
    }
    d.res_mma = 0;
    if (s.residual && !s.scatter2 && !s.scale && s.Cout % 64 == 0 && d.bn % 64 == 0 && s.ldc % 8 == 0) {
      const uint64_t rd[2] = {(uint64_t)s.Cout, (uint64_t)d.npix};
      const uint64_t rs[1] = {(uint64_t)s.ldc * 2};
      const uint32_t rb[2] = {64, (uint32_t)BM};
      rc = encode_tiled_bf16(&d.tmRes, s.residual, 2, rd, rs, rb);
      if (rc != DSLB_OK) {
        delete h;
        return rc;
      }
      d.res_mma = 1;
      d.residual = nullptr;
      any_ident = true;
    }
    d.aux_kind = 0;
    if (d.staged && s.Cout % 64 == 0 && d.bn > 64 && s.ldc % 8 == 0 && (d.residual || s.relu_mask)) {
      const void* aux = d.residual ? d.residual : s.relu_mask;
      d.aux_kind = d.residual ? 1 : 2;
      const uint64_t ad[2] = {(uint64_t)s.Cout, (uint64_t)d.npix};
      const uint64_t as[1] = {(uint64_t)s.ldc * 2};
      const uint32_t ab[2] = {64, (uint32_t)BM};
      rc = encode_tiled_bf16(&d.tmAux, aux, 2, ad, as, ab);
      if (rc != DSLB_OK) {
        delete h;
        return rc;
      }
      any_aux = true;
    }
    if (d.aux_kind == 0 && s.gnb_x && d.staged && s.Cout % 64 == 0 && d.bn > 64 && getenv("DSLB_GNB_NO_TMA") == nullptr) {
      // the pre-norm tile of the GroupNorm backward sums travels like a residual tile: [128 px][64 ch] by TMA into the
      // staging slab, read in place by the thread that then overwrites the same bytes with its output
      d.aux_kind = 3;
      const uint64_t ad[2] = {(uint64_t)s.Cout, (uint64_t)d.npix};
      const uint64_t as[1] = {(uint64_t)s.Cout * 2};
      const uint32_t ab[2] = {64, (uint32_t)BM};
      rc = encode_tiled_bf16(&d.tmAux, s.gnb_x, 2, ad, as, ab);
      if (rc != DSLB_OK) {
        delete h;
        return rc;
      }
      any_aux = true;
    }
    if (s.gn_stats && s.scatter2) {   // (fp32 direct-store outputs accumulate the statistics in the same chunk loop)
      set_error("conv seg %d: gn_stats needs a non-scattered output", i);
      delete h;
      return DSLB_EINVAL;
    }
    flops += 2.0 * (double)d.npix * s.Cout * s.Cin * s.R * s.S;
  }
  h->nseg = nseg;
  h->total_tiles = tiles;
  // residual / mask tiles by TMA need the second staging slab, paid for with one operand stage
  // plans without TMA-loaded operands whose tiles are short (few k-iterations, narrow tiles: store / epilogue bound)
  // use the same 3-stage + 2-slab layout to double-buffer their TMA stores
  bool short_tiles = true;
  for (int i = 0; i < nseg; ++i) {
    const ConvSegDev& d = h->seg[i];
    if (!d.staged || (long long)d.taps * d.cin_chunks * d.bn >= 16 * 256) short_tiles = false;
  }
  // every tile lean-epilogue eligible, <= 16 k-iterations, and wide enough to occupy at least 3 of the 4 groups?
  bool fast4 = getenv("DSLB_NO_FAST4") == nullptr;
  int bn_widest = 0;
  for (int i = 0; i < nseg && fast4; ++i) {
    const ConvSegDev& d = h->seg[i];
    const dslb_conv_seg_t& s = segs[i];
    const bool relu_ok = d.relu_nch <= 0 || d.relu_nch >= d.cout;
    const bool mask_ok = s.relu_mask == nullptr || d.aux_kind == 2;
    if (!d.staged || d.stats || d.scale || d.residual || s.cout_pad != s.Cout || d.cout % d.bn != 0 || !relu_ok ||
        !mask_ok || d.taps * d.cin_chunks > 16 || d.bn > 256)
      fast4 = false;
    bn_widest = d.bn > bn_widest ? d.bn : bn_widest;
  }
  if (bn_widest < 192) fast4 = false;
  // pair mode (two 128-pixel tiles per work item sharing B): compute-heavy plans without TMA-loaded epilogue operands
  // or identity tile (their smem goes to the 64 KiB stages), whose wave quantisation does not eat the gain
  bool pair = !fast4 && !any_aux && !any_ident;
  {
    const char* e = getenv("DSLB_PAIR");
    long long tiles1 = 0, tiles2 = 0;
    for (int i = 0; i < nseg && pair; ++i) {
      const ConvSegDev& d = h->seg[i];
      if (d.taps * d.cin_chunks < 18 || d.bn < 128 || d.scatter2) pair = false;
      tiles1 += (long long)d.m_tiles * d.n_tiles;
      tiles2 += (long long)((d.m_tiles + 1) / 2) * d.n_tiles;
    }
    if (pair) {
      const int sms = num_sms();
      const double w1 = (double)((tiles1 + sms - 1) / sms);          // waves of 1 tile-time
      const double w2 = (double)((tiles2 + sms - 1) / sms) * 2.0;    // waves of 2 tile-times
      if (w2 * 0.80 > w1) pair = false;                              // assume a pair-mode tile-time is ~0.8x
    }
    if (e) pair = pair && e[0] == '1';
    else pair = false;   // opt-in until validated on the GPU
  }
  if (pair) {   // re-number the tiles: every segment gets an even number of 128-pixel row tiles
    int t = 0;
    for (int i = 0; i < nseg; ++i) {
      ConvSegDev& d = h->seg[i];
      d.m_tiles = (d.m_tiles + 1) & ~1;
      d.tile_begin = t;
      t += d.m_tiles * d.n_tiles;
    }
    tiles = t;
    h->total_tiles = tiles;
  }
  // CTA pairs (cta_group::2): compute-heavy plans whose every segment is one 256-wide n-tile with a staged bf16 output and
  // no residual (the FCOSHead tower layers + their dgrads, the FPN output / lateral convs, the 256- and 512-wide backbone
  // convs): M = 256 per MMA halves the weight bytes each SM pulls per k-iteration. DSLB_CTA2=0 switches it off.
  bool cta2 = !fast4 && !pair && !any_ident && !(getenv("DSLB_CTA2") != nullptr && getenv("DSLB_CTA2")[0] == '0');
  {
    int kmin = 8;   // k-iterations per tile from which the pair pays (DSLB_CTA2_KMIN overrides, for experiments)
    if (const char* e = getenv("DSLB_CTA2_KMIN")) kmin = atoi(e);
    for (int i = 0; i < nseg && cta2; ++i) {
      const ConvSegDev& d = h->seg[i];
      if (!d.staged || d.bn != 256 || d.scatter2 || d.residual || d.taps * d.cin_chunks < kmin) cta2 = false;
    }
  }
  if (cta2) {   // every segment gets an even number of 128-pixel row tiles: a pair never straddles two segments
    int t = 0;
    for (int i = 0; i < nseg; ++i) {
      ConvSegDev& d = h->seg[i];
      const dslb_conv_seg_t& sg = segs[i];
      d.m_tiles = (d.m_tiles + 1) & ~1;
      d.tile_begin = t;
      t += d.m_tiles * d.n_tiles;
      // each CTA of the pair loads its own half of the weight tile: box of bn / 2 rows
      const uint64_t wd[3] = {(uint64_t)sg.Cin, (uint64_t)sg.cout_pad, (uint64_t)(sg.R * sg.S)};
      const uint64_t ws[2] = {(uint64_t)sg.Cin * 2, (uint64_t)sg.cout_pad * sg.Cin * 2};
      const uint32_t wb[3] = {64, (uint32_t)d.bn / 2, 1};
      const int rc2 = encode_tiled_bf16(&d.tmB, sg.w, 3, wd, ws, wb);
      if (rc2 != DSLB_OK) {
        delete h;
        return rc2;
      }
    }
    tiles = t;
    h->total_tiles = tiles;
  }
  h->cta2 = cta2 ? 1 : 0;
  h->pair = pair ? 1 : 0;
  h->any_aux = any_aux ? 1 : 0;
  h->has_ident = any_ident ? 1 : 0;
  h->nbuf = (any_aux || any_ident || short_tiles || fast4) ? 2 : 1;
  {
    int bn_max = 16;
    for (int i = 0; i < nseg; ++i) bn_max = h->seg[i].bn > bn_max ? h->seg[i].bn : bn_max;
    if (cta2) bn_max /= 2;   // each CTA of a pair stages half of the weight tile
    h->b_stride = ((bn_max * BK * 2 + 1023) / 1024) * 1024;  // tiles stay 1024-byte aligned (128B swizzle atoms)
    const int budget = CONV_SMEM - 1024 - h->nbuf * OUT_BYTES - (any_ident ? IDENT_BYTES : 0) - BAR_BYTES - STAT_BYTES;
    int nst = budget / ((pair ? 2 : 1) * A_BYTES + h->b_stride);
    h->nstages = nst > MAX_STAGES ? MAX_STAGES : nst;
  }
  if (!any_aux)
    for (int i = 0; i < nseg; ++i) h->seg[i].aux_kind = 0;

  dslb_conv_plan* plan = new (std::nothrow) dslb_conv_plan();
  if (!plan) {
    delete h;
    set_error("out of host memory");
    return DSLB_ENOMEM;
  }
  plan->dev = h;
  plan->total_tiles = tiles;
  plan->flops = flops;
  plan->fast4 = fast4 ? 1 : 0;
  plan->cta2 = cta2 ? 1 : 0;
  static bool attr_set = false;
  if (!attr_set) {
    DSLB_CHECK_CUDA(
        cudaFuncSetAttribute(conv_igemm_cta2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_SMEM));
    DSLB_CHECK_CUDA(
        cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_SMEM));
    DSLB_CHECK_CUDA(
        cudaFuncSetAttribute(conv_igemm_fast4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_SMEM));
    attr_set = true;
  }
  *out = plan;
  return DSLB_OK;
}

extern "C" int dslb_conv_plan_run(const dslb_conv_plan_t* plan, void* stream) {
  DSLB_CHECK_ARG(plan && (plan->dev || plan->halo), "dslb_conv_plan_run: null plan");
  if (plan->halo) return halo_plan_run(plan->halo, stream);
  const int work = plan->dev->pair ? plan->total_tiles / 2 : plan->total_tiles;
  const int grid = work < num_sms() ? work : num_sms();
  if (plan->cta2) {
    // clusters of two CTAs (same TPC): even grid, total_tiles is even by construction
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(grid & ~1));
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = CONV_SMEM;
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
    DSLB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_cta2_kernel, *plan->dev));
    DSLB_CHECK_CUDA(cudaGetLastError());
    return DSLB_OK;
  }
  if (plan->fast4)
    DSLB_CHECK_CUDA(launch_pdl(conv_igemm_fast4_kernel, dim3(grid), dim3(640), CONV_SMEM, (cudaStream_t)stream, *plan->dev));
  else
    DSLB_CHECK_CUDA(launch_pdl(conv_igemm_kernel, dim3(grid), dim3(384), CONV_SMEM, (cudaStream_t)stream, *plan->dev));
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

extern "C" void dslb_conv_plan_destroy(dslb_conv_plan_t* plan) {
  if (!plan) return;
  if (plan->halo) halo_plan_destroy(plan->halo);
  delete plan->dev;
  delete plan;
}

extern "C" double dslb_conv_plan_flops(const dslb_conv_plan_t* plan) { return plan ? plan->flops : 0.0; }
