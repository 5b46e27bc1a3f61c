// 3x3 stride-1 implicit-GEMM convolution for NARROW layers (Cin, Cout in {64, 128}: the bottleneck conv2 of ResNet
// layer1 / layer2, resnet.py:262-301, and its data gradient) with HALO tiles instead of im2col tiles.
//
// Why: with 64 channels one im2col k-iteration is a 128 x 64 x 64 MMA (64 tensor-core clocks) fed by a 16 KiB A tile +
// an 8 KiB B tile pulled from L2 — 9 taps re-read the same pixels nine times (464 MB of L2->SM traffic for 69 MB of
// HBM traffic at C2, profiles/r01d_notes.md), so the kernel is L2->SM bandwidth bound at ~300 TF/s. Here an output
// tile is a 16 x 8 pixel PATCH and its input is loaded ONCE per k-chunk as three shifted copies of the 18-row halo:
//
//     A_dx[yy][x] = in[y0 - 1 + yy][x0 - 1 + dx + x],   yy in [0,18), x in [0,8), dx in {0,1,2}
//
// one tiled TMA box (64 ch x 8 w x 18 h, out-of-bounds = the conv's zero padding) each, landing as 144 K-major rows of
// 128 B (row = yy*8 + x, 128B swizzle). The A operand of tap (dy, dx) is then copy dx at row offset dy*8, i.e. at byte
// offset dy*1024 — every tcgen05 descriptor start stays on a 1024-byte swizzle-atom boundary, no per-row phase games.
// L2->SM traffic per tile: 3 x 18 KiB instead of 9 x 16 KiB, and the 64 x 64 weights (72 KiB for all nine taps) stay
// RESIDENT in shared memory for the whole kernel (128-channel layers stream their 16 KiB weight tiles through a ring).
//
//   Warp roles   warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-11 epilogue (2 groups x 4).
//   Epilogue     the lean path of conv_epilogue.cuh: + shift, ReLU in the bf16 conversion, & mask (dgrad: TMA-loaded
//                into the staging slab) -> 128B-swizzled slab -> ONE 4-D TMA store (64 ch x 8 w x 16 h) per slab, which
//                also clips patches that hang over the right / bottom image edge.
//   Scheduling   persistent, grid = min(#patches, #SMs), TMEM accumulators double-buffered.
#include <new>
#include <stdlib.h>
#include <string.h>

#include "common.h"
#include "conv_epilogue.cuh"
#include "conv_halo.h"
#include "ptx.cuh"

namespace dslb {

constexpr int HP_ROWS = 16, HP_COLS = 8;              // output patch
constexpr int HALO_ROWS = HP_ROWS + 2;                // 18 input rows per copy
constexpr int COPY_BYTES = HALO_ROWS * HP_COLS * 128; // 18 432 = 18 swizzle atoms
constexpr int A_STAGE = 3 * COPY_BYTES;               // 55 296
constexpr int NA = 2;                                 // halo stages
constexpr int SLAB = 128 * 128;                       // staging slab: 128 px x 64 ch bf16
constexpr int HALO_BAR_BYTES = 256;
constexpr int HALO_SMEM_MAX = 232448;

struct alignas(128) HaloParams {
  CUtensorMap tmA;    // input  [N][H][W][Cin]: box 64 x 8 x 18 x 1
  CUtensorMap tmB;    // weights [9][cout][Cin]: box 64 x bn x 1
  CUtensorMap tmY;    // output [N][H][W][ldc]: box 64 x 8 x 16 x 1
  CUtensorMap tmAux;  // ReLU mask, same geometry as tmY (aux == 2)
  const float* shift;
  int tiles_x, tiles_per_img, total_tiles;
  int cin_chunks, bn;
  int relu, aux;      // aux: 0 none, 2 mask tile by TMA
  int b_resident;     // all 9 * cin_chunks weight tiles stay in shared memory
  int nb;             // weight ring depth (streamed mode)
};

__device__ __forceinline__ void tma_load_4d(const void* desc, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* desc, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(desc)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

struct HaloTile {
  int n, y0, x0;
};
__device__ __forceinline__ HaloTile halo_tile(const HaloParams& P, int t) {
  HaloTile h;
  h.n = t / P.tiles_per_img;
  const int rem = t - h.n * P.tiles_per_img;
  const int ty = rem / P.tiles_x;
  h.y0 = ty * HP_ROWS;
  h.x0 = (rem - ty * P.tiles_x) * HP_COLS;
  return h;
}

__global__ void __launch_bounds__(384, 1) conv3x3_halo_kernel(const __grid_constant__ HaloParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int bn = P.bn, kch = P.cin_chunks;
  const int b_tile = bn * 128;                                    // one (tap, k-chunk) weight tile
  const int nb = P.b_resident ? 9 * kch : P.nb;
  uint8_t* const sA = smem;
  uint8_t* const sB = sA + NA * A_STAGE;
  uint8_t* const sOut = sB + nb * b_tile;
  const int nslab = bn / 64;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(sOut + nslab * SLAB);
  uint64_t* const a_full = bars;            // [NA]
  uint64_t* const a_empty = bars + 2;       // [NA]
  uint64_t* const b_full = bars + 4;        // [<= 8] streamed ring; [0] = "resident weights landed"
  uint64_t* const b_empty = bars + 12;      // [<= 8]
  uint64_t* const tfull = bars + 20;        // [2]
  uint64_t* const tempty = bars + 22;       // [2]
  uint64_t* const auxfull = bars + 24;      // [2] per slab
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = P.total_tiles;

  pdl_launch_dependents();
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NA; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);   // 8 epilogue warps
      mbar_init(&auxfull[i], 1);
    }
    fence_mbar_init();
  } else if (warp == 2) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // ------------------------------------------------------------------ producer
      if (P.b_resident) {
        mbar_expect_tx(&b_full[0], 9 * kch * b_tile);
        for (int kc = 0; kc < kch; ++kc)
          for (int tap = 0; tap < 9; ++tap) tma_load_3d(&P.tmB, &b_full[0], sB + (kc * 9 + tap) * b_tile, kc * 64, 0, tap);
      }
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const HaloTile h = halo_tile(P, t);
        for (int kc = 0; kc < kch; ++kc) {
          mbar_wait(&a_empty[as], aph ^ 1);
          mbar_expect_tx(&a_full[as], A_STAGE);
          for (int dx = 0; dx < 3; ++dx)
            tma_load_4d(&P.tmA, &a_full[as], sA + as * A_STAGE + dx * COPY_BYTES, kc * 64, h.x0 - 1 + dx, h.y0 - 1, h.n);
          if (++as == NA) {
            as = 0;
            aph ^= 1;
          }
          if (!P.b_resident) {
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&b_empty[bs], bph ^ 1);
              mbar_expect_tx(&b_full[bs], b_tile);
              tma_load_3d(&P.tmB, &b_full[bs], sB + bs * b_tile, kc * 64, 0, tap);
              if (++bs == nb) {
                bs = 0;
                bph ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ------------------------------------------------------------------ MMA issuer
      const uint32_t idesc = make_idesc_bf16(128, bn, 0, 0);
      if (P.b_resident) {
        mbar_wait(&b_full[0], 0);
        tc_fence_after();
      }
      int as = 0, bs = 0, it = 0;
      uint32_t aph = 0, bph = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 128;
        for (int kc = 0; kc < kch; ++kc) {
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + as * A_STAGE);
          for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap - dy * 3;
            uint32_t b_base;
            if (P.b_resident) {
              b_base = smem_u32(sB + (kc * 9 + tap) * b_tile);
            } else {
              mbar_wait(&b_full[bs], bph);
              tc_fence_after();
              b_base = smem_u32(sB + bs * b_tile);
            }
            const uint32_t a_tap = a_base + dx * COPY_BYTES + dy * (HP_COLS * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem, make_sdesc(a_tap + k * 32, 16, 1024), make_sdesc(b_base + k * 32, 16, 1024), idesc,
                        (kc | tap | k) != 0);
            if (!P.b_resident) {
              umma_commit(&b_empty[bs]);
              if (++bs == nb) {
                bs = 0;
                bph ^= 1;
              }
            }
          }
          umma_commit(&a_empty[as]);
          if (++as == NA) {
            as = 0;
            aph ^= 1;
          }
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else if (warp >= 4) {
    // -------------------------------------------------------------------- epilogue: 2 groups x 4 warps
    const int ew = warp & 3;            // TMEM lane quadrant
    const int eg = (warp - 4) >> 2;     // group 0 / 1
    const int et = ew * 32 + lane;      // patch pixel (row of the tile) owned by this thread
    // bn = 128: group g owns channels [64g, 64g + 64) and slab g. bn = 64: the two groups split the 64 channels of
    // slab 0 (32 each) and hand it over with 256-thread barriers.
    const bool shared = (bn == 64);
    const int cbeg = shared ? eg * 32 : eg * 64;
    const int cend = shared ? cbeg + 32 : cbeg + 64;
    const int slab_i = shared ? 0 : eg;
    uint8_t* const slab = sOut + slab_i * SLAB;
    const uint32_t slab_row = smem_u32(slab) + et * 128;
    const uint32_t sw = et & 7;
    const bool leader = (ew == 0 && lane == 0) && (!shared || eg == 0);
    const int nthr = shared ? 256 : 128;
    const int bar_a = shared ? 1 : 1 + eg, bar_b = shared ? 3 : 3 + eg;
    uint64_t* const auxbar = &auxfull[slab_i];
    uint32_t aux_par = 0;
    auto aux_issue = [&](int tn) {
      if (tn >= total || P.aux != 2) return;
      const HaloTile h = halo_tile(P, tn);
      mbar_expect_tx(auxbar, SLAB);
      tma_load_4d(&P.tmAux, auxbar, slab, slab_i * 64, h.x0, h.y0, h.n);
    };
    if (leader) aux_issue(blockIdx.x);
    const float* __restrict__ shp = P.shift;
    const int variant = (shp ? 6 : 0) + P.aux * 2 + (P.relu ? 1 : 0);
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const HaloTile h = halo_tile(P, t);
      const int acc = it & 1;
      if (P.aux == 2) {
        mbar_wait(auxbar, aux_par);   // the mask tile is in the slab (and the previous store has drained)
        aux_par ^= 1;
      } else {
        asm volatile("bar.sync %0, %1;" ::"r"(bar_a), "r"(nthr) : "memory");   // slab free: leader drained the store
      }
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * 128;
      fast_dispatch<true>(variant, taddr, cbeg, cend, 0, slab_row, sw, shp);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      fence_proxy_async();
      asm volatile("bar.sync %0, %1;" ::"r"(bar_b), "r"(nthr) : "memory");
      if (leader) {
        tma_store_4d(&P.tmY, slab, slab_i * 64, h.x0, h.y0, h.n);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        aux_issue(t + gridDim.x);
      }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace dslb

using namespace dslb;

struct dslb_halo_plan {
  HaloParams p;
  int smem;
  int grid;
};

bool dslb::halo_eligible(const dslb_conv_seg_t& s) {
  if (getenv("DSLB_NO_HALO")) return false;
  if (s.R != 3 || s.S != 3 || s.stride != 1 || s.pad != 1) return false;
  if (!(s.Cin == 64 || s.Cin == 128) || !(s.Cout == 64 || s.Cout == 128) || s.cout_pad != s.Cout) return false;
  if (s.out_fp32 || s.scatter2 || s.gn_stats || s.gnb_x || s.scale || s.residual) return false;
  if (!(s.relu_nch == 0 || s.relu_nch >= s.Cout) || s.ldc % 8 != 0) return false;
  if (s.W < HP_COLS || s.H < 1) return false;
  return true;
}

int dslb::halo_plan_create(const dslb_conv_seg_t& s, dslb_halo_plan** out) {
  dslb_halo_plan* h = new (std::nothrow) dslb_halo_plan();
  if (!h) {
    set_error("out of host memory");
    return DSLB_ENOMEM;
  }
  memset(h, 0, sizeof(*h));
  HaloParams& P = h->p;
  const uint64_t ad[4] = {(uint64_t)s.Cin, (uint64_t)s.W, (uint64_t)s.H, (uint64_t)s.N};
  const uint64_t as[3] = {(uint64_t)s.Cin * 2, (uint64_t)s.W * s.Cin * 2, (uint64_t)s.H * s.W * s.Cin * 2};
  const uint32_t ab[4] = {64, HP_COLS, HALO_ROWS, 1};
  int rc = encode_tiled_bf16(&P.tmA, s.x, 4, ad, as, ab);
  const uint64_t wd[3] = {(uint64_t)s.Cin, (uint64_t)s.cout_pad, 9};
  const uint64_t ws[2] = {(uint64_t)s.Cin * 2, (uint64_t)s.cout_pad * s.Cin * 2};
  const uint32_t wb[3] = {64, (uint32_t)s.Cout, 1};
  if (rc == DSLB_OK) rc = encode_tiled_bf16(&P.tmB, s.w, 3, wd, ws, wb);
  const uint64_t yd[4] = {(uint64_t)s.Cout, (uint64_t)s.W, (uint64_t)s.H, (uint64_t)s.N};
  const uint64_t ys[3] = {(uint64_t)s.ldc * 2, (uint64_t)s.W * s.ldc * 2, (uint64_t)s.H * s.W * s.ldc * 2};
  const uint32_t yb[4] = {64, HP_COLS, HP_ROWS, 1};
  if (rc == DSLB_OK) rc = encode_tiled_bf16(&P.tmY, s.y, 4, yd, ys, yb);
  if (rc == DSLB_OK && s.relu_mask) rc = encode_tiled_bf16(&P.tmAux, s.relu_mask, 4, yd, ys, yb);
  if (rc != DSLB_OK) {
    delete h;
    return rc;
  }
  P.shift = s.shift;
  P.tiles_x = cdiv(s.W, HP_COLS);
  P.tiles_per_img = P.tiles_x * cdiv(s.H, HP_ROWS);
  P.total_tiles = P.tiles_per_img * s.N;
  P.cin_chunks = s.Cin / 64;
  P.bn = s.Cout;
  P.relu = s.relu_nch >= s.Cout ? 1 : 0;
  P.aux = s.relu_mask ? 2 : 0;
  const int b_tile = P.bn * 128;
  const int fixed = 1024 + NA * A_STAGE + (P.bn / 64) * SLAB + HALO_BAR_BYTES;
  const int room = HALO_SMEM_MAX - fixed;
  if (9 * P.cin_chunks * b_tile <= room) {
    P.b_resident = 1;
    P.nb = 0;
    h->smem = fixed + 9 * P.cin_chunks * b_tile;
  } else {
    P.b_resident = 0;
    P.nb = room / b_tile > 8 ? 8 : room / b_tile;
    if (P.nb < 2) {
      delete h;
      set_error("halo conv: no room for the weight ring");
      return DSLB_EINVAL;
    }
    h->smem = fixed + P.nb * b_tile;
  }
  h->grid = P.total_tiles < num_sms() ? P.total_tiles : num_sms();
  static bool attr_set = false;
  if (!attr_set) {
    DSLB_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_SMEM_MAX));
    attr_set = true;
  }
  *out = h;
  return DSLB_OK;
}

int dslb::halo_plan_run(const dslb_halo_plan* h, void* stream) {
  DSLB_CHECK_CUDA(launch_pdl(conv3x3_halo_kernel, dim3(h->grid), dim3(384), (size_t)h->smem, (cudaStream_t)stream, h->p));
  DSLB_CHECK_CUDA(cudaGetLastError());
  return DSLB_OK;
}

void dslb::halo_plan_destroy(dslb_halo_plan* h) { delete h; }
