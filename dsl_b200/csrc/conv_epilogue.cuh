// Lean epilogue shared by the implicit-GEMM conv kernels (conv_igemm.cu, conv_halo.cu): TMEM accumulators -> + shift ->
// + residual / & mask (TMA-loaded into the staging slab) -> bf16 (ReLU fused into the conversion) -> 128B-swizzled
// staging slab, addressed through 32-bit shared-space ld / st.
#pragma once
#include "ptx.cuh"

namespace dslb {

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// two fp32 -> packed bf16x2 (lo = a, hi = b), round to nearest even; the .relu form clamps negatives to +0, which is
// exactly relu-then-round
__device__ __forceinline__ uint32_t cvt_bf16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ uint32_t cvt_relu_bf16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// 0xffff in each half where the bf16 half of m is > 0, else 0
__device__ __forceinline__ uint32_t gt0_mask_bf16x2(uint32_t m) {
  uint32_t r;
  const uint32_t z = 0u;
  asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(m), "r"(z));
  return r;
}

// One 16-channel chunk of the lean epilogue: fp32 accumulators (already loaded from TMEM) -> + shift -> + residual
// (AUX 1) -> bf16x2 (ReLU fused into the conversion) -> & mask (AUX 2) -> the thread's two 16-byte slots of the
// 128B-swizzled staging row. `aux` = this thread's 32 bytes of the TMA-loaded residual / mask tile.
template <bool SHIFT, int AUX, bool RELU>
__device__ __forceinline__ void fast_chunk_math(const uint32_t (&rr)[16], const float4 (&sh)[4], const uint4 (&aux)[2],
                                                uint32_t ad0, uint32_t ad1) {
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rr[j]);
  if (SHIFT) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[4 * j] += sh[j].x;
      v[4 * j + 1] += sh[j].y;
      v[4 * j + 2] += sh[j].z;
      v[4 * j + 3] += sh[j].w;
    }
  }
  const uint32_t aw[8] = {aux[0].x, aux[0].y, aux[0].z, aux[0].w, aux[1].x, aux[1].y, aux[1].z, aux[1].w};
  if (AUX == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[2 * j] += bf16_lo(aw[j]);
      v[2 * j + 1] += bf16_hi(aw[j]);
    }
  }
  uint32_t o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = RELU ? cvt_relu_bf16x2(v[2 * j], v[2 * j + 1]) : cvt_bf16x2(v[2 * j], v[2 * j + 1]);
  if (AUX == 2) {
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] &= gt0_mask_bf16x2(aw[j]);
  }
  sts128(ad0, o[0], o[1], o[2], o[3]);
  sts128(ad1, o[4], o[5], o[6], o[7]);
}

// Lean epilogue over the channel range [cbeg, cend) of one round (r0 = first channel of the round's 128-channel
// window), two 16-channel chunks per iteration so both TMEM loads, the shift loads and the slab loads are in flight
// before the single tcgen05.wait.
template <bool SHIFT, int AUX, bool RELU, bool PAIR>
__device__ __forceinline__ void fast_chunks(uint32_t taddr, int cbeg, int cend, int r0, uint32_t slab_row, uint32_t sw,
                                            const float* __restrict__ shp) {
  int c0 = cbeg;
  for (; PAIR && c0 + 32 <= cend; c0 += 32) {
    uint32_t ra[16], rb[16];
    tmem_ld16(taddr + c0, ra);
    tmem_ld16(taddr + c0 + 16, rb);
    const uint32_t chx = ((c0 - r0) & 63) >> 3;  // 16-byte slot of the chunk's first 8 channels in the 128-byte row
    const uint32_t ad0 = slab_row + ((chx ^ sw) << 4), ad1 = slab_row + (((chx + 1) ^ sw) << 4);
    const uint32_t ad2 = slab_row + (((chx + 2) ^ sw) << 4), ad3 = slab_row + (((chx + 3) ^ sw) << 4);
    uint4 xa[2], xb[2];
    if (AUX) {
      xa[0] = lds128(ad0);
      xa[1] = lds128(ad1);
      xb[0] = lds128(ad2);
      xb[1] = lds128(ad3);
    }
    float4 sa[4], sb[4];
    if (SHIFT) {
      const float4* sp = reinterpret_cast<const float4*>(shp + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sa[j] = __ldg(sp + j);
        sb[j] = __ldg(sp + 4 + j);
      }
    }
    tmem_ld_wait();
    fast_chunk_math<SHIFT, AUX, RELU>(ra, sa, xa, ad0, ad1);
    fast_chunk_math<SHIFT, AUX, RELU>(rb, sb, xb, ad2, ad3);
  }
  for (; c0 < cend; c0 += 16) {
    uint32_t ra[16];
    tmem_ld16(taddr + c0, ra);
    const uint32_t chx = ((c0 - r0) & 63) >> 3;
    const uint32_t ad0 = slab_row + ((chx ^ sw) << 4), ad1 = slab_row + (((chx + 1) ^ sw) << 4);
    uint4 xa[2];
    if (AUX) {
      xa[0] = lds128(ad0);
      xa[1] = lds128(ad1);
    }
    float4 sa[4];
    if (SHIFT) {
      const float4* sp = reinterpret_cast<const float4*>(shp + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) sa[j] = __ldg(sp + j);
    }
    tmem_ld_wait();
    fast_chunk_math<SHIFT, AUX, RELU>(ra, sa, xa, ad0, ad1);
  }
}

template <bool PAIR>
__device__ __forceinline__ void fast_dispatch(int variant, uint32_t taddr, int cbeg, int cend, int r0, uint32_t slab_row,
                                              uint32_t sw, const float* __restrict__ shp) {
#define DSLB_FAST(S_, A_, R_) fast_chunks<S_, A_, R_, PAIR>(taddr, cbeg, cend, r0, slab_row, sw, shp)
  switch (variant) {  // (shift ? 6 : 0) + aux_kind * 2 + relu
    case 0: DSLB_FAST(false, 0, false); break;
    case 1: DSLB_FAST(false, 0, true); break;
    case 2: DSLB_FAST(false, 1, false); break;
    case 3: DSLB_FAST(false, 1, true); break;
    case 4: DSLB_FAST(false, 2, false); break;
    case 5: DSLB_FAST(false, 2, true); break;
    case 6: DSLB_FAST(true, 0, false); break;
    case 7: DSLB_FAST(true, 0, true); break;
    case 8: DSLB_FAST(true, 1, false); break;
    case 9: DSLB_FAST(true, 1, true); break;
    case 10: DSLB_FAST(true, 2, false); break;
    default: DSLB_FAST(true, 2, true); break;
  }
#undef DSLB_FAST
}

}  // namespace dslb
