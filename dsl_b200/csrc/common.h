// Host-side helpers shared by the C-ABI translation units: thread-local error text, CUDA error checks,
// and TMA tensor-map encoding through the driver entry point (so libdslb.so has no link-time dependency on
// libcuda.so and loads on a GPU-less build box).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <utility>

#include "../../include/dslb.h"

namespace dslb {

void set_error(const char* fmt, ...);

#define DSLB_CHECK_ARG(cond, ...)  \
  do {                             \
    if (!(cond)) {                 \
      dslb::set_error(__VA_ARGS__); \
      return DSLB_EINVAL;          \
    }                              \
  } while (0)

#define DSLB_CHECK_CUDA(expr)                                                                   \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      dslb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));    \
      return DSLB_ECUDA;                                                                        \
    }                                                                                           \
  } while (0)

int num_sms();

// bf16 tiled tensor map (rank 2..4), 128B swizzle. dims/strides innermost first; strides in bytes for dims 1..
int encode_tiled_bf16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                      const uint32_t* box);
// same with the shared-memory swizzle chosen by the caller: 0 = none (dense box rows), 128 = 128B swizzle
int encode_tiled_bf16_swz(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                          const uint32_t* box, int swizzle_bytes);
// bf16 im2col tensor map on an NHWC tensor [N][H][W][C]: 64 channels x `pixels` output pixels per load.
int encode_im2col_bf16(CUtensorMap* tm, const void* base, int N, int H, int W, int C, int R, int S, int stride,
                       int pad, int pixels);

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// true unless DSLB_NO_PDL is set: launch the tensor-core kernels with programmatic stream serialization
bool pdl_enabled();

// Launch `kernel` so that it may begin (up to its griddepcontrol.wait) before the previous kernel of the stream has
// finished. The kernel must call pdl_wait() before touching anything a predecessor wrote.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace dslb
