/*
 * dslb.h — C ABI of libdslb.so: the B200 (sm_100a) kernels behind DSL's dense teacher-student hot path.
 *
 * Nothing like this exists in the reference (chenbinghui1/DSL is pure Python + torch/cuDNN + two mmcv ops,
 * SURVEY.md §2b); each entry point below names the reference Python it replaces (paths relative to the
 * reference root). Conventions:
 *   - plain pointers + sizes, no torch types; every pointer is DEVICE memory unless it says "host";
 *   - the caller owns every buffer (inputs, outputs, workspaces); nothing is allocated per call except by
 *     the *_plan_create functions, which own a small device block holding TMA descriptors;
 *   - asynchronous on the cudaStream_t passed as `void* stream`; no hidden synchronisation;
 *   - return 0 on success, a negative DSLB_E* code otherwise; dslb_last_error() gives a thread-local text.
 * Activations are NHWC bf16 ("pixel-major": [N*H*W][C]); conv weights are packed by dslb_pack_weights into
 * [taps][Cout_pad][Cin] bf16 (K-major rows for the tensor-core B operand).
 */
#ifndef DSLB_H_
#define DSLB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSLB_OK 0
#define DSLB_EINVAL (-1)  /* bad argument / unsupported shape */
#define DSLB_ECUDA (-2)   /* CUDA runtime / driver error      */
#define DSLB_ENOMEM (-3)

#define DSLB_MAX_SEGS 10
/* GroupNorm statistics records: one (sum, sumsq) pair of doubles per (image, group), each record padded to its own
 * 256-byte block so that concurrent L2 atomics from different tiles do not serialise on one cache line. */
#define DSLB_GN_STAT_STRIDE 32

const char* dslb_last_error(void);
int dslb_version(void);

/* ------------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 tensor cores (fprop, and dgrad expressed as fprop on dY with
 * transformed weights). Replaces every nn.Conv2d on the path: mmdet/models/backbones/resnet.py:262-301,
 * mmdet/models/necks/fpn.py:151-202, mmdet/models/dense_heads/anchor_free_head.py:197-217,
 * mmdet/models/dense_heads/fcos_head.py:139-168 (+ their autograd dgrad).
 * One launch executes up to DSLB_MAX_SEGS independent convs ("segments": e.g. the 5 FPN levels x 2 towers
 * of one FCOSHead layer) through one persistent tile scheduler.
 * Epilogue, per output element (n,p,q,c), in this order:
 *     v = acc * scale[c] + shift[c];  v += residual;  if (c < relu_nch) v = max(v,0);
 *     if (relu_mask) v = relu_mask > 0 ? v : 0;
 *     gn_stats[n][c / gn_cpg][0..1] += (v, v*v)   (fp64 atomics on the bf16-rounded output: GroupNorm statistics)
 * ---------------------------------------------------------------------------------------------------- */
typedef struct dslb_conv_seg {
  const void* x;         /* bf16 NHWC [N][H][W][Cin], Cin % 64 == 0                                   */
  const void* w;         /* bf16 packed [R*S][cout_pad][Cin]                                          */
  void* y;               /* output, pixel-major rows of ldc elements (bf16, or fp32 if out_fp32)      */
  const void* residual;  /* bf16, same indexing as y (may alias y = accumulate), or NULL              */
  const void* relu_mask; /* bf16, same indexing as y, or NULL                                         */
  const float* scale;    /* [Cout] or NULL (=1)                                                       */
  const float* shift;    /* [Cout] or NULL (=0)                                                       */
  double* gn_stats;      /* [N][Cout/gn_cpg][DSLB_GN_STAT_STRIDE] ([0]=sum,[1]=sumsq), pre-zeroed, or NULL */
  int32_t N, H, W, Cin;
  int32_t Cout;          /* real output channels                                                      */
  int32_t cout_pad;      /* rows per tap in w; multiple of 16; tiles of <=256                         */
  int32_t R, S, stride, pad;
  int32_t ldc;           /* elements between consecutive output pixels in y/residual/relu_mask        */
  int32_t out_fp32;      /* 0: y is bf16, 1: y is fp32                                                */
  int32_t relu_nch;      /* ReLU on channels c < relu_nch (0 = none)                                  */
  int32_t gn_cpg;        /* channels per GroupNorm group, 8 or 16 (only with gn_stats)                */
  int32_t scatter2;      /* 1: write output pixel (p,q) at (2p,2q) of an [N][Hs][Ws] map (dgrad of a  */
  int32_t Hs, Ws;        /*    stride-2 1x1 conv); y must then be pre-zeroed or accumulated           */
} dslb_conv_seg_t;

typedef struct dslb_conv_plan dslb_conv_plan_t;
int dslb_conv_plan_create(const dslb_conv_seg_t* segs, int nseg, dslb_conv_plan_t** out);
int dslb_conv_plan_run(const dslb_conv_plan_t* plan, void* stream);
void dslb_conv_plan_destroy(dslb_conv_plan_t* plan);
/* algorithmic FLOPs (2*MACs on real channels) of one run of the plan */
double dslb_conv_plan_flops(const dslb_conv_plan_t* plan);

/* ------------------------------------------------------------------------------------------------------
 * Weight gradient of the same convs (autograd wgrad of nn.Conv2d): dW[tap][co][ci] += sum_pix dY * X.
 * dw is fp32 packed [R*S][dw_rows][Cin] and is ACCUMULATED into with atomics (split-K over pixels), so
 * segments that share `dw` (the 5 FPN levels of a shared FCOSHead conv) sum naturally; zero it first.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct dslb_wgrad_seg {
  const void* x;   /* bf16 NHWC [N][H][W][Cin], Cin % 64 == 0                                         */
  const void* dy;  /* bf16 pixel-major [N*Ho*Wo][ldy], ldy % 64 == 0, channels >= Cout are zero       */
  float* dw;       /* fp32 [R*S][dw_rows][Cin]                                                        */
  int32_t N, H, W, Cin;
  int32_t Cout, ldy, dw_rows;
  int32_t R, S, stride, pad;
} dslb_wgrad_seg_t;

typedef struct dslb_wgrad_plan dslb_wgrad_plan_t;
int dslb_wgrad_plan_create(const dslb_wgrad_seg_t* segs, int nseg, dslb_wgrad_plan_t** out);
int dslb_wgrad_plan_run(const dslb_wgrad_plan_t* plan, void* stream);
void dslb_wgrad_plan_destroy(dslb_wgrad_plan_t* plan);
double dslb_wgrad_plan_flops(const dslb_wgrad_plan_t* plan);

#ifdef __cplusplus
}
#endif
#endif /* DSLB_H_ */
