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
int dslb_version(void);   /* 103 = this header (102 + GroupNorm backward sums in the conv epilogue: gnb_* / gsums; dslb_sgd_ema_step) */

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
 *     gn_stats[n][c / gn_cpg][0..1] += (v, v*v)   (GroupNorm statistics: per-tile fp32 partial sums of the fp32 v, i.e.
 *                                                  before the bf16 rounding of the store, added with fp64 atomics)
 * ---------------------------------------------------------------------------------------------------- */
typedef struct dslb_conv_seg {
  const void* x;         /* bf16 NHWC [N][H][W][Cin], Cin % 64 == 0                                   */
  const void* w;         /* bf16 packed [R*S][cout_pad][Cin]                                          */
  void* y;               /* output, pixel-major rows of ldc elements (bf16, or fp32 if out_fp32)      */
  const void* residual;  /* bf16, same indexing as y (may alias y = accumulate), or NULL              */
  const void* relu_mask; /* bf16, same indexing as y, or NULL                                         */
  const float* scale;    /* [Cout] or NULL (=1)                                                       */
  const float* shift;    /* [Cout] or NULL (=0)                                                       */
  double* gn_stats;      /* [N][Cout/gn_cpg][DSLB_GN_STAT_STRIDE] ([0]=sum,[1]=sumsq), pre-zeroed, or NULL (bf16 or fp32 y) */
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
  /* GroupNorm BACKWARD sums in the epilogue (gnb_x != NULL; excludes gn_stats): this conv is the dgrad whose output dz
   * is the gradient w.r.t. relu(GroupNorm(x)) of the layer below (mmcv ConvModule conv -> GN -> ReLU,
   * anchor_free_head.py:95-139). Per image n and group g, with xhat = (x - mean) * rstd and
   * dy = bf16(dz) * [xhat * gamma + beta > 0], the epilogue adds
   *     gnb_sums[n][g][0] += sum gamma * dy,     gnb_sums[n][g][1] += sum gamma * dy * xhat
   * (fp64 atomics), which dslb_gn_bwd then uses instead of running its own reduction pass over x and dz. */
  const void* gnb_x;       /* bf16 pre-norm map [N*Ho*Wo][Cout] (dense), Cout % 16 == 0, cout_pad == Cout        */
  const float* gnb_mr;     /* [N][Cout/gn_cpg][4]: (mean, rstd, -, -) as the forward apply left them              */
  const float* gnb_gamma;  /* [Cout]                                                                             */
  const float* gnb_beta;   /* [Cout]                                                                             */
  double* gnb_sums;        /* [N][Cout/gn_cpg][DSLB_GN_STAT_STRIDE], pre-zeroed                                   */
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

/* ------------------------------------------------------------------------------------------------------
 * HBM-bound glue around the convs (each replaces the torch ops named).
 * ---------------------------------------------------------------------------------------------------- */
/* NCHW fp32 -> NHWC bf16 with channels zero-padded to Cpad (input of the plugin modules: the reference's tensors
 * are NCHW fp32, mmdet/models/detectors/single_stage.py:136-141). */
int dslb_nchw_to_nhwc_bf16(const float* x, void* y, int N, int C, int H, int W, int Cpad, void* stream);
/* pixel-major rows of `ld` elements (bf16, or fp32 if x_is_fp32) -> NCHW fp32, first C channels. */
int dslb_nhwc_to_nchw_f32(const void* x, float* y, int N, int C, int H, int W, int ld, int x_is_fp32, void* stream);
/* The whole stem in one kernel: 7x7/2 pad-3 conv (3 -> 64) + frozen BatchNorm + ReLU (resnet.py:597-610,630-637) from
 * the NCHW fp32 image to NHWC bf16 [N][Ho][Wo][64]. The im2col tile is assembled in shared memory and fed to
 * tcgen05.mma; w is the fp32 OIHW master weight [64][3][7][7], bn_* the frozen BatchNorm tensors [64]. */
size_t dslb_stem_workspace_bytes(int N, int H, int W); /* NHWC4 bf16 copy of the image: N*H*W*8 bytes, 16-byte aligned */
int dslb_stem_conv(const float* img, const float* w, const float* bn_gamma, const float* bn_beta, const float* bn_mean,
                   const float* bn_var, float eps, void* workspace, void* out, int N, int H, int W, void* stream);
/* nn.MaxPool2d(3, 2, 1) (resnet.py:611), NHWC bf16. */
int dslb_maxpool3x3s2(const void* x, void* y, int N, int H, int W, int C, void* stream);
/* Scale-invariant extra input of the DSL runner (mmdet/runner/hooks/semi_epoch_based_runner.py:186-204): out (C,H,W)
 * fp32 = zeros with F.interpolate(img, (int(H/2), int(W/2)), mode="bilinear") of img (C,H,W) in the top-left corner. */
int dslb_si_half_image(const float* img, float* out, int C, int H, int W, void* stream);
/* FPN top-down: dst += nearest_upsample(src) (necks/fpn.py:163-172) and its backward w.r.t. src. */
int dslb_upsample_add(void* dst, const void* src, int N, int H, int W, int h, int w, int C, void* stream);
int dslb_upsample_add_bwd(void* dsrc, const void* ddst, int N, int H, int W, int h, int w, int C, void* stream);
/* bf16 elementwise over n elements (n % 8 == 0): mode 0 y=relu(x); 1 y=(m>0)?x:0 (ReLU backward); 2 y=x+m. */
int dslb_relu_family(const void* x, const void* m, void* y, long long n, int mode, void* stream);

/* GroupNorm(groups) + ReLU after a conv (mmcv ConvModule, anchor_free_head.py:95-139; norm_cfg fcos_head.py:82),
 * driven by the statistics the conv epilogue accumulated. Up to DSLB_MAX_SEGS maps per launch. */
typedef struct dslb_gn_seg {
  const void* x;       /* bf16 pre-norm conv output [N*HW][C]                                            */
  void* y;             /* fwd: bf16 relu(gn(x)); bwd: bf16 gradient w.r.t. x                             */
  const void* dz;      /* bwd only: bf16 gradient w.r.t. the post-ReLU output                            */
  const double* stats; /* [N][groups][DSLB_GN_STAT_STRIDE]                                               */
  const float* gamma;  /* [C]                                                                            */
  const float* beta;   /* [C]                                                                            */
  double* red;         /* bwd only: [N][C][2] scratch, pre-zeroed (sum dy, sum dy*xhat)                  */
  float* dbias;        /* bwd only: [C], += gradient of the conv bias in front of the norm               */
  float* mr;           /* [N][groups][4] fp32 scratch: (mean, rstd) written by the forward apply and re-used by
                          the backward, which adds its two per-group reduction constants                       */
  int32_t N, HW;
  double* gsums;       /* bwd only, or NULL: [N][groups][DSLB_GN_STAT_STRIDE] group sums a conv epilogue accumulated
                          (dslb_conv_seg_t::gnb_sums). When EVERY segment has them dslb_gn_bwd skips its reduction pass:
                          one apply launch computes the two group constants from gsums and also produces `red`.   */
} dslb_gn_seg_t;
int dslb_gn_apply_relu(const dslb_gn_seg_t* segs, int nseg, int C, int groups, float eps, void* stream);
/* same result, faster: driven by the block table of dslb_gn_bwd_plan (C must be 256) */
int dslb_gn_apply_relu_tab(const dslb_gn_seg_t* segs, int nseg, int C, int groups, float eps, const int* blk_tab_dev,
                           int nblocks, void* stream);
/* Accurate ("bf16x3", split-bf16) FCOSHead mode, inference only: north_star asks for FCOSHead outputs within 1e-3 of the
 * fp32 reference (fcos_head.py:118-168 runs in fp32). A value v travels as hi = bf16(v), lo = bf16(v - hi); a conv over
 * the 3C-channel rows [hi | lo | hi] against packed weights [w_hi | w_hi | w_lo] accumulates hi*w_hi + lo*w_hi + hi*w_lo
 * in fp32 on the tensor core (operand precision ~2^-17). The tower maps stay fp32 (conv segments with out_fp32 = 1 AND
 * gn_stats), and this apply turns them into the next layer's split operand:
 *     y[pix][3C] = split(relu(GroupNorm(x)))   with x fp32 [N*HW][C] (seg.x), y bf16 (seg.y). */
int dslb_gn_apply_relu_split(const dslb_gn_seg_t* segs, int nseg, int C, int groups, float eps, void* stream);
/* bf16 map [npix][C] -> split operand [npix][3C] = [x | 0 | x] (the FPN outputs entering the accurate head). */
int dslb_bf16_to_split(const void* x, void* y, long long npix, int C, void* stream);
/* backward: dslb_gn_bwd_blocks -> size of the block table; dslb_gn_bwd_plan fills a HOST table of 2*blocks ints the
 * caller copies to the device once; dslb_gn_bwd runs the reduce + apply passes (C must be 256). */
int dslb_gn_bwd_blocks(const dslb_gn_seg_t* segs, int nseg);
int dslb_gn_bwd_plan(const dslb_gn_seg_t* segs, int nseg, int* blk_tab_host);
int dslb_gn_bwd(const dslb_gn_seg_t* segs, int nseg, int C, int groups, float eps, const int* blk_tab_dev, int nblocks,
                void* stream);
/* dgamma[c] += sum_n red[n][c][1]; dbeta[c] += sum_n red[n][c][0] */
int dslb_gn_bwd_params(const double* red, float* dgamma, float* dbeta, int N, int C, void* stream);

/* OIHW fp32 master weights -> packed bf16 [R*S][rows_pad][cols_pad] (zero padded), optionally scaled per output
 * channel (frozen-BN fold). transpose=1 builds the dgrad operand (180-degree rotated taps, in/out swapped). */
int dslb_pack_weight(const float* w, void* out, int O, int I, int R, int S, int rows_pad, int cols_pad,
                     const float* oscale, int transpose, void* stream);
/* packed fp32 wgrad [R*S][rows][I] -> OIHW fp32 (x oscale[o]); accumulate=1 adds into g. */
int dslb_unpack_wgrad(const float* dw, float* g, int O, int I, int R, int S, int rows, const float* oscale,
                      int accumulate, void* stream);
/* frozen BatchNorm2d in eval mode folded to y = x*scale + shift (resnet.py:647-656). */
int dslb_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* scale,
                 float* shift, int C, void* stream);
/* ---- multi-tensor versions: ONE launch for every conv of a network (the derived-operand refresh after an optimizer /
 * EMA step, and packed-wgrad -> OIHW gradient views at the end of the backward). A plan copies its descriptor table
 * to the device once; pointers must stay valid for the plan's lifetime. */
typedef struct dslb_pack_desc {
  const float* w;          /* OIHW fp32 master weight                                                          */
  void* out;               /* packed bf16 block [taps][rows_pad][cols_pad] (mode 2: [1][rows_pad][cols_pad])    */
  const float* bn_gamma;   /* frozen BatchNorm to fold in (all four, [O]) or NULL                               */
  const float* bn_beta;
  const float* bn_mean;
  const float* bn_var;
  float* scale_out;        /* [O] folded BN scale / shift for the conv epilogue (written by modes 0 and 2)      */
  float* shift_out;
  int32_t O, I, R, S;
  int32_t rows_pad, cols_pad;
  int32_t row_off, col_off; /* where the real block starts inside the packed block (shared operands)            */
  int32_t mode;            /* 0 fprop [tap][o][i]; 1 dgrad [tap][i][o], taps rotated 180 deg; 2 stem [o][(r*S+s)*I+i] */
  int32_t fill_padding;    /* 1: rewrite the zero padding too; 0: touch only the real sub-block                 */
  float bn_eps;
  int32_t w_ld;            /* input channels per output-channel row of `w` in memory (0 = I). With w pointing at input
                              channel i0 of a wider OIHW tensor and w_ld = its full width, an input-channel SLICE is packed
                              (RLA_Bottleneck.conv1 acts on cat(x, h), resnet_rla.py:108-110: one operand per half)  */
} dslb_pack_desc_t;
typedef struct dslb_unpack_desc {
  const float* dw;         /* packed fp32 wgrad [R*S][rows][I]                                                  */
  float* g;                /* OIHW fp32 gradient, overwritten: g[o][i][r][s] = dw[r*S+s][row_off+o][i] * bn scale */
  const float* bn_gamma;   /* folded BN (chain rule through w*scale) or NULL                                     */
  const float* bn_var;
  int32_t O, I, R, S;
  int32_t rows, row_off;
  float bn_eps;
  int32_t dw_ld;           /* columns per packed row of dw (0 = I): > I when the conv ran on a channel-padded input   */
  int32_t g_ld;            /* input channels per output-channel row of g in memory (0 = I): slice of a wider OIHW tensor */
} dslb_unpack_desc_t;
typedef struct dslb_table_plan dslb_table_plan_t;
int dslb_pack_plan_create(const dslb_pack_desc_t* descs_host, int n, dslb_table_plan_t** out);
int dslb_unpack_plan_create(const dslb_unpack_desc_t* descs_host, int n, dslb_table_plan_t** out);
int dslb_table_plan_run(const dslb_table_plan_t* plan, void* stream);
void dslb_table_plan_destroy(dslb_table_plan_t* plan);
/* Epilogue constants of the fused conv_reg + conv_centerness predictor (fcos_head.py:157-166): for level l,
 * rc_scale[l][0..3] = scales[l] * level_mult[l] (level_mult = stride in eval mode, 1 in training), rc_scale[l][4..7] = 1,
 * rc_shift[l][j] = bias5[j] * rc_scale[l][j] with bias5 = (conv_reg.bias[0..3], conv_centerness.bias); scale_vals[l] =
 * scales[l]. `scales` points at scales.0.scale; consecutive levels are scale_stride floats apart. */
int dslb_fcos_regctr_affine(const float* scales, int scale_stride, const float* reg_bias, const float* ctr_bias,
                            const float* level_mult, float* rc_scale, float* rc_shift, float* scale_vals, int nlevels,
                            void* stream);
/* ------------------------------------------------------------------------------------------------------
 * RLA_ResNet (mmdet/models/backbones/resnet_rla.py; the backbone of configs/fcos_semi/RLA_*.py:3-13).
 * The convs of RLA_Bottleneck (:71-137) and the shared conv_out / recurrent_conv (:259-260, 303-311) run through
 * dslb_conv_plan_* / dslb_wgrad_plan_*; the recurrent state h (rla_channel = 32) is stored as bf16 NHWC with
 * 64-channel rows (channels 32..63 zero) so it is a regular tensor-core operand.
 * dslb_rla_state_fwd:  hb = tanh(BatchNorm_eval(pre)),  pre = h_old' + y_out  (resnet_rla.py:306-310), where
 *   h_old' = AvgPool2d(2,2)(h_old) in the first block of a strided stage (pool = 1, :93-95,131-132; h_old is then
 *   [N][2Ho][2Wo][64]) and h_old otherwise ([N][Ho][Wo][64]); y_out, hb: [N][Ho][Wo][64].
 * dslb_rla_state_bwd:  autograd of the above. d_hb = gradient w.r.t. hb; writes d_pre (gradient w.r.t. pre = w.r.t.
 *   y_out = w.r.t. an un-pooled h_old) and, with pool = 1, dh_old = the pooled gradient 0.25 * d_pre spread over each
 *   2x2 source block; dgamma / dbeta (both or neither; [32] fp32) are ACCUMULATED into.
 * ---------------------------------------------------------------------------------------------------- */
int dslb_rla_state_fwd(const void* h_old, const void* y_out, const float* bn_gamma, const float* bn_beta,
                       const float* bn_mean, const float* bn_var, float eps, void* hb, int N, int Ho, int Wo, int pool,
                       void* stream);
int dslb_rla_state_bwd(const void* d_hb, const void* hb, const void* h_old, const void* y_out, const float* bn_gamma,
                       const float* bn_mean, const float* bn_var, float eps, void* d_pre, void* dh_old, float* dgamma,
                       float* dbeta, int N, int Ho, int Wo, int pool, void* stream);
/* Gradients of BatchNorm affine parameters that are trainable while the statistics are frozen (RLA_ResNet: norm_eval=True
 * keeps running stats, but only frozen_stages' parameters have requires_grad=False, resnet_rla.py:361-375,389-399).
 * The BatchNorm is folded into its conv (W' = W * gamma / sigma), so for output channel c
 *     dgamma[c] = (<dW'[c], W[c]> - mean[c] * dbeta[c]) / sqrt(var[c] + eps),   dbeta[c] = sum_pix dy[c] (dslb_colsum)
 * with dW' the packed fp32 weight gradient [R*S][rows][dw_ld] of the folded conv and W the OIHW master weight. A conv on
 * a concatenated input (RLA_Bottleneck.conv1 on cat(x, h)) passes one (dW', W-slice) piece per part. One launch per plan. */
typedef struct dslb_bn_grad_desc {
  const float* dw0;        /* piece 0: packed wgrad, rows0 rows of dw_ld0 columns per tap                           */
  const float* w0;         /*          master weight (pointer to the slice's first input channel)                    */
  const float* dw1;        /* piece 1 or NULL                                                                        */
  const float* w1;
  const float* mean;       /* [O] running_mean                                                                       */
  const float* var;        /* [O] running_var                                                                        */
  const float* dbeta;      /* [O] already reduced                                                                    */
  float* dgamma;           /* [O] overwritten                                                                        */
  int32_t O, R, S;
  int32_t I0, dw_ld0, w_ld0, rows0;   /* slice width; columns per dw row (0 = I0); input channels per w row (0 = I0) */
  int32_t I1, dw_ld1, w_ld1, rows1;
  float bn_eps;
} dslb_bn_grad_desc_t;
typedef struct dslb_bn_grad_plan dslb_bn_grad_plan_t;
int dslb_bn_grad_plan_create(const dslb_bn_grad_desc_t* descs_host, int n, dslb_bn_grad_plan_t** out);
int dslb_bn_grad_plan_run(const dslb_bn_grad_plan_t* plan, void* stream);
void dslb_bn_grad_plan_destroy(dslb_bn_grad_plan_t* plan);
/* y[n][2p][2q][:] = x[n][p][q][:], every other pixel of the [N][H][W][C] bf16 map zero: turns the data gradient of a
 * stride-2 conv into a stride-1 tensor-core dgrad over the zero-upsampled dY (FPN P6/P7 convs, necks/fpn.py:192-201). */
int dslb_zero_upsample2(const void* x, void* y, int N, int h, int w, int H, int W, int C, void* stream);
/* Small glue so that no framework kernel sits inside the captured step: buffer clears (a memset node in the CUDA graph),
 * dst[idx[i]] = src[i] (the reg / centerness bias and Scale gradients into the flat gradient buffer), fp64 loss sums ->
 * the fp32 scalars the detector returns (detectors/base.py:201-206 logs them). */
int dslb_zero(void* p, size_t bytes, void* stream);
int dslb_scatter_f32(float* dst, const int64_t* idx, const float* src, int n, void* stream);
int dslb_f64_to_f32(const double* src, float* dst, int n, void* stream);
/* out[c] += sum_p x[p][c] over a pixel-major bf16 matrix (conv bias gradient). */
int dslb_colsum(const void* x, float* out, long long npix, int ld, int C, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Dense pseudo-label path: FCOSHead.loss (mmdet/models/dense_heads/fcos_head.py:170-338) =
 * get_points (:550-560, anchor_free_head.py:287-321) + get_targets/_get_target_single (:562-705) for the
 * (pseudo-)GT boxes and for the ignore boxes (:208-215) + ignore / unlabeled weights (:217-235, :297-307) +
 * centerness_target (:707-726) + FocalLoss (losses/focal_loss.py:11-56) + GIoULoss (losses/iou_loss.py:85-102,
 * 329-366; core/bbox/iou_calculators/iou2d_calculator.py:214-260; core/bbox/transforms.py:119-162) + sigmoid BCE
 * (losses/cross_entropy_loss.py:73-112) + the scale-invariant soft loss (:312-333), forward AND backward.
 * Flat point order everywhere = the reference's: level-major, then image, then row-major (y, x).
 * ---------------------------------------------------------------------------------------------------- */
typedef struct dslb_fcos_level {
  const float* cls;     /* fp32 logits, pixel-major [B*h*w][ld_cls]                                       */
  const float* regctr;  /* fp32 [B*h*w][8]: 0-3 bbox_pred = relu(scale*conv_reg), 4 centerness logit      */
  void* dcls_bf16;      /* out (or NULL): bf16 d loss/d logits [B*h*w][ld_dcls] (pad columns untouched)   */
  float* dcls_f32;      /* out (or NULL): fp32 d loss/d logits [B*h*w][C]                                 */
  void* dregctr_bf16;   /* out (or NULL): bf16 [B*h*w][ld_dreg]: 0-3 d loss/d conv_reg, 4 d/d ctr, 5-7 =0 */
  float* dregctr_f32;   /* out (or NULL): fp32 [B*h*w][8]: 0-3 d loss/d bbox_pred, 4 d/d ctr logit        */
  int32_t h, w, stride;
  int32_t ld_cls, ld_dcls, ld_dreg;
  float rr_lo, rr_hi;   /* regress range of the level                                                     */
  float scale;          /* value of the level's Scale parameter (chain rule through relu(scale*x))        */
  float cs_radius;      /* float(stride * center_sample_radius)                                           */
} dslb_fcos_level_t;

/* Target assignment. gt_boxes [sumG][4] fp32, gt_labels [sumG] int64, gt_off [B+1] int32 (all device); ig_* the
 * same for the ignore boxes or NULL. Images [0,n_labeled) are labeled (weight 1), the others x loss_weight.
 * Outputs (flat point order): labels int64 (num_classes = background), bbox_targets [P][4], weights [P] (focal
 * weight: ignore mask x unlabeled weight), ctr_targets [P] (0 on non-positives); counts[0] += num_pos,
 * counts[1] += sum ctr_targets (fp64, pre-zeroed). labels / bbox_targets are bit-exact with the reference. */
int dslb_fcos_targets(const dslb_fcos_level_t* levels, int nlevels, int B, int num_classes, const float* gt_boxes,
                      const int64_t* gt_labels, const int32_t* gt_off, const float* ig_boxes, const int32_t* ig_off,
                      int center_sampling, int norm_on_bbox, float loss_weight, int n_labeled, int64_t* labels,
                      float* bbox_targets, float* weights, float* ctr_targets, double* counts, void* stream);
/* norm[0] = max(counts[0]/world_size, 1), norm[1] = max(counts[1]/world_size, 1e-6): the reduce_mean'd normalisers
 * (fcos_head.py:266,273-274); `counts` holds the SUM over ranks (all-reduce it between the two calls). */
int dslb_fcos_norm(const double* counts, float world_size, float* norm, void* stream);
/* Losses + gradients. loss_sums[0..3] += loss_cls, loss_bbox, loss_centerness, loss_sisoft (fp64, pre-zeroed);
 * dscale[l] += d loss / d scales[l].scale (or NULL). si_weight: 0 = off, else the (warm-up adjusted) soft weight.
 * level_scales: device array [nlevels] of the Scale values (NULL = use levels[l].scale). */
int dslb_fcos_loss(const dslb_fcos_level_t* levels, int nlevels, int B, int num_classes, const int64_t* labels,
                   const float* bbox_targets, const float* weights, const float* ctr_targets, const float* norm,
                   float alpha, float gamma, float loss_weight, int n_labeled, float si_weight,
                   const float* level_scales, double* loss_sums, float* dscale, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Flat-buffer parameter kernels. The detector's parameters live in one flat fp32 buffer (nn.Parameters are views).
 * ---------------------------------------------------------------------------------------------------- */
/* EMA teacher update, SemiEpochBasedRunner.EMA (mmdet/runner/hooks/semi_epoch_based_runner.py:392-406):
 * teacher = student*c_student + teacher*c_teacher with c_student = float(1-keep_rate), c_teacher = float(keep_rate);
 * in place, every product and the sum rounded separately like the reference's fp32 expression. */
int dslb_ema_update(float* teacher, const float* student, long long n, float c_student, float c_teacher, void* stream);
/* *out += sum g^2 (fp64, pre-zeroed): global L2 norm for clip_grad_norm_ (mmcv OptimizerHook, cfg grad_clip 35). */
int dslb_sq_norm(const float* g, long long n, double* out, void* stream);
/* coef[0] = min(max_norm/(sqrt(*sqnorm)+1e-6), 1) (1 if max_norm <= 0); coef[1] = sqrt(*sqnorm). */
int dslb_clip_coef(const double* sqnorm, float max_norm, float* coef, void* stream);
/* torch.optim.SGD step (momentum, dampening 0) with the clip coefficient read from device memory (or NULL):
 * d = g*coef[0] + wd*p; buf = first_step ? d : momentum*buf + d; p -= lr*lr_scale[0]*buf (lr_scale: device scalar
 * for the warm-up / step schedule, or NULL = 1, so a captured CUDA graph can be replayed with a changing LR).
 * cfg optimizer: lr .01, momentum .9, wd 1e-4, paramwise bias_lr_mult 2 / bias_decay_mult 0 (one call per region). */
int dslb_sgd_step(float* p, const float* g, float* buf, long long n, const float* coef, const float* lr_scale, float lr,
                  float momentum, float weight_decay, int first_step, void* stream);
/* the same step with the EMA teacher of these parameters updated in the same pass (the arithmetic of dslb_ema_update on
 * the NEW p: teacher = p*c_student + teacher*c_teacher, products and sum rounded separately; EMAOWNHook per-iteration
 * mode, runner/hooks/semi_epoch_based_runner.py:398-404): one read of the student weights less per step. */
int dslb_sgd_ema_step(float* p, const float* g, float* buf, long long n, const float* coef, const float* lr_scale,
                      float lr, float momentum, float weight_decay, int first_step, float* teacher, float c_student,
                      float c_teacher, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Teacher decode + score gate (FCOSHead._get_bboxes, fcos_head.py:406-527; multiclass_nms gate,
 * mmdet/core/post_processing/bbox_nms.py:34-67). Per level.
 * ---------------------------------------------------------------------------------------------------- */
/* out[pt] = max_c sigmoid(cls[pt][c]) * sigmoid(centerness[pt])  — the key of the per-level top-nms_pre. */
int dslb_fcos_point_scores(const float* cls, const float* regctr, float* out, long long npts, int C, int ld_cls,
                           void* stream);
/* Per-level top-K of the point scores (`max_scores.topk(nms_pre)`, fcos_head.py:452-460), all levels that need it in
 * ONE launch: for level i, sel[i] [B][K[i]] int64 receives the indices (inside the level) of the K[i] largest of
 * scores[i] [B][n[i]] per image. The set equals torch.topk's whenever the K-th score is unique; ties at the cut go to
 * the lower point index; the order inside sel is unspecified. scores / sel / n / K are HOST arrays of nlevels (<= 8)
 * entries holding device pointers / sizes. */
int dslb_fcos_topk_points(const float* const* scores, int64_t* const* sel, const int32_t* n, const int32_t* K,
                          int nlevels, int B, void* stream);
/* For the K selected points of each image (sel [B][K] int64 point indices inside the level, NULL = the first K):
 * decode + clip to img_hw[n] = (H, W) + divide by scale_factor[n][4] (NULL = no rescale); every class with raw
 * sigmoid score > score_thr appends (box, score*centerness, label, point_offset+point) at an atomically claimed slot
 * of image n (counts[n], pre-zeroed; slots >= cap are dropped but still counted, and *overflow — a sticky device flag,
 * may be NULL — is set to 1: the reference has no cap, so a caller must treat a set flag as an error). Order within an
 * image is unspecified. */
int dslb_fcos_decode_gate(const float* cls, const float* regctr, const int64_t* sel, int B, int K, int C, int h, int w,
                          int stride, int ld_cls, const float* img_hw, const float* scale_factor, float score_thr,
                          int point_offset, float* out_boxes, float* out_scores, int32_t* out_labels,
                          int32_t* out_points, int32_t* counts, int cap, int32_t* overflow, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Teacher post-processing on the device (the reference does this on the host: D2H, numpy, JSON files).
 * ---------------------------------------------------------------------------------------------------- */
/* multiclass_nms -> mmcv batched_nms (mmdet/core/post_processing/bbox_nms.py:78-94): class-aware greedy NMS with the
 * class-offset trick (boxes + label * (max_coordinate + 1), fp32), IoU > iou_thr suppresses, survivors in descending
 * score order, first max_det kept. Inputs are the gated candidates of dslb_fcos_decode_gate ([B][cap] slots, counts[n]
 * valid, any order; cap a power of two <= 8192). Ties in score are broken by ascending (point * num_classes + label).
 * dets [B][max_det][5] = (x1, y1, x2, y2, score); det_labels [B][max_det]; det_count [B]. */
size_t dslb_nms_workspace_bytes(int B, int cap);
int dslb_multiclass_nms(const float* boxes, const float* scores, const int32_t* labels, const int32_t* points,
                        const int32_t* counts, int B, int cap, int num_classes, float iou_thr, int max_det,
                        void* workspace, size_t ws_bytes, float* dets, int32_t* det_labels, int32_t* det_count,
                        void* stream);
/* Detections -> pseudo ground truth for the student, i.e. the rule chain the reference spreads over
 * UnlabelPredHook (mmdet/runner/hooks/unlabel_pred_hook.py:20-38 parse_det_results: score >= infer_score_thr, int()
 * truncation, round(score, 6); :142-165 per-class nms(iou, score_threshold=0.1) over range(0, len(id2cat)-1) = all C
 * classes: the reference's category file ends with a background entry) and SemiCOCODataset._parse_ann_info (mmdet/datasets/
 * semicoco.py:220-269: boxes without overlap with the image or thinner than 1 px dropped; ignore_lo <= score <
 * thr_class[c] -> ignore region, every other box -> GT). Output order = the reference's (class ascending, score
 * descending). Writes the packed box lists + offsets that dslb_fcos_targets consumes. max_det <= 128. */
int dslb_pseudo_labels(const float* dets, const int32_t* det_labels, const int32_t* det_count, const double* thr_class,
                       const float* img_wh, int B, int max_det, int num_classes, double infer_score_thr, float nms_iou,
                       double ignore_lo, int max_boxes, float* gt_boxes, int64_t* gt_labels, int32_t* gt_off,
                       float* ig_boxes, int32_t* ig_off, void* stream);
/* Same, plus the on-device statistics of the adaptive-threshold rule (adathres(), mmdet/runner/hooks/
 * unlabel_pred_hook.py:295-343): every box the hook would have written to its per-image JSON (i.e. alive after the
 * per-class NMS, BEFORE the dataset's geometry filter) with score >= 0.3 (stat_prev NULL: no history file yet) or
 * >= stat_prev[class] (last epoch's threshold; -inf = class absent from the history) adds 1 to stat_cnt[class] and its
 * score to stat_cum[class]. The caller zeroes the two [num_classes] accumulators at the start of an epoch and, with
 * several ranks, sums them over ranks before dslb_adathres_finalize (the reference's rank 0 reads every file). */
int dslb_pseudo_labels_stats(const float* dets, const int32_t* det_labels, const int32_t* det_count,
                             const double* thr_class, const float* img_wh, int B, int max_det, int num_classes,
                             double infer_score_thr, float nms_iou, double ignore_lo, int max_boxes, float* gt_boxes,
                             int64_t* gt_labels, int32_t* gt_off, float* ig_boxes, int32_t* ig_off, int64_t* stat_cnt,
                             double* stat_cum, const double* stat_prev, void* stream);
/* dslb_pseudo_labels, plus the list the hook writes to the image's JSON file (save_results2file,
 * unlabel_pred_hook.py:142-175: alive after the per-class NMS, BEFORE the dataset's geometry filter and threshold rule),
 * in file order (class ascending, score descending): saved_boxes [B][max_det][4] (integer-valued), saved_scores
 * [B][max_det] (the fp32 value the JSON carries), saved_labels [B][max_det], saved_count [B]. For hand-over to the
 * reference's SemiCOCODataset through dsl_b200/formats.py. No statistics are accumulated by this call. */
int dslb_pseudo_labels_saved(const float* dets, const int32_t* det_labels, const int32_t* det_count,
                             const double* thr_class, const float* img_wh, int B, int max_det, int num_classes,
                             double infer_score_thr, float nms_iou, double ignore_lo, int max_boxes, float* gt_boxes,
                             int64_t* gt_labels, int32_t* gt_off, float* ig_boxes, int32_t* ig_off, float* saved_boxes,
                             float* saved_scores, int32_t* saved_labels, int32_t* saved_count, void* stream);
/* adathres() tail (unlabel_pred_hook.py:344-361) in fp64: mean = sum(cnt) / #{c: cnt_c > 0};
 * thr_out[c] = clip((cum_c / mean)^gamma1 * base, lo, hi); weight_out[c] = (mean / cum_c)^gamma2 (the reference writes
 * the weights but never reads them). Classes never counted: thr_out = absent_thr (SemiCOCODataset's default),
 * weight_out = 0, prev_out = -inf. prev_out (nullable) is next epoch's stat_prev. */
int dslb_adathres_finalize(const int64_t* stat_cnt, const double* stat_cum, int num_classes, double gamma1, double gamma2,
                           double base, double lo, double hi, double absent_thr, double* thr_out, double* weight_out,
                           double* prev_out, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Device-side view geometry of the data pipeline: the BOX part of Resize -> PatchShuffle -> RandomFlip (horizontal) as
 * the reference's train pipelines chain them (configs/fcos_semi/*.py:70-92; mmdet/datasets/pipelines/
 * transforms.py:249-257 _resize_bboxes, :2168-2248 PatchShuffle incl. the split of a box that straddles the cut,
 * :397-429 bbox_flip), per image, in fp32 exactly like the NumPy float32 expressions. Boxes are packed over images
 * with offsets off[B+1]; a box can become two, so outputs hold up to 2x the inputs (truncated at max_out).
 * dslb_pad_batch: zero-padded batch assembly (Pad + collate): out[b] (C,H,W) <- imgs[b] (C,h_b,w_b) top-left.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct dslb_view {
  float sx, sy;            /* Resize scale_factor (w_scale, h_scale)                                             */
  int32_t img_w, img_h;    /* img_shape after the resize: clip range, PatchShuffle / flip extent                 */
  int32_t clip;            /* Resize.bbox_clip_border                                                            */
  int32_t ps_mode;         /* PatchShuffle: 0 off, 1 'flip' (vertical cut at column ps_crop), 2 'flop' (row)     */
  int32_t ps_crop;         /* min(int(round(extent * PS_place)), extent); 0 or extent = no-op, as in the reference */
  int32_t flip;            /* RandomFlip, direction 'horizontal'                                                 */
} dslb_view_t;
size_t dslb_view_boxes_workspace_bytes(int max_in);
int dslb_view_boxes(const float* boxes, const int64_t* labels, const int32_t* off, const dslb_view_t* views_dev, int B,
                    int max_in, int max_out, void* workspace, size_t ws_bytes, float* out_boxes, int64_t* out_labels,
                    int32_t* out_off, void* stream);
/* Scale-invariant extra sample (mmdet/runner/hooks/semi_epoch_based_runner.py:186-204: gt_bboxes.append(gt_bboxes[-1] / 2),
 * same labels): in a packed list with offsets off[0..B], image B-1's boxes x scale are appended as image B and off[B+1]
 * is written (truncated at max_boxes). labels may be NULL (ignore lists). off must have room for B + 2 entries. */
int dslb_append_scaled_boxes(float* boxes, int64_t* labels, int32_t* off, int B, float scale, int max_boxes, void* stream);
int dslb_pad_batch(const float* const* imgs_dev, const int32_t* hw_dev, float* out, int B, int C, int H, int W,
                   void* stream);

/* Pixel side of the same pipelines (mmdet/datasets/pipelines/transforms.py:218-247 _resize_img -> mmcv.imrescale ->
 * cv2.resize INTER_LINEAR on uint8; :2180-2199 PatchShuffle; :374-383 RandomFlip -> mmcv.imflip; :668-683 Normalize ->
 * mmcv.imnormalize; :729-741 Pad -> mmcv.impad_to_multiple; formatting.py DefaultFormatBundle HWC -> CHW; collate):
 * B uint8 HWC 3-channel device images (srcs_dev[b], src_h x src_w, channel order as cv2.imread gives it) -> out
 * [B][3][H][W] fp32, zero outside img_h x img_w (the kernel writes every element of out; rows / columns beyond H, W
 * are cropped). Bit-exact with the CPU pipeline: OpenCV's 11-bit fixed-point bilinear taps, float32 mean subtraction,
 * multiplication by the double reciprocal of the float32 std rounded once. mean / std: 3 host floats, indexed by
 * OUTPUT channel (after the BGR -> RGB swap when to_rgb). img_h / img_w come from mmcv.rescale_size on the host. */
typedef struct dslb_image_view {
  int32_t src_h, src_w;    /* source image                                                                        */
  int32_t img_h, img_w;    /* size after Resize (img_shape)                                                       */
  int32_t ps_mode;         /* PatchShuffle: 0 off, 1 'flip' (columns), 2 'flop' (rows), as dslb_view_t            */
  int32_t ps_crop;         /* min(int(round(extent * PS_place)), extent); 0 or extent = no-op                     */
  int32_t flip;            /* RandomFlip, direction 'horizontal'                                                  */
  int32_t reserved;        /* 0                                                                                   */
} dslb_image_view_t;
int dslb_view_images(const uint8_t* const* srcs_dev, const dslb_image_view_t* views_dev, int B, const float* mean,
                     const float* std, int to_rgb, float* out, int H, int W, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Standalone LOSSES-registry kernels (the training step uses the fused dslb_fcos_loss; these answer configs that build
 * the loss modules by themselves). One pass each: loss_elem (nullable) = weighted element-wise loss, *loss_sum
 * (nullable, fp64, ACCUMULATED: zero it first) += its sum, d* (nullable) = gradient of the weighted element-wise loss
 * w.r.t. the prediction. Reduction / avg_factor / loss_weight (mmdet/models/losses/utils.py:27-54) are scalars for
 * the caller.
 *   dslb_sigmoid_focal_loss  losses/focal_loss.py:11-56,59-102; labels int64 in [0, C], C = background; weight [N]
 *   dslb_giou_loss           losses/iou_loss.py:85-102 + iou2d_calculator.py:214-260 (aligned GIoU, eps); weight [n]
 *   dslb_bce_with_logits     losses/cross_entropy_loss.py:73-112 (use_sigmoid=True, no class_weight); weight [n]
 * ---------------------------------------------------------------------------------------------------- */
int dslb_sigmoid_focal_loss(const float* logits, const int64_t* labels, const float* weight, long long N, int C,
                            float alpha, float gamma, float* loss_elem, double* loss_sum, float* dlogits, void* stream);
int dslb_giou_loss(const float* pred, const float* target, const float* weight, long long n, float eps, float* loss_elem,
                   double* loss_sum, float* dpred, void* stream);
int dslb_bce_with_logits(const float* x, const float* target, const float* weight, long long n, float* loss_elem,
                         double* loss_sum, float* dx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DSLB_H_ */
