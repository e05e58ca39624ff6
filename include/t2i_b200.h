/* t2i_b200 -- C ABI of the B200-native wgancls hot path.
 *
 * The reference (crisbodnar/text-to-image) has no FFI: its hot path is Python calling
 * TensorFlow-1.4 library ops.  Each entry point below replaces the TF op(s) behind one reference
 * call site (cited per function, paths relative to the reference root).  The binding a
 * maintainer adds is a ctypes stub; text-to-image_b200/_lib.py is that stub.
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer owned by the caller (PyTorch tensors in our host code);
 *    the library never allocates or frees user-visible memory.
 *  - All work is enqueued on the caller's `stream` (a cudaStream_t passed as void*); no internal
 *    synchronisation, CUDA-graph capturable.
 *  - Return 0 on success, a negative t2i_status otherwise; t2i_last_error() gives the text.
 *    Nothing throws or aborts across the ABI.  There is no CPU fallback.
 *  - Activations are NHWC "bf16 planes": np = 1 -> one bf16 tensor (throughput mode);
 *    np = 2 -> value = hi + lo, two bf16 tensors plane_stride elements apart (parity mode: every
 *    tensor-core product is expanded into hi*hi + lo*hi + hi*lo, ~2^-17 relative error).
 */
#ifndef T2I_B200_H
#define T2I_B200_H
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    T2I_OK = 0,
    T2I_ERR_BAD_ARG = -1,     /* shape / alignment / unsupported configuration */
    T2I_ERR_CUDA = -2,        /* a CUDA runtime or driver call failed          */
    T2I_ERR_UNSUPPORTED = -3  /* device is not sm_100                          */
} t2i_status;

enum { T2I_CONV_S1 = 0, T2I_CONV_K4S2 = 1, T2I_DECONV_K4S2 = 2 };
enum { T2I_ACT_NONE = 0, T2I_ACT_LRELU = 1, T2I_ACT_RELU = 2 };
enum { T2I_MASK_NONE = 0, T2I_MASK_LRELU = 1, T2I_MASK_RELU = 2 };
enum { T2I_W_NK = 0, T2I_W_KN = 1 };

/* View of an NHWC activation tensor stored as bf16 planes. */
typedef struct {
    void* ptr;              /* plane 0, element (n=0,h=0,w=0,channel 0 of the buffer)            */
    long long plane_stride; /* elements between plane 0 and plane 1 (ignored when np == 1)      */
    int n, h, w, c;         /* logical extent of the view; c = channels in the window           */
    int pitch;              /* channels per pixel in memory (>= coff + c, multiple of 8)        */
    int coff;               /* first channel of the window inside the pixel (multiple of 8)     */
} t2i_act;

/* Implicit-GEMM convolution on tcgen05 tensor cores (TMA-fed, accumulators in TMEM):
 *   y = act( conv(x, w) + bias + add ) * mask'(mask)
 * mode T2I_CONV_S1   : k x k (k = 1, 3 or 4) stride 1, SAME (k = 4: TF's asymmetric padding, 1 before / 2 after).
 *                      flip = 1 correlates with the taps mirrored (input-gradient).  A dense layer is k = 1, h = w = 1.
 * mode T2I_CONV_K4S2 : 4x4 stride 2 SAME (y is h/2 x w/2); also the input-gradient of a deconv.
 * mode T2I_DECONV_K4S2: 4x4 stride 2 SAME transposed conv as four 2x2 sub-pixel phases
 *                      (y is 2h x 2w); also the input-gradient of a 4x4/s2 conv.
 * w: packed bf16 planes [np][taps][w_rows][w_cols], cols contiguous, tap = kh*k+kw.
 *    w_layout T2I_W_NK: rows = output channels, cols = contraction (a layer's packed forward weights
 *    used forward);  T2I_W_KN: rows = contraction, cols = output channels (the SAME packed forward
 *    weights used for the layer's input-gradient: MN-major B operand, no transposed copy).
 * Replaces Conv2D/Conv2DBackpropInput/MatMul + BiasAdd + LeakyRelu behind utils/ops.py:61,69,87
 * (called from models/wgancls/model.py:135-160,174-219) and their tf.gradients counterparts.
 */
typedef struct {
    int mode, k, flip, np;
    t2i_act x;
    const void* w;
    long long w_plane_stride;
    int w_rows, w_cols, w_layout;
    t2i_act y;
    const float* bias; /* [y.c] or NULL */
    t2i_act add;       /* ptr NULL = none; same pixel grid as y */
    t2i_act mask;      /* ptr NULL = none; same pixel grid as y */
    int act, mask_kind;
    /* Optional per-channel statistics of the output values v (fp32, after act / mask, before the bf16
     * rounding), accumulated (+=, fp32 atomics; the caller zeroes) over the samples [0, stat_n) and the
     * output channels [0, stat_c) (0 = all):  stat_sum[c] += sum v;  stat_sq[c] += sum v^2;
     * stat_dot[c] += sum v * stat_x (a tensor on the output grid).  They fuse what would be separate passes
     * over y: BatchNorm batch statistics (utils/ops.py:20-29) of a conv output, bias gradients (sum of an
     * input-gradient over pixels) and the two reductions of the BatchNorm backward.  stat_sq and stat_dot
     * are mutually exclusive; any pointer may be NULL. */
    float* stat_sum;
    float* stat_sq;
    float* stat_dot;
    t2i_act stat_x;
    int stat_n, stat_c;
    /* first output channel inside the weight matrix: y's c channels are output channels [w_n0, w_n0 + y.c) of w
     * (an output-channel window, e.g. the two halves of a concat gradient with different epilogues); bias,
     * add / mask / stat_x and the stat_* vectors are indexed by the WINDOW's channels. */
    int w_n0;
    /* The 3-channel ends (d_net's first conv, model.py:135; the input gradient of g_net's last transposed conv, :218):
     * x_img != NULL makes the A operand the 4x4 / stride-2 SAME patch matrix of the 3-channel image [x.n][x.h][x.w][3]
     * (row = output pixel, column (kh*4 + kw)*3 + c, 48 columns), assembled in shared memory -- no patch matrix in HBM.
     * x_img points at the image's padded bf16 rows (t2i_img_to_rows: planes [np][x.n][x.h][pitch], x.plane_stride
     * elements apart).  mode = T2I_CONV_K4S2, k = 4, x.ptr ignored, x.c = 48; w is [np][1][w_rows][w_cols] with 64
     * (48 used) along the contraction; y is [x.n][x.h/2][x.w/2][c].  Everything else (bias, act, mask, stat_*) as above. */
    const void* x_img;
} t2i_conv_gemm_desc;
int t2i_conv_gemm(const t2i_conv_gemm_desc* d, void* stream);
/* Development aid (a library built with -DT2I_TIMELINE_BUILD, T2I_TIMELINE=1 in the environment; the default build
 * records nothing): %globaltimer stamps CTA 0 of the last 32 t2i_conv_gemm launches left at the hand-over points of its
 * first 64 tiles -- 32 regions (one per launch, round robin) of [64 tiles][8 events], row 63 holding kernel-level
 * stamps (csrc/conv_gemm.cu, tools/conv_timeline.py); count <= 32 * 512 values from region 0 on; synchronises. */
int t2i_debug_timeline(unsigned long long* host_dst, int count);

/* Weight gradient of the same three conv forms on tcgen05 (both operands MN-major, contraction
 * over pixels, split-K with fp32 atomic accumulation into dw, which the caller zeroes):
 *   dw[tap][co][ci] += sum_pixels dy[pixel][co] * x[pixel shifted by tap][ci]
 * x: the layer's forward input (or its tangent for the gradient-penalty second-order term),
 * dy: gradient at the layer's output.  Replaces Conv2DBackpropFilter / MatMul(grad) reached
 * through tf.train.AdamOptimizer.minimize at models/wgancls/model.py:94-106.
 */
typedef struct {
    int mode, k, np;
    t2i_act x, dy;
    float* dw; /* fp32 [taps][cout][cin] */
    int cout, cin;
    int split_k; /* 0 = choose */
} t2i_wgrad_desc;
int t2i_wgrad_gemm(const t2i_wgrad_desc* d, void* stream);

/* ---- HBM-bound kernels (elementwise.cu); see each definition for the reference lines ---- */

/* fp32 -> bf16 planes, optional per-row scale (rows x cols, row-major). */
int t2i_to_planes(const float* src, void* dst, long long plane_stride, int np, long long rows, int cols,
                  const float* row_scale, void* stream);
/* bf16 planes -> fp32 */
int t2i_from_planes(const void* src, long long plane_stride, int np, float* dst, long long n, void* stream);

/* ---- the 3-channel ends as direct kernels (img_gemm.cu; see also t2i_conv_gemm_desc.x_img) ----
 *
 * t2i_deconv_img: transpose of a 4x4 / stride-2 SAME convolution whose image side has 3 channels, on tcgen05 with the
 * overlap-add (col2im) done on chip:
 *     out[n, 2p-1+kh, 2q-1+kw, c] = bias3[c] + sum over the <= 4 patches (p, q) and ci of a[n, p, q, ci] * W[(kh*4+kw)*3 + c][ci]
 * a: bf16 planes [n][h][w][c <= 256], w a power of two <= 128, h*w a multiple of 128.  W: bf16 planes, 64 (48 used) along
 * the (tap, channel) axis: w_layout T2I_W_NK = [64][K] (g_net's last transposed conv, model.py:218, forward) or
 * T2I_W_KN = [K][64] (d_net's first conv, model.py:135, used for its input gradient).  out: fp32 NHWC [n][2h][2w][3].
 * w9 / b9 / img (all or none): additionally img = tanh(conv3x3(out, w9 HWIO [3][3][3][3]) + b9), model.py:219-221. */
int t2i_deconv_img(const t2i_act* a, const void* w, long long w_plane_stride, int w_rows, int w_cols, int w_layout, int np,
                   const float* bias3, float* out, const float* w9, const float* b9, float* img, void* stream);
/* fp32 NHWC 3-channel image [n][h][w][3] (optionally scaled per sample) -> padded bf16 rows, planes [np][n][h][pitch],
 * pitch = 4 * ceil((3w + 6) / 4) entries, entry j = x[j - 3] (zeros outside: the left / right SAME padding).  6 bytes per
 * pixel; this is the form in which the image-patch producers stream an image (csrc/img_patch.cuh). */
int t2i_img_to_rows(const float* img, int n, int h, int w, const float* sample_scale, void* rows, long long plane_stride,
                    int np, void* stream);
/* Weight gradient of a 4x4 / stride-2 conv between a 3-channel image [n][h][w][3], given as padded bf16 rows (patches
 * assembled on chip), and a bf16-plane tensor on the [n][h/2][w/2] grid, on tcgen05, dw += (the caller zeroes):
 *   img_side 1 (the image is the conv INPUT, model.py:135):        dw[co][64] (48 used) += sum other[pixel][co] * patch[pixel][:]
 *   img_side 2 (the image is the gradient at a deconv OUTPUT, :218): dw[64][ci] (48 used) += sum patch[pixel][:] * other[pixel][ci] */
int t2i_wgrad_img(const void* img_rows, long long plane_stride, int n, int h, int w, const t2i_act* other, int img_side,
                  int np, float* dw, int cout, int cin, void* stream);
/* y[r][o] = act(sum_k x[r][k] * w[o][k] + bias[o]) in fp32 (the conditioning head, model.py:113-114). */
int t2i_dense_f32(const float* x, int rows, int cin, const float* w, const float* bias, int cout, int act, float* y,
                  void* stream);
/* dst[r][:] = row_scale[r] * src[r][:], fp32 (gradient-penalty tangent seed, model.py:62-65). */
int t2i_scale_rows(const float* src, const float* row_scale, float* dst, long long rows, int cols, void* stream);

/* 3-channel image <-> 4x4/s2 patch matrix [n*(h/2)*(w/2), 64] (col = (kh*4+kw)*3 + c, 48 used).
 * im2col feeds d_net's first conv (model.py:135) and the input-gradient of g_net's last deconv
 * (model.py:218); col2im is their transpose. */
int t2i_im2col_k4s2_c3(const float* img, int n, int h, int w, const float* sample_scale, void* col,
                       long long plane_stride, int np, void* stream);
int t2i_col2im_k4s2_c3(const void* col, long long plane_stride, int np, int n, int h, int w,
                       const float* bias3, float* img, void* stream);

/* 3-channel image -> 3x3/s1 SAME patch matrix [n*h*w, 32] (col = (kh*3+kw)*3 + c, 27 used): StackGAN stage-II's
 * first generator conv (models/stackgan/stageII/model.py:139).  t2i_tanh_c3_fwd / _bwd: the 3 leading channels of an
 * 8-channel planes tensor (the padded output of the last 3x3 conv, :172) -> tanh -> fp32 NHWC image, and
 * dlogit = dy * (1 - y^2) back into 8-channel planes (channels 3..7 zero). */
int t2i_im2col_k3s1_c3(const float* img, int n, int h, int w, void* col, long long plane_stride, int np, void* stream);
int t2i_tanh_c3_fwd(const void* logits8, long long plane_stride, int np, float* img, long long pixels, void* stream);
int t2i_tanh_c3_bwd(const float* img, const float* dimg, void* dlogits8, long long plane_stride, int np, long long pixels,
                    void* stream);

/* g_net's last conv 3->3 k3 s1 + tanh (model.py:219-221), direct; and its backward
 * (dw, db, dx_sum are accumulated: dx_sum[c] += sum of dx over pixels = bias gradient of the
 * preceding transposed conv; may be NULL).  The backward entry launches one kernel per requested result:
 * dx != NULL -> the input gradient (+ dx_sum); dw != NULL (with db, x) -> the weight / bias gradients; a
 * caller may ask for the two separately (the input gradient is on the critical path, the other is not). */
int t2i_conv3x3_c3_tanh_fwd(const float* x, const float* w, const float* b, float* y, int n, int h, int wd,
                            void* stream);
int t2i_conv3x3_c3_tanh_bwd(const float* x, const float* w, const float* y, const float* dy, float* dx,
                            float* dw, float* db, float* dx_sum, int n, int h, int wd, void* stream);

/* per-channel sum over rows [row_begin,row_end) of a [rows, pitch] planes tensor -> out[c] (+=) */
int t2i_colsum(const void* src, long long plane_stride, int np, long long rows, int c, int pitch, int coff,
               float* out, void* stream);

/* Training-mode batch norm (utils/ops.py:7-29 via model.py:176-216): statistics, apply, backward.
 * sums: caller-owned fp32 scratch [2*c], zero on entry and left zero on return (no memset per call). */
int t2i_bn_stats(const void* x, long long plane_stride, int np, long long rows, int c, float* sums, float* mean,
                 float* rstd, float* var, float eps, void* stream);
int t2i_bn_apply(const void* x, long long x_ps, const float* mean, const float* rstd, const float* gamma,
                 const float* beta, const void* residual, long long r_ps, void* y, long long y_ps, int np,
                 long long rows, int c, int relu, void* stream);
int t2i_bn_bwd_reduce(const void* dy, long long dy_ps, const void* x, long long x_ps, const float* mean,
                      const float* rstd, int np, long long rows, int c, float* dgamma, float* dbeta,
                      void* stream);
int t2i_bn_bwd_apply(const void* dy, long long dy_ps, const void* x, long long x_ps, const float* mean,
                     const float* rstd, const float* gamma, const float* dgamma, const float* dbeta, void* dx,
                     long long dx_ps, int np, long long rows, int c, void* stream);
/* The fused forms the training step uses.  t2i_bn_apply_train: sums = [sum x | sum x^2] (2*c floats, produced by
 * t2i_conv_gemm's stat_sum / stat_sq); derives mean / biased variance / rstd (written out for the backward
 * pass), applies gamma / beta (+ residual) (+ ReLU), and when moving_mean / moving_var are given also steps the
 * moving statistics (UPDATE_OPS of the G run, model.py:98,102).  t2i_bn_bwd_fused: dbeta = sum dy and
 * dot = sum dy * x were accumulated by the kernel that produced dy (stat_sum / stat_dot; dot_normalised = 1:
 * dot already is sum dy * xhat); dgamma[c] += out_scale * rstd * (dot - mean * dbeta);
 * dbeta_out[c] += out_scale * dbeta (optional); dx as t2i_bn_bwd_apply; dx_sum[c] += sum dx (optional: the bias
 * gradient of the conv in front of the BatchNorm).
 * stat_rows (0 = rows): number of values per channel behind the sums when they were all-reduced over the
 * data-parallel ranks (synchronised BatchNorm = the reference's whole-batch statistics, utils/ops.py:20-29);
 * out_scale = 1 / world then keeps the later gradient all-reduce(sum) exact.
 * relu: 0 none, 1 ReLU, 2 LeakyReLU(0.2).  y_pitch / dy_pitch (0 = c): channels per pixel in memory of y / dy when
 * they are the leading channels of a wider buffer (the discriminator's concat buffer).  affine_scale multiplies gamma
 * and beta (StackGAN stage-II adds a BatchNorm output to itself, models/stackgan/stageII/model.py:117; pass
 * out_scale = affine_scale backward so that dgamma / dbeta are those of the unscaled parameters). */
int t2i_bn_apply_train(const void* x, long long x_ps, const float* sums, float eps, const float* gamma,
                       const float* beta, const void* residual, long long r_ps, void* y, long long y_ps, int np,
                       long long rows, int c, int relu, float* mean, float* rstd, float* var, float* moving_mean,
                       float* moving_var, float decay, long long stat_rows, int y_pitch, float affine_scale, void* stream);
int t2i_bn_bwd_fused(const void* dy, long long dy_ps, const void* x, long long x_ps, const float* mean,
                     const float* rstd, const float* gamma, const float* dot, const float* dbeta, float* dgamma,
                     float* dbeta_out, float out_scale, int dot_normalised, void* dx, long long dx_ps, float* dx_sum,
                     int np, long long rows, int c, long long stat_rows, int dy_pitch, float affine_scale, void* stream);
int t2i_bn_update_moving(float* moving_mean, float* moving_var, const float* mean, const float* var,
                         long long rows, int c, float decay, void* stream);
/* dst = dy * act'(y)  (relu / lrelu masks on post-activation values) */
int t2i_act_bwd(const void* dy, long long dy_ps, const void* y, long long y_ps, void* dst, long long dst_ps,
                int np, long long n, int mask_kind, void* stream);

/* d_net embedding path (model.py:150-155): tile [s,c] -> channels [coff,coff+c) of [s,4,4,pitch];
 * reduce sums the 16 positions back. */
int t2i_embed_tile(const void* e, long long e_ps, void* cat, long long cat_ps, int np, int s, int c, int pitch,
                   int coff, int hw, void* stream);
int t2i_embed_reduce(const void* dcat, long long dcat_ps, void* de, long long de_ps, int np, int s, int c,
                     int pitch, int coff, int hw, void* stream);

/* d_net output conv 4x4/s4 VALID 1024->1 (model.py:160) as a per-sample dot product. */
int t2i_dout_fwd(const void* a, long long a_ps, int np, const float* w, const float* b, float* logit, int s,
                 int k, void* stream);
/* da = seed[s] * w * lrelu'(a);  dw += sum_s seed[s] * a_or_tangent[s];  db += sum seed[0..s_bias) */
int t2i_dout_bwd_data(const void* a, long long a_ps, int np, const float* w, const float* seed, void* da,
                      long long da_ps, int s, int k, void* stream);
int t2i_dout_bwd_weight(const void* a, long long a_ps, int np, const float* seed, float* dw, float* db,
                        int s, int s_bias, int k, void* stream);

/* x_hat = eps*G + (1-eps)*x (model.py:53) */
int t2i_gp_interp(const float* g, const float* x, const float* eps, float* xhat, int n, int per_sample,
                  void* stream);
/* slopes / one-sided penalty / second-order seed coefficient (model.py:62-70,88-91):
 *   slope[b] = ||grad[b,:]||_2 ; *pen_sum += sum_b max(0,slope-1)^2 ;
 *   coef[b] = weight * 2 * max(0,slope-1) / slope * inv_global_batch     (0 when slope <= 1) */
int t2i_gp_penalty(const float* grad, int n, int per_sample, float weight, float inv_global_batch,
                   float* slope, float* coef, float* pen_sum, void* stream);

/* conditioning augmentation (model.py:108-127): ms = [mean | log_sigma] (post-LeakyReLU), an fp32 [b,2*ce] tensor
 * produced by t2i_dense_f32 (log_sigma feeds exp(): it is kept out of bf16 storage).
 * fwd: zc[b, z_dim + j] = mean + exp(log_sigma) * tn_eps ; zc[b, :z_dim] = z ; *kl_sum += KL terms.
 * bwd: dms = lrelu'(ms) * ( [dc | dc*eps*exp(ls)] + kl_scale * [mean | exp(2 ls) - 1] ),
 *      kl_scale = kl_coeff / (global_batch * ce). */
int t2i_ca_fwd(const float* ms, const float* z, const float* tn_eps, void* zc, long long zc_ps,
               int np, int b, int z_dim, int ce, float* kl_sum, void* stream);
int t2i_ca_bwd(const float* ms, const void* dzc, long long dzc_ps, const float* tn_eps,
               void* dms, long long dms_ps, int np, int b, int z_dim, int ce, float kl_scale, void* stream);

/* Scalars of the two runs (model.py:79-92,100).  The per-rank sums live at the tail of the flat
 * gradient buffer so that the single NCCL allreduce of the step also reduces them.
 *   d sums: [0] sum Dg, [1] sum Dx, [2] sum Dxmi, [3] sum Dxmi^2, [4] sum pen(x_hat), [5] sum pen(cond)
 *   g sums: [0] sum Dg, [1] sum KL terms
 * t2i_d_seeds writes the per-sample backward seeds of the 4 segments [fake|real|mismatch|x_hat]:
 *   1/B, -(1+kt)/B, kt/B, 1  (B = global batch).  t2i_d_scalars also applies kt -= kt_lr * dkt. */
enum {
    T2I_S_D_LOSS = 0, T2I_S_D_LOSS_REAL, T2I_S_D_LOSS_FAKE, T2I_S_D_LOSS_MISMATCH, T2I_S_WDIST, T2I_S_WDIST2,
    T2I_S_REG_LOSS, T2I_S_BALANCE_LOSS, T2I_S_REAL_GP, T2I_S_REAL_GP2, T2I_S_KT, T2I_S_KT_GRAD,
    T2I_S_G_LOSS, T2I_S_G_KL_LOSS, T2I_S_COUNT = 16
};
int t2i_d_seeds(const float* kt, float* seed, int b, float inv_global_batch, void* stream);
int t2i_d_sums(const float* logit, int b, float* sums, void* stream);
int t2i_d_scalars(const float* sums, float* kt, float* scalars, int global_batch, float gp_weight, float kt_lr,
                  void* stream);
int t2i_g_sums(const float* logit_fake, int b, float* sums, void* stream);
/* StackGAN stage-I (models/stackgan/stageI/trainer.py:21-44): seed[i] = weight * (sigmoid(logit) - label) / B, the
 * backward seed of weight * mean sigmoid_cross_entropy_with_logits(logit, label); *loss_sum += sum of the
 * cross-entropy terms.  t2i_s1_scalars: sums = [CE(syn,0), CE(real,0.9), CE(mismatch,0), CE(syn,1), KL terms] ->
 * which = 0: scalars[0..3] = D_loss, D_synthetic_loss, D_real_match_loss, D_real_mismatch_loss;
 * which = 1: scalars[4..6] = G_loss, G_gan_loss, G_kl_loss. */
int t2i_ce_seeds(const float* logit, int n, float label, float weight, float inv_global_batch, float* seed,
                 float* loss_sum, void* stream);
int t2i_s1_scalars(const float* sums, float* scalars, int global_batch, int ce, float alpha, float kl_coeff, int which,
                   void* stream);
int t2i_g_scalars(const float* sums, float* scalars, int global_batch, int ce, float kl_coeff, void* stream);

/* weights: fp32 master [taps][cout][cin] -> bf16 planes, same layout (fwd) and/or transposed
 * (bwd [taps][cin][cout]); either destination may be NULL.  The engine only needs fwd (to_planes
 * on the flat buffer / fused into t2i_adam_tf); kept for layouts that want a K-major transpose. */
int t2i_pack_weight(const float* w, int taps, int cout, int cin, void* fwd, long long fwd_ps, void* bwd,
                    long long bwd_ps, int np, void* stream);

/* TF-form Adam on a flat fp32 buffer (tf.train.AdamOptimizer at model.py:94-96,103-105):
 * m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; theta -= lr_t m / (sqrt(v) + eps),
 * lr_t = lr sqrt(1-b2^t)/(1-b1^t) is read from DEVICE memory (*lr_t_dev) so that a captured CUDA graph
 * replays with the current value.  beta1 == 0 skips the m slot entirely (m == g).  grad_scale
 * multiplies g first.  packed (optional): the updated theta is also written as bf16 planes
 * [np][n] (the tensor-core copy of the weights), fusing the re-pack into the optimizer pass. */
int t2i_adam_tf(float* theta, const float* grad, float* m, float* v, long long n, const float* lr_t_dev,
                float beta1, float beta2, float eps, float grad_scale, void* packed, long long packed_ps, int np,
                void* stream);

/* ---- conditional progressive-growing WGAN (models/pggan/pggan.py; op wrappers utils/ops.py:74-81,100-111) ----
 * layer_norm = tf.contrib.layers.layer_norm(begin_norm_axis=1, begin_params_axis=-1): per-SAMPLE statistics over the
 * m = rows * c values of a sample, gamma / beta per channel (rank 2: rows = 1, c = features).
 *   t2i_ln_stats:      sums[n][2] += [sum x | sum x^2]   (the caller zeroes sums)
 *   t2i_ln_apply:      y = act((x - mean_n) * rstd_n * gamma[c] + beta[c]), rstd = rsqrt(biased var + eps); relu 0 / 1
 *   t2i_ln_bwd_reduce: dsums[n][2] += [sum g | sum g * xhat], g = dy * gamma; dgamma += sum dy * xhat; dbeta += sum dy
 *   t2i_ln_bwd_apply:  dx = rstd_n * (g - mean_m(g) - xhat * mean_m(g * xhat)); dx_sum[c] += sum dx (the bias gradient
 *                      of the conv / dense layer in front; may be NULL)
 * dy is the gradient at the layer-norm OUTPUT with the activation's derivative already applied (the producing
 * input-gradient GEMM masks it in its epilogue). */
int t2i_ln_stats(const void* x, long long ps, int np, int n, long long m, float* sums, void* stream);
int t2i_ln_apply(const void* x, long long x_ps, const float* sums, float eps, const float* gamma, const float* beta,
                 void* y, long long y_ps, int np, int n, long long rows, int c, int relu, void* stream);
int t2i_ln_bwd_reduce(const void* dy, long long dy_ps, const void* x, long long x_ps, const float* sums, float eps,
                      const float* gamma, float* dsums, float* dgamma, float* dbeta, int np, int n, long long rows, int c,
                      void* stream);
int t2i_ln_bwd_apply(const void* dy, long long dy_ps, const void* x, long long x_ps, const float* sums, float eps,
                     const float* gamma, const float* dsums, void* dx, long long dx_ps, float* dx_sum, int np, int n,
                     long long rows, int c, void* stream);
/* NHWC planes.  upscale2x (utils/ops.py:109-111 resize_nearest_neighbor x2): y[n,i,j,:] = scale * x[n,i/2,j/2,:],
 * x is h x w.  pool2x (utils/ops.py:100-101 tf.nn.pool AVG 2 with scale = 1/4): y[n,p,q,:] = scale * sum of the 2x2
 * block, x is h x w (even).  Each is the transpose of the other (pool backward = upscale2x(scale 1/4), upscale
 * backward = pool2x(scale 1)).  upscale2x's optional mask (shaped like y, post-activation values; mask_kind as
 * t2i_conv_gemm_desc.mask_kind) multiplies by the activation derivative of the layer in front of the pool. */
int t2i_upscale2x(const void* x, long long x_ps, void* y, long long y_ps, int np, int n, int h, int w, int c, float scale,
                  const void* mask, long long m_ps, int mask_kind, void* stream);
int t2i_pool2x(const void* x, long long x_ps, void* y, long long y_ps, int np, int n, int h, int w, int c, float scale,
               void* stream);
/* out = ab[0] * x + ab[1] * z over n values (z NULL: out = ab[0] * x); ab in DEVICE memory: the fade-in blend
 * alpha * x + (1 - alpha) * x_iden of pggan.py:268,313 and its backward, capturable in CUDA graphs. */
int t2i_axpby(const void* x, long long x_ps, const void* z, long long z_ps, void* out, long long o_ps, int np, long long n,
              const float* ab, void* stream);
/* dst[row, d_coff + k] = src[row, s_coff + k] for k < c: channel window of one pitched planes buffer into another
 * (image part of the discriminator's concat buffer, pggan.py:318-322, and of its gradient). */
int t2i_copy_window(const void* src, long long s_ps, int s_pitch, int s_coff, void* dst, long long d_ps, int d_pitch,
                    int d_coff, int np, long long rows, int c, void* stream);
/* fp32 NHWC 3-channel image <-> planes with 8 channels (3..7 zero): the operand of from_rgb's 1x1 conv (pggan.py:343-345)
 * and the output of to_rgb's (pggan.py:367-371).  sample_scale (optional, [n]) multiplies sample i. */
int t2i_img_to_c8(const float* img, int n, long long pix_per_sample, const float* sample_scale, void* dst, long long ps,
                  int np, void* stream);
int t2i_c8_to_img(const void* src, long long ps, int np, float* img, long long pixels, void* stream);

const char* t2i_last_error(void);
int t2i_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long t2i_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
