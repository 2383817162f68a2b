/* tcvom_b200 -- C ABI of the B200 (sm_100a) kernels behind TCVOM's GCA+TAM hot path.
 *
 * Every entry point takes plain device pointers, explicit dimensions and a CUDA stream
 * (pass torch.cuda.current_stream().cuda_stream); returns 0 on success or a negative
 * error code (never throws, never exits); tcv_last_error() returns the message for the
 * calling thread.  The library owns no device memory: all workspaces are passed in.
 *
 * Data layout ("split-bf16 NHWC"): an activation tensor [N,H,W,C] is stored as two bf16
 * planes, hi at `base` and lo at `base + N*H*W*C` (elements), value = float(hi)+float(lo).
 * Boundary tensors that the reference exposes (images, trimaps, alpha, TAM logits) are
 * fp32 in the reference's own layouts.
 *
 * Reference interfaces replaced (reference checkout, commit f5fa07a):
 *   tcv_preprocess_eval   models/model.py:360-387   EvalModel.preprocess
 *   tcv_postprocess_eval  models/model.py:413-424   EvalModel.forward tail (where/compositing)
 *   tcv_sn_fold_pack      models/GCA/ops.py:38-45   SpectralNorm._noupdate_u_v (W_bar / u^T W v)
 *   tcv_conv2d            nn.Conv2d / nn.ConvTranspose2d + BatchNorm2d + ReLU/LeakyReLU/+residual
 *                         call sites: GCA/encoders/resnet_enc.py:33-49,129-145,
 *                         res_gca_enc.py:20-33,47-55,57-90, GCA/decoders/resnet_dec.py:43-59,
 *                         VMN/VMN_GCA.py:26-49, VMN/VMN_model.py:13-15
 *   tcv_preprocess_train  models/model.py:54-92     FullModel.preprocess + make_trimap
 *   tcv_losses_vmd        models/model.py:94-127,285-345  L_im / L_af / L_tc forward (+ utils/loss_func.py:9-22)
 *   tcv_avgpool2          nn.AvgPool2d(2,2)          resnet_enc.py:112
 *   tcv_gca_*             models/GCA/ops.py:106-229  GuidedCxtAtten.forward
 *   tcv_tam_attend        models/VMN/VMN_model.py:18-68  FeatureAggregationModule.forward
 *   tcv_nchw_to_split / tcv_split_to_nchw   layout adapters for the operator seams
 */
#ifndef TCVOM_B200_H
#define TCVOM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* tcv_stream_t; /* cudaStream_t */

#define TCV_OK 0
#define TCV_ERR_INVALID (-1)
#define TCV_ERR_CUDA (-2)
#define TCV_ERR_UNSUPPORTED (-3)

#define TCV_ACT_NONE 0
#define TCV_ACT_RELU 1
#define TCV_ACT_LEAKY02 2
#define TCV_ACT_TANH01 3 /* (tanh(x)+1)/2 -- VMN_GCA.py:47 */
#define TCV_ACT_LEAKY001 4 /* nn.LeakyReLU() default slope 0.01 -- FBA/models.py:266-300 */
#define TCV_ACT_CLAMP01 5  /* .clamp(0, 1) -- VMN_DIM.py:135 */
#define TCV_ACT_RELU6 6    /* nn.ReLU6 -- models/Index/hlconv.py:39, net.py:47 */

#define TCV_PAD_ZERO 0
#define TCV_PAD_REFLECT 1

#define TCV_MAX_TAPS 16

int tcv_version(void);
const char* tcv_last_error(void);
/* number of kernels launched by this library since load (for bench.py's gpu_launches) */
long long tcv_launch_count(void);

/* Generic gather-form convolution with fused epilogue.
 *   out[n, gy*oy_mul+oy_off, gx*ox_mul+ox_off, co] = epi( sum_t sum_ci
 *        in[n, gy*stride+dy[t], gx*stride+dx[t], ci] * w[wtap[t]][ci][co] )
 *   epi(a) : t = a*s1[co]+b1[co]; t += res1[n, oy>>res1_shift, ox>>res1_shift, co];
 *            t = act(t); t = t*s2[co]+b2[co]; t += res2[n,oy,ox,co]
 * An ordinary conv has gh=oh, gw=ow, oy_mul=1; one phase of the 4x4/stride-2 transposed conv
 * has gh=ih, gw=iw, oy_mul=2, oy_off=phase and 4 taps.  Null s1/b1/s2/b2/res* are skipped. */
typedef struct {
  const void* x;      /* split-bf16 NHWC [n, ih, iw, cin]                                  */
  long long x_plane;  /* elements between the hi and lo plane of x (0: n*ih*iw*cin)          */
  long long x_img_stride; /* elements between consecutive images of x (0: ih*iw*cin)         */
  int n, ih, iw, cin; /* cin multiple of 8                                                   */
  const float* w;     /* fp32 [wtaps][cin][cout] (wtaps >= max(wtap)+1)                      */
  const void* w_tc;   /* optional bf16 hi/lo [2][w_tc_taps][cout][cin] copy of w (tcv_pack_weight_tc);
                         when present and the shape qualifies, the tcgen05 implicit-GEMM path runs */
  int w_tc_taps;
  const void* w_tc_fold; /* optional, cin == 8 3x3 only: bf16 hi/lo [2][3][cout][32] (tcv_pack_weight_fold):
                            horizontal taps folded into K for the narrow-layer tcgen05 kernel             */
  int ntaps;
  int dy[TCV_MAX_TAPS], dx[TCV_MAX_TAPS];
  int wtap[TCV_MAX_TAPS]; /* weight slice used by tap t (identity for an ordinary conv)      */
  int stride;
  int pad_mode;
  void* y;            /* split-bf16 NHWC [n, oh, ow, cout] (may be null if y_f32 is given)   */
  float* y_f32;       /* optional fp32 NHWC copy of the output                               */
  int oh, ow, cout;
  int gh, gw;
  int oy_mul, oy_off, ox_mul, ox_off;
  const float* s1;
  const float* b1;
  const void* res1;   /* split-bf16 NHWC [n, oh>>res1_shift, ow>>res1_shift, cout]           */
  long long res1_plane; /* elements between hi and lo plane of res1 (0: contiguous default)  */
  int res1_shift;
  int act;
  const float* s2;
  const float* b2;
  const void* res2;   /* split-bf16 NHWC [n, oh, ow, cout]                                   */
  long long res2_plane; /* elements between hi and lo plane of res2 (0: contiguous default)  */
  /* optional per-channel statistics of the OUTPUT (the value written to y), accumulated by the epilogue of the CTA-pair
   * tcgen05 kernel (tcv_conv2d_path(d) == 4 with stats == NULL; any other shape rejects a non-null `stats`):
   *   stats[copy][(img % stats_groups) * cout + co][0..1] += sum, sum of squares over the pixels of image img
   * (fp64, caller-zeroed).  `stats_copies` (>= 1) accumulator copies [copies][stats_groups][cout][2] spread the atomics of
   * the ~10^3 tiles of a layer over distinct addresses (same-address fp64 atomics serialise in L2: with one copy they cost
   * more than the pass they replace); the consumer sums the copies (tcv_gn_finalize_acc).  GroupNorm (layers_WS.py:26-27,
   * stats_groups = n) reads its sums from here instead of a separate pass over y. */
  double* stats;
  int stats_groups;
  int stats_copies;
} tcv_conv_desc;

int tcv_conv2d(const tcv_conv_desc* d, tcv_stream_t stream);
/* cudaMemsetAsync(p, 0, bytes) on the stream (a recorded plan zeroes its accumulators with it) */
int tcv_zero_bytes(void* p, long long bytes, tcv_stream_t stream);
/* which kernel tcv_conv2d dispatches this descriptor to: 3 = narrow-layer tcgen05 conv (un-swizzled
 * halo tile), 2 = persistent shared-halo tcgen05 conv,
 * 1 = tcgen05 implicit GEMM (one tap per K block), 0 = CUDA-core gather conv (no launch) */
int tcv_conv2d_path(const tcv_conv_desc* d);
/* selects the highest tensor-core conv kernel generation tcv_conv2d may use (1, 2 or 3; default 3);
 * returns the previous value.  For A/B measurements and tests. */
int tcv_set_conv_tc_version(int v);
/* measurement switches for kernel bring-up (bit 0: skip MMAs, 1: skip epilogue memory ops, 2/3: load
 * activations / weights only once; bits 4..7: further epilogue switches).  Results are WRONG when any of bits 0..7
 * is set; default 0.  Bit 8 (256) is a tuning switch with unchanged results: N = 64 instead of N = 128 tiles in the shared-halo conv kernel when that
 * shortens the persistent schedule.  Returns the old value. */
int tcv_set_debug_flags(int flags);

/* sigma = u^T W v  (W viewed [rows, cols], rows = w_bar.shape[0]); then packs W/sigma into the
 * kernel layout fp32 [ntaps][cin_pad][cout].  `transposed` != 0: w_bar is [cin,cout,kh,kw]
 * (ConvTranspose2d), else [cout,cin,kh,kw].  u == NULL: plain conv weight (sigma = 1).
 * Tap order in the packed tensor is (kh, kw) raster.  sigma_out (1 float, device) optional. */
int tcv_sn_fold_pack(const float* w_bar, const float* u, const float* v, int cout, int cin, int kh,
                     int kw, int transposed, int cin_pad, float* packed, float* sigma_out,
                     tcv_stream_t stream);

/* bf16 hi/lo re-layout of a packed fp32 weight for the tcgen05 path:
 * packed fp32 [taps][cin][cout] -> w_tc bf16 [2][taps][cout][cin] (plane 0 = hi, plane 1 = lo) */
int tcv_pack_weight_tc(const float* packed, int taps, int cin, int cout, void* w_tc, tcv_stream_t stream);

/* folded 3x3 weights of an 8-channel-input conv: packed fp32 [9][8][cout] -> bf16 [2][3][cout][32] with
 * w_fold[dy][co][dxi*8 + c] = packed[dy*3 + dxi][c][co] (dxi < 3), zero for dxi == 3 */
int tcv_pack_weight_fold(const float* packed, int cout, void* w_fold, tcv_stream_t stream);

/* scale = gamma / sqrt(var+eps), shift = beta - mean*scale (eval BatchNorm2d as an affine) */
int tcv_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                int c, float* scale, float* shift, tcv_stream_t stream);

/* EvalModel.preprocess.  imgs fp32 [F,3,H,W] BGR 0..255, tris fp32 [F,1,H,W].
 * x8: split-bf16 NHWC [F,H,W,8] = (normalised RGB, one-hot{bg,unknown,fg}, 0, 0);
 * trimask: fp32 [F,H,W] (0/1).  dilate <= 0: no dilation; else max-pool radius; tmp = 2*F*H*W bytes. */
int tcv_preprocess_eval(const float* imgs, const float* tris, int frames, int h, int w, int dilate,
                        void* x8, float* trimask, uint8_t* tmp, tcv_stream_t stream);

/* same, for uint8 frames / trimaps (the reference casts with .float(), models/model.py:366,368; moving
 * uint8 over PCIe is 4x cheaper) */
int tcv_preprocess_eval_u8(const uint8_t* imgs, const uint8_t* tris, int frames, int h, int w, int dilate,
                           void* x8, float* trimask, uint8_t* tmp, tcv_stream_t stream);

/* alpha[f] = trimask ? pred : tri/255 for inner frames, 0 for the first/last frame of each
 * sample.  pred fp32 [B,H,W] (centre frames only when S==3: index b*(S-2)+(s-1)). */
int tcv_postprocess_eval(const float* pred, const float* tris, const float* trimask, int batch,
                         int frames, int h, int w, float* alphas, tcv_stream_t stream);

/* FullModel.preprocess + make_trimap (models/model.py:54-92, TRIMAP_CHANNEL == 3).
 * a fp32 [B,S,1,H,W], fg/bg fp32 [B,S,3,H,W] (BGR 0..255); radii int32 [B] (device): max-pool dilation
 * radius per sample (model.py:62 draws it per sample on the host when DILATION_KERNEL is None).
 * Outputs: x8 split-bf16 NHWC [B*S,H,W,8]; trimask/gts/tris_vis fp32 [B,S,1,H,W]; fgs/bgs/imgs fp32
 * [B,S,3,H,W] (RGB, /255).  tmp = 2*B*S*H*W bytes. */
int tcv_preprocess_train(const float* a, const float* fg, const float* bg, int batch, int frames_per_sample, int h,
                         int w, float eps, const int* radii, void* x8, float* trimask, float* gts, float* fgs,
                         float* bgs, float* imgs, float* tris_vis, uint8_t* tmp, tcv_stream_t stream);

/* Forward of the FullModel_VMD losses (models/model.py:94-127 L_im, :285-323 L_af, :326-345 L_tc) and
 * the visualisation tensors.  pred fp32 [B,S-2,1,H,W] (inner frames); attb/attf fp32 [B,S-2,w*w,H*W/64],
 * small_mask uint8 [B,S-2,H*W/64] (NULL for the plain FullModel: L_af = 0).  Outputs alphas fp32
 * [B,S,1,H,W], comps fp32 [B,S,3,H,W] (clamped, zero end frames), losses fp32 [5] =
 * (L_alpha, L_comp=0, L_grad=0, L_dt, L_att).  Workspaces: gt8 fp32 [B,S,H/8,W/8], acc double [6*S]. */
int tcv_losses_vmd(const float* pred, const float* trimask, const float* gts, const float* fgs, const float* bgs,
                   const float* attb, const float* attf, const uint8_t* small_mask, int batch, int frames_per_sample,
                   int h, int w, int window, float att_thres, float label_smooth, float att_multiplier,
                   float* alphas, float* comps, float* gt8, double* acc, float* losses, tcv_stream_t stream);

int tcv_postprocess_eval_u8(const float* pred, const uint8_t* tris, const float* trimask, int batch,
                            int frames, int h, int w, float* alphas, tcv_stream_t stream);

int tcv_avgpool2(const void* x, int n, int h, int w, int c, void* y, tcv_stream_t stream);

/* y [n, h+2, w+2, c] = x [n, h, w, c] with a 1-pixel reflect border (nn.ReflectionPad2d(1),
 * res_gca_enc.py:20-33), split-bf16 NHWC */
int tcv_pad_reflect1(const void* x, int n, int h, int w, int c, void* y, tcv_stream_t stream);

/* unknown[n, y, x] = x8[n, y*8, x*8, channel 4]  (res_gca_enc.py:71) ; fp32 [n, h/8, w/8] */
int tcv_unknown_os8(const void* x8, int n, int h, int w, float* unknown, tcv_stream_t stream);

/* ---- guided contextual attention (GCA/ops.py:106-229), per image, P = (h/2)*(w/2) patches,
 * h,w = OS8 feature size, P_pad = P rounded up to 64.
 *  prep:    g split-bf16 [n,h/2,w/2,64] (guidance_conv output at stride 2), unknown fp32 [n,h,w]
 *           -> Q [n,P,576], Kn [n,P,576] (= Q/max(|Q|,1e-4) * per-key scale): fp32 when
 *              bf16_split == 0, else bf16 planes [bf16_split][n][P][576] (2: hi/lo, 3: hi/mid/lo);
 *              mm fp32 [n,P] ; scales fp32 [n,2] = (unknown_scale, known_scale)
 *  values:  feat split-bf16 [n,h,w,128] -> Vt [n,2048,P_pad] (row = (ty*4+tx)*128+c);
 *           mode 0 fp32, 1 bf16, 2 split-bf16 planes [2][n][2048][P_pad], 3 fp16
 *  softmax: S fp32 [n,P,P_pad]: P = softmax_p(S - 1e4*[q==p]*mm[p]), pad cols = 0; mode 0 writes
 *           fp32 in place, modes 1/2/3 write bf16 / split-bf16 planes / fp16 [n,P,P_pad] into P_out
 *  fold:    O fp32 [n,P,2048] -> Y split-bf16 [n,h,w,128] = fold(O; k4,s2,p1)/4            */
int tcv_gca_prep(const void* g, const float* unknown, int n, int h, int w, void* Q, void* Kn,
                 float* mm, float* scales, int bf16_split, tcv_stream_t stream);
int tcv_gca_values(const void* feat, int n, int h, int w, void* Vt, int mode, tcv_stream_t stream);
int tcv_gca_softmax(float* S, const float* mm, int n, int P, int P_pad, void* P_out, int mode,
                    tcv_stream_t stream);
int tcv_gca_fold(const float* O, int n, int h, int w, void* Y, tcv_stream_t stream);

/* ---- shift-sum form of the aggregation  fold(A.V; k4,s2,p1)/4  (GCA/ops.py:112-118,204; tcvom_b200/csrc/gca.cu):
 * with hh = h/2, ww = w/2, the key grid padded to (hh+1) x (ww+1) (Pk positions, ld = Pk rounded up to 64),
 *     Y[2my+ry-1][2mx+rx-1][c] = 1/4 * sum_p' A2[m][p'] * F_r[p'][c],    A2[m][p'] = sum_{a in {0,1}^2} A[m-a][p'-a]
 * -- one [Pk x ld].[ld x 512] GEMM instead of [P x P].[P x 2048] + overlap-add (3.8x fewer FLOPs, exact).
 *  prep_grid:     like tcv_gca_prep(bf16_split = 2), but Kn rows live on the padded key grid: Kn bf16 planes [2][n][Pk][576],
 *                 row py*(ww+1)+px, zero rows at px == ww / py == hh.  S = Q.Kn^T is then fp32 [n][P][ld].
 *  values_parity: feat split-bf16 [n,h,w,128] -> Ft planes [2][n][512][ld], row (ry*2+rx)*128+c, column p' on the padded grid
 *                 = feat_reflect[2p'y+ry-1][2p'x+rx-1][c]; columns >= Pk zero
 *  rowstats:      stats fp32 [n][P][2] = (max, 1/sum exp) of S[q][:] - 1e4*[key == q]*mm[q] over the real key columns;
 *                 normalise != 0: S is overwritten by A = softmax (zeros at the pad columns)
 *  shift_add:     A2 planes [2][n][Pk][ld] from the normalised A [n][P][ld] (rowstats with normalise): 4 shifted row reads
 *  softmax_shift: A2 planes [2][n][Pk][ld]: A2[m][j] = sum_a softmax(S)[m-a][j - (ay*(ww+1)+ax)] (pad columns zero),
 *                 from S and stats (rowstats without normalise): same result as shift_add, exponentials in the consumer
 *  unfold_parity: O2 fp32 [n][Pk][512] -> Y split-bf16 [n,h,w,128] (each output pixel read from exactly one entry, /4) */
int tcv_gca_prep_grid(const void* g, const float* unknown, int n, int h, int w, void* Q, void* Kn, float* mm,
                      float* scales, tcv_stream_t stream);
int tcv_gca_values_parity(const void* feat, int n, int h, int w, int ld, void* Ft, tcv_stream_t stream);
int tcv_gca_rowstats(float* S, const float* mm, int n, int h, int w, int ld, float* stats, int normalise,
                     tcv_stream_t stream);
int tcv_gca_shift_add(const float* A, int n, int h, int w, int ld, void* A2, tcv_stream_t stream);
int tcv_gca_softmax_shift(const float* S, const float* stats, const float* mm, int n, int h, int w, int ld, void* A2,
                          tcv_stream_t stream);
int tcv_gca_unfold_parity(const float* O2, int n, int h, int w, void* Y, tcv_stream_t stream);

/* C[b] = A[b] * B[b]^T, fp32 row-major, A [M,K] lda, B [N,K] ldb, C [M,N] ldc, K % 8 == 0,
 * batch strides in elements. */
int tcv_gemm_tn_f32(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb,
                    int ldc, long long strideA, long long strideB, long long strideC, int batch,
                    tcv_stream_t stream);

/* C[b] = A[b] * B[b]^T on the tensor cores (tcgen05, fp32 accumulate in TMEM).  A bf16 [batch][M][K]
 * (nsplit == 3: hi plane at A, lo plane a_plane elements later; products Ahi.Bhi+Ahi.Blo+Alo.Bhi;
 * nsplit == 6: three planes hi/mid/lo, products hh+hm+mh+hl+lh+mm, ~24 mantissa bits per operand),
 * B bf16 [batch][N][K] likewise; C fp32 or (out_bf16) bf16 [batch][M][ldc].  K % 64 == 0.
 * in_fp16 != 0 (nsplit == 1 only): A and B hold IEEE half instead of bf16. */
int tcv_gemm_tn_tc(const void* A, long long a_plane, const void* B, long long b_plane, void* C, int M, int N,
                   int K, long long ldc, long long c_batch_stride, int batch, int nsplit, int out_bf16,
                   int in_fp16, tcv_stream_t stream);

/* ---- temporal attention module core (VMN_model.py:27-68) after the q/k/v convolutions.
 * q, v, kb, kf: split-bf16 NHWC [B,H,W,C] (contiguous); mask fp32 full resolution
 * [mh,mw] per sample with per-sample stride mask_stride (elements), sampled at
 * (y*mh/H, x*mw/W) (nearest, VMN_model.py:22).
 * out split-bf16 [B,H,W,C] = v + m*(agg_b+agg_f); attb/attf fp32 [B,win*win,H*W] raw scaled
 * logits zeroed outside the mask; small_mask uint8 [B,H,W]. */
int tcv_tam_attend(const void* q, const void* v, const void* kb, const void* kf, const float* mask,
                   long long mask_stride, int mh, int mw, int batch, int h, int w, int c, int window,
                   void* out, float* attb, float* attf, uint8_t* small_mask, tcv_stream_t stream);

/* layout adapters for the operator-level seams (fp32 NCHW <-> split-bf16 NHWC, c_pad >= c);
 * plane = elements between the hi and lo plane of the split tensor (0: n*h*w*c_pad) */
int tcv_nchw_to_split(const float* x, int n, int c, int h, int w, int c_pad, void* y, long long y_plane,
                      tcv_stream_t stream);
int tcv_split_to_nchw(const void* x, int n, int c, int h, int w, int c_pad, long long x_plane, float* y,
                      tcv_stream_t stream);


/* =====================================================================================================
 * Training path (train_ddp.py:52-65): train-mode forward pieces and the backward kernels.
 *
 * Reference semantics replaced:
 *   tcv_sn_power_iter       models/GCA/ops.py:25-36   SpectralNorm._update_u_v (one power iteration per call)
 *   tcv_bn_*                nn.BatchNorm2d in .train() (batch statistics, running-stat update, autograd backward);
 *                           nn.SyncBatchNorm when the caller all-reduces the `sums` between the two halves
 *   tcv_conv2d_wgrad        autograd of nn.Conv2d / nn.ConvTranspose2d w.r.t. the weight
 *                           (the data gradient is tcv_conv2d itself on transposed weights, tcv_transpose_packed)
 *   tcv_weight_grad_unpack  autograd through W_bar / sigma (ops.py:35-36), u and v constants
 *   tcv_gca_*_bwd           autograd of GuidedCxtAtten.forward (ops.py:106-229)
 *   tcv_tam_attend_bwd      autograd of FeatureAggregationModule.forward (VMN_model.py:27-68)
 *   tcv_losses_vmd_bwd      autograd of L_im / L_tc / L_af (models/model.py:94-127, 285-345)
 * Gradients of activations use the same split-bf16 NHWC layout as the activations.
 * ===================================================================================================== */

/* images (b, j), j < group: to[b*g_to + j + off_to] (+)= from[b*g_from + j + off_from]; n_pairs = B*group images
 * of img_elems elements (multiple of 8) per plane.  Gathers the centre / previous / next frames of every
 * sample for the decoder tail (VMN_model.py:107-110) and scatters their gradients back. */
int tcv_copy_images(const void* from, long long from_plane, void* to, long long to_plane, long long img_elems,
                    int n_pairs, int group, int g_from, int off_from, int g_to, int off_to, int accumulate,
                    tcv_stream_t stream);
/* y += x (split-bf16, elems per plane, multiple of 8) */
int tcv_add_split(const void* x, long long x_plane, void* y, long long y_plane, long long elems,
                  tcv_stream_t stream);
/* out[c] += sum over pixels of x[pixel][c]  (bias gradients); c a power of two in [8, 512] */
int tcv_channel_sum(const void* x, long long x_plane, long long pixels, int c, float* out, tcv_stream_t stream);
/* y[n,h/2,w/2,c] = scale * sum of the 2x2 block of x (scale 0.25: AvgPool2d; 1: gradient of nearest x2) */
int tcv_pool2_scaled(const void* x, int n, int h, int w, int c, float scale, void* y, tcv_stream_t stream);
/* y[n,2h,2w,c] = scale * x[n,h,w,c] replicated 2x2 (scale 0.25: gradient of AvgPool2d(2)) */
int tcv_upsample2_scaled(const void* x, int n, int h, int w, int c, float scale, void* y, tcv_stream_t stream);
/* gradient of tcv_pad_reflect1: dx [n,h,w,c] from dy [n,h+2,w+2,c] */
int tcv_pad_reflect1_bwd(const void* dy, int n, int h, int w, int c, void* dx, tcv_stream_t stream);
/* gradient of (tanh(z)+1)/2 given the output: dz8[i][0] = dpred[i]*2*pred[i]*(1-pred[i]), channels 1..7 = 0
 * (split-bf16 [pixels][8], the 8-channel padding the conv kernels need) */
int tcv_tanh01_bwd(const float* pred, const float* dpred, long long pixels, void* dz8, tcv_stream_t stream);
/* alpha head of the decoder (resnet_dec.py:80 conv2 = Conv2d(32, 1, 3, padding=1, bias) ; VMN_GCA.py:46-47 (tanh+1)/2) in one
 * HBM-bound pass: x split-bf16 NHWC [n,h,w,32] (lo plane x_plane elements later; 0 = contiguous), wt fp32 [9][32] (tap-major,
 * tap = ky*3+kx), bias fp32 [1] or NULL -> pred fp32 [n,h,w].  Inference path; training keeps the padded 32-channel
 * tensor-core form (tcv_conv2d + tcv_head_tanh01) whose gradients run on the conv kernels. */
int tcv_head_conv_tanh01(const void* x, long long x_plane, int n, int h, int w, const float* wt, const float* bias,
                         float* pred, tcv_stream_t stream);

/* Alpha head on the tensor-core conv path: decoder.conv2 (32 -> 1, VMN_GCA.py:46) is run as a 32 -> 32 conv whose
 * output channels 1..31 have zero weights; pred[i] = (tanh(x[i][0]) + 1) / 2 (VMN_GCA.py:47) and the gradient
 * dz[i][0] = dpred[i]*2*pred[i]*(1-pred[i]), dz[i][1..c-1] = 0.  x / dz split-bf16 [pixels][c]. */
int tcv_head_tanh01(const void* x, long long x_plane, long long pixels, int c, float* pred, tcv_stream_t stream);
int tcv_head_tanh01_bwd(const float* pred, const float* dpred, long long pixels, int c, void* dz, tcv_stream_t stream);
/* fp32 [count] -> split-bf16 planes */
int tcv_f32_to_split(const float* x, long long count, void* y, long long y_plane, tcv_stream_t stream);
/* split-bf16 planes -> fp32 [count] (hi + lo) */
int tcv_split_to_f32(const void* x, long long x_plane, long long count, float* y, tcv_stream_t stream);

/* packed fp32 [taps][cin][cout] -> [taps][cout_pad][cin] (rows co >= cout zero): the weight of the data-gradient
 * convolution in tcv_conv2d's own layout */
int tcv_transpose_packed(const float* packed, int taps, int cin, int cout, int cout_pad, float* out,
                         tcv_stream_t stream);

/* SpectralNorm power iteration for `n` layers in one launch (one 8-CTA cluster per layer; rows <= 512, cols <= 8192).  Layer i is called
 * `calls` times per step (once per frame, VMN_model.py:93-98,107-110); call k uses (u_k, v_k, sigma_k) obtained by
 * k+1 iterations from the stored u, v.  The final u, v are written back to the module's buffers. */
typedef struct {
  const float* w_bar; /* torch layout, viewed [rows, cols], rows = shape[0] */
  int rows, cols;
  float* u;           /* [rows] in/out */
  float* v;           /* [cols] in/out */
  int calls;
  float* u_hist;      /* [calls][rows] */
  float* v_hist;      /* [calls][cols] */
  float* sigma;       /* [calls] */
  float* inv_sigma;   /* [calls] */
} tcv_sn_desc;
int tcv_sn_power_iter(const tcv_sn_desc* descs_device, int n, tcv_stream_t stream);

/* train-mode BatchNorm fused with what surrounds it.  t0 = z * inv_sigma[g], g = image % groups.
 *   mode 1:  y = act( BN(t0) + up(res1) ) + res2        (conv -> BN -> [+identity] -> act [-> + skip])
 *   mode 2:  y = BN( act(t0) )                          (conv -> ReLU -> BN, res_gca_enc.py:20-33,47-55)
 * Statistics are per (group, channel): the reference calls the encoder/decoder once per frame. */
typedef struct {
  const void* z; long long z_plane;
  int n, h, w, c;          /* c a power of two in [8, 512] */
  int groups;
  const float* inv_sigma;  /* [groups] or NULL */
  int mode, act;
  const float* gamma; const float* beta;
  float* mean; float* invstd;   /* [groups][c] */
  const void* res1; long long res1_plane; int res1_shift;
  const void* res2; long long res2_plane;
  void* y; long long y_plane;
} tcv_bn_desc;
/* sums[g][c] = (sum t, sum t^2) of the BatchNorm input (double, zeroed here) */
int tcv_bn_stats(const tcv_bn_desc* d, double* sums, tcv_stream_t stream);
/* mean / invstd from (all-reduced) sums; count = elements per (group, channel) over all ranks; running statistics
 * updated once per group in order (momentum; variance scaled by u/(u-1), u = unbiased_count -- it differs from
 * count only where a 1x1 conv + BatchNorm is evaluated before the nearest x2 upsample it commutes with) when
 * running_mean != NULL */
int tcv_bn_finalize(const double* sums, double count, double unbiased_count, int groups, int c, float eps,
                    float momentum, float* mean, float* invstd, float* running_mean, float* running_var,
                    tcv_stream_t stream);
int tcv_bn_apply(const tcv_bn_desc* d, tcv_stream_t stream);
/* e = dL/d(BN output): mode 1 dy * act'(BN(t0)+res1) (written to e; it is also the gradient of res1 before
 * down-pooling), mode 2 dy itself (e may be NULL).  sums[g][c] = (sum e, sum e*xhat), zeroed here.
 * Mode 1 WITHOUT res1 may pass e == NULL as well: nothing but tcv_bn_bwd_apply reads e then, and it can recompute it from dy
 * (e_is_dy) -- one full-tensor write less per such layer. */
int tcv_bn_bwd_reduce(const tcv_bn_desc* d, const void* dy, long long dy_plane, void* e, long long e_plane,
                      double* sums, tcv_stream_t stream);
/* dgamma[c] += sum_g sums[g][c][1], dbeta[c] += sum_g sums[g][c][0] (local sums, before any all-reduce) */
int tcv_bn_param_grads(const double* sums, int groups, int c, float* dgamma, float* dbeta, tcv_stream_t stream);
/* dz = dL/dz from e and the (all-reduced) sums; zdot[g] += sum z*dz (NULL to skip; feeds the sigma gradient).
 * e_is_dy != 0 (mode 1 without res1): `e` holds dy and e = dy * act'(BN(t0)) is recomputed here */
int tcv_bn_bwd_apply(const tcv_bn_desc* d, const void* e, long long e_plane, const double* sums, double count,
                     void* dz, long long dz_plane, double* zdot, int e_is_dy, tcv_stream_t stream);
/* zdot[g] += sum over images of group g of z*dz  (layers whose spectral-norm conv is not followed by BatchNorm) */
int tcv_group_dot(const void* z, long long z_plane, const void* dz, long long dz_plane, int n, long long img_elems,
                  int groups, double* zdot, tcv_stream_t stream);

/* dw[wtap[t]][ci][co] += sum over the descriptor's compute grid of x[n, gy*stride+dy[t], gx*stride+dx[t], ci] *
 * dz[n, gy*oy_mul+oy_off, gx*ox_mul+ox_off, co]; d is the forward descriptor (x, taps, stride, padding, grid);
 * dz split-bf16 [n, oh, ow, dz_c]; dw fp32 [wtaps][cin][dz_c]. */
int tcv_conv2d_wgrad(const tcv_conv_desc* d, const void* dz, long long dz_plane, int dz_c, float* dw,
                     tcv_stream_t stream);
/* Tensor-core weight gradient of the stride-1 gather convolutions (3x3, 1x1, the phases of the 4x4/stride-2 deconv).
 *  transpose_pad: split-bf16 NHWC [n,h,w,c], sub-sampled (gy*mul+off_y, gx*mul+off_x) -> channel-major planes with a
 *                 zero ring: xt[ch][(img*(gh+2)+gy+1)*row_stride + gx+1+shift], gh = h/mul, row_stride >= gw+2;
 *                 rows of ktot elements (ktot % 8 == 0, ktot >= n*(gh+2)*row_stride), lo plane xt_plane elements
 *                 after the hi plane.  shift in {-1,0,1} bakes a horizontal tap offset into the copy.
 *  wgrad_tc:      dw[wtap[t]][ci][co] += sum_p xt[ci][p + dy[t]*row_stride + dx[t]] * zt[co][p]:
 *                 split-K tcgen05 GEMMs (bf16x3), one per tap or -- for <= 64 output channels -- one per call with the
 *                 (up to 3) taps stacked along N; partial fp32 [nsplit][cin][3*cout] workspace.
 *                 dy[t]*row_stride + dx[t] must be a multiple of 8 (TMA needs 16-byte aligned inner coordinates):
 *                 use row_stride % 8 == 0, dx = 0 and a zt copy shifted by the tap's horizontal offset. */
int tcv_transpose_pad(const void* x, long long x_plane, int n, int h, int w, int c, int mul, int off_y, int off_x,
                      int row_stride, int shift, void* xt, long long xt_plane, long long ktot, tcv_stream_t stream);
int tcv_wgrad_tc(const void* xt, long long xt_plane, const void* zt, long long zt_plane, int cin, int cout,
                 long long ktot, int row_stride, int ntaps, const int* dy, const int* dx, const int* wtap,
                 float* partial, int nsplit, float* dw, int dw_cout, tcv_stream_t stream);

/* Same weight gradient, read straight from the NHWC tensors: MN-major tcgen05 operands (a TMA box of 64 channels x
 * TWxTH pixels is already the canonical [K = pixel][MN = channel] SWIZZLE_128B layout), filter taps as (W, H) shifts
 * of the x box with TMA zero fill, deconv phases through a traversal stride of 2 on dz; fp32 atomics into dw.
 * d = the forward descriptor (stride 1 or 2 -- x is then read with a traversal stride --, zero padding or a pre-padded
 * input); dz split-bf16 [n, oh, ow, dz_c]. */
int tcv_conv2d_wgrad_nhwc_tc(const tcv_conv_desc* d, const void* dz, long long dz_plane, int dz_c, float* dw,
                             tcv_stream_t stream);

/* packed gradient fp32 [taps][cin_pad][cout_pad] -> torch layout ([cout,cin,kh,kw], or [cin,cout,kh,kw] when
 * transposed), minus the spectral-norm term sum_k (zdot[k]/sigma[k]) * u_k v_k^T when calls > 0. */
int tcv_weight_grad_unpack(const float* dw, int cout, int cin, int kh, int kw, int transposed, int cin_pad,
                           int cout_pad, const float* u_hist, const float* v_hist, const float* sigma,
                           const double* zdot, int calls, float* grad, tcv_stream_t stream);

/* C[b][m][n] (+)= sum_k A[b][m*sam + k*sak] * B[b][n*sbn + k*sbk], fp32, one of (sam, sak) and one of
 * (sbn, sbk) equal to 1 (CUDA cores; the exact path of the attention backward GEMMs) */
int tcv_gemm_f32_strided(const float* A, long long sam, long long sak, const float* B, long long sbn, long long sbk,
                         float* C, long long ldc, int M, int N, int K, long long strideA, long long strideB,
                         long long strideC, int batch, int accumulate, tcv_stream_t stream);

/* guided contextual attention backward (see tcv_gca_* above for the forward tensors)
 *  fold_bwd:    dY split [n,h,w,128] -> dO fp32 [n,P,2048] = unfold(dY)/4 ; delta fp32 [n,P] = rowsum(dO*O)
 *  softmax_bwd: dS = A*(dA - delta) in place on dA fp32 [n,P,P_pad] (pad columns zeroed); A fp32 [n,P,P_pad]
 *  values_bwd:  dV fp32 [n,P,2048] -> dfeat split [n,h,w,128] (gradient of the 4x4/stride-2 reflect-padded patches)
 *  prep_bwd:    dQ fp32 [n,P,576] += gradient through Kn = Q/max(|Q|,1e-4)*scale (dKn fp32 [n,P,576], Q fp32,
 *               mm, scales); then dg split [n,h/2,w/2,64] = gradient of the 3x3 reflect-padded patches */
int tcv_gca_fold_bwd(const void* dY, const float* O, int n, int h, int w, float* dO, float* delta, void* dO_split,
                     tcv_stream_t stream);
int tcv_gca_softmax_bwd(const float* A, float* dA, const float* delta, int n, int P, int P_pad, void* dS_split,
                        tcv_stream_t stream);
/* out[b][c][r] = in[b][r][c] on both planes of a split-bf16 matrix (rows x cols, row stride ld_in, batch stride bs_in
 * elements); output rows of ld_out >= rows elements, zero-filled beyond `rows`.  Produces the K-major operands of the
 * tensor-core GEMMs of the attention backward (dO_split / dS_split above are optional split-bf16 copies, planes
 * n*P*2048 resp. n*P*P_pad elements apart). */
int tcv_transpose_planes(const void* in, long long in_plane, int rows, int cols, long long ld_in, long long bs_in,
                         void* out, long long out_plane, long long ld_out, long long bs_out, int batch,
                         tcv_stream_t stream);
int tcv_gca_values_bwd(const float* dV, int n, int h, int w, void* dfeat, tcv_stream_t stream);
int tcv_gca_prep_bwd(float* dQ, const float* dKn, const float* Q, const float* mm, const float* scales, int n,
                     int h, int w, void* dg, tcv_stream_t stream);

/* TAM backward.  dout split [B,H,W,C]; dattb/dattf fp32 [B,win*win,H*W] gradients of the returned logits
 * (NULL: none).  dq split; dkb/dkf fp32 [B,H,W,C] accumulators (zeroed here); the gradient of v is dout itself. */
int tcv_tam_attend_bwd(const void* q, const void* kb, const void* kf, const float* mask, long long mask_stride,
                       int mh, int mw, int batch, int h, int w, int c, int window, const void* dout,
                       const float* dattb, const float* dattf, void* dq, float* dkb, float* dkf,
                       tcv_stream_t stream);

/* gradients of the losses: gl fp32 [5] (device) = dL/d(L_alpha, L_comp, L_grad, L_dt, L_att); acc = the
 * workspace tcv_losses_vmd filled in the forward.  dpred fp32 [B,S-2,1,H,W]; dattb/dattf fp32 [B,S-2,w*w,H*W/64]. */
int tcv_losses_vmd_bwd(const float* pred, const float* trimask, const float* gts, const float* attb,
                       const float* attf, const uint8_t* small_mask, const float* gt8, const double* acc,
                       const float* gl, int batch, int frames_per_sample, int h, int w, int window, float att_thres,
                       float label_smooth, float att_multiplier, float* dpred, float* dattb, float* dattf,
                       tcv_stream_t stream);


/* =====================================================================================================
 * FBA base network (config 5 "FBA+TAM forward").  The convolutions, the TAM and the
 * eval pre-processing reuse the entry points above; what FBA adds is below.
 *
 * Reference semantics replaced:
 *   tcv_ws_pack               models/FBA/layers_WS.py:13-23    Conv2d.forward (weight standardisation)
 *   tcv_gn_*                  models/FBA/layers_WS.py:26-27, models/FBA/models.py:239-243   nn.GroupNorm(32, C)
 *   tcv_space_to_depth2, tcv_s2d_pack_stem   models/FBA/resnet_GN_WS.py:98 (the 7x7 stem conv, with tcv_conv2d)
 *   tcv_maxpool3s2            models/FBA/resnet_GN_WS.py:102   nn.MaxPool2d(3, 2, 1) (indices unused by the decoder)
 *   tcv_adaptive_avgpool      models/FBA/models.py:264         nn.AdaptiveAvgPool2d(scale) of the pyramid pooling
 *   tcv_bilinear              models/VMN/VMN_FBA.py:27-30,36,41,46   F.interpolate(bilinear, align_corners=False)
 *   tcv_copy_channels         torch.cat along channels          VMN_FBA.py:31,38,43
 *   tcv_fba_encode_inputs     models/model.py:366-368,379-386  EvalModel.preprocess, TRIMAP_CHANNEL == 8
 *   tcv_fba_edt_cols/rows     utils/utils.py:12-39             dt() + trimap_transform (exact Euclidean transform)
 *   tcv_fba_cat_inputs        models/VMN/VMN_FBA.py:47         cat(x, conv_out[-6][:, :3], img, two_chan_trimap)
 *   tcv_fba_fusion            models/VMN/VMN_FBA.py:51-57, models/FBA/models.py:246-255
 *   tcv_postprocess_eval_fba  models/model.py:426-446          EvalModel.forward tail for method 'fba'
 * All activations are split-bf16 NHWC; channel counts, channel offsets and row strides are multiples of 8.
 * ===================================================================================================== */

/* w fp32 [cout,cin,kh,kw] (torch layout) -> packed fp32 [kh*kw][cin_pad][cout_pad] (tcv_conv2d's weight layout; rows
 * ci >= cin and columns co >= cout are zero).  standardize != 0: per output channel (w - mean) / (sqrt(var + 1e-12)
 * + 1e-5) with the unbiased variance over cin*kh*kw elements (layers_WS.py:16-21). */
int tcv_ws_pack(const float* w, int cout, int cin, int kh, int kw, int standardize, int cin_pad, int cout_pad,
                float* packed, tcv_stream_t stream);

/* GroupNorm over x split-bf16 [n][pixels][c] (dense), c/8 a power of two <= 256:
 *  stats:    sums[n][c][2] (double, zeroed here) = per (image, channel) sum and sum of squares
 *  finalize: per (image, group of c/groups channels) mean / biased variance -> per (image, channel)
 *            scale = gamma*invstd, shift = beta - mean*scale  (fp32 [n][c] each)
 *  apply:    y[n][pixel][y_off + ch] (row stride y_c elements, hi/lo planes y_plane apart; 0: n*pixels*y_c)
 *            = act( x*scale + shift + res ), res split-bf16 [n][pixels][c] or NULL; scale / shift are read with 16-byte
 *            loads: 16-byte aligned pointers (any cudaMalloc / torch allocation; c is a multiple of 8) */
int tcv_gn_stats(const void* x, long long x_plane, int n, long long pixels, int c, double* sums, tcv_stream_t stream);
int tcv_gn_finalize(const double* sums, int n, long long pixels, int c, int groups, const float* gamma,
                    const float* beta, float eps, float* scale, float* shift, tcv_stream_t stream);
/* tcv_gn_finalize over `copies` accumulator copies sums[copies][n][c][2] (tcv_conv_desc.stats); clear != 0: the copies
 * are zeroed after they have been read, so a recorded plan needs no memset before its next replay */
int tcv_gn_finalize_acc(double* sums, int copies, int clear, int n, long long pixels, int c, int groups, const float* gamma,
                        const float* beta, float eps, float* scale, float* shift, tcv_stream_t stream);
int tcv_gn_apply(const void* x, long long x_plane, int n, long long pixels, int c, const float* scale,
                 const float* shift, const void* res, long long res_plane, int act, void* y, long long y_plane,
                 int y_c, int y_off, tcv_stream_t stream);

/* y [n, (h-1)/2+1, (w-1)/2+1, c] = 3x3 / stride 2 / pad 1 max pooling of x [n,h,w,c] */
int tcv_maxpool3s2(const void* x, int n, int h, int w, int c, void* y, tcv_stream_t stream);

/* y dense [n,s,s,c] = adaptive average pooling (bins [floor(i*h/s), ceil((i+1)*h/s))) of channels
 * [x_off, x_off+c) of x [n,h,w,x_c]; c % 64 == 0 */
int tcv_adaptive_avgpool(const void* x, long long x_plane, int n, int h, int w, int c, int x_c, int x_off, int s,
                         void* y, tcv_stream_t stream);

/* y[n,oh,ow, y_off..y_off+c) (row stride y_c) = bilinear resize (align_corners=False, torch semantics) of x dense
 * [n,ih,iw,c] */
int tcv_bilinear(const void* x, int n, int ih, int iw, int c, void* y, long long y_plane, int oh, int ow, int y_c,
                 int y_off, tcv_stream_t stream);

/* y[p][y_off + ch] = x[p][x_off + ch], ch < c, p < pixels (rows of x_c / y_c elements; planes 0: pixels*row) */
int tcv_copy_channels(const void* x, long long x_plane, int x_c, int x_off, void* y, long long y_plane, int y_c,
                      int y_off, int c, long long pixels, tcv_stream_t stream);

/* The 7x7 / stride-2 / pad-3 stem (resnet_GN_WS.py:98) as a 4x4 / stride-1 convolution (taps -2..1) over the 2x2
 * space-to-depth image, so that it runs as ONE 16-tap tcgen05 launch of tcv_conv2d:
 *  space_to_depth2: x [n,h,w,c] (h, w even) -> y dense [n,h/2,w/2,4c], y[.., (py*2+px)*c + ch] = x[2Y+py, 2X+px, ch]
 *  s2d_pack_stem:   packed 7x7 weights fp32 [49][cin_pad][cout] -> fp32 [16][4*cin_pad][cout], tap (ty+2)*4+(tx+2),
 *                   row (py*2+px)*cin_pad + c = w[ky = 2ty+py+3][kx = 2tx+px+3][c] (zero outside the 7x7 support) */
int tcv_space_to_depth2(const void* x, long long x_plane, int n, int h, int w, int c, void* y, tcv_stream_t stream);
int tcv_s2d_pack_stem(const float* w49, int cin_pad, int cout, float* out, tcv_stream_t stream);
/* the same rewrite for any k x k / stride-2 convolution with zero padding `pad` (pad = 0: pre-padded input):
 * T x T / stride-1 taps t0 .. t0+T-1, t0 = floor(-pad/2), T = floor((k-1-pad)/2) - t0 + 1 (k3 p1: taps -1..0; k3 p0: 0..1).
 * src fp32 [k*k][cin_src][cout_src] -> out fp32 [T*T][4*cin_dst][cout_dst] (zero rows / columns for the padding) */
int tcv_s2d_pack(const float* src, int k, int pad, int cin_src, int cout_src, int cin_dst, int cout_dst, float* out,
                 tcv_stream_t stream);

/* EvalModel.preprocess for 'fba'.  imgs [F,3,H,W] BGR 0..255 and tris [F,1,H,W] (fp32, or uint8 when is_u8) ->
 * x16 split-bf16 [F,H,W,16]: ch 0..2 normalised RGB, 9 = (tri*1/255 == 0), 10 = (tri*1/255 == 1), 11..13 RGB/255
 * (the decoder's `img` extra; the stem's weights for channels 11..15 are zero), 14..15 zero.  Channels 3..8 are
 * written by tcv_fba_edt_rows. */
int tcv_fba_encode_inputs(const void* imgs, const void* tris, int is_u8, int frames, int h, int w, void* x16,
                          tcv_stream_t stream);
/* exact squared Euclidean distance to the nearest seed pixel of channel 9 (k = 0) / 10 (k = 1) of x16, separable:
 *  cols: g int32 [F][2][H][W] = vertical distance to the nearest seed of the same column (1 << 20: none)
 *  rows: d2 = min_x' (x-x')^2 + g[y][x']^2 ; x16 channel 3+3k+j = exp(-(sqrt(d2))^2 / (2*(f_j*320)^2)),
 *        f = (0.02, 0.08, 0.16) (utils.py:34-37); 0 when the image holds no seed of that kind, and exact 0 instead
 *        of values below 1e-12 for d^2 > 145000 (the search radius is capped there) */
int tcv_fba_edt_cols(const void* x16, int frames, int h, int w, int* g, tcv_stream_t stream);
int tcv_fba_edt_rows(const int* g, int frames, int h, int w, void* x16, tcv_stream_t stream);

/* channels [y_off, y_off+8) of y = x16 channels (0,1,2, 11,12,13, 9,10); [y_off+8, y_off+32) = 0.
 * x16_plane: elements between the hi and lo plane of x16 (0: pixels*16) */
int tcv_fba_cat_inputs(const void* x16, long long x16_plane, long long pixels, void* y, long long y_plane, int y_c,
                       int y_off, tcv_stream_t stream);

/* o8 split-bf16 [n,H,W,8] (7 used) + x16 of the same frames (first image at x16, images x16_img_stride elements
 * apart) -> pred fp32 [n,7,H,W]: alpha = clamp(o0,0,1), F = sigmoid(o1..3), B = sigmoid(o4..6), then fba_fusion */
int tcv_fba_fusion(const void* o8, const void* x16, long long x16_plane, long long x16_img_stride, int n, int h, int w,
                   float* pred, tcv_stream_t stream);

/* pred fp32 [B*(S-2),7,H,W] (inner frames) -> alphas [B,S,1,H,W], Fs, Bs [B,S,3,H,W]: where(trimask, pred, tri/255
 * resp. RGB/255), zeros for the first / last frame of every sample.  imgs/tris as in tcv_fba_encode_inputs. */
int tcv_postprocess_eval_fba(const float* pred, const void* imgs, const void* tris, int is_u8, const float* trimask,
                             int batch, int frames, int h, int w, float* alphas, float* Fs, float* Bs,
                             tcv_stream_t stream);

/* =====================================================================================================
 * DIM base network behind the TAM operator (models/VMN/VMN_DIM.py).  Convolutions
 * (3x3 / 5x5 / 7x7 as chains of <= 3x3 tap groups of tcv_conv2d), eval BatchNorm, the TAM and the eval pre/post-
 * processing reuse the entry points above; what DIM adds:
 *   tcv_maxpool2_idx     VMN_DIM.py:14,20,28,36,44   nn.MaxPool2d((2,2), stride=2, return_indices=True): split-bf16 NHWC
 *                        [n,h,w,c] -> pooled [n,h/2,w/2,c] + idx uint8 [n,h/2,w/2,c] = position of the FIRST maximum in
 *                        the window in row-major order (ky*2 + kx), torch's tie rule
 *   tcv_maxunpool2       VMN_DIM.py:82-96,113-131    nn.MaxUnpool2d((2,2), stride=2): [n,h/2,w/2,c] + idx -> [n,h,w,c], zeros
 *                        everywhere but the recorded position
 *   tcv_dim_fix_inputs   models/model.py:366-368,392 (TRIMAP_CHANNEL == 1): channel 3 of the 8-channel input tensor written
 *                        by tcv_preprocess_eval := tri / 255, channels 4..7 := 0  (tris fp32 or uint8 [F,1,H,W]) */
int tcv_maxpool2_idx(const void* x, int n, int h, int w, int c, void* y, uint8_t* idx, tcv_stream_t stream);
int tcv_maxunpool2(const void* x, const uint8_t* idx, int n, int h, int w, int c, void* y, tcv_stream_t stream);
int tcv_dim_fix_inputs(const void* tris, int is_u8, int frames, int h, int w, void* x8, tcv_stream_t stream);
/* alpha head of the DIM decoder (VMN_DIM.py:97,135: alpha_pred = Conv2d(64, 1, 5, padding=2) then .clamp(0, 1)) in one HBM-
 * bound pass: x split-bf16 NHWC [n,h,w,64], wt fp32 [25][64] (tap-major, tap = ky*5+kx), bias fp32 [1] or NULL ->
 * pred fp32 [n,h,w] */
int tcv_head_conv5_clamp01(const void* x, long long x_plane, int n, int h, int w, const float* wt, const float* bias,
                           float* pred, tcv_stream_t stream);

/* =====================================================================================================
 * IndexNet base network behind the TAM operator (models/Index, models/VMN/VMN_Index.py).
 * Dense convolutions (1x1, 3x3, 4x4 / stride 2, 5x5 as tap-group chains), eval BatchNorm + ReLU6 (TCV_ACT_RELU6), the
 * ASPP pooling branch (tcv_adaptive_avgpool, tcv_bilinear, tcv_copy_channels), the TAM and the eval pre/post-processing
 * reuse the entry points above; what IndexNet adds (split-bf16 NHWC, channel counts multiples of 8):
 *   tcv_dwconv3x3     net.py:42-44,52-54, hlaspp.py:40  depthwise 3x3 (dilation = padding = dil) + BatchNorm affine +
 *                     activation.  wt fp32 [9][c]; border fp32 [c] or NULL = value an out-of-image tap reads: the
 *                     reference's InvertedResidual pads the block input and runs its 1x1 expansion + BN + ReLU6 over the
 *                     padded tensor (net.py:62-83), so the depthwise conv sees relu6(BN shift) there, not zero.  wt / scale /
 *                     shift / border are read with 16-byte loads (16-byte aligned pointers)
 *   tcv_index_finish  hlindex.py:155-166  four branch outputs [n,h2,w2,c] -> idx_en = softmax over the branches of
 *                     sigmoid(branch), idx_de = sigmoid(branch), both [n,2*h2,2*w2,c] (pixel shuffle: branch k lands
 *                     on sub-pixel (k / 2, k % 2))
 *   tcv_index_pool    net.py:193-194 etc.  masked = idx_en * x [n,h,w,c], pooled = 4 * avg_pool2(masked) [n,h/2,w/2,c]
 *   tcv_index_upcat   hldecoder.py:121-127  cat [n,h,w,cat_c] = [ idx * nearest_up(dec)[:dec_real] | low[:low_real] | 0 ];
 *                     idx == NULL: dec is taken as is (up must be 0); idx_plane / low_plane: elements between the hi and lo
 *                     plane (0: dense; idx / low may be image slices of a larger tensor) */
int tcv_dwconv3x3(const void* x, int n, int h, int w, int c, int dil, const float* wt, const float* scale, const float* shift,
                  const float* border, int act, void* y, tcv_stream_t stream);
int tcv_index_finish(const void* b0, const void* b1, const void* b2, const void* b3, int n, int h2, int w2, int c,
                     void* idx_en, void* idx_de, tcv_stream_t stream);
int tcv_index_pool(const void* x, const void* idx_en, int n, int h, int w, int c, void* masked, void* pooled,
                   tcv_stream_t stream);
int tcv_index_upcat(const void* dec, int dec_c, int dec_real, int up, const void* idx, int idx_c, long long idx_plane,
                    const void* low, int low_c, long long low_plane, int low_real, int n, int h, int w, int cat_c, void* cat,
                    tcv_stream_t stream);

/* Evaluation metrics of one frame (calc_metric.py:22-46,74-98) in one pass over the unknown
 * region 0 < tri < 255.  alpha / gt / tri (and the NEXT frame's next_alpha / next_gt, or NULL) uint8 [h, w] as read from
 * the PNGs, flow fp32 [h, w, 2] current -> next in pixels with NaN = invalid (NULL: no warped metric).
 * out double[7] = (pixel_count, sum |a-g|, sum (a-g)^2, sum ((a-ha)-(g-hg))^2, sum |(a-g)-(pa-pg)|, sum |(a-g)^2-(pa-pg)^2|,
 * flow_pixel_count) with a = alpha / 255 etc. and pa / pg the next frame sampled bilinearly at x + flow (zero padding):
 * mSAD = out[1]/out[0], MSE = out[2]/out[0], SSDA = sqrt(out[2]), dtSSD = sqrt(out[3]), MESSDdt_fix = out[4],
 * MESSDdt = out[5]. */
int tcv_frame_metrics(const uint8_t* alpha, const uint8_t* gt, const uint8_t* tri, const uint8_t* next_alpha,
                      const uint8_t* next_gt, const float* flow, int h, int w, double* out, tcv_stream_t stream);

/* C[b] = A[b] . B[b]^T (bf16x3, fp32 out) on the CTA-pair kernel (gemm_tc2.cu) with either operand
 *   K-major  (x_mn == 0): split-bf16 [batch][rows][ld], the reduction index contiguous (K <= ld), or
 *   MN-major (x_mn != 0): split-bf16 [batch][K][ld], the ROW index contiguous (rows <= ld); any K (TMA zero-fills the tail):
 * the operand of a "transposed" product is read as its producer left it -- the attention backward (dF = A2^T.dO2,
 * dKn = dS^T.Q, dA2 = dO2.F^T, dQ = dS.Kn) needs no transposed copies.  M >= 512, N >= 256; planes / strides in elements. */
int tcv_gemm_tc_ex(const void* A, long long a_plane, long long a_ld, long long a_batch_stride, int a_mn, const void* B,
                   long long b_plane, long long b_ld, long long b_batch_stride, int b_mn, float* C, int M, int N, int K,
                   long long ldc, long long c_batch_stride, int batch, tcv_stream_t stream);

/* ---- training side of the shift-sum aggregation (csrc/gca_train2.cu; autograd of GCA/ops.py:112-118,204):
 *  unfold_parity_bwd:  dY split-bf16 [n,h,w,128] -> dO2 planes [2][n][Pk][512] = gradient of tcv_gca_unfold_parity (x 1/4)
 *  values_parity_bwd:  dF fp32 [n][Pk][512] -> dfeat split-bf16 [n,h,w,128] = gradient of tcv_gca_values_parity (the reflect
 *                      border makes rows / columns 1 and h-2 / w-2 receive two contributions) */
int tcv_gca_unfold_parity_bwd(const void* dY, int n, int h, int w, void* dO2, tcv_stream_t stream);
int tcv_gca_values_parity_bwd(const float* dF, int n, int h, int w, void* dfeat, tcv_stream_t stream);
/* fused backward of tcv_gca_rowstats(normalise) + tcv_gca_shift_add on the padded key grid: A fp32 [n][P][ld] (probabilities,
 * zeros at the pad keys), dA2 fp32 [n][Pk][ld] -> dS split-bf16 planes [2][n][P][ld] = A * (gather(dA2) - <A, gather(dA2)>) */
int tcv_gca_softmax_bwd_grid(const float* A, const float* dA2, int n, int h, int w, int ld, void* dS, tcv_stream_t stream);
/* tcv_gca_prep_bwd with the key gradient on the padded grid of tcv_gca_prep_grid: dKn_grid fp32 [n][Pk][576] */
int tcv_gca_prep_bwd_grid(float* dQ, const float* dKn_grid, const float* Q, const float* mm, const float* scales, int n,
                          int h, int w, void* dg, tcv_stream_t stream);

/* ---- SyncBatchNorm statistic exchange over NVLink peer memory (train_ddp.py:273 nn.SyncBatchNorm; csrc/peer_reduce.cu).
 * In-place sum of `count` doubles over `world` ranks of one node in ONE kernel on the caller's stream.  `peers_dev`: device
 * array of `world` pointers, entry r = rank r's symmetric buffer (tcv_peer_buffer_bytes(slot_doubles) bytes, zero-filled
 * before the first call, mapped into this process: torch.distributed._symmetric_memory or CUDA IPC provide that).
 * `epoch`: 1, 2, 3, ... -- the same on every rank for the same call.  The sum runs in rank order: all ranks get identical
 * bits.  A peer that never arrives makes the kernel trap after ~2 s instead of hanging. */
int tcv_peer_allreduce_f64(double* data, int count, void* const* peers_dev, int rank, int world,
                           unsigned long long epoch, long long slot_doubles, tcv_stream_t stream);
int tcv_peer_buffer_bytes(long long slot_doubles, long long* bytes);

#ifdef __cplusplus
}
#endif
#endif
