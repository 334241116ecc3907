/* rnagan_b200.h -- C ABI of the B200-native RNA-GAN hot path (librnagan_b200.so).
 *
 * The reference (gevaertlab/RNA-GAN) has NO native/FFI boundary: its hot path is Python calling ATen
 * (nn.ConvTranspose2d / nn.Conv2d / nn.BatchNorm2d / nn.Linear / autograd / torch.optim.Adam).  Each entry point
 * below therefore cites the reference call site whose ATen work it replaces.  All pointers are DEVICE pointers
 * borrowed from the caller (the library never allocates persistent memory), every call is asynchronous on the
 * given stream, and the return value is 0 on success, a negative RG_E* code for bad arguments, or a positive
 * cudaError_t.  rg_last_error() returns a human-readable message for the last failure on the calling thread.
 *
 * Layout vocabulary
 *   "link"      one stride-2 4x4 connection between a low-resolution NHWC tensor lo[B,H,W,Cp] and a
 *               high-resolution NHWC tensor hi[B,2H,2W,Cs] with the torch weight W[Cp][Cs][4][4] (fp32):
 *               nn.Conv2d(Cs->Cp,4,2,1) has weight [Cout=Cp][Cin=Cs][4][4]   (critic,    torchgan DCGANDiscriminator)
 *               nn.ConvTranspose2d(Cp->Cs,4,2,1) has weight [Cin=Cp][Cout=Cs][4][4] (generator, src/dcgan.py:52)
 *               so both directions of both networks are the same three contractions: DOWN, UP, WGRAD.
 *   w_down      bf16 [Cp][16*Cs]      k = (kh*4+kw)*Cs + s
 *               (rg_conv_up reads the same buffer as an MN-major B operand: one packed copy per link)
 *   w_up        bf16 [4][Cs_pad][4*Cp] phase = (y&1)*2+(x&1), k = tap*Cp + p, Cs_pad = max(Cs,16) rounded to 16
 *               (only for the <=8-channel image variant rg_conv_up_img)
 *   activations bf16 NHWC; images fp32 NCHW (the reference's tensor layout at the module boundary).
 */
#ifndef RNAGAN_B200_H
#define RNAGAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* rg_stream_t; /* cudaStream_t */

#define RG_EINVAL (-1)   /* bad shape / alignment / unsupported size */
#define RG_EARCH (-2)    /* device is not sm_100 */
#define RG_EDRIVER (-3)  /* cuTensorMapEncodeTiled unavailable or failed */
#define RG_EWORKSPACE (-4) /* caller-supplied workspace too small */
#define RG_ENOTFOUND (-5) /* rg_lmdb_get: no such key */

int rg_version(void);
const char* rg_last_error(void);
/* 0 when the current device can run the library (compute capability 10.x), RG_EARCH otherwise. */
int rg_check_device(void);
/* number of CUDA kernels this library has launched in the calling process (bench.py's gpu_launches) */
long long rg_launch_count(void);
/* profiling aid (tools/gemm_prof.py): when non-NULL, forward tile-engine launches write per-CTA clock64 totals of
 * their producer / MMA / epilogue roles into buf[grid][12]; NULL (the default) disables it. */
void rg_debug_set_prof(long long* buf);

/* ---- weight packing (after every optimizer step) --------------------------------------------------------- */
/* fp32 W[Cp][Cs][4][4] -> bf16 w_down and/or w_up (either may be NULL).  Replaces nothing in the reference: it is
 * the derived operand layout for rg_conv_down / rg_conv_up. */
int rg_pack_link(const float* W, void* w_down, void* w_up, int Cp, int Cs, rg_stream_t st);
/* generator layer 0, nn.ConvTranspose2d(E, C0, 4, 1, 0) weight [E][C0][4][4] (src/dcgan.py:38-40)
 * -> bf16 [16*C0][E]  (row = tap*C0 + co), the B operand of rg_gemm_nt. */
int rg_pack_proj(const float* W, void* w_proj, int E, int C0, rg_stream_t st);
/* bf16 w_down[Cp][16*Cs] -> bf16 w_up[4][Cs_pad][4*Cp] (the K-major copy narrow layers keep for rg_conv_up) */
int rg_pack_up_from_down(const void* w_down, void* w_up, int Cp, int Cs, rg_stream_t st);
/* image-side (3-channel) link, K padded to 64: bf16 w_col[Cp][64], k = (kh*4+kw)*4 + c. */
int rg_pack_edge(const float* W, void* w_col, int Cp, int Cimg, rg_stream_t st);
/* fp32 [rows][cols] -> bf16 [rows][cols_pad] (zero padded), nn.Linear weights (src/betaVAE.py:31,76). */
int rg_cast_pad_bf16(const float* src, void* dst, int rows, int cols, int cols_pad, rg_stream_t st);

/* ---- dense contractions on tcgen05 ----------------------------------------------------------------------- */
/* lo[b,i,j,p] = sum_{kh,kw,s} hi[b,2i-1+kh,2j-1+kw,s] * W[p,s,kh,kw]
 * critic forward nn.Conv2d(4,2,1) (torchgan DCGANDiscriminator; args src/histopathology_gan.py:186-192) and
 * generator dgrad (autograd of src/dcgan.py:52). */
/* Optional fused elementwise backward in the epilogue of an input-gradient contraction (rg_conv_down / rg_conv_up /
 * rg_gemm_nt_ld with a bf16 output; pass NULL for none): `aux` is a bf16 tensor with the layout of the output.
 *   mode 1  the layer below has no BatchNorm: out = acc * LeakyReLU'(aux), aux = its stored activation h
 *           (autograd of nn.LeakyReLU(0.2) after the critic's first Conv2d [tg])
 *   mode 2  BatchNorm2d + LeakyReLU below: aux = the pre-BN activation a; out = du = acc * LeakyReLU'(scale*a + shift),
 *           and with stats_ws the per-CTA partials become S(du), S(du * xhat) with xhat = (a - mean) * rstd, i.e.
 *           rg_bn_bwd_reduce without a pass over dh and a; finish with rg_reduce_partials + rg_bn_bwd_apply(du_in=1). */
typedef struct rg_epilogue_aux {
  const void* aux;
  int mode;
  const float* mean;
  const float* rstd;
  const float* scale;
  const float* shift;
  float slope;
} rg_epilogue_aux;
/* out[k][c] = sum over the rg_stats_parts() rows of stats_ws[part][k][c] (fixed order), KC = K*C values */
int rg_reduce_partials(const float* stats_ws, int KC, float* out, rg_stream_t st);

/* stats_ws (all three bf16 convolutions; may be NULL): fp32 [rg_stats_parts()][2][C_out] of rg_stats_ws_bytes(C_out)
 * bytes.  When given, the epilogue also accumulates per-channel sums and sums of squares of the STORED (bf16-rounded)
 * outputs, one row per CTA in a fixed order (deterministic) -- the nn.BatchNorm2d batch statistics
 * without a second pass over the activation; finish with rg_bn_finalize_partials.  Needs C_out % 64 == 0. */
size_t rg_stats_ws_bytes(int C);
int rg_stats_parts(void);
int rg_conv_down(const void* hi, const void* w_down, void* lo, int B, int H, int W, int Cs, int Cp, float* stats_ws,
                 const rg_epilogue_aux* aux, rg_stream_t st);
/* hi[b,y,x,s] = sum lo[b,i,j,p] * W[p,s,kh,kw] over y=2i-1+kh, x=2j-1+kw
 * generator forward nn.ConvTranspose2d(4,2,1) (src/dcgan.py:52) and critic dgrad. */
/* w: either w_up (w_is_down=0; K-major B, best for Cs <= 128 where the extra packed copy is tiny) or w_down
 * (w_is_down=1; read as an MN-major B operand, so large layers keep a single packed copy). */
/* w_is_down = 2: w is the merged-phase operand of rg_pack_up9_from_down (Cs == 64, at least two 128-pixel M tiles): one
 * tile then contracts all four output phases, fetching the 9 distinct shifted input tiles once instead of 16 times. */
int rg_conv_up(const void* lo, const void* w, int w_is_down, void* hi, int B, int H, int W, int Cp, int Cs,
               float* stats_ws, const rg_epilogue_aux* aux, rg_stream_t st);
size_t rg_up9_elems(int Cp);
int rg_pack_up9_from_down(const void* w_down, void* w_up9, int Cp, int Cs, rg_stream_t st);
/* same as rg_conv_up for Cs<=16 image channels, fp32 NCHW output, optional bias + tanh
 * (generator last layer, src/dcgan.py:82; critic layer-0 dgrad). */
int rg_conv_up_img(const void* lo, const void* w_up, float* img, const float* bias, int act_tanh, int B, int H,
                   int W, int Cp, int Cimg, rg_stream_t st);
/* dW[p,s,kh,kw] = beta*dW + alpha*(*alpha_dev)*sum_{b,i,j} lo[b,i,j,p] * hi[b,2i-1+kh,2j-1+kw,s]  (fp32)
 * autograd wgrad of both conv kinds.  ws: fp32 scratch of at least rg_conv_wgrad_ws_bytes(). alpha_dev may be NULL.
 * native_layout = 0: dW is a contiguous torch tensor [Cp][Cs][4][4]; 1: dW is the physical memory of a
 * torch.channels_last tensor of that shape, i.e. [Cp][kh][kw][Cs] -- the engine's own layout (one coalesced store per
 * accumulator row, no split-K scratch when the pixel count fits one unit, and the same element order as w_down). */
size_t rg_conv_wgrad_ws_bytes(int B, int H, int W, int Cp, int Cs);
int rg_conv_wgrad(const void* lo, const void* hi, float* dW, void* ws, size_t ws_bytes, int B, int H, int W, int Cp,
                  int Cs, float alpha, const float* alpha_dev, float beta, int native_layout, rg_stream_t st);
/* generator layer 0 wgrad: dW[e,c,kh,kw] = sum_b z[b,e] * da0[b,kh,kw,c]. */
size_t rg_proj_wgrad_ws_bytes(int B, int E, int C0);
int rg_proj_wgrad(const void* z, const void* da0, float* dW, void* ws, size_t ws_bytes, int B, int E, int C0,
                  float alpha, const float* alpha_dev, float beta, int native_layout, rg_stream_t st);
/* C[M,N] = act((A[M,K] . Bw[N,K]^T) * col_scale + col_shift); A, Bw bf16 row-major (K multiple of 64), C bf16 or
 * fp32 with leading dimension ldc.  out_f32 is a flag word: bit 0 = fp32 output, bit 1 = tanh instead of LeakyReLU
 * (fp32 output only; the betaVAE decoder's Linear + Tanh, src/betaVAE.py:92).  nn.Linear(+eval BatchNorm1d+LeakyReLU) of the encoder (src/betaVAE.py:29-36),
 * generator layer 0 (src/dcgan.py:38-40), image-side im2col GEMMs. */
int rg_gemm_nt(const void* A, const void* Bw, void* C, int M, int N, int K, int ldc, const float* col_scale,
               const float* col_shift, float slope, int out_f32, rg_stream_t st);
/* same with explicit leading dimensions (K, N arbitrary: TMA zero-fills the k tail) */
int rg_gemm_nt_ld(const void* A, int lda, const void* Bw, int ldb, void* C, int M, int N, int K, int ldc,
                  const float* col_scale, const float* col_shift, float slope, int out_f32, rg_stream_t st);
/* bf16 C[M,N] = A[M,K] . Bw[N,K]^T with fused statistics and / or the fused elementwise backward (see rg_epilogue_aux):
 * the generator's image-side input gradient feeding its last BatchNorm (autograd of src/dcgan.py:82). */
int rg_gemm_nt_bwd(const void* A, int lda, const void* Bw, int ldb, void* C, int M, int N, int K, int ldc,
                   float* stats_ws, const rg_epilogue_aux* aux, rg_stream_t st);
/* C[M,N] = act((A[M,K] . Bw[K,N]) * col_scale + col_shift) with Bw row-major [K][N]: the input gradient of an
 * nn.Linear whose weight is [out=K][in=N] (autograd of src/betaVAE.py:31,76,86; betaVAE training, config 5). */
int rg_gemm_nn(const void* A, int lda, const void* Bw, int ldb, void* C, int M, int N, int K, int ldc,
               const float* col_scale, const float* col_shift, float slope, int out_f32, rg_stream_t st);
/* C[M,N] (fp32, row-major) = beta*C + alpha*(*alpha_dev) * sum_r A[r,M]^T B[r,N]; A,B bf16 row-major [R][M],[R][N]. */
size_t rg_gemm_tn_ws_bytes(int R, int M, int N);
int rg_gemm_tn(const void* A, const void* Bm, float* C, void* ws, size_t ws_bytes, int R, int M, int N, float alpha,
               const float* alpha_dev, float beta, rg_stream_t st);
int rg_gemm_tn_ld(const void* A, int lda, const void* Bm, int ldb, float* C, void* ws, size_t ws_bytes, int R, int M,
                  int N, float alpha, const float* alpha_dev, float beta, rg_stream_t st);

/* ---- HBM-bound kernels (rg_ops.cu) ----------------------------------------------------------------------- */
/* Activations are bf16 NHWC viewed as [M rows][C channels]; per-channel vectors are fp32 [C].  Reductions need a
 * scratch buffer of rg_reduce_ws_bytes(M, C) bytes. */
size_t rg_reduce_ws_bytes(int M, int C);
/* nn.BatchNorm2d training forward (11 instances): sums[0][C]=sum a, sums[1][C]=sum a^2 */
int rg_bn_stats(const void* a, int M, int C, void* ws, size_t ws_bytes, float* sums, rg_stream_t st);
/* mean/rstd/scale=gamma*rstd/shift=beta-mean*scale; running stats with momentum and unbiased variance; counter += 1 */
int rg_bn_finalize(const float* sums, const float* gamma, const float* beta, int M, int C, float eps, float momentum,
                   float* running_mean, float* running_var, int64_t* num_batches_tracked, float* mean, float* rstd,
                   float* scale, float* shift, rg_stream_t st);
/* same from the per-CTA partial sums a convolution epilogue left in stats_ws (fixed-order reduction over the
 * rg_stats_parts() rows); sums_out (optional) receives sums[0][C], sums[1][C] */
int rg_bn_finalize_partials(const float* stats_ws, const float* gamma, const float* beta, int M, int C, float eps,
                            float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                            float* sums_out, float* mean, float* rstd, float* scale, float* shift, rg_stream_t st);
/* h = LeakyReLU(scale*a + shift) */
int rg_bn_act(const void* a, const float* scale, const float* shift, float slope, void* h, int M, int C,
              rg_stream_t st);
/* backward of LeakyReLU(BN(a)): sums[0]=S(du), sums[1]=S(du*xhat), du = dh*lrelu'(u) */
int rg_bn_bwd_reduce(const void* dh, const void* a, const float* mean, const float* rstd, const float* scale,
                     const float* shift, float slope, int M, int C, void* ws, size_t ws_bytes, float* sums,
                     rg_stream_t st);
/* da = scale*(du - S(du)/M - xhat*S(du*xhat)/M) (+ add); du_out optional (kept for the gradient-penalty pass) */
int rg_bn_bwd_apply(const void* dh, const void* a, const void* add, const float* mean, const float* rstd,
                    const float* scale, const float* shift, float slope, const float* sums, int M, int C, void* da,
                    void* du_out, rg_stream_t st);
/* dgamma = acc_gamma*dgamma + S(du*xhat); dbeta = acc_beta*dbeta + S(du) */
int rg_bn_param_grads(const float* sums, float* dgamma, float* dbeta, int C, float acc_gamma, float acc_beta,
                      rg_stream_t st);
/* da = dh * lrelu'(h) for the BatchNorm-free first critic layer */
int rg_lrelu_bwd(const void* dh, const void* h, float slope, void* da, int M, int C, rg_stream_t st);
/* out[c] = acc*out[c] + sum_rows x[row][c]  (bias gradients); tmp: fp32 [C] */
int rg_col_sum(const void* x, int M, int C, void* ws, size_t ws_bytes, float* tmp, float* out, float acc,
               rg_stream_t st);
/* double backward of BatchNorm inside the gradient penalty (autograd.grad(create_graph=True), src/wgan_loss.py:34-41;
 * formulas SURVEY.md Appendix C): q[0]=S(ggI), q[1]=S(ggI*xhat), q[2]=S(ggI*gO) */
int rg_bn_gp_reduce(const void* ggI, const void* a, const void* gO, const float* mean, const float* rstd, int M, int C,
                    void* ws, size_t ws_bytes, float* q, rg_stream_t st);
int rg_bn_gp_apply(const void* ggI, const void* a, const void* gO, const float* mean, const float* rstd,
                   const float* gamma, const float* scale, const float* shift, float slope, const float* s,
                   const float* q, int M, int C, void* A_dh, void* A_a, float* dgamma, float dgamma_acc,
                   rg_stream_t st);
/* latent = standardise_0(noise + z), unbiased std (src/wgan_loss.py:105-106, src/gan_utils.py:215-216);
 * z_rows == 1 broadcasts one profile (generate_images). Outputs bf16 and/or fp32 (either may be NULL). */
int rg_latent_prep(const float* noise, const float* z, int B, int E, int z_rows, void* lat_bf16, float* lat_f32,
                   rg_stream_t st);
/* im2col of an fp32 NCHW image for the 3-channel 4x4/s2/p1 link: col bf16 [B*(S/2)^2][64], k = tap*4 + c.
 * mode 0: x*mul; mode 1: (eps*x + (1-eps)*y)*mul (src/wgan_loss.py:377); mode 2: x*(1-y^2)*mul (tanh backward). */
int rg_im2col_img(const float* x, const float* y, int mode, const float* eps_dev, const float* mul_dev, int B,
                  int Cimg, int S, void* col, float* mixed_out, rg_stream_t st);
int rg_img_channel_sum(const float* x, const float* y, int mode, int B, int Cimg, int S, float* partial_ws,
                       int partial_len, float* out, float acc, rg_stream_t st);
/* image-side transposed conv in "dgrad form": col[pix][tap*Cimg+c] (fp32, from rg_gemm_nt with w_colT) is folded back
 * onto the 2x larger image: img[b,c,y,x] = act(bias[c] + sum of the 4 taps hitting (y,x)); fp32 NCHW output.
 * Generator last layer (ConvTranspose2d(64,3,4,2,1)+Tanh, src/dcgan.py:82) and the critic's layer-0 dgrad.
 * act_tanh is a flag word: bit 0 = tanh, bit 1 = write the synthesis output instead, (v + 1) / 2 as fp32 NHWC
 * [B, 2H, 2W, Cimg] (src/gan_utils.py:236-241), which saves the separate un-normalise + permute pass; bit 2 = write
 * uint8 NHWC tiles trunc(255 * (v + 1) / 2) instead (`img` then points to B*2H*2W*Cimg bytes: what
 * src/generate_tissue_images.py:127-129 computes on the host before cv2.imwrite), bit 3 = with bit 2, reversed channel
 * order (cv2's BGR). */
int rg_col2im_img(const float* col, int ldc, const float* bias, int act_tanh, int B, int Cimg, int H, int W, float* img,
                  rg_stream_t st);
/* ---- fused image-side convolutions (csrc/rg_img.cu): no materialised im2col / col2im buffer --------------------------
 * The 64-channel side is bf16 NHWC [B][S/2][S/2][64]; the image side fp32 NCHW [B][Cimg][S][S] (Cimg <= 4); weights are
 * the nn.Module parameters themselves, fp32 [64][Cimg][4][4] (Conv2d(Cimg,64).weight is [64][Cimg][4][4];
 * ConvTranspose2d(64,Cimg).weight is [64][Cimg][4][4] too), read directly -- no packed copies.
 *
 * rg_img_conv_up: ConvTranspose2d(64, Cimg, 4, 2, 1) (+bias, Tanh) = the generator's output block (torchgan DCGANGenerator
 * last block; src/dcgan.py:82 keeps the line as a comment) and the input gradient of the critic's first Conv2d
 * (autograd.grad w.r.t. the interpolate, src/wgan_loss.py:34-41).  `flags` as rg_col2im_img: bit 0 tanh, bit 1 fp32 NHWC
 * (v+1)/2 (src/gan_utils.py:236-241), bit 2 uint8 NHWC trunc(255*(v+1)/2) (src/generate_tissue_images.py:127-129),
 * bit 3 reversed channel order with bit 2.  H, W: low-resolution side (powers of two >= 8). */
int rg_img_conv_up(const void* lo, const void* wfrag, const float* bias, int flags, int B, int H, int Wd, int Cp,
                   int Cimg, void* out, const float* bn_scale, const float* bn_shift, float bn_slope, rg_stream_t st);
/* bn_scale / bn_shift (both or neither): `lo` is then the PRE-BatchNorm activation and h = lrelu(scale*a + shift, bn_slope)
 * (nn.BatchNorm2d + LeakyReLU of the generator's last hidden block) is applied to the staged tile on the fly, so forward
 * passes that are not differentiated (synthesis, the generator pass of the critic / penalty steps) never write h. */
/* wfrag: the weight W fp32 [64][Cimg][4][4] re-arranged into per-lane mma.sync B fragments (bf16), rg_img_conv_up_pack_bytes()
 * bytes; refresh after every change of W (12 KB, one tiny launch). */
size_t rg_img_conv_up_pack_bytes(void);
int rg_img_conv_up_pack(const float* W, int Cp, int Cimg, void* wfrag, rg_stream_t st);
/* rg_img_conv_down: Conv2d(Cimg, 64, 4, 2, 1) (+bias, LeakyReLU(slope); slope 1 = none) = the critic's first block
 * (torchgan DCGANDiscriminator) and the input gradient of the generator's output block.  The image operand is
 * transformed while it is staged: mode 0: x * mul_dev[0] (mul_dev may be NULL); mode 1: eps_dev[0]*x + (1-eps_dev[0])*y,
 * the gradient-penalty interpolate (src/wgan_loss.py:376-380); mode 2: x * (1 - y^2), the Tanh backward with y the
 * generator output.  mask_src (optional, bf16 like out): out *= (mask_src > 0 ? 1 : mask_slope), the LeakyReLU backward
 * mask of the double-backward sweep.  S: image side (power of two >= 16). */
int rg_img_conv_down(const float* x, const float* y, int mode, const float* eps_dev, const float* mul_dev,
                     const float* W, const float* bias, float slope, const void* mask_src, float mask_slope, int B,
                     int Cimg, int S, int Cp, void* out, rg_stream_t st);
/* rg_img_conv_wgrad: dW[p][c][kh][kw] = acc*dW + sum act[b,y,x,p] * img'[b,c,2y-1+kh,2x-1+kw] with img' the image
 * operand transformed as in rg_img_conv_down -- the weight gradient of either layer (Conv2d: act = d(out);
 * ConvTranspose2d: act = the layer input, img' = d(pre-activation)); dbias (optional, Cimg <= 3):
 * dbias[p] = acc_bias*dbias + sum act[..,p] (the Conv2d bias gradient).  Two-stage, fixed order (bit-reproducible).
 * ws: rg_img_conv_wgrad_ws_bytes() bytes of scratch. */
size_t rg_img_conv_wgrad_ws_bytes(void);
int rg_img_conv_wgrad(const void* act, const float* x, const float* y, int mode, const float* eps_dev,
                      const float* mul_dev, int B, int Cimg, int S, int Cp, void* ws, size_t ws_bytes, float* dW,
                      float acc, float* dbias, float acc_bias, rg_stream_t st);
/* W[Cp][Cimg][4][4] -> bf16 w_colT[rows][Cp], row n = tap*Cimg + c (rows beyond 16*Cimg zero) */
int rg_pack_edge_t(const float* W, void* w_colT, int Cp, int Cimg, int rows, rg_stream_t st);
int rg_unpack_edge_grad(const float* dcol, float* dW, int Cp, int Cimg, float acc, rg_stream_t st);
/* critic head `disc` = Conv2d(C,1,4,1,0)+LeakyReLU on a 4x4 map (torchgan DCGANDiscriminator): w_head[k=tap*C+c] */
int rg_pack_head(const float* W, float* w_head, int C, rg_stream_t st);
int rg_head_fwd(const void* h5, const float* w_head, int B, int K, float slope, float* a6, float* out, rg_stream_t st);
int rg_head_bwd_data(const float* a6, const float* dout, float dout_const, const float* w_head, int B, int K,
                     float slope, float* da6, void* dh5, rg_stream_t st);
int rg_head_wgrad(const float* da6, const void* x, int B, int K, int C, float* dW, float acc, rg_stream_t st);
/* loss_out[0] = mean(sign_a*a) + mean(sign_b*b)  (src/wgan_loss.py:24-29) */
int rg_wgan_loss(const float* a, float sign_a, const float* b, float sign_b, int B, float* loss_out, rg_stream_t st);
/* out3 = {(||g||-1)^2, lambda*2*(||g||-1)/||g||, ||g||}: whole-batch Frobenius norm (src/wgan_loss.py:43) */
int rg_gp_norm(const float* g, size_t n, float lambd, float* partial_ws, int partial_len, float* out3, rg_stream_t st);
/* torch.optim.Adam(lr, betas, eps; no weight decay/amsgrad) over a table of tensors (src/histopathology_gan.py:252,257),
 * with the optional WGAN weight clamp (src/wgan_loss.py:213-215) fused in.  rg_adam_build_table fills a HOST table
 * and returns the number of chunks (>0); copy it to the device and pass it to rg_adam_step. */
int rg_adam_table_bytes(int num_chunks);
/* shadows (may be NULL, entries may be NULL): per tensor a bf16 buffer of the same element count that receives the
 * updated parameters in the same element order -- for a channels_last conv weight that IS the packed w_down operand,
 * so the optimiser step re-emits the GEMM operands and no pack kernel runs afterwards. */
int rg_adam_build_table(void* const* params, void* const* grads, void* const* ms, void* const* vs,
                        void* const* shadows, const int64_t* sizes, int num_tensors, int chunk_elems, void* table_host,
                        int max_chunks);
/* the same with PITCHED shadows: tensor i is a dense [rows][shadow_cols[i]] matrix whose bf16 copy has rows of
 * shadow_pitch[i] >= shadow_cols[i] elements (nn.Linear weights whose K is padded for the 16-byte TMA stride rule, e.g.
 * 19198 -> 19200 genes); pad columns are never written.  shadow_cols / shadow_pitch may be NULL (all dense). */
int rg_adam_build_table_pitched(void* const* params, void* const* grads, void* const* ms, void* const* vs,
                                void* const* shadows, const int* shadow_cols, const int* shadow_pitch,
                                const int64_t* sizes, int num_tensors, int chunk_elems, void* table_host,
                                int max_chunks);
/* grad_scale multiplies every gradient first (1/world_size after a SUM all-reduce; 1.0 otherwise) */
int rg_adam_step(const void* table_dev, int num_chunks, float lr, float beta1, float beta2, float eps, int step,
                 int do_clamp, float clamp_lo, float clamp_hi, float grad_scale, rg_stream_t st);
/* the same step with the learning rate and the step count read from DEVICE memory, dyn = {lr, step}: the call first
 * advances dyn[1] by one, then runs the update with bias corrections 1 - beta^dyn[1].  Nothing in the launch arguments
 * changes between steps, so a captured CUDA graph of a whole optimiser step can be replayed. */
int rg_adam_step_dyn(const void* table_dev, int num_chunks, float* dyn, float beta1, float beta2, float eps,
                     int do_clamp, float clamp_lo, float clamp_hi, float grad_scale, rg_stream_t st);
int rg_clamp(float* p, size_t n, float lo, float hi, rg_stream_t st);
/* data-parallel gradient exchange through NVSwitch multicast (NVLink SHARP): in place, over the MULTICAST mapping `mc` of
 * a symmetric buffer (every rank's copy at the same offset), the calling rank reduces floats [offset, offset+n) --
 * multimem.ld_reduce (fp32 sum inside the switch) then multimem.st (broadcast to every copy).  Ranks call it for disjoint
 * slices between two barriers; replaces the NCCL all-reduce of G/D gradients (src/histopathology_gan.py runs one process;
 * SURVEY.md 8e shards by batch).  max_ctas bounds its CTAs of 128 threads (0: 128); the CTAs are small enough to co-reside with a tile-engine CTA. */
int rg_nvls_allreduce(float* mc, size_t offset, size_t n, int max_ctas, rg_stream_t st);
/* data-parallel gradient exchange (what DistributedDataParallel's all-reduce would do around the
 * reference): out[i] = sum over r < nparts of src[r * stride + i], r ascending -- the reduction step of the peer-to-peer
 * exchange, after every rank's copy of this rank's slice has been pulled into src by the copy engines.  Fixed order,
 * so the result is identical on reruns and independent of arrival order.  n and stride in floats, multiples of 4. */
int rg_slices_sum(const float* src, int nparts, size_t stride, size_t n, float* out, rg_stream_t st);
/* data path: uint8 HWC tiles [B][S][S][C] as the patch LMDBs hold them -> fp32 NCHW in [-1, 1];
 * cv2 BGR->RGB (swap_rb), permute(2,0,1), ConvertImageDtype(float) and Normalize(0.5, 0.5) of src/read_data.py:341-343
 * and src/histopathology_gan.py:106-109 in one pass, bit-identical to those CPU transforms.  Moving uint8 instead of
 * fp32 over PCIe cuts the per-batch host->device bytes by 4. */
int rg_tiles_u8_to_nchw(const void* tiles, float* img, int B, int C, int S, int swap_rb, rg_stream_t st);
/* (x+1)/2 and NCHW -> NHWC fp32 (src/gan_utils.py:236-241) */
int rg_tiles_to_unit_nhwc(const float* img, float* out, int B, int C, int S, rg_stream_t st);

/* ---- tile data path, host side (csrc/rg_data.cu -- no device work) -----------------------------------
 * What PatchRNADataset gets from the C extensions `lmdb` and `lz4framed` (src/read_data.py:284-371):
 * `lmdb.open(path, subdir=False, readonly=True, lock=False)` + `txn.get(key)` + `txn.stat()['entries']` (:314-320,
 * :346-351) and `lz4framed.decompress(value)` (:318, :332).
 * rg_lmdb_open: memory-maps one LMDB file read-only; NULL on failure (see rg_last_error()).
 * rg_lmdb_get: *val points into the mapping (valid until rg_lmdb_close); 0, RG_ENOTFOUND or RG_EINVAL (corrupt tree). */
void* rg_lmdb_open(const char* path);
void rg_lmdb_close(void* db);
int rg_lmdb_stat(void* db, unsigned long long* entries, unsigned* page_size, unsigned* depth);
int rg_lmdb_get(void* db, const void* key, size_t key_len, const void** val, size_t* val_len);
/* one LZ4 frame -> dst; returns the bytes written, -1 for a malformed frame, -2 when `cap` is too small (retry with a
 * larger buffer); *content_size = the frame header's content size field or -1 */
long long rg_lz4f_decompress(const void* src, size_t n, void* dst, size_t cap, long long* content_size);

/* ---- betaVAE training step (config 5; src/betaVAE.py:96-115, 145-162, 216-236) ---------------------------- */
/* dst = bf16(src * mul * scale), zero pad columns: train-mode Dropout(0.5) with the caller's keep mask (mul may be NULL) */
int rg_mul_cast_pad_bf16(const float* src, const float* mul, float scale, void* dst, int rows, int cols, int cols_pad,
                         rg_stream_t st);
/* mulv fp32 [B][2Z] = (z_mean | z_logvar): z = mu + eps*exp(lv/2) (bf16); partial sums of 1+lv-mu^2-exp(lv).
 * Returns the number of partials written (>0) or a negative error. */
int rg_vae_reparam(const float* mulv, const float* eps, int B, int Z, void* z, float* partial, int partial_len,
                   rg_stream_t st);
/* out = tanh(pre); d_pre = gscale*(out-x)*(1-out^2) (bf16 [B][ldp]); partial sums of (out-x)^2; returns #partials */
int rg_vae_recon(const float* pre, int ldp, const float* x, int B, int F, float gscale, void* dpre, float* partial,
                 int partial_len, rg_stream_t st);
/* dcat bf16 [B][2Z] = (dz + kscale*mu | dz*eps*0.5*exp(lv/2) - 0.5*kscale*(1-exp(lv))), kscale = beta/B */
int rg_vae_latent_grad(const void* dz, const float* mulv, const float* eps, int B, int Z, float kscale, void* dcat,
                       rg_stream_t st);
/* out3 = {total, reconstruction, kl}: MSE + beta * mean_b(-0.5*sum_j(...)) (betaVAEloss) */
int rg_vae_loss_finalize(const float* p_sse, int n1, const float* p_kld, int n2, int B, int F, float beta, float* out3,
                         rg_stream_t st);

/* ---- resize-conv generator DCGANUpGenerator (src/dcgan.py:8-99) -------------------------------------------- */
/* nn.Upsample(x2, bilinear, align_corners=False) + nn.ReflectionPad2d(1) (src/dcgan.py:48-49,78-79) and its adjoint:
 * bf16 NHWC [B,H,W,C] <-> [B,2H+2,2W+2,C] */
int rg_upsample2x_reflectpad(const void* h, void* u, int B, int H, int W, int C, rg_stream_t st);
int rg_upsample2x_reflectpad_bwd(const void* du, void* dh, int B, int H, int W, int C, rg_stream_t st);
/* nn.Conv2d(Cin, Cout, 3, 1, 0) on the padded tensor u [B,Ho+2,Wo+2,Cin] (src/dcgan.py:50-51,80-81):
 * w3 bf16 [Cout_pad][9*Cin] from rg_pack_conv3; forward (+bias), image-channel forward (fp32 NCHW), input gradient on
 * the padded grid (weights read MN-major from w3), weight gradient in the torch layout. */
int rg_pack_conv3(const float* W, void* w3, int Cout, int Cin, int rows, rg_stream_t st);
int rg_conv3x3(const void* u, const void* w3, void* out, const float* bias, int B, int Ho, int Wo, int Cin, int Cout,
               float* stats_ws, rg_stream_t st);
int rg_conv3x3_img(const void* u, const void* w3, float* img, const float* bias, int B, int Ho, int Wo, int Cin,
                   int Cimg, rg_stream_t st);
int rg_conv3x3_dgrad(const void* da, const void* w3, void* du, int B, int Ho, int Wo, int Cin, int Cout, rg_stream_t st);
size_t rg_conv3x3_wgrad_ws_bytes(int B, int Ho, int Wo, int Cin, int Cout);
int rg_conv3x3_wgrad(const void* da, const void* u, float* dW, void* ws, size_t ws_bytes, int B, int Ho, int Wo,
                     int Cin, int Cout, float beta, rg_stream_t st);
/* backward of the last Conv2d(64, 3, 3) of the resize-conv generator: dW fp32 [Cimg][64][3][3] and (optional) du bf16
 * [B,S+2,S+2,64] from dout fp32 NCHW [B,Cimg,S,S]; ws of rg_upg_last_ws_bytes() bytes */
size_t rg_upg_last_ws_bytes(int B, int S, int C, int Cimg);
int rg_upg_last_bwd(const void* u, const float* dout, const float* W, int B, int S, int C, int Cimg, float* dW, void* du,
                    void* ws, size_t ws_bytes, rg_stream_t st);

#ifdef __cplusplus
}
#endif
#endif /* RNAGAN_B200_H */
