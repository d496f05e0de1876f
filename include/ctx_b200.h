/* ctx_b200.h — C-ABI of libctx_b200.so: the context-translation encoder/decoder hot path on B200 (sm_100a).
 *
 * The reference (wyndwarrior/imitation_from_observation) has no FFI: its boundary for this path is a Python
 * graph object evaluated with sess.run.  Every dense op of that graph is a TensorFlow call
 * site; each entry point below replaces one of those call sites (cited per function, paths relative to the
 * reference root).  The Python host (imitation_from_observation_b200/) binds these with ctypes; the binding a
 * reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - All tensors are dense fp32 (unless stated), NHWC, in DEVICE memory owned by the caller.  The library
 *    allocates nothing persistent and keeps no state except the per-thread last-error string.
 *  - Every function enqueues work on `stream` (a cudaStream_t passed as void*) and returns immediately:
 *    0 = ok, negative = error (message via ctx_last_error()).  No internal synchronisation.
 *  - Preconditions (shape limits, alignment for the tensor-core path) are checked and reported, never UB.
 *  - There is NO CPU fallback: without a CUDA device every compute entry point fails with CTX_ERR_CUDA.
 *  - Gradient outputs documented "+=" ACCUMULATE into caller-zeroed buffers (split-K partial sums are
 *    reduced with fp32 atomics), so a flat gradient buffer is zeroed once per step.
 */
#ifndef CTX_B200_H_
#define CTX_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTX_B200_VERSION 200

enum {
  CTX_OK = 0,
  CTX_ERR_ARG = -1,      /* bad shape / null pointer / unsupported combination */
  CTX_ERR_CUDA = -2,     /* CUDA runtime error (launch failure, no device)      */
  CTX_ERR_UNSUPPORTED = -3
};

/* Which implementation a GEMM-shaped entry point uses. */
enum {
  CTX_IMPL_AUTO = 0,     /* tcgen05 kernel, 3-pass TF32 split (fp32-faithful, ~2^-21 per product) when the  */
                         /* shape qualifies, else the SIMT kernel                                            */
  CTX_IMPL_SIMT = 1,     /* fp32 CUDA-core kernel (exact fp32 products)                                      */
  CTX_IMPL_TC = 2,       /* force tcgen05, single-pass TF32 (operands must be TF32-rounded); error if the    */
                         /* shape is not eligible                                                            */
  CTX_IMPL_TC3 = 3,      /* force tcgen05, 3-pass split; error if the shape is not eligible                  */
  CTX_IMPL_FAST = 4,     /* single-pass TF32 when the shape qualifies, else SIMT                             */
  CTX_IMPL_P16X3 = 5,    /* reported by ctx_last_impl() after a 16-bit-planes call with three products ...   */
  CTX_IMPL_BF16X3 = 5,   /* ... (round-1 name)                                                               */
  CTX_IMPL_P16X1 = 6     /* ... and with one product (fp16 planes); neither is a selector                    */
};

/* Geometry of one SAME conv / transposed conv.  "big" is the conv input (= transposed-conv output) grid,
 * "small" the conv output (= transposed-conv input) grid:  P = ceil(H/stride), pad_* = TF SAME low pads. */
typedef struct ctx_conv_geom {
  int32_t N;             /* images                                                         */
  int32_t H, W;          /* big grid                                                       */
  int32_t P, Q;          /* small grid                                                     */
  int32_t kh, kw;
  int32_t stride;
  int32_t pad_h, pad_w;  /* rows/cols of zero padding above / left of the big grid         */
} ctx_conv_geom;

/* Fused epilogue of every GEMM-shaped kernel; applied per output element v (flat index i, column n):
 *   v += bias[n]                      if bias
 *   v  = max(v, slope*v)              if act == 1                 (lrelu, arm_shaping.py:18-19)
 *   v *= (act_ref[i] >= 0 ? 1:slope)  if act_ref                  (lrelu' in backward, from the saved output)
 *   v += out[i]                       if accumulate
 *   v  = round_to_tf32(v)             if round_tf32               (so the next tcgen05 consumer is unbiased)
 */
typedef struct ctx_epilogue {
  const float* bias;
  const float* act_ref;
  float slope;
  int32_t act;
  int32_t accumulate;
  int32_t round_tf32;
} ctx_epilogue;

int ctx_version(void);
const char* ctx_last_error(void);
/* sm count, compute capability; CTX_ERR_CUDA without a device. */
int ctx_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);
/* 1 if the tcgen05 path can run on the current device (sm_100). */
int ctx_tc_available(void);
/* Number of kernels this library has launched in this process (all streams); bench.py's "gpu_launches". */
int64_t ctx_launch_count(void);
/* Host only: CRC-32C (Castagnoli) of n bytes, continuing from `crc` (0 to start) — the checksum TensorFlow stores per
 * tensor in a tensor-bundle checkpoint (tf.train.Saver V2: scripts/train_script.py:189, rllab/sampler/base.py:144-145). */
uint32_t ctx_crc32c(const void* data, uint64_t n, uint32_t crc);

/* Which implementation the last GEMM-shaped call of this thread used: CTX_IMPL_SIMT / CTX_IMPL_TC / CTX_IMPL_TC3. */
int ctx_last_impl(void);

/* ---- convolutions ------------------------------------------------------------------------------------- */

/* y[N,P,Q,Co] = epi( sum_{r,s,c} x[n, p*st+r-pad_h, q*st+s-pad_w, c] * w[r,s,c,co] )
 * Replaces tf.nn.conv2d(SAME)+bias_add (+lrelu) — gym/envs/mujoco/arm_shaping.py:21-32 — and is the
 * data-gradient of conv2d_transpose (x = dY of the deconv, w = its [kh,kw,Cout,Cin] filter read as HWIO). */
int ctx_conv2d_fprop(const float* x, int32_t C, const float* w, float* y, int32_t Co,
                     const ctx_conv_geom* g, const ctx_epilogue* epi, int32_t impl, void* stream);

/* Same, writing into a WIDER NHWC tensor: y points at the first output channel inside it and y_ld >= Co is its channel
 * count, so that a branch of an Inception block lands directly in the block's tf.concat(axis=3) result
 * (nets/inception_v3.py:177 ...).  No act_ref / accumulate in this form. */
int ctx_conv2d_fprop_ld(const float* x, int32_t C, const float* w, float* y, int32_t y_ld, int32_t Co,
                        const ctx_conv_geom* g, const ctx_epilogue* epi, int32_t impl, void* stream);

/* y[N,H,W,Cn] = epi( sum_{p,q,r,s,k : p*st+r-pad_h == h, q*st+s-pad_w == w} x[n,p,q,k] * w[r,s,cn,k] )
 * with x the channel concatenation [x1 (C1 ch, batch N) ‖ x2 (C2 ch, batch x2_batch, image n % x2_batch)];
 * x2 may be NULL (C2 = 0).  w is [kh,kw,Cn,C1+C2].
 * Replaces tf.nn.conv2d_transpose(SAME)+bias_add over tf.concat([h, skip], 3) — arm_shaping.py:62-85,
 * 1323-1329 (the concat is never materialised) — and is the data-gradient of conv2d (x1 = dY, w = HWIO). */
int ctx_conv2d_transpose(const float* x1, int32_t C1, const float* x2, int32_t C2, int32_t x2_batch,
                         const float* w, float* y, int32_t Cn,
                         const ctx_conv_geom* g, const ctx_epilogue* epi, int32_t impl, void* stream);

/* dw[kh,kw,Ca,Cb1+Cb2] += sum_{n,p,q} big[n, p*st+r-pad_h, q*st+s-pad_w, a] * small[n,p,q,b]
 * with small = [small1 (Cb1 ch) ‖ small2 (Cb2 ch, batch small2_batch, image n % small2_batch)].
 * Filter gradient of conv2d (big = x, small = dY -> HWIO) and of conv2d_transpose (big = dY, small = x ->
 * [kh,kw,Cout,Cin]); the autodiff the reference gets from AdamOptimizer.minimize (scripts/train_script.py:128). */
int ctx_conv2d_wgrad(const float* big, int32_t Ca, const float* small1, int32_t Cb1,
                     const float* small2, int32_t Cb2, int32_t small2_batch,
                     float* dw, const ctx_conv_geom* g, int32_t impl, void* stream);

/* ---- dense layers --------------------------------------------------------------------------------------- */

/* C[M,N] = epi( opA(A) * opB(B) ), row-major with leading dimensions (elements).
 * transA = 0: A is [M,K] (lda >= K); 1: A is stored [K,M] (lda >= M).  transB likewise for [K,N] / [N,K].
 * Replaces tf.matmul(x, Matrix) + bias — arm_shaping.py:48-59 — and its two gradient GEMMs. */
int ctx_gemm(int32_t transA, int32_t transB, int32_t M, int32_t N, int32_t K,
             const float* A, int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc,
             const ctx_epilogue* epi, int32_t impl, void* stream);

/* ---- convolutions on 16-bit operand planes (csrc/tc_h16.cu) ------------------------------------------------------
 * A "planes" tensor is [nplanes][shape] of fp16 or bf16: plane 0 = hi = rn16(s*x), plane 1 = lo = rn16(s*x - hi), with s
 * a power-of-two scale (1 unless stated).  hi + lo carries 22 (fp16) / 16 (bf16) significant bits of the fp32 value.
 *   nprod = 3: hi*hi + hi*lo + lo*hi on the 16-bit tensor cores with fp32 accumulation — fp32-grade results (max-norm
 *              error 5e-6..9e-6 per GEMM, measured) at half the tensor-core time of the 3-pass TF32 split and with no operand
 *              splitting inside the GEMM: the FORWARD pass.
 *   nprod = 1: hi*hi with fp16 planes — fp16 has the TF32 significand (1 + 10 bits), rounded to nearest by the producer of
 *              the planes, so this is single-pass-TF32 accuracy at twice the tensor-core rate and half the operand bytes:
 *              the BACKWARD pass.  Gradients are brought into fp16's range by a per-tensor power-of-two scale
 *              (ctx_grad_planes); the GEMM divides its result by *a_scale.
 * Same layer semantics as ctx_conv2d_fprop / ctx_conv2d_transpose / ctx_conv2d_wgrad (arm_shaping.py:21-32, 62-85,
 * 1323-1329 and their autodiff).  Epilogue: bias, lrelu, act_ref mask, accumulate (no round_tf32).  y is written in fp32
 * and, when y_planes is not null, also as y_planes_n planes for the next layer.  Shapes: channels % 64 == 0, stride 1 or
 * 2, output (phase) grids that tile exactly into 128 pixels (64 for the filter gradient) — ask ctx_*_p16_eligible;
 * ineligible shapes return CTX_ERR_UNSUPPORTED (use the fp32-input entry points). */
enum { CTX_P16_BF16 = 0, CTX_P16_F16 = 1 };

typedef struct ctx_p16 {
  int32_t dtype;             /* CTX_P16_F16 | CTX_P16_BF16: element type of every planes operand of the call          */
  int32_t nprod;             /* 3 | 1 (1 needs fp16)                                                                   */
  const float* a_scale;      /* device scalar or NULL: the gradient operand's planes hold (*a_scale) * x; result /= it  */
  uint32_t* amax_bits;       /* device word or NULL: atomicMax of the bit patterns of |y| (input of ctx_grad_planes)    */
} ctx_p16;

/* v = scale * x[i]; planes[0][i] = rn16(v); planes[1][i] = rn16(v - planes[0][i]) (nplanes == 2) for i < n; fp16 saturates.
 * scale: a power of two.  Filters are ~N(0, 0.02): at scale 1 the lo part of an fp16 split falls into fp16's subnormals and
 * hi + lo keeps 19 bits instead of 22; the forward pass splits its filters at scale 2^8 and passes a device scalar 2^8 as
 * ctx_p16.a_scale (which divides the result). */
int ctx_split_planes(const float* x, int64_t n, float scale, void* planes, int32_t nplanes, int32_t dtype, void* stream);

/* w [T][R][Cc] fp32 -> planes [nplanes][T][Cc][R]: a conv filter [kh*kw][Cin][Cout] (arm_shaping.py:25) becomes the per-tap
 * K-major matrix [kh*kw*Cout][Cin] ctx_conv2d_fprop_p16 reads; a transposed-conv filter [kh*kw][Cout][Cin] becomes the
 * [kh*kw*Cin][Cout] matrix its data gradient (a conv fprop over dY) reads.  (Used as stored, a transposed-conv filter is
 * already K-major for its forward pass and a conv filter for its data gradient: ctx_split_planes.) */
int ctx_split_planes_transpose(const float* w, int32_t T, int32_t R, int32_t Cc, float scale, void* planes, int32_t nplanes,
                               int32_t dtype, void* stream);

int ctx_conv2d_fprop_p16_eligible(int32_t C, int32_t Co, const ctx_conv_geom* g);

/* x_planes [.][N,H,W,C], w_planes [.][kh*kw*Co][C], y [N,P,Q,Co] fp32, y_planes [y_planes_n][N,P,Q,Co] or NULL.
 * Also the data gradient of a transposed conv (x = planes of dY scaled by *cfg->a_scale, w from
 * ctx_split_planes_transpose of the [kh,kw,Cout,Cin] filter). */
int ctx_conv2d_fprop_p16(const void* x_planes, int32_t C, const void* w_planes, float* y, void* y_planes,
                         int32_t y_planes_n, int32_t Co, const ctx_conv_geom* g, const ctx_epilogue* epi,
                         const ctx_p16* cfg, void* stream);

int ctx_conv2d_transpose_p16_eligible(int32_t C1, int32_t C2, int32_t x2_batch, int32_t Cn, const ctx_conv_geom* g);

/* x1_planes [.][N,P,Q,C1], x2_planes [.][x2_batch,P,Q,C2] or NULL (image n reads x2 image n % x2_batch: the fused
 * skip concat), w_planes [.][kh*kw*Cn][C1+C2], y [N,H,W,Cn] fp32, y_planes [y_planes_n][N,H,W,Cn] or NULL.
 * Also the data gradient of a conv (x1 = planes of dY, w = planes of the HWIO filter as stored). */
int ctx_conv2d_transpose_p16(const void* x1_planes, int32_t C1, const void* x2_planes, int32_t C2, int32_t x2_batch,
                             const void* w_planes, float* y, void* y_planes, int32_t y_planes_n, int32_t Cn,
                             const ctx_conv_geom* g, const ctx_epilogue* epi, const ctx_p16* cfg, void* stream);

/* Stride-2 transposed conv with few output channels (4 * Cn == 256 or 512, e.g. d_h3: 256 -> 64 channels) as ONE GEMM over
 * all four stride phases: rows = input pixels, columns = (phase, channel), K = (distinct input offset, source, channel) —
 * the input is fetched once per distinct offset (9 for a 5x5 filter) instead of once per tap (25), in 256-wide CTA-pair
 * tiles.  The filter must first be packed (per parameter update) into wm_planes
 * [nplanes][ctx_conv2d_transpose_merged_rows(Cn, g)][Ck]: row (offset, phase, co) holds the weights of the one tap phase
 * `phase` uses at that offset, or zeros.  Same semantics as ctx_conv2d_transpose_p16. */
int ctx_conv2d_transpose_merged_p16_eligible(int32_t C1, int32_t C2, int32_t x2_batch, int32_t Cn,
                                             const ctx_conv_geom* g);
int64_t ctx_conv2d_transpose_merged_rows(int32_t Cn, const ctx_conv_geom* g);
int ctx_conv2d_transpose_merged_pack_p16(const float* w, int32_t Ck, int32_t Cn, const ctx_conv_geom* g, float scale,
                                         void* wm_planes, int32_t nplanes, int32_t dtype, void* stream);
int ctx_conv2d_transpose_merged_p16(const void* x1_planes, int32_t C1, const void* x2_planes, int32_t C2,
                                    int32_t x2_batch, const void* wm_planes, float* y, void* y_planes, int32_t y_planes_n,
                                    int32_t Cn, const ctx_conv_geom* g, const ctx_epilogue* epi, const ctx_p16* cfg,
                                    void* stream);

int ctx_conv2d_wgrad_p16_eligible(int32_t Ca, int32_t Cb1, int32_t Cb2, int32_t small2_batch, const ctx_conv_geom* g);

/* ctx_conv2d_wgrad from planes: big_planes [.][N,H,W,Ca], small1_planes [.][N,P,Q,Cb1], small2_planes
 * [.][small2_batch,P,Q,Cb2] or NULL; dw [kh,kw,Ca,Cb1+Cb2] fp32 +=.  Pixels are the GEMM's K dimension: both operands are
 * MN-major tiles, split-K over 64-pixel blocks, fp32 vector reductions.  *cfg->a_scale is the scale of whichever operand
 * is the gradient (big for a transposed-conv filter, small1 for a conv filter); the sums are divided by it. */
int ctx_conv2d_wgrad_p16(const void* big_planes, int32_t Ca, const void* small1_planes, int32_t Cb1,
                         const void* small2_planes, int32_t Cb2, int32_t small2_batch, float* dw,
                         const ctx_conv_geom* g, const ctx_p16* cfg, void* stream);

/* Gradient -> operand planes, one pass over a raw fp32 gradient x [rows, ldx] (columns 0..C-1 are used):
 *   v[r,c] = x[r*ldx + c] * (ref ? (ref[r*C + c] >= 0 ? 1 : slope) : 1)          lrelu' from the saved output
 *   s      = 2^k with amax * s in [2^14, 2^15), amax = the float whose bits are *amax_bits (0 -> s = 1; bf16: s = 1);
 *            *scale = s                                                           (device scalar, read by the consumers)
 *   planes[0][r*C + c] = rn16(s * v),  planes[1][...] = rn16(s * v - planes[0])   (nplanes == 2)
 *   db[c]             += sum_r v[r,c]                                             if db      (tf.nn.bias_add's gradient)
 *   out_f32[r*C + c]   = v                                                        if out_f32 (fp32 consumers: dense layers)
 * Replaces lrelu' + bias gradient + operand rounding of the backward pass the reference gets from
 * AdamOptimizer.minimize (scripts/train_script.py:128). */
int ctx_grad_planes(const float* x, int64_t rows, int32_t C, int32_t ldx, const float* ref, float slope,
                    const uint32_t* amax_bits, float* scale, void* planes, int32_t nplanes, int32_t dtype,
                    float* db, float* out_f32, void* stream);

/* *amax_bits = max(*amax_bits, bits(max_i |x[i]|)) — for gradients produced by kernels without the amax epilogue. */
int ctx_amax(const float* x, int64_t n, uint32_t* amax_bits, void* stream);

/* dskip[B*pix, C2] (+)= (sum_g dcat[(g*B*pix + row), C1 + c]) * lrelu'(skip_ref): the skip half of ctx_split_skip_grad
 * (the state half goes through ctx_grad_planes). */
int ctx_skip_grad_sum(const float* dcat, int32_t G, int64_t rows, int32_t C1, int32_t C2, const float* skip_ref,
                      float slope, float* dskip, int32_t accumulate, void* stream);

/* ---- the image-side layers: convolutions over a few-channel input, patch matrix built in shared memory (csrc/tc_few.cu) ----
 * The reference's first encoder layer  h0 = lrelu(conv2d(image, 5x5, stride 2, 3 -> 64))  (arm_shaping.py:1283) and the
 * backward pass of its mirror image, the last decoder layer d_h4 (128 -> 3, arm_shaping.py:1329): no im2col buffer in HBM,
 * the input is read once and the result written once.  x is [N,H,W,C] fp32, or uint8 frames normalised on the fly
 * ((u/255 - 0.5) * 2 in three fp32 roundings, base.py:117-119); kw * C <= 16, kh <= 8, stride 1 or 2, 64 or 128 output
 * channels, output grids that tile into 128 pixels — ask ctx_conv2d_few_eligible.  Operands are fp16: nprod = 3 splits the
 * input and the filter into (hi, lo) (fp32-grade, the forward pass), nprod = 1 uses the hi parts (backward).  When
 * x_amax_bits is given x is a gradient: it is scaled by the power of two 2^k with amax * 2^k in [2^14, 2^15) before the
 * fp16 conversion and the result is divided by it. */
int ctx_conv2d_few_eligible(int32_t C, int32_t Co, const ctx_conv_geom* g);

/* Bytes of the packed filter image for ctx_conv2d_few_fprop; ctx_conv2d_few_pack fills it from the HWIO filter
 * [kh,kw,C,Co] (once per parameter update).  The image is the kernel's shared-memory operand tile, swizzle included. */
int64_t ctx_conv2d_few_pack_bytes(int32_t Co, const ctx_conv_geom* g, int32_t nplanes);
int ctx_conv2d_few_pack(const float* w, int32_t C, int32_t Co, const ctx_conv_geom* g, void* wimg, int32_t nplanes,
                        void* stream);

/* y [N,P,Q,Co] fp32 = conv2d(SAME)(x, w) + bias (+ lrelu); y_planes [y_planes_n][N,P,Q,Co] fp16 or NULL; epilogue: bias and
 * lrelu only; amax_bits (or NULL): atomicMax of the bit patterns of |y|.  Also the data gradient of a transposed conv with
 * few OUTPUT channels (x = dY [N,H,W,Cn], the [kh,kw,Cn,Ck] filter read as HWIO, nprod = 1, x_amax_bits set). */
int ctx_conv2d_few_fprop(const void* x, int32_t x_is_u8, int32_t C, const void* wimg, float* y, void* y_planes,
                         int32_t y_planes_n, int32_t Co, const ctx_conv_geom* g, const ctx_epilogue* epi, int32_t nprod,
                         const uint32_t* x_amax_bits, uint32_t* amax_bits, void* stream);

/* dw [kh,kw,C,Cb1+Cb2] fp32 += sum over pixels of patches(x) (x) [small1 ‖ small2]: the filter gradient of the conv
 * (x = its input, small1 = one fp16 plane of dY scaled by *r_scale) and of the few-output-channel transposed conv (x = dY
 * with x_amax_bits, small1 / small2 = fp16 hi planes of [state ‖ skip], small2 read at image n % small2_batch).  small*
 * are [rows, Cb*] fp16, Cb1 and Cb2 multiples of 64, Cb1 + Cb2 = 64 or 128.  One product; the sums are divided by
 * *r_scale when given, else by the scale x_amax_bits implies. */
int ctx_conv2d_few_wgrad(const void* x, int32_t x_is_u8, int32_t C, const void* small1_plane, int32_t Cb1,
                         const void* small2_plane, int32_t Cb2, int32_t small2_batch, float* dw, const ctx_conv_geom* g,
                         const uint32_t* x_amax_bits, const float* r_scale, void* stream);

/* The reference's last decoder layer  d_h4 = conv2d_transpose([state ‖ skip], 5x5, stride 2, 128 -> 3)  (arm_shaping.py:1329):
 * channel contraction on the tensor cores (G[pixel][tap, co] = cat[pixel][:] . W[tap, co, :]), col2im as a gather from shared
 * memory — no atomics, y written once.  x1_planes [2][N,P,Q,C1], x2_planes [2][x2_batch,P,Q,C2] or NULL: fp16 (hi, lo)
 * planes, three products.  wimg from ctx_conv2d_transpose_few_pack of the [kh,kw,Cn,C1+C2] filter.  Epilogue: bias, lrelu.
 * With `loss` the reconstruction losses of arm_shaping.py:1352-1353 are evaluated in the same pass, as ctx_l2_loss_fwd_bwd
 * would on y: image n belongs to decoder pass n / pass_images, d = y (+ resid) - tgt (both indexed by the image within
 * its pass), loss[pass] += sum d^2 / 2, dpred = weight[pass] * d, *dpred_amax = max(.., bits(max |dpred|)). */
typedef struct ctx_l2_fused {
  const float* tgt;          /* [pass_images,H,W,Cn] */
  const float* resid;        /* [pass_images,H,W,Cn] or NULL (ContextAEInception2's residual output, arm_shaping.py:1889) */
  const float* weight;       /* [passes] or NULL */
  float* loss;               /* [passes], += */
  float* dpred;              /* [N,H,W,Cn] or NULL */
  uint32_t* dpred_amax;      /* or NULL */
  int32_t pass_images;
} ctx_l2_fused;
int ctx_conv2d_transpose_few_eligible(int32_t C1, int32_t C2, int32_t x2_batch, int32_t Cn, const ctx_conv_geom* g);
int64_t ctx_conv2d_transpose_few_pack_bytes(int32_t Ck, const ctx_conv_geom* g);
int ctx_conv2d_transpose_few_pack(const float* w, int32_t Ck, int32_t Cn, const ctx_conv_geom* g, void* wimg, void* stream);
int ctx_conv2d_transpose_few(const void* x1_planes, int32_t C1, const void* x2_planes, int32_t C2, int32_t x2_batch,
                             const void* wimg, float* y, int32_t Cn, const ctx_conv_geom* g, const ctx_epilogue* epi,
                             const ctx_l2_fused* loss, void* stream);

/* ---- round-1 names of the planes entry points: bf16 planes, three products, two result planes -------------------- */

/* planes[0][i] = bf16_rn(x[i]); planes[1][i] = bf16_rn(x[i] - planes[0][i]) for i < n. */
int ctx_split_bf16(const float* x, int64_t n, void* planes, void* stream);
int ctx_split_bf16_transpose(const float* w, int32_t T, int32_t R, int32_t Cc, void* planes, void* stream);
int ctx_conv2d_fprop_bf16x3_eligible(int32_t C, int32_t Co, const ctx_conv_geom* g);
int ctx_conv2d_fprop_bf16x3(const void* x_planes, int32_t C, const void* w_planes, float* y, void* y_planes, int32_t Co,
                            const ctx_conv_geom* g, const ctx_epilogue* epi, void* stream);
int ctx_conv2d_transpose_bf16x3_eligible(int32_t C1, int32_t C2, int32_t x2_batch, int32_t Cn, const ctx_conv_geom* g);
int ctx_conv2d_transpose_bf16x3(const void* x1_planes, int32_t C1, const void* x2_planes, int32_t C2, int32_t x2_batch,
                                const void* w_planes, float* y, void* y_planes, int32_t Cn, const ctx_conv_geom* g,
                                const ctx_epilogue* epi, void* stream);
int ctx_conv2d_transpose_merged_bf16x3_eligible(int32_t C1, int32_t C2, int32_t x2_batch, int32_t Cn,
                                                const ctx_conv_geom* g);
int ctx_conv2d_transpose_merged_pack(const float* w, int32_t Ck, int32_t Cn, const ctx_conv_geom* g, void* wm_planes,
                                     void* stream);
int ctx_conv2d_transpose_merged_bf16x3(const void* x1_planes, int32_t C1, const void* x2_planes, int32_t C2,
                                       int32_t x2_batch, const void* wm_planes, float* y, void* y_planes, int32_t Cn,
                                       const ctx_conv_geom* g, const ctx_epilogue* epi, void* stream);

/* ---- pointwise / reductions ------------------------------------------------------------------------------ */

/* y = (float(x) * (1/255) - 0.5) * 2 in fp32 — rllab/sampler/base.py:117-119, scripts/train_script.py:102-104. */
int ctx_u8_normalize(const uint8_t* x, float* y, int64_t n, void* stream);

/* out[b, :] = float(bank[frame_idx[b], :]) for b < n — the trainer's batch assembly traindata[arange(B) % nlen, choice]
 * (scripts/train_script.py:154-159, 169-173) on a demo bank resident in HBM: `bank` is [bank_frames, frame_elems] fp32
 * or uint8, frame_idx (device, int64) holds flat frame numbers t * N + video.  A uint8 bank is mapped through lut256
 * (device, 256 floats), which the host fills with the reference's own pixel transform (arm_shaping.py:12-16) so that the
 * result is bit-identical to gathering the float bank.  An index outside the bank yields NaN frames. */
int ctx_gather_frames(const void* bank, int32_t bank_is_u8, const float* lut256, int64_t bank_frames,
                      int64_t frame_elems, const int64_t* frame_idx, int32_t n, float* out, void* stream);

/* out (+)= dy * (y_ref >= 0 ? 1 : slope) — derivative of arm_shaping.py:18-19 from the saved output. */
int ctx_lrelu_bwd(const float* dy, const float* y_ref, float* out, int64_t n, float slope, int32_t accumulate,
                  void* stream);

/* db[C] += column sums of dy[M,C] (ld = ldy) — gradient of tf.nn.bias_add (arm_shaping.py:30,80). */
int ctx_bias_grad(const float* dy, int64_t M, int32_t C, int32_t ldy, float* db, void* stream);

/* Pixel reconstruction losses, arm_shaping.py:1352-1353 (tf.nn.l2_loss = sum(t^2)/2), and their gradient.
 * pred is [G*B, per_img] (G decoder passes stacked on the batch axis), tgt is [B, per_img] (image b % B),
 * resid optional [B, per_img] added to pred first (ContextAEInception2 residual output, :1890-1891).
 * loss[g] += 0.5 * sum_{b,i} (pred[g*B+b,i] + resid[b,i] - tgt[b,i])^2 ;  dpred = weight[g] * (pred+resid-tgt)
 * (weight[g] in {0,1} selects the terms of the ablated loss, ablations_code/ablations.py:175-182). */
int ctx_l2_loss_fwd_bwd(const float* pred, const float* tgt, const float* resid, int32_t G, int64_t B,
                        int64_t per_img, const float* weight, float* loss, float* dpred, void* stream);

/* Feature-alignment loss, arm_shaping.py:1345:  loss += scale * inv_count * sum (a-b)^2 ;
 * da (+)= grad_weight * 2*scale*inv_count*(a-b) ; db (+)= -that.  inv_count = 1/(global_batch*F) so that
 * data-parallel ranks sum to the global mean.  a,b,da,db are [rows, F] with leading dimensions. */
int ctx_mse_scaled_fwd_bwd(const float* a, int32_t lda, const float* b, int32_t ldb, int64_t rows, int32_t F,
                           float scale, float inv_count, float grad_weight, float* loss,
                           float* da, int32_t ldda, float* db, int32_t lddb, int32_t accumulate,
                           void* stream);

/* Split the data-gradient of a decoder level, dcat [G*B, pix, C1+C2], into
 *   dstate[G*B, pix, C1] = dcat[..., :C1] * lrelu'(state_ref)           (state_ref may be NULL: no mask)
 *   dskip [B,   pix, C2] (+)= (sum_g dcat[g*B+b, pix, C1:]) * lrelu'(skip_ref)
 * — the gradient of tf.concat([h, skip], 3) (arm_shaping.py:1323-1329) with the skip shared by G passes.
 * round_tf32: round both outputs to the nearest TF32 (they are the next tensor-core operands). */
int ctx_split_skip_grad(const float* dcat, int32_t G, int64_t B, int64_t pix, int32_t C1, int32_t C2,
                        const float* state_ref, const float* skip_ref, float slope,
                        float* dstate, float* dskip, int32_t accumulate_skip, int32_t round_tf32, void* stream);

/* out[rows, C] = a[rows, C] + b[rows % b_rows, C]   (residual output of ContextAEInception2). */
int ctx_add_bcast(const float* a, const float* b, float* out, int64_t rows, int64_t b_rows, int64_t C,
                  void* stream);

/* Patch matrix of a SAME conv over a FEW-channel NHWC tensor (the 3-channel image side of the first encoder conv and
 * the last decoder layer, arm_shaping.py:1283,1329), so those layers run as dense tensor-core GEMMs (K = kh*kw*C):
 *   out[(n,p,q), (r*kw+s)*C + c] = x[n, p*st+r-pad_h, q*st+s-pad_w, c];  0 outside the image and in columns
 *   kh*kw*C .. ld-1.   out is [N*P*Q, ld], ld >= kh*kw*C (ld % 4 == 0 for the tensor-core GEMM).
 * round_tf32: round the gathered values to the nearest TF32 (operand of a single-pass tensor-core GEMM).
 * (TMA cannot gather a 12-byte pixel; every layer with >= 32 channels stays im2col-free.) */
int ctx_im2col(const float* x, int32_t C, float* out, int32_t ld, const ctx_conv_geom* g, int32_t round_tf32,
               void* stream);

/* slim.max_pool2d (mode 0) / slim.avg_pool2d (mode 1) over an NHWC tensor — nets/inception_v3.py:119,134,158 ...
 * g gives the window (kh, kw, stride), the low padding (pad_h, pad_w; 0 for VALID) and the output grid (P, Q).
 * Average pooling divides by the number of IN-RANGE window elements (TensorFlow's SAME semantics). */
int ctx_pool2d(const float* x, float* y, int32_t C, const ctx_conv_geom* g, int32_t mode, void* stream);

/* dst[r, c_off + c] = src[r, c] for r < rows: tf.concat(axis=3) of NHWC tensors into a wider one
 * (nets/inception_v3.py:177 and every Mixed_* block). */
int ctx_copy_channels(const float* src, int32_t Cs, float* dst, int32_t Cd, int32_t c_off, int64_t rows, void* stream);

/* TF-style Adam on flat buffers (scripts/train_script.py:128; epsilon outside the bias correction):
 *   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ; p -= lr_t * m / (sqrt(v) + eps)
 * lr_t = lr*sqrt(1-b2^t)/(1-b1^t) is computed by the caller.  grad_scale multiplies g first (1.0 normally).
 * p_tf32 (optional, may be NULL): receives round_to_nearest_tf32(p) — the shadow copy the single-pass tensor-core
 * GEMMs read, so that the hardware's operand truncation cannot bias them. */
int ctx_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr_t, float b1, float b2,
                  float eps, float grad_scale, float* p_tf32, void* stream);

/* The same update with the step-dependent rate on the DEVICE, so that a whole training step can be replayed as a CUDA graph:
 * ctx_adam_schedule does  *step += 1;  *lr_t = lr * sqrt(1 - b2^step) / (1 - b1^step)  (float64 arithmetic, as the host
 * formula), ctx_adam_step_dev reads *lr_t. */
int ctx_adam_schedule(int64_t* step, float lr, float b1, float b2, float* lr_t, void* stream);
int ctx_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_t, float b1, float b2,
                      float eps, float grad_scale, float* p_tf32, void* stream);

/* y = round_to_nearest_tf32(x) (keeps fp32 storage) — operand preparation for the tcgen05 path. */
int ctx_round_tf32(const float* x, float* y, int64_t n, void* stream);

/* out[c, r] = round_tf32?(in[r, c]) for a [rows, cols] matrix, `batch` matrices back to back. */
int ctx_transpose_round(const float* in, float* out, int32_t batch, int32_t rows, int32_t cols,
                        int32_t round_tf32, void* stream);

/* nn_err — scripts/train_script.py:148 (integer, bit-exact):
 *   idx[b] = argmin_a mean_i (tgt[a,i] - out[b,i])^2  (first minimum);  err = sum_b |idx[b] - (b % nlen)|.
 * dist_ws is a caller-provided [B,B] fp32 workspace; idx is [B] int64; err one int64. */
int ctx_nn_err(const float* tgt, const float* out, int32_t B, int64_t per_img, int32_t nlen,
               float* dist_ws, int64_t* idx, int64_t* err, void* stream);

/* Feature-distance reward — rllab/sampler/base.py:243-257 — for `paths` trajectories of T frames each:
 *   cost[p,j] = sum_k (means[j,k]-z[p*T+j,k])^2 + scale * sum_i (mean_imgs[j,i]-x[p*T+j,i])^2
 *   rewards[p*rew_stride + 2j+1] -= (double)cost[p,j] * j*j       (rewards are float64, like path['rewards'])
 * feat_weight / img_weight in {0,1} implement ablation_type nofeat / noimage (base.py:246-249).
 * cost (optional, may be NULL) receives the fp32 costs [paths, T]. */
int ctx_reward_ours(const float* z, const float* x, const float* means, const float* mean_imgs,
                    int32_t paths, int32_t T, int32_t F, int64_t per_img, float scale,
                    float feat_weight, float img_weight, double* rewards, int64_t rew_stride, float* cost,
                    void* stream);

/* Inception-feature reward — rllab/sampler/base.py:184-189:
 *   d = means[j,i]-feat[p*T+j,i]; d = 0 where std == 0; m[p,j] = mean_i d^2/(std[j,i]+1e-5);
 *   rewards[p*rew_stride + 2j+1] -= (double)m[p,j] * j*j */
int ctx_reward_inception(const float* feat, const float* means, const float* std, int32_t paths, int32_t T,
                         int64_t per_frame, double* rewards, int64_t rew_stride, float* cost, void* stream);

/* The advantage / return tail of a TRPO batch (rllab/sampler/base.py:258-266) for `paths` trajectories laid end to end;
 * offsets[p] .. offsets[p + 1] are the steps of path p (int64, device).  All float64, as the reference:
 *   deltas     = rewards + discount * append(baselines, 0)[1:] - baselines
 *   advantages = discount_cumsum(deltas, discount * gae_lambda)      (rllab/misc/special.py:107-111)
 *   returns    = discount_cumsum(rewards, discount)
 * bit-identical to the numpy / scipy.signal.lfilter evaluation (one multiply and one add per step, no FMA). */
int ctx_gae_tail(const double* rewards, const double* baselines, const int64_t* offsets, int32_t paths, double discount,
                 double gae_lambda, double* advantages, double* returns, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTX_B200_H_ */
