// HBM-bound pointwise and reduction kernels of the hot path: image normalisation, lrelu', bias gradients,
// the three losses (+ their gradients), skip-gradient split, fused Adam, TF32 operand preparation, nn_err and
// the two reward distances.  All are grid-stride, 128-bit vectorised where the layout allows, with
// warp-shuffle + one atomic per block for reductions.
#include <cuda_bf16.h>
#include "common.cuh"

namespace ctx {

static inline int grid_for(int64_t work_items, int block, int per_sm = 8) {
  int sms = sm_count();
  if (sms <= 0) sms = 1;
  int64_t g = cdiv(work_items, block);
  int64_t cap = (int64_t)sms * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; result valid in thread 0.  blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ float block_sum(float v) {
  __shared__ float red[32];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (wid == 0) v = warp_sum(v);
  return v;
}

// ---- u8 -> [-1,1] ------------------------------------------------------------------------------------------
__global__ void u8_normalize_kernel(const uint8_t* __restrict__ x, float* __restrict__ y, int64_t n) {
  const float inv = 1.0f / 255.0f;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n / 4;
  if ((((uintptr_t)x) & 3) == 0 && (((uintptr_t)y) & 15) == 0) {
    for (int64_t j = i; j < n4; j += stride) {
      uchar4 u = reinterpret_cast<const uchar4*>(x)[j];
      float4 o;
      // three separate fp32 roundings, exactly as the TF graph (no FMA contraction)
      o.x = __fmul_rn(__fsub_rn(__fmul_rn((float)u.x, inv), 0.5f), 2.0f);
      o.y = __fmul_rn(__fsub_rn(__fmul_rn((float)u.y, inv), 0.5f), 2.0f);
      o.z = __fmul_rn(__fsub_rn(__fmul_rn((float)u.z, inv), 0.5f), 2.0f);
      o.w = __fmul_rn(__fsub_rn(__fmul_rn((float)u.w, inv), 0.5f), 2.0f);
      reinterpret_cast<float4*>(y)[j] = o;
    }
    for (int64_t j = n4 * 4 + i; j < n; j += stride)
      y[j] = __fmul_rn(__fsub_rn(__fmul_rn((float)x[j], inv), 0.5f), 2.0f);
  } else {
    for (int64_t j = i; j < n; j += stride)
      y[j] = __fmul_rn(__fsub_rn(__fmul_rn((float)x[j], inv), 0.5f), 2.0f);
  }
}

// ---- demo-bank gather: out[b] = float(bank[frame_idx[b]]) ----------------------------------------------------------------
// The trainer's batch assembly  traindata[arange(B) % nlen, choice]  (train_script.py:154-159) on a bank that lives in
// HBM.  blockIdx.y = batch row, blockIdx.x strides over the frame.  A uint8 bank is mapped through a 256-entry table the
// host fills with the reference's own arithmetic (arm_shaping.py:12-16: x / 127.5 - 1 in float64, cast to fp32 by the
// feed), so the result is bit-identical to feeding the float bank.  A frame index outside the bank writes NaN (loud).
template <bool kU8>
__global__ void gather_frames_kernel(const void* __restrict__ bank, const float* __restrict__ lut, int64_t bank_frames,
                                     int64_t frame_elems, const int64_t* __restrict__ frame_idx,
                                     float* __restrict__ out) {
  __shared__ float s_lut[256];
  if (kU8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = lut[i];
    __syncthreads();
  }
  const int64_t f = frame_idx[blockIdx.y];
  float* dst = out + (int64_t)blockIdx.y * frame_elems;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, step = (int64_t)gridDim.x * blockDim.x;
  if (f < 0 || f >= bank_frames) {
    for (int64_t j = t0; j < frame_elems; j += step) dst[j] = __int_as_float(0x7fc00000);
    return;
  }
  if (kU8) {
    const uint8_t* src = reinterpret_cast<const uint8_t*>(bank) + f * frame_elems;
    if ((((uintptr_t)src) & 3) == 0 && (((uintptr_t)dst) & 15) == 0 && frame_elems % 4 == 0) {
      for (int64_t j = t0; j < frame_elems / 4; j += step) {
        uchar4 u = reinterpret_cast<const uchar4*>(src)[j];
        reinterpret_cast<float4*>(dst)[j] = make_float4(s_lut[u.x], s_lut[u.y], s_lut[u.z], s_lut[u.w]);
      }
    } else {
      for (int64_t j = t0; j < frame_elems; j += step) dst[j] = s_lut[src[j]];
    }
  } else {
    const float* src = reinterpret_cast<const float*>(bank) + f * frame_elems;
    if ((((uintptr_t)src) & 15) == 0 && (((uintptr_t)dst) & 15) == 0 && frame_elems % 4 == 0) {
      for (int64_t j = t0; j < frame_elems / 4; j += step)
        reinterpret_cast<float4*>(dst)[j] = reinterpret_cast<const float4*>(src)[j];
    } else {
      for (int64_t j = t0; j < frame_elems; j += step) dst[j] = src[j];
    }
  }
}

// ---- lrelu backward ------------------------------------------------------------------------------------------
__global__ void lrelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ yref, float* out,
                                 int64_t n, float slope, int accumulate) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float v = dy[i] * (yref[i] >= 0.f ? 1.f : slope);
    if (accumulate) v += out[i];
    out[i] = v;
  }
}

// ---- bias gradient: column sums of dy[M, C] --------------------------------------------------------------------
// block = 32 x 8 threads: threadIdx.x walks 32 consecutive columns (coalesced), threadIdx.y strides rows.
__global__ void bias_grad_kernel(const float* __restrict__ dy, int64_t M, int C, int ldy, float* db,
                                 int64_t rows_per_block) {
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  float s = 0.f;
  if (c < C)
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) s += dy[r * ldy + c];
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += part[j][threadIdx.x];
    atomicAdd(db + c, t);
  }
}

// ---- pixel reconstruction losses + gradient ------------------------------------------------------------------------
__global__ void l2_loss_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                               const float* __restrict__ resid, int64_t per_pass, const float* __restrict__ weight,
                               float* loss, float* __restrict__ dpred) {
  // blockIdx.y = decoder pass g; elements of one pass are [B*per_img] and line up with tgt / resid
  const int g = blockIdx.y;
  const float wg = weight ? weight[g] : 1.f;
  const float* pg = pred + (int64_t)g * per_pass;
  float* dg = dpred ? dpred + (int64_t)g * per_pass : nullptr;
  float s = 0.f;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < per_pass; i += stride) {
    float d = pg[i] + (resid ? resid[i] : 0.f) - tgt[i];
    s += d * d;
    if (dg) dg[i] = wg * d;
  }
  s = block_sum(s);
  if (threadIdx.x == 0) atomicAdd(loss + g, 0.5f * s);
}

// ---- feature alignment loss + gradient --------------------------------------------------------------------------------
__global__ void mse_scaled_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                  int64_t rows, int F, float scale_inv, float gcoef, float* loss, float* da,
                                  int ldda, float* db, int lddb, int accumulate) {
  const int64_t n = rows * F;
  float s = 0.f;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const int64_t r = i / F;
    const int c = (int)(i - r * F);
    const float d = a[r * lda + c] - b[r * ldb + c];
    s += d * d;
    const float gv = gcoef * d;
    if (da) { float* q = da + r * ldda + c; *q = accumulate ? *q + gv : gv; }
    if (db) { float* q = db + r * lddb + c; *q = accumulate ? *q - gv : -gv; }
  }
  s = block_sum(s);
  if (threadIdx.x == 0 && loss) atomicAdd(loss, scale_inv * s);
}

// ---- decoder-level data-gradient split ------------------------------------------------------------------------------
__global__ void split_skip_grad_kernel(const float* __restrict__ dcat, int G, int64_t B, int64_t pix, int C1, int C2,
                                       const float* __restrict__ state_ref, const float* __restrict__ skip_ref,
                                       float slope, float* __restrict__ dstate, float* dskip, int accumulate_skip,
                                       int rnd) {
  const int Ct = C1 + C2;
  const int64_t n_state = (int64_t)G * B * pix * C1;
  const int64_t n_skip = B * pix * C2;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = i; j < n_state; j += stride) {
    const int64_t row = j / C1;
    const int c = (int)(j - row * C1);
    float v = dcat[row * Ct + c];
    if (state_ref) v *= (state_ref[j] >= 0.f ? 1.f : slope);
    dstate[j] = rnd ? round_tf32(v) : v;
  }
  for (int64_t j = i; j < n_skip; j += stride) {
    const int64_t row = j / C2;                         // row in [0, B*pix)
    const int c = (int)(j - row * C2);
    float v = 0.f;
    for (int g = 0; g < G; ++g) v += dcat[((int64_t)g * B * pix + row) * Ct + C1 + c];
    if (skip_ref) v *= (skip_ref[j] >= 0.f ? 1.f : slope);
    v = accumulate_skip ? dskip[j] + v : v;
    dskip[j] = rnd ? round_tf32(v) : v;
  }
}

// 128-bit version (C1 % 4 == 0, C2 % 4 == 0): one thread per float4 of a concatenated row; the G passes of a row are
// handled by the same thread, so every byte of dcat is read exactly once.
__global__ void split_skip_grad_vec_kernel(const float4* __restrict__ dcat, int G, int64_t rows, int C1q, int C2q,
                                           const float4* __restrict__ state_ref, const float4* __restrict__ skip_ref,
                                           float slope, float4* __restrict__ dstate, float4* dskip,
                                           int accumulate_skip, int rnd, FastDiv divCt) {
  const int Ctq = C1q + C2q;
  const int64_t total = rows * Ctq;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int64_t row;
    int c;
    if (total < (1ll << 31)) { int r32; divCt.divmod((int)i, r32, c); row = r32; }
    else { row = i / Ctq; c = (int)(i - row * Ctq); }
    if (c < C1q) {
      for (int g = 0; g < G; ++g) {
        const int64_t r = (int64_t)g * rows + row;
        float4 v = dcat[r * Ctq + c];
        if (state_ref) {
          const float4 a = state_ref[r * C1q + c];
          v.x *= (a.x >= 0.f ? 1.f : slope); v.y *= (a.y >= 0.f ? 1.f : slope);
          v.z *= (a.z >= 0.f ? 1.f : slope); v.w *= (a.w >= 0.f ? 1.f : slope);
        }
        if (rnd) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
        dstate[r * C1q + c] = v;
      }
    } else {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int g = 0; g < G; ++g) {
        const float4 t = dcat[((int64_t)g * rows + row) * Ctq + c];
        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
      }
      const int64_t j = row * C2q + (c - C1q);
      if (skip_ref) {
        const float4 a = skip_ref[j];
        v.x *= (a.x >= 0.f ? 1.f : slope); v.y *= (a.y >= 0.f ? 1.f : slope);
        v.z *= (a.z >= 0.f ? 1.f : slope); v.w *= (a.w >= 0.f ? 1.f : slope);
      }
      if (accumulate_skip) { const float4 o = dskip[j]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
      if (rnd) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
      dskip[j] = v;
    }
  }
}

__global__ void add_bcast_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                 int64_t n, int64_t nb) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = a[i] + b[i % nb];
}

// ---- y = (accumulate ? y : 0) + bias[c]: initialises the output of the scatter-epilogue transposed conv -----------
__global__ void bias_init_kernel(float* __restrict__ y, const float* __restrict__ bias, int64_t n, int C,
                                 int accumulate) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float v = accumulate ? y[i] : 0.f;
    if (bias) v += __ldg(bias + (int)(i % C));
    y[i] = v;
  }
}

int bias_init(float* y, const float* bias, int64_t n, int C, int accumulate, cudaStream_t stream) {
  if (accumulate && !bias) return CTX_OK;
  bias_init_kernel<<<grid_for(n, 256), 256, 0, stream>>>(y, bias, n, C, accumulate);
  CTX_LAUNCH_CHECK("bias_init");
  return CTX_OK;
}

// ---- patch matrix of a SAME conv over a few-channel NHWC tensor (C*kh*kw <= ld) ------------------------------------
// out[(n,p,q), (r*kw+s)*C + c] = x[n, p*st+r-pad_h, q*st+s-pad_w, c]  (0 outside the image and in columns >= kh*kw*C).
// One thread per output element, column index fastest: stores are fully coalesced, loads hit L1/L2 (every input
// element is read kh*kw/st^2 times from cache, once from HBM).
__global__ void im2col_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int ld, int C, int H,
                              int W, int kw, int K, int st, int padh, int padw, int rnd, FastDiv divLd, FastDiv divC,
                              FastDiv divKw, FastDiv divQ, FastDiv divP) {
  // Thread t owns patch column k = t (its tap (r, s) and channel c are decoded once); the block walks R rows, whose
  // pixel decode is the same for every thread.  Stores are coalesced along k; loads hit L1/L2.
  constexpr int R = 64;                                     // rows per block
  const int k = threadIdx.x;
  if (k >= ld) return;
  int tap, c, r, s;
  divC.divmod(k < K ? k : 0, tap, c);
  divKw.divmod(tap, r, s);
  const int row0 = blockIdx.x * R;
  const int nrow = min(R, rows - row0);
  float* o = out + (int64_t)row0 * ld + k;
  int t, q, n, pp;
  divQ.divmod(row0, t, q);
  divP.divmod(t, n, pp);
  const int Q = (int)divQ.d, P = (int)divP.d;
  for (int j = 0; j < nrow; ++j) {
    float v = 0.f;
    if (k < K) {
      const int h = pp * st + r - padh, w = q * st + s - padw;
      if ((unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W)
        v = __ldg(x + (((int64_t)n * H + h) * W + w) * C + c);
      if (rnd) v = round_tf32(v);
    }
    o[(int64_t)j * ld] = v;
    if (++q == Q) { q = 0; if (++pp == P) { pp = 0; ++n; } }
  }
}

// ---- pooling (slim max_pool2d / avg_pool2d, nets/inception_v3.py:119-413) -----------------------------------------
// mode 0: max over the in-range window elements; mode 1: average over the IN-RANGE elements (TF's SAME average pooling
// divides by the number of valid positions, not by kh*kw).  One thread per output element, channel fastest.
__global__ void pool2d_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t total, int C, int H, int W,
                              int P, int Q, int kh, int kw, int st, int padh, int padw, int mode) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int c = (int)(i % C);
    int64_t t = i / C;
    const int q = (int)(t % Q); t /= Q;
    const int pp = (int)(t % P);
    const int64_t n = t / P;
    const int h0 = pp * st - padh, w0 = q * st - padw;
    float acc = mode == 0 ? -3.402823466e38f : 0.f;
    int cnt = 0;
    for (int r = 0; r < kh; ++r) {
      const int h = h0 + r;
      if ((unsigned)h >= (unsigned)H) continue;
      for (int s = 0; s < kw; ++s) {
        const int w = w0 + s;
        if ((unsigned)w >= (unsigned)W) continue;
        const float v = __ldg(x + ((n * H + h) * W + w) * C + c);
        acc = mode == 0 ? fmaxf(acc, v) : acc + v;
        ++cnt;
      }
    }
    y[i] = mode == 0 ? acc : (cnt > 0 ? acc / (float)cnt : 0.f);
  }
}

// 128-bit version (C % 4 == 0): one thread per 4 channels of an output pixel; 32-bit index arithmetic
__global__ void pool2d_vec_kernel(const float4* __restrict__ x, float4* __restrict__ y, int total, int Cq, int H, int W,
                                  int kh, int kw, int st, int padh, int padw, int mode, FastDiv divC, FastDiv divQ,
                                  FastDiv divP) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int t, c, q, n, pp;
    divC.divmod(i, t, c);
    divQ.divmod(t, t, q);
    divP.divmod(t, n, pp);
    const int h0 = pp * st - padh, w0 = q * st - padw;
    float4 acc = mode == 0 ? make_float4(-3.402823466e38f, -3.402823466e38f, -3.402823466e38f, -3.402823466e38f)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
    int cnt = 0;
    for (int r = 0; r < kh; ++r) {
      const int h = h0 + r;
      if ((unsigned)h >= (unsigned)H) continue;
      for (int s = 0; s < kw; ++s) {
        const int w = w0 + s;
        if ((unsigned)w >= (unsigned)W) continue;
        const float4 v = __ldg(x + ((int64_t)(n * H + h) * W + w) * Cq + c);
        if (mode == 0) { acc.x = fmaxf(acc.x, v.x); acc.y = fmaxf(acc.y, v.y); acc.z = fmaxf(acc.z, v.z); acc.w = fmaxf(acc.w, v.w); }
        else { acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        ++cnt;
      }
    }
    if (mode == 1) {
      acc.x = cnt > 0 ? acc.x / (float)cnt : 0.f; acc.y = cnt > 0 ? acc.y / (float)cnt : 0.f;
      acc.z = cnt > 0 ? acc.z / (float)cnt : 0.f; acc.w = cnt > 0 ? acc.w / (float)cnt : 0.f;
    }
    y[i] = acc;
  }
}

// dst[r, c_off + c] = src[r, c]: channel concatenation of NHWC tensors (tf.concat(axis=3), inception_v3.py:177 ...)
__global__ void copy_channels_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t rows, int Cs,
                                     int Cd, int coff) {
  const int64_t total = rows * Cs;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int64_t r = i / Cs;
    const int c = (int)(i - r * Cs);
    dst[r * Cd + coff + c] = src[i];
  }
}

__global__ void copy_channels_vec_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int total, int Csq,
                                         int Cdq, int coffq, FastDiv divCs) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int r, c;
    divCs.divmod(i, r, c);
    dst[(int64_t)r * Cdq + coffq + c] = __ldg(src + i);
  }
}

// ---- fused TF-Adam ----------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr_t, float b1, float b2, float eps,
                            float gscale, float* __restrict__ p_tf32, const float* __restrict__ lr_t_dev) {
  if (lr_t_dev) lr_t = __ldg(lr_t_dev);                  // (CUDA-graph replays: the step-dependent rate lives on the device)
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n / 4;
  // TF's ApplyAdam functor, term for term (fp32 scalars):  m += (g - m)*(1 - b1);  v += (g*g - v)*(1 - b2);
  // var -= (m*alpha) / (sqrt(v) + eps)
  const float ob1 = 1.f - b1, ob2 = 1.f - b2;
  for (int64_t j = i; j < n4; j += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[j];
    float4 gg = reinterpret_cast<const float4*>(g)[j];
    float4 mm = reinterpret_cast<float4*>(m)[j];
    float4 vv = reinterpret_cast<float4*>(v)[j];
#define CTX_ADAM1(c)                                               \
    {                                                              \
      float gr = gg.c * gscale;                                    \
      mm.c = __fadd_rn(mm.c, __fmul_rn(__fsub_rn(gr, mm.c), ob1));                  \
      vv.c = __fadd_rn(vv.c, __fmul_rn(__fsub_rn(__fmul_rn(gr, gr), vv.c), ob2));   \
      pp.c = __fsub_rn(pp.c, __fdiv_rn(__fmul_rn(mm.c, lr_t), __fadd_rn(__fsqrt_rn(vv.c), eps)));  \
    }
    CTX_ADAM1(x) CTX_ADAM1(y) CTX_ADAM1(z) CTX_ADAM1(w)
#undef CTX_ADAM1
    reinterpret_cast<float4*>(p)[j] = pp;
    reinterpret_cast<float4*>(m)[j] = mm;
    reinterpret_cast<float4*>(v)[j] = vv;
    if (p_tf32)
      reinterpret_cast<float4*>(p_tf32)[j] =
          make_float4(round_tf32(pp.x), round_tf32(pp.y), round_tf32(pp.z), round_tf32(pp.w));
  }
  for (int64_t j = n4 * 4 + i; j < n; j += stride) {
    float gr = g[j] * gscale;
    float mm = __fadd_rn(m[j], __fmul_rn(__fsub_rn(gr, m[j]), ob1));
    float vv = __fadd_rn(v[j], __fmul_rn(__fsub_rn(__fmul_rn(gr, gr), v[j]), ob2));
    m[j] = mm; v[j] = vv;
    p[j] = __fsub_rn(p[j], __fdiv_rn(__fmul_rn(mm, lr_t), __fadd_rn(__fsqrt_rn(vv), eps)));
    if (p_tf32) p_tf32[j] = round_tf32(p[j]);
  }
}

// ---- TF32 operand preparation --------------------------------------------------------------------------------------
__global__ void round_tf32_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) y[i] = round_tf32(x[i]);
}

// out[b][c][r] = in[b][r][c]; 32x32 smem tile, block (32, 8)
__global__ void transpose_round_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols,
                                       int do_round) {
  __shared__ float tile[32][33];
  const int64_t boff = (int64_t)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? in[boff + (int64_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) {
      float v = tile[threadIdx.x][j];
      out[boff + (int64_t)c * rows + r] = do_round ? round_tf32(v) : v;
    }
  }
}

// ---- nn_err -------------------------------------------------------------------------------------------------------
// dist[a, b] = mean_i (tgt[a,i] - out[b,i])^2 ; one block per (a, b) pair, direct form (no ||a||^2-2ab trick,
// so rounding behaves like the reference's reduce_mean of squared differences).
__global__ void nn_dist_kernel(const float* __restrict__ tgt, const float* __restrict__ out, int B, int64_t per_img,
                               float* __restrict__ dist) {
  const int a = blockIdx.x, b = blockIdx.y;
  const float* ta = tgt + (int64_t)a * per_img;
  const float* ob = out + (int64_t)b * per_img;
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < per_img; i += blockDim.x) {
    float d = ta[i] - ob[i];
    s = fmaf(d, d, s);
  }
  s = block_sum(s);
  if (threadIdx.x == 0) dist[(int64_t)a * B + b] = s / (float)per_img;
}

// idx[b] = first argmin_a dist[a,b];  err = sum_b |idx[b] - b % nlen|   (single block)
__global__ void nn_argmin_kernel(const float* __restrict__ dist, int B, int nlen, int64_t* __restrict__ idx,
                                 int64_t* __restrict__ err) {
  __shared__ unsigned long long total;
  if (threadIdx.x == 0) total = 0ull;
  __syncthreads();
  unsigned long long local = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float best = dist[b];
    int bi = 0;
    for (int a = 1; a < B; ++a) {
      float d = dist[(int64_t)a * B + b];
      if (d < best) { best = d; bi = a; }
    }
    idx[b] = bi;
    int diff = bi - (b % nlen);
    local += (unsigned long long)(diff < 0 ? -diff : diff);
  }
  atomicAdd(&total, local);
  __syncthreads();
  if (threadIdx.x == 0) *err = (int64_t)total;
}

// ---- rewards ---------------------------------------------------------------------------------------------------------
// one block per (path, frame j)
__global__ void reward_ours_kernel(const float* __restrict__ z, const float* __restrict__ x,
                                   const float* __restrict__ means, const float* __restrict__ mean_imgs, int T, int F,
                                   int64_t per_img, float scale, float fw, float iw, double* rewards,
                                   int64_t rew_stride, float* cost) {
  const int p = blockIdx.y, j = blockIdx.x;
  const int64_t row = (int64_t)p * T + j;
  float sf = 0.f, si = 0.f;
  if (fw != 0.f) {
    const float* zz = z + row * F;
    const float* mm = means + (int64_t)j * F;
    for (int k = threadIdx.x; k < F; k += blockDim.x) { float d = mm[k] - zz[k]; sf = fmaf(d, d, sf); }
  }
  if (iw != 0.f) {
    const float* xx = x + row * per_img;
    const float* mi = mean_imgs + (int64_t)j * per_img;
    for (int64_t k = threadIdx.x; k < per_img; k += blockDim.x) { float d = mi[k] - xx[k]; si = fmaf(d, d, si); }
  }
  sf = block_sum(sf);
  si = block_sum(si);
  if (threadIdx.x == 0) {
    float c = fw * sf + iw * (scale * si);
    if (cost) cost[row] = c;
    if (rewards) rewards[(int64_t)p * rew_stride + 2 * j + 1] -= (double)c * (double)(j * j);
  }
}

__global__ void reward_inception_kernel(const float* __restrict__ feat, const float* __restrict__ means,
                                        const float* __restrict__ stdv, int T, int64_t per_frame, double* rewards,
                                        int64_t rew_stride, float* cost) {
  const int p = blockIdx.y, j = blockIdx.x;
  const int64_t row = (int64_t)p * T + j;
  const float* f = feat + row * per_frame;
  const float* mu = means + (int64_t)j * per_frame;
  const float* sd = stdv + (int64_t)j * per_frame;
  float s = 0.f;
  for (int64_t k = threadIdx.x; k < per_frame; k += blockDim.x) {
    float sdv = sd[k];
    float d = (sdv == 0.f) ? 0.f : (mu[k] - f[k]);
    s += d * d / (sdv + 1e-5f);
  }
  s = block_sum(s);
  if (threadIdx.x == 0) {
    float m = s / (float)per_frame;
    if (cost) cost[row] = m;
    if (rewards) rewards[(int64_t)p * rew_stride + 2 * j + 1] -= (double)m * (double)(j * j);
  }
}

}  // namespace ctx

// ==== C-ABI ============================================================================================================
using namespace ctx;

extern "C" int ctx_u8_normalize(const uint8_t* x, float* y, int64_t n, void* stream) {
  CTX_REQUIRE(n >= 0, "u8_normalize: negative size");
  if (n == 0) return CTX_OK;
  CTX_REQUIRE(x && y, "u8_normalize: null pointer");
  u8_normalize_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n);
  CTX_LAUNCH_CHECK("u8_normalize");
  return CTX_OK;
}

extern "C" int ctx_gather_frames(const void* bank, int32_t bank_is_u8, const float* lut256, int64_t bank_frames,
                                 int64_t frame_elems, const int64_t* frame_idx, int32_t n, float* out, void* stream) {
  CTX_REQUIRE(n >= 0 && frame_elems >= 0 && bank_frames >= 0, "gather_frames: negative size");
  if (n == 0 || frame_elems == 0) return CTX_OK;
  CTX_REQUIRE(bank && frame_idx && out, "gather_frames: null pointer");
  CTX_REQUIRE(!bank_is_u8 || lut256, "gather_frames: a uint8 bank needs the 256-entry value table");
  CTX_REQUIRE(n <= 65535, "gather_frames: at most 65535 frames per call");
  int64_t vec = (frame_elems + 3) / 4;
  int gx = (int)std::min<int64_t>(cdiv(vec, 256), 64);
  dim3 grid(gx, n);
  if (bank_is_u8)
    gather_frames_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(bank, lut256, bank_frames, frame_elems, frame_idx, out);
  else
    gather_frames_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(bank, lut256, bank_frames, frame_elems, frame_idx, out);
  CTX_LAUNCH_CHECK("gather_frames");
  return CTX_OK;
}

extern "C" int ctx_lrelu_bwd(const float* dy, const float* y_ref, float* out, int64_t n, float slope,
                             int32_t accumulate, void* stream) {
  CTX_REQUIRE(dy && y_ref && out && n >= 0, "lrelu_bwd: bad arguments");
  if (n == 0) return CTX_OK;
  lrelu_bwd_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, y_ref, out, n, slope, accumulate);
  CTX_LAUNCH_CHECK("lrelu_bwd");
  return CTX_OK;
}

extern "C" int ctx_bias_grad(const float* dy, int64_t M, int32_t C, int32_t ldy, float* db, void* stream) {
  CTX_REQUIRE(dy && db && M >= 0 && C > 0 && ldy >= C, "bias_grad: bad arguments");
  if (M == 0) return CTX_OK;
  int sms = sm_count();
  if (sms <= 0) return cuda_fail(cudaErrorNoDevice, "bias_grad");
  int gx = (int)cdiv(C, 32);
  int64_t want_y = cdiv((int64_t)sms * 4, gx);
  int64_t rows_per_block = cdiv(M, want_y);
  if (rows_per_block < 64) rows_per_block = 64;
  int gy = (int)cdiv(M, rows_per_block);
  bias_grad_kernel<<<dim3(gx, gy), dim3(32, 8), 0, (cudaStream_t)stream>>>(dy, M, C, ldy, db, rows_per_block);
  CTX_LAUNCH_CHECK("bias_grad");
  return CTX_OK;
}

extern "C" int ctx_l2_loss_fwd_bwd(const float* pred, const float* tgt, const float* resid, int32_t G, int64_t B,
                                   int64_t per_img, const float* weight, float* loss, float* dpred, void* stream) {
  CTX_REQUIRE(pred && tgt && loss && G > 0 && B >= 0 && per_img > 0, "l2_loss: bad arguments");
  if (B == 0) return CTX_OK;
  int64_t per_pass = B * per_img;
  int gx = grid_for(per_pass, 256, 4);
  l2_loss_kernel<<<dim3(gx, G), 256, 0, (cudaStream_t)stream>>>(pred, tgt, resid, per_pass, weight, loss, dpred);
  CTX_LAUNCH_CHECK("l2_loss");
  return CTX_OK;
}

extern "C" int ctx_mse_scaled_fwd_bwd(const float* a, int32_t lda, const float* b, int32_t ldb, int64_t rows,
                                      int32_t F, float scale, float inv_count, float grad_weight, float* loss,
                                      float* da, int32_t ldda, float* db, int32_t lddb, int32_t accumulate,
                                      void* stream) {
  CTX_REQUIRE(a && b && rows >= 0 && F > 0, "mse_scaled: bad arguments");
  if (rows == 0) return CTX_OK;
  mse_scaled_kernel<<<grid_for(rows * F, 256, 4), 256, 0, (cudaStream_t)stream>>>(
      a, lda, b, ldb, rows, F, scale * inv_count, grad_weight * 2.f * scale * inv_count, loss, da, ldda, db, lddb,
      accumulate);
  CTX_LAUNCH_CHECK("mse_scaled");
  return CTX_OK;
}

extern "C" int ctx_split_skip_grad(const float* dcat, int32_t G, int64_t B, int64_t pix, int32_t C1, int32_t C2,
                                   const float* state_ref, const float* skip_ref, float slope, float* dstate,
                                   float* dskip, int32_t accumulate_skip, int32_t round_tf32, void* stream) {
  CTX_REQUIRE(dcat && dstate && dskip && G > 0 && C1 > 0 && C2 > 0, "split_skip_grad: bad arguments");
  if (B == 0 || pix == 0) return CTX_OK;
  auto al = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
  if (C1 % 4 == 0 && C2 % 4 == 0 && al(dcat) && al(dstate) && al(dskip) && al(state_ref) && al(skip_ref)) {
    const int64_t rows = B * pix;
    const int Ctq = (C1 + C2) / 4;
    split_skip_grad_vec_kernel<<<grid_for(rows * Ctq, 256, 16), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(dcat), G, rows, C1 / 4, C2 / 4, reinterpret_cast<const float4*>(state_ref),
        reinterpret_cast<const float4*>(skip_ref), slope, reinterpret_cast<float4*>(dstate),
        reinterpret_cast<float4*>(dskip), accumulate_skip, round_tf32, FastDiv(Ctq));
  } else {
    int64_t n = (int64_t)G * B * pix * C1;
    split_skip_grad_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        dcat, G, B, pix, C1, C2, state_ref, skip_ref, slope, dstate, dskip, accumulate_skip, round_tf32);
  }
  CTX_LAUNCH_CHECK("split_skip_grad");
  return CTX_OK;
}

extern "C" int ctx_add_bcast(const float* a, const float* b, float* out, int64_t rows, int64_t b_rows, int64_t C,
                             void* stream) {
  CTX_REQUIRE(a && b && out && rows >= 0 && b_rows > 0 && C > 0, "add_bcast: bad arguments");
  if (rows == 0) return CTX_OK;
  add_bcast_kernel<<<grid_for(rows * C, 256), 256, 0, (cudaStream_t)stream>>>(a, b, out, rows * C, b_rows * C);
  CTX_LAUNCH_CHECK("add_bcast");
  return CTX_OK;
}

extern "C" int ctx_im2col(const float* x, int32_t C, float* out, int32_t ld, const ctx_conv_geom* g,
                          int32_t round_tf32, void* stream) {
  CTX_REQUIRE(g != nullptr && x && out, "im2col: null pointer");
  CTX_REQUIRE(C > 0 && g->kh > 0 && g->kw > 0 && g->stride > 0 && ld >= g->kh * g->kw * C,
              "im2col: ld (%d) must be >= kh*kw*C (%d)", ld, g->kh * g->kw * C);
  const int64_t rows = (int64_t)g->N * g->P * g->Q;
  if (rows == 0) return CTX_OK;
  CTX_REQUIRE(rows < (1ll << 31) - 64 && ld <= 4096, "im2col: too many rows (%lld) or ld too large", (long long)rows);
  CTX_REQUIRE(ld <= 1024, "im2col: ld (%d) must be <= 1024", ld);
  im2col_kernel<<<(unsigned)cdiv(rows, 64), ((ld + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(
      x, out, (int)rows, ld, C, g->H, g->W, g->kw, g->kh * g->kw * C, g->stride, g->pad_h, g->pad_w, round_tf32,
      FastDiv(ld),
      FastDiv(C), FastDiv(g->kw), FastDiv(g->Q), FastDiv(g->P));
  CTX_LAUNCH_CHECK("im2col");
  return CTX_OK;
}

extern "C" int ctx_pool2d(const float* x, float* y, int32_t C, const ctx_conv_geom* g, int32_t mode, void* stream) {
  CTX_REQUIRE(x && y && g && C > 0 && (mode == 0 || mode == 1), "pool2d: bad arguments");
  CTX_REQUIRE(g->kh > 0 && g->kw > 0 && g->stride > 0 && g->P > 0 && g->Q > 0 && g->H > 0 && g->W > 0 &&
                  g->pad_h >= 0 && g->pad_w >= 0, "pool2d: bad geometry");
  const int64_t total = (int64_t)g->N * g->P * g->Q * C;
  if (total == 0) return CTX_OK;
  if (C % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && total / 4 < (1ll << 31)) {
    const int tq = (int)(total / 4);
    pool2d_vec_kernel<<<grid_for(tq, 256, 16), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), tq, C / 4, g->H, g->W, g->kh, g->kw, g->stride,
        g->pad_h, g->pad_w, mode, FastDiv(C / 4), FastDiv(g->Q), FastDiv(g->P));
  } else
    pool2d_kernel<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(x, y, total, C, g->H, g->W, g->P, g->Q, g->kh,
                                                                              g->kw, g->stride, g->pad_h, g->pad_w, mode);
  CTX_LAUNCH_CHECK("pool2d");
  return CTX_OK;
}

extern "C" int ctx_copy_channels(const float* src, int32_t Cs, float* dst, int32_t Cd, int32_t c_off, int64_t rows,
                                 void* stream) {
  CTX_REQUIRE(src && dst && Cs > 0 && Cd > 0 && c_off >= 0 && c_off + Cs <= Cd && rows >= 0,
              "copy_channels: bad arguments (Cs=%d Cd=%d c_off=%d)", Cs, Cd, c_off);
  if (rows == 0) return CTX_OK;
  if (Cs % 4 == 0 && Cd % 4 == 0 && c_off % 4 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0 &&
      rows * (Cs / 4) < (1ll << 31)) {
    const int tq = (int)(rows * (Cs / 4));
    copy_channels_vec_kernel<<<grid_for(tq, 256, 16), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst), tq, Cs / 4, Cd / 4, c_off / 4, FastDiv(Cs / 4));
  } else
    copy_channels_kernel<<<grid_for(rows * Cs, 256, 16), 256, 0, (cudaStream_t)stream>>>(src, dst, rows, Cs, Cd, c_off);
  CTX_LAUNCH_CHECK("copy_channels");
  return CTX_OK;
}

extern "C" int ctx_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr_t, float b1,
                             float b2, float eps, float grad_scale, float* p_tf32, void* stream) {
  CTX_REQUIRE(p && g && m && v && n >= 0, "adam_step: bad arguments");
  CTX_REQUIRE(((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                  ((uintptr_t)v % 16 == 0) && ((uintptr_t)p_tf32 % 16 == 0),
              "adam_step: buffers must be 16-byte aligned");
  if (n == 0) return CTX_OK;
  adam_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr_t, b1, b2, eps,
                                                                            grad_scale, p_tf32, nullptr);
  CTX_LAUNCH_CHECK("adam_step");
  return CTX_OK;
}

__global__ void adam_schedule_kernel(int64_t* step, float lr, float b1, float b2, float* lr_t) {
  const int64_t t = *step + 1;
  *step = t;
  *lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, (double)t)) / (1.0 - pow((double)b1, (double)t)));
}

extern "C" int ctx_adam_schedule(int64_t* step, float lr, float b1, float b2, float* lr_t, void* stream) {
  CTX_REQUIRE(step && lr_t, "adam_schedule: null pointer");
  adam_schedule_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step, lr, b1, b2, lr_t);
  CTX_LAUNCH_CHECK("adam_schedule");
  return CTX_OK;
}

extern "C" int ctx_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_t, float b1,
                                 float b2, float eps, float grad_scale, float* p_tf32, void* stream) {
  CTX_REQUIRE(p && g && m && v && lr_t && n >= 0, "adam_step_dev: bad arguments");
  CTX_REQUIRE(((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                  ((uintptr_t)v % 16 == 0) && ((uintptr_t)p_tf32 % 16 == 0),
              "adam_step_dev: buffers must be 16-byte aligned");
  if (n == 0) return CTX_OK;
  adam_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, 0.f, b1, b2, eps, grad_scale, p_tf32,
                                                                            lr_t);
  CTX_LAUNCH_CHECK("adam_step_dev");
  return CTX_OK;
}

extern "C" int ctx_round_tf32(const float* x, float* y, int64_t n, void* stream) {
  CTX_REQUIRE(x && y && n >= 0, "round_tf32: bad arguments");
  if (n == 0) return CTX_OK;
  round_tf32_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n);
  CTX_LAUNCH_CHECK("round_tf32");
  return CTX_OK;
}

extern "C" int ctx_transpose_round(const float* in, float* out, int32_t batch, int32_t rows, int32_t cols,
                                   int32_t do_round, void* stream) {
  CTX_REQUIRE(in && out && batch > 0 && rows > 0 && cols > 0, "transpose_round: bad arguments");
  CTX_REQUIRE(batch <= 65535 && cdiv(rows, 32) <= 65535, "transpose_round: grid too large");
  dim3 grid((unsigned)cdiv(cols, 32), (unsigned)cdiv(rows, 32), (unsigned)batch);
  transpose_round_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(in, out, rows, cols, do_round);
  CTX_LAUNCH_CHECK("transpose_round");
  return CTX_OK;
}

extern "C" int ctx_nn_err(const float* tgt, const float* out, int32_t B, int64_t per_img, int32_t nlen,
                          float* dist_ws, int64_t* idx, int64_t* err, void* stream) {
  CTX_REQUIRE(tgt && out && dist_ws && idx && err && B > 0 && per_img > 0 && nlen > 0, "nn_err: bad arguments");
  CTX_REQUIRE(B <= 65535, "nn_err: batch too large");
  nn_dist_kernel<<<dim3(B, B), 128, 0, (cudaStream_t)stream>>>(tgt, out, B, per_img, dist_ws);
  CTX_LAUNCH_CHECK("nn_dist");
  nn_argmin_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(dist_ws, B, nlen, idx, err);
  CTX_LAUNCH_CHECK("nn_argmin");
  return CTX_OK;
}

extern "C" int ctx_reward_ours(const float* z, const float* x, const float* means, const float* mean_imgs,
                               int32_t paths, int32_t T, int32_t F, int64_t per_img, float scale,
                               float feat_weight, float img_weight, double* rewards, int64_t rew_stride,
                               float* cost, void* stream) {
  CTX_REQUIRE(paths >= 0 && T > 0 && F > 0 && per_img > 0, "reward_ours: bad arguments");
  CTX_REQUIRE((feat_weight == 0.f || (z && means)) && (img_weight == 0.f || (x && mean_imgs)),
              "reward_ours: null input for an enabled term");
  CTX_REQUIRE(rewards || cost, "reward_ours: no output requested");
  CTX_REQUIRE(!rewards || rew_stride >= 2 * (int64_t)T, "reward_ours: rew_stride must cover indices 2j+1");
  CTX_REQUIRE(paths <= 65535, "reward_ours: too many paths per call");
  if (paths == 0) return CTX_OK;
  reward_ours_kernel<<<dim3(T, paths), 256, 0, (cudaStream_t)stream>>>(
      z, x, means, mean_imgs, T, F, per_img, scale, feat_weight, img_weight, rewards, rew_stride, cost);
  CTX_LAUNCH_CHECK("reward_ours");
  return CTX_OK;
}

extern "C" int ctx_reward_inception(const float* feat, const float* means, const float* std, int32_t paths,
                                    int32_t T, int64_t per_frame, double* rewards, int64_t rew_stride, float* cost,
                                    void* stream) {
  CTX_REQUIRE(feat && means && std && paths >= 0 && T > 0 && per_frame > 0, "reward_inception: bad arguments");
  CTX_REQUIRE(rewards || cost, "reward_inception: no output requested");
  CTX_REQUIRE(!rewards || rew_stride >= 2 * (int64_t)T, "reward_inception: rew_stride must cover indices 2j+1");
  CTX_REQUIRE(paths <= 65535, "reward_inception: too many paths per call");
  if (paths == 0) return CTX_OK;
  reward_inception_kernel<<<dim3(T, paths), 256, 0, (cudaStream_t)stream>>>(feat, means, std, T, per_frame, rewards,
                                                                            rew_stride, cost);
  CTX_LAUNCH_CHECK("reward_inception");
  return CTX_OK;
}

// ---- advantage / return tail of a TRPO batch (rllab/sampler/base.py:258-266) ------------------------------------------------
// One thread per path walks its steps backwards: y[t] = x[t] + d * y[t + 1], one fp64 multiply and one fp64 add per step in
// the order scipy.signal.lfilter([1], [1, -d], x[::-1])[::-1] (rllab/misc/special.py:107-111) performs them — no FMA
// contraction — so the result is bit-identical to the reference's.
namespace ctx {
__global__ void gae_tail_kernel(const double* __restrict__ rewards, const double* __restrict__ baselines,
                                const int64_t* __restrict__ offsets, int paths, double discount, double gae_lambda,
                                double* __restrict__ advantages, double* __restrict__ returns) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= paths) return;
  const int64_t t0 = offsets[p], t1 = offsets[p + 1];
  const double dl = __dmul_rn(discount, gae_lambda);
  double adv = 0.0, ret = 0.0, b_next = 0.0;                    // np.append(path_baselines, 0): the baseline after the last step
  for (int64_t t = t1 - 1; t >= t0; --t) {
    const double r = rewards[t], b = baselines[t];
    const double delta = __dsub_rn(__dadd_rn(r, __dmul_rn(discount, b_next)), b);
    adv = (t == t1 - 1) ? delta : __dadd_rn(delta, __dmul_rn(dl, adv));
    ret = (t == t1 - 1) ? r : __dadd_rn(r, __dmul_rn(discount, ret));
    advantages[t] = adv;
    returns[t] = ret;
    b_next = b;
  }
}
}  // namespace ctx

extern "C" int ctx_gae_tail(const double* rewards, const double* baselines, const int64_t* offsets, int32_t paths,
                            double discount, double gae_lambda, double* advantages, double* returns, void* stream) {
  CTX_REQUIRE(paths >= 0, "gae_tail: negative path count");
  if (paths == 0) return CTX_OK;
  CTX_REQUIRE(rewards && baselines && offsets && advantages && returns, "gae_tail: null pointer");
  ctx::gae_tail_kernel<<<(paths + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rewards, baselines, offsets, paths, discount,
                                                                             gae_lambda, advantages, returns);
  CTX_LAUNCH_CHECK("gae_tail");
  return CTX_OK;
}
