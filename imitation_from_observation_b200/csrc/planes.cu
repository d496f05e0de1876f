// HBM-bound producers of the 16-bit operand planes the tcgen05 kernels of tc_h16.cu read: plane 0 = hi = rn16(s * x),
// plane 1 = lo = rn16(s * x - hi)  (fp16 or bf16; fp16 conversions saturate instead of overflowing).
//   ctx_split_planes / ctx_split_planes_transpose   activations made by other kernels, filters after every Adam update
//   ctx_grad_planes                                  raw fp32 gradient -> lrelu' mask -> scaled planes + bias gradient
//   ctx_amax, ctx_skip_grad_sum                      helpers of the backward chain
// Each kernel moves every byte once, 128 bits per access.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace ctx {

static inline int planes_grid(int64_t work_items, int block, int per_sm = 8) {
  int64_t blocks = (work_items + block - 1) / block;
  int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

__device__ __forceinline__ uint32_t p_pack16(float a, float b, bool f16) {
  if (f16) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
  }
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
}
__device__ __forceinline__ float p_lo16(uint32_t w, bool f16) {
  return f16 ? __half2float(__ushort_as_half((unsigned short)(w & 0xFFFFu))) : __uint_as_float(w << 16);
}
__device__ __forceinline__ float p_hi16(uint32_t w, bool f16) {
  return f16 ? __half2float(__ushort_as_half((unsigned short)(w >> 16))) : __uint_as_float(w & 0xFFFF0000u);
}
__device__ __forceinline__ uint16_t p_one16(float a, bool f16) { return (uint16_t)(p_pack16(a, 0.f, f16) & 0xFFFFu); }

// (hi, lo) words of four consecutive values
__device__ __forceinline__ void p_split4(const float4 v, bool f16, uint2& hi, uint2& lo) {
  hi.x = p_pack16(v.x, v.y, f16);
  hi.y = p_pack16(v.z, v.w, f16);
  lo.x = p_pack16(v.x - p_lo16(hi.x, f16), v.y - p_hi16(hi.x, f16), f16);
  lo.y = p_pack16(v.z - p_lo16(hi.y, f16), v.w - p_hi16(hi.y, f16), f16);
}

// ---- x -> planes ---------------------------------------------------------------------------------------------------------------
__global__ void split_planes_kernel(const float* __restrict__ x, int64_t n, uint16_t* __restrict__ planes, int nplanes, int f16,
                                    float scale) {
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, step = (int64_t)gridDim.x * blockDim.x;
  uint16_t* hi = planes;
  uint16_t* lo = planes + n;
  const bool h = f16 != 0;
  if (n % 4 == 0 && (((uintptr_t)x) & 15) == 0 && (((uintptr_t)planes) & 7) == 0) {
    for (int64_t j = i0; j < n / 4; j += step) {
      uint2 a, b;
      float4 v = reinterpret_cast<const float4*>(x)[j];
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      p_split4(v, h, a, b);
      reinterpret_cast<uint2*>(hi)[j] = a;
      if (nplanes == 2) reinterpret_cast<uint2*>(lo)[j] = b;
    }
  } else {
    for (int64_t j = i0; j < n; j += step) {
      const float v = x[j] * scale;
      const uint16_t a = p_one16(v, h);
      hi[j] = a;
      if (nplanes == 2) lo[j] = p_one16(v - p_lo16(a, h), h);
    }
  }
}

// w [T][R][Cc] fp32 -> planes [nplanes][T][Cc][R]; 32 x 32 tiles through shared memory, both global accesses coalesced
__global__ void split_planes_transpose_kernel(const float* __restrict__ w, int T, int R, int Cc, uint16_t* __restrict__ planes,
                                              int nplanes, int f16, float scale) {
  __shared__ float tile[32][33];
  const int t = blockIdx.z, r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const float* src = w + (int64_t)t * R * Cc;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < Cc) ? src[(int64_t)r * Cc + c] * scale : 0.f;
  }
  __syncthreads();
  const int64_t plane = (int64_t)T * R * Cc;
  uint16_t* hi = planes + (int64_t)t * R * Cc;
  const bool h = f16 != 0;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < Cc && r < R) {
      const float v = tile[threadIdx.x][i];
      const uint16_t a = p_one16(v, h);
      hi[(int64_t)c * R + r] = a;
      if (nplanes == 2) hi[plane + (int64_t)c * R + r] = p_one16(v - p_lo16(a, h), h);
    }
  }
}

// ---- gradient -> planes ----------------------------------------------------------------------------------------------------------
// s = 2^k with amax * s in [2^14, 2^15): fp16's largest finite value is 65504, its smallest normal 2^-14, so every element
// within 2^-28 of the tensor's maximum keeps the full 11-bit significand.
__device__ __forceinline__ float p16_scale_for(uint32_t amax_bits) {
  const int e = (int)((amax_bits >> 23) & 0xFFu);        // biased exponent of amax: amax in [2^(e-127), 2^(e-126))
  if (amax_bits == 0u || e == 0 || e == 255) return 1.f;
  int k = 14 - (e - 127);                                 // amax * 2^k in [2^14, 2^15)
  k = k > 126 ? 126 : (k < -126 ? -126 : k);
  return __uint_as_float((uint32_t)(k + 127) << 23);
}

// block (64, 4): threadIdx.x walks 64 column quads (256 columns, coalesced 128-bit accesses), threadIdx.y strides the rows
// of the block's slab; column sums are reduced over threadIdx.y in shared memory, one atomic per column and block.  A tensor
// with fewer than 256 channels puts `rp` = 64 / (C / 4) consecutive rows side by side on the 64 x-lanes (C = 64: four rows
// per pass), so every lane moves bytes whatever the channel count.
__global__ void grad_planes_kernel(const float* __restrict__ x, int64_t rows, int C, int ldx, const float* __restrict__ ref,
                                   float slope, const uint32_t* __restrict__ amax_bits, float* scale_out,
                                   uint16_t* __restrict__ planes, int nplanes, int f16, float* db, float* __restrict__ out_f32,
                                   int64_t rows_per_block) {
  __shared__ float4 part[4][64];
  const bool h = f16 != 0;
  float s = 1.f;
  if (h && amax_bits) s = p16_scale_for(*amax_bits);
  if (scale_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0) *scale_out = s;
  const int Cq = C >> 2;
  const int rp = Cq < 64 ? 64 / Cq : 1;                  // rows side by side on the x-lanes
  const int sub = rp > 1 ? (int)threadIdx.x / Cq : 0;
  const int cq = rp > 1 ? (int)threadIdx.x - sub * Cq : (int)(blockIdx.x * 64 + threadIdx.x);   // column quad
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t plane = rows * (int64_t)C;
  if (cq < Cq && sub < rp) {
    for (int64_t r = r0 + (int64_t)threadIdx.y * rp + sub; r < r1; r += 4 * rp) {
      float4 v = *reinterpret_cast<const float4*>(x + r * ldx + 4 * cq);
      const int64_t o = r * C + 4 * cq;
      if (ref) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(ref + o));
        v.x *= (a.x >= 0.f ? 1.f : slope); v.y *= (a.y >= 0.f ? 1.f : slope);
        v.z *= (a.z >= 0.f ? 1.f : slope); v.w *= (a.w >= 0.f ? 1.f : slope);
      }
      sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = v;
      if (planes) {
        uint2 a, b;
        p_split4(make_float4(v.x * s, v.y * s, v.z * s, v.w * s), h, a, b);
        *reinterpret_cast<uint2*>(planes + o) = a;
        if (nplanes == 2) *reinterpret_cast<uint2*>(planes + plane + o) = b;
      }
    }
  }
  if (db) {
    part[threadIdx.y][threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.y == 0 && cq < Cq && sub == 0) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < 4; ++j)
        for (int q = 0; q < rp; ++q) {
          const float4 u = part[j][threadIdx.x + q * Cq];
          t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
      atomicAdd(db + 4 * cq, t.x); atomicAdd(db + 4 * cq + 1, t.y);
      atomicAdd(db + 4 * cq + 2, t.z); atomicAdd(db + 4 * cq + 3, t.w);
    }
  }
}

__global__ void amax_kernel(const float* __restrict__ x, int64_t n, uint32_t* amax_bits) {
  uint32_t m = 0u;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, step = (int64_t)gridDim.x * blockDim.x;
  if (n % 4 == 0 && (((uintptr_t)x) & 15) == 0) {
    for (int64_t j = i0; j < n / 4; j += step) {
      const float4 v = reinterpret_cast<const float4*>(x)[j];
      m = max(max(m, __float_as_uint(v.x) & 0x7FFFFFFFu), __float_as_uint(v.y) & 0x7FFFFFFFu);
      m = max(max(m, __float_as_uint(v.z) & 0x7FFFFFFFu), __float_as_uint(v.w) & 0x7FFFFFFFu);
    }
  } else {
    for (int64_t j = i0; j < n; j += step) m = max(m, __float_as_uint(x[j]) & 0x7FFFFFFFu);
  }
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m) atomicMax(amax_bits, m);
}

// one thread per float4 of a skip row; the G passes of a row are summed by the same thread
__global__ void skip_grad_sum_kernel(const float4* __restrict__ dcat, int G, int64_t rows, int C1q, int C2q,
                                     const float4* __restrict__ skip_ref, float slope, float4* dskip, int accumulate,
                                     FastDiv divC2) {
  const int Ctq = C1q + C2q;
  const int64_t total = rows * C2q;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int64_t row;
    int c;
    if (total < (1ll << 31)) { int r32; divC2.divmod((int)i, r32, c); row = r32; }
    else { row = i / C2q; c = (int)(i - row * C2q); }
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int g = 0; g < G; ++g) {
      const float4 t = dcat[((int64_t)g * rows + row) * Ctq + C1q + c];
      v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    if (skip_ref) {
      const float4 a = skip_ref[i];
      v.x *= (a.x >= 0.f ? 1.f : slope); v.y *= (a.y >= 0.f ? 1.f : slope);
      v.z *= (a.z >= 0.f ? 1.f : slope); v.w *= (a.w >= 0.f ? 1.f : slope);
    }
    if (accumulate) { const float4 o = dskip[i]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
    dskip[i] = v;
  }
}

}  // namespace ctx

using namespace ctx;

static int dtype_ok(int32_t dtype) { return dtype == CTX_P16_F16 || dtype == CTX_P16_BF16; }

extern "C" int ctx_split_planes(const float* x, int64_t n, float scale, void* planes, int32_t nplanes, int32_t dtype,
                                void* stream) {
  CTX_REQUIRE(n >= 0, "split_planes: negative size");
  CTX_REQUIRE((nplanes == 1 || nplanes == 2) && dtype_ok(dtype), "split_planes: nplanes must be 1 or 2, dtype F16 or BF16");
  if (n == 0) return CTX_OK;
  CTX_REQUIRE(x && planes, "split_planes: null pointer");
  split_planes_kernel<<<planes_grid(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(x, n, (uint16_t*)planes, nplanes,
                                                                                   dtype == CTX_P16_F16, scale);
  CTX_LAUNCH_CHECK("split_planes");
  return CTX_OK;
}

extern "C" int ctx_split_planes_transpose(const float* w, int32_t T, int32_t R, int32_t Cc, float scale, void* planes,
                                          int32_t nplanes, int32_t dtype, void* stream) {
  CTX_REQUIRE(T >= 0 && R >= 0 && Cc >= 0, "split_planes_transpose: negative size");
  CTX_REQUIRE((nplanes == 1 || nplanes == 2) && dtype_ok(dtype), "split_planes_transpose: nplanes must be 1 or 2, dtype F16 or BF16");
  if (T == 0 || R == 0 || Cc == 0) return CTX_OK;
  CTX_REQUIRE(w && planes, "split_planes_transpose: null pointer");
  CTX_REQUIRE(T <= 65535 && cdiv(R, 32) <= 65535, "split_planes_transpose: grid too large");
  dim3 grid((unsigned)cdiv(Cc, 32), (unsigned)cdiv(R, 32), (unsigned)T);
  split_planes_transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(w, T, R, Cc, (uint16_t*)planes, nplanes,
                                                                               dtype == CTX_P16_F16, scale);
  CTX_LAUNCH_CHECK("split_planes_transpose");
  return CTX_OK;
}

extern "C" int ctx_split_bf16(const float* x, int64_t n, void* planes, void* stream) {
  return ctx_split_planes(x, n, 1.f, planes, 2, CTX_P16_BF16, stream);
}

extern "C" int ctx_split_bf16_transpose(const float* w, int32_t T, int32_t R, int32_t Cc, void* planes, void* stream) {
  return ctx_split_planes_transpose(w, T, R, Cc, 1.f, planes, 2, CTX_P16_BF16, stream);
}

extern "C" int ctx_grad_planes(const float* x, int64_t rows, int32_t C, int32_t ldx, const float* ref, float slope,
                               const uint32_t* amax_bits, float* scale, void* planes, int32_t nplanes, int32_t dtype,
                               float* db, float* out_f32, void* stream) {
  CTX_REQUIRE(rows >= 0 && C > 0 && ldx >= C, "grad_planes: bad shape");
  CTX_REQUIRE(dtype_ok(dtype) && (!planes || nplanes == 1 || nplanes == 2), "grad_planes: nplanes must be 1 or 2, dtype F16 or BF16");
  CTX_REQUIRE(C % 4 == 0 && ldx % 4 == 0, "grad_planes: C and ldx must be multiples of 4");
  auto al = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
  CTX_REQUIRE(x && al(x) && al(ref) && al(out_f32) && (((uintptr_t)planes) & 7) == 0, "grad_planes: pointers must be 16-byte aligned");
  if (rows == 0) return CTX_OK;
  const int gx = (int)cdiv(C / 4, 64);
  // enough row slabs to fill the machine a few times over, at least 16 rows each (the block reduces over 4 row lanes)
  int64_t gy = cdiv((int64_t)(sm_count() > 0 ? sm_count() : 148) * 8, gx);
  const int rp = C / 4 < 64 ? 64 / (C / 4) : 1;        // rows a block covers per pass of its 64 x-lanes (see the kernel)
  if (gy > cdiv(rows, 16 * rp)) gy = cdiv(rows, 16 * rp);
  if (gy < 1) gy = 1;
  if (gy > 65535) gy = 65535;
  const int64_t rpb = cdiv(rows, gy);
  gy = cdiv(rows, rpb);
  grad_planes_kernel<<<dim3((unsigned)gx, (unsigned)gy), dim3(64, 4), 0, (cudaStream_t)stream>>>(
      x, rows, C, ldx, ref, slope, amax_bits, scale, (uint16_t*)planes, nplanes, dtype == CTX_P16_F16, db, out_f32, rpb);
  CTX_LAUNCH_CHECK("grad_planes");
  return CTX_OK;
}

extern "C" int ctx_amax(const float* x, int64_t n, uint32_t* amax_bits, void* stream) {
  CTX_REQUIRE(n >= 0 && amax_bits, "amax: bad arguments");
  if (n == 0) return CTX_OK;
  CTX_REQUIRE(x, "amax: null pointer");
  amax_kernel<<<planes_grid(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(x, n, amax_bits);
  CTX_LAUNCH_CHECK("amax");
  return CTX_OK;
}

extern "C" int ctx_skip_grad_sum(const float* dcat, int32_t G, int64_t rows, int32_t C1, int32_t C2, const float* skip_ref,
                                 float slope, float* dskip, int32_t accumulate, void* stream) {
  CTX_REQUIRE(dcat && dskip && G > 0 && C1 >= 0 && C2 > 0 && rows >= 0, "skip_grad_sum: bad arguments");
  CTX_REQUIRE(C1 % 4 == 0 && C2 % 4 == 0, "skip_grad_sum: channel counts must be multiples of 4");
  auto al = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
  CTX_REQUIRE(al(dcat) && al(dskip) && al(skip_ref), "skip_grad_sum: pointers must be 16-byte aligned");
  if (rows == 0) return CTX_OK;
  skip_grad_sum_kernel<<<planes_grid(rows * (C2 / 4), 256, 16), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(dcat), G, rows, C1 / 4, C2 / 4, reinterpret_cast<const float4*>(skip_ref), slope,
      reinterpret_cast<float4*>(dskip), accumulate, FastDiv(C2 / 4));
  CTX_LAUNCH_CHECK("skip_grad_sum");
  return CTX_OK;
}
