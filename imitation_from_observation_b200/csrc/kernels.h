// Internal (non-ABI) declarations shared between the translation units of libctx_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "common.cuh"

namespace ctx {

// simt_gemm.cu — exact-fp32 CUDA-core implicit GEMMs
int simt_conv2d_fprop(const float* x, int C, const float* w, float* y, int y_ld, int Co, const ctx_conv_geom& g,
                      const Epi& epi, cudaStream_t stream);
int simt_conv2d_transpose(const float* x1, int C1, const float* x2, int C2, int x2_batch, const float* w,
                          float* y, int Cn, const ctx_conv_geom& g, const Epi& epi, cudaStream_t stream);
int simt_conv2d_wgrad(const float* big, int Ca, const float* s1, int Cb1, const float* s2, int Cb2,
                      int s2_batch, float* dw, const ctx_conv_geom& g, cudaStream_t stream);
int simt_gemm(int tA, int tB, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
              int ldc, const Epi& epi, cudaStream_t stream);

// tc_gemm.cu — tcgen05 / TMEM / TMA implicit GEMMs (TF32 operands, fp32 accumulate)
bool tc_compiled();
int tc_conv2d_fprop_eligible(int C, int Co, const ctx_conv_geom& g);
int tc_conv2d_fprop(const float* x, int C, const float* w, float* y, int y_ld, int Co, const ctx_conv_geom& g,
                    const Epi& epi, int passes, cudaStream_t stream);
int tc_conv2d_transpose_eligible(int C1, int C2, int x2_batch, int Cn, const ctx_conv_geom& g);
int tc_conv2d_transpose(const float* x1, int C1, const float* x2, int C2, int x2_batch, const float* w,
                        float* y, int Cn, const ctx_conv_geom& g, const Epi& epi, int passes, cudaStream_t stream);
int tc_conv2d_transpose_scatter_eligible(int C1, int C2, int x2_batch, int Cn, const ctx_conv_geom& g);
int tc_conv2d_transpose_scatter(const float* x1, int C1, const float* x2, int C2, int x2_batch, const float* w,
                                float* y, int Cn, const ctx_conv_geom& g, int passes, cudaStream_t stream);
int tc_conv2d_wgrad_eligible(int Ca, int Cb1, int Cb2, int s2_batch, const ctx_conv_geom& g);
int tc_conv2d_wgrad(const float* big, int Ca, const float* s1, int Cb1, const float* s2, int Cb2, int s2_batch,
                    float* dw, const ctx_conv_geom& g, int passes, cudaStream_t stream);
int tc_gemm_eligible(int tA, int tB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                     const float* C, int ldc);
int tc_gemm(int tA, int tB, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
            int ldc, const Epi& epi, int passes, cudaStream_t stream);

// tc_h16.cu — implicit-GEMM convolutions on 16-bit operand planes (fp16 / bf16, kind::f16, fp32 accumulate):
// nprod = 3 (hi*hi + hi*lo + lo*hi, fp32-grade: the forward pass) or nprod = 1 (fp16 hi planes = the TF32 significand,
// rounded to nearest: the backward pass)
struct H16Cfg {
  int f16;                  // 1 = fp16 planes, 0 = bf16 planes
  int nprod;                // 1 | 3
  const float* a_scale;     // device scalar or null: the (gradient) operand planes hold s * x; the result is divided by s
  uint32_t* amax_bits;      // device word or null: atomicMax of the bit patterns of |result|
};
// tensor map (f16: 0 = bf16, 1 = fp16, 2 = fp32), zero fill, 128-byte swizzle; dims / box innermost first; strides in ELEMENTS
// for dims 1..rank-1
int h16_make_map(CUtensorMap* out, int f16, const void* base, int rank, const int64_t* dims, const int64_t* strides,
                 const int* box);
int h16_conv2d_fprop_eligible(int C, int Co, const ctx_conv_geom& g);
int h16_conv2d_fprop(const void* xp, int C, const void* wp, float* y, void* yp, int yp_n, int Co, const ctx_conv_geom& g,
                     const Epi& epi, const H16Cfg& cfg, cudaStream_t stream);
int h16_conv2d_transpose_eligible(int C1, int C2, int x2_batch, int Cn, const ctx_conv_geom& g);
int h16_conv2d_transpose(const void* x1p, int C1, const void* x2p, int C2, int x2_batch, const void* wp, float* y, void* yp,
                         int yp_n, int Cn, const ctx_conv_geom& g, const Epi& epi, const H16Cfg& cfg, cudaStream_t stream);
int h16_conv2d_transpose_merged_eligible(int C1, int C2, int x2_batch, int Cn, const ctx_conv_geom& g);
int64_t h16_conv2d_transpose_merged_rows(int Cn, const ctx_conv_geom& g);
int h16_conv2d_transpose_merged_pack(const float* w, int Ck, int Cn, const ctx_conv_geom& g, float scale, void* planes,
                                     int nplanes, int f16, cudaStream_t stream);
int h16_conv2d_transpose_merged(const void* x1p, int C1, const void* x2p, int C2, int x2_batch, const void* wmp, float* y,
                                void* yp, int yp_n, int Cn, const ctx_conv_geom& g, const Epi& epi, const H16Cfg& cfg,
                                cudaStream_t stream);
int h16_conv2d_wgrad_eligible(int Ca, int Cb1, int Cb2, int s2_batch, const ctx_conv_geom& g);
int h16_conv2d_wgrad(const void* big, int Ca, const void* s1, int Cb1, const void* s2, int Cb2, int s2_batch, float* dw,
                     const ctx_conv_geom& g, const H16Cfg& cfg, cudaStream_t stream);

// pointwise.cu
int bias_init(float* y, const float* bias, int64_t n, int C, int accumulate, cudaStream_t stream);

}  // namespace ctx
