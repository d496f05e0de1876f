// Internal (non-ABI) declarations shared between the translation units of libctx_b200.so.
#pragma once
#include <cuda_bf16.h>
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

// tc_bf16x3.cu — forward convolutions from pre-split bf16 (hi, lo) planes (kind::f16, three products, fp32 accumulate)
int bf3_conv2d_fprop_eligible(int C, int Co, const ctx_conv_geom& g);
int bf3_conv2d_fprop(const __nv_bfloat16* xp, int C, const __nv_bfloat16* wp, float* y, __nv_bfloat16* yp, int Co,
                     const ctx_conv_geom& g, const Epi& epi, cudaStream_t stream);
int bf3_conv2d_transpose_eligible(int C1, int C2, int x2_batch, int Cn, const ctx_conv_geom& g);
int bf3_conv2d_transpose(const __nv_bfloat16* x1p, int C1, const __nv_bfloat16* x2p, int C2, int x2_batch,
                         const __nv_bfloat16* wp, float* y, __nv_bfloat16* yp, int Cn, const ctx_conv_geom& g,
                         const Epi& epi, cudaStream_t stream);
int bf3_conv2d_transpose_merged_eligible(int C1, int C2, int x2_batch, int Cn, const ctx_conv_geom& g);
int64_t bf3_conv2d_transpose_merged_rows(int Cn, const ctx_conv_geom& g);
int bf3_conv2d_transpose_merged_pack(const float* w, int Ck, int Cn, const ctx_conv_geom& g, __nv_bfloat16* planes,
                                     cudaStream_t stream);
int bf3_conv2d_transpose_merged(const __nv_bfloat16* x1p, int C1, const __nv_bfloat16* x2p, int C2, int x2_batch,
                                const __nv_bfloat16* wmp, float* y, __nv_bfloat16* yp, int Cn, const ctx_conv_geom& g,
                                const Epi& epi, cudaStream_t stream);

// pointwise.cu
int bias_init(float* y, const float* bias, int64_t n, int C, int accumulate, cudaStream_t stream);

}  // namespace ctx
