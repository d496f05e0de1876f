// tcgen05 path — placeholder until the tensor-core kernels land: nothing is eligible, so CTX_IMPL_AUTO uses
// the SIMT kernels and CTX_IMPL_TC reports CTX_ERR_UNSUPPORTED.
#include "kernels.h"
namespace ctx {
bool tc_compiled() { return false; }
int tc_conv2d_fprop_eligible(int, int, const ctx_conv_geom&) { return 0; }
int tc_conv2d_fprop(const float*, int, const float*, float*, int, const ctx_conv_geom&, const Epi&, cudaStream_t) { return CTX_ERR_UNSUPPORTED; }
int tc_conv2d_transpose_eligible(int, int, int, int, const ctx_conv_geom&) { return 0; }
int tc_conv2d_transpose(const float*, int, const float*, int, int, const float*, float*, int, const ctx_conv_geom&, const Epi&, cudaStream_t) { return CTX_ERR_UNSUPPORTED; }
int tc_conv2d_wgrad_eligible(int, int, int, int, const ctx_conv_geom&) { return 0; }
int tc_conv2d_wgrad(const float*, int, const float*, int, const float*, int, int, float*, const ctx_conv_geom&, cudaStream_t) { return CTX_ERR_UNSUPPORTED; }
int tc_gemm_eligible(int, int, int, int, int, const float*, int, const float*, int, const float*, int) { return 0; }
int tc_gemm(int, int, int, int, int, const float*, int, const float*, int, float*, int, const Epi&, cudaStream_t) { return CTX_ERR_UNSUPPORTED; }
}
