// Shared host/device helpers for libctx_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/ctx_b200.h"

namespace ctx {

// ---- error reporting -------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define CTX_REQUIRE(cond, ...)                       \
  do {                                               \
    if (!(cond)) {                                   \
      ::ctx::set_error(__VA_ARGS__);                 \
      return CTX_ERR_ARG;                            \
    }                                                \
  } while (0)

#define CTX_CUDA(call)                                               \
  do {                                                               \
    cudaError_t e__ = (call);                                        \
    if (e__ != cudaSuccess) return ::ctx::cuda_fail(e__, #call);     \
  } while (0)

#define CTX_LAUNCH_CHECK(name)                                       \
  do {                                                               \
    cudaError_t e__ = cudaGetLastError();                            \
    if (e__ != cudaSuccess) return ::ctx::cuda_fail(e__, name);      \
    ::ctx::count_launch();                                           \
  } while (0)

void count_launch();   // every successful kernel launch of this library (ctx_launch_count)

int sm_count();   // cached; <=0 when no device

// ---- fast integer division by a runtime-constant divisor (valid for 0 <= n < 2^31) -----------------------
struct FastDiv {
  uint32_t d, mul, shr;
  FastDiv() : d(1), mul(0), shr(0) {}
  explicit FastDiv(int32_t div) {
    d = (uint32_t)div;
    if (d <= 1) { d = 1; mul = 0; shr = 0; return; }
    uint32_t lg = 0;
    while ((1u << lg) < d) ++lg;                 // ceil(log2 d)
    uint32_t p = 31 + lg;
    mul = (uint32_t)(((1ull << p) + d - 1) / d);
    shr = p - 32;
  }
  __host__ __device__ __forceinline__ int32_t div(int32_t n) const {
#ifdef __CUDA_ARCH__
    return d == 1 ? n : (int32_t)(__umulhi((uint32_t)n, mul) >> shr);
#else
    return d == 1 ? n : (int32_t)((((uint64_t)(uint32_t)n * mul) >> 32) >> shr);
#endif
  }
  __host__ __device__ __forceinline__ void divmod(int32_t n, int32_t& q, int32_t& r) const {
    q = div(n);
    r = n - q * (int32_t)d;
  }
};

// ---- device epilogue -------------------------------------------------------------------------------------
__device__ __forceinline__ float round_tf32(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

struct Epi {
  const float* bias;
  const float* act_ref;
  float slope;
  int act, accumulate, round;

  __device__ __forceinline__ float apply(float v, int n, int64_t i, const float* out) const {
    if (bias) v += __ldg(bias + n);
    if (act) v = fmaxf(v, slope * v);
    if (act_ref) v *= (__ldg(act_ref + i) >= 0.f ? 1.f : slope);
    if (accumulate) v += out[i];
    if (round) v = round_tf32(v);
    return v;
  }
};

inline Epi make_epi(const ctx_epilogue* e) {
  Epi r;
  if (e) {
    r.bias = e->bias; r.act_ref = e->act_ref; r.slope = e->slope;
    r.act = e->act; r.accumulate = e->accumulate; r.round = e->round_tf32;
  } else {
    r.bias = nullptr; r.act_ref = nullptr; r.slope = 0.2f; r.act = 0; r.accumulate = 0; r.round = 0;
  }
  return r;
}

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace ctx
