// C-ABI entry points for the GEMM-shaped ops + library-wide helpers (errors, device info).
#include <string.h>
#include "common.cuh"
#include "kernels.h"

namespace ctx {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error in %s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return CTX_ERR_CUDA;
}

static unsigned long long g_launches = 0;
static thread_local int g_last_impl = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

// impl -> (use tensor cores?, passes, error if not eligible?)
struct ImplChoice { bool tc; int passes; bool force; };
static ImplChoice impl_choice(int impl) {
  switch (impl) {
    case CTX_IMPL_SIMT: return {false, 0, false};
    case CTX_IMPL_TC: return {true, 1, true};
    case CTX_IMPL_TC3: return {true, 3, true};
    case CTX_IMPL_FAST: return {true, 1, false};
    default: return {true, 3, false};          // CTX_IMPL_AUTO
  }
}

int sm_count() {
  static int cached_dev = -1;
  static int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

static int check_geom(const ctx_conv_geom* g, const char* who) {
  CTX_REQUIRE(g != nullptr, "%s: null geometry", who);
  CTX_REQUIRE(g->N >= 0 && g->H > 0 && g->W > 0 && g->P > 0 && g->Q > 0 && g->kh > 0 && g->kw > 0 && g->stride > 0 &&
                  g->pad_h >= 0 && g->pad_w >= 0,
              "%s: bad geometry N=%d H=%d W=%d P=%d Q=%d k=%dx%d stride=%d pad=%d,%d", who, g->N, g->H, g->W, g->P,
              g->Q, g->kh, g->kw, g->stride, g->pad_h, g->pad_w);
  // every small-grid pixel must map inside the padded big grid (SAME geometry)
  CTX_REQUIRE((g->P - 1) * g->stride - g->pad_h < g->H && (g->Q - 1) * g->stride - g->pad_w < g->W,
              "%s: small grid %dx%d does not fit big grid %dx%d at stride %d", who, g->P, g->Q, g->H, g->W, g->stride);
  return CTX_OK;
}

}  // namespace ctx

using namespace ctx;

extern "C" int ctx_version(void) { return CTX_B200_VERSION; }

extern "C" const char* ctx_last_error(void) { return g_err; }

extern "C" int ctx_device_info(int32_t* sms, int32_t* major, int32_t* minor) {
  int dev = 0;
  CTX_CUDA(cudaGetDevice(&dev));
  int a = 0, b = 0, c = 0;
  CTX_CUDA(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
  CTX_CUDA(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
  CTX_CUDA(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
  if (sms) *sms = a;
  if (major) *major = b;
  if (minor) *minor = c;
  return CTX_OK;
}

extern "C" int ctx_last_impl(void) { return g_last_impl; }

// CRC-32C (Castagnoli, reflected polynomial 0x82F63B78), host only: checksums of TF tensor-bundle checkpoints
// (tf_checkpoint.py).  Slicing-by-8 tables; ~1 GB/s, enough for the 190 MB of parameters.
extern "C" uint32_t ctx_crc32c(const void* data, uint64_t n, uint32_t crc) {
  static uint32_t T[8][256];
  static bool ready = false;
  if (!ready) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : (c >> 1);
      T[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) T[t][i] = (T[t - 1][i] >> 8) ^ T[0][T[t - 1][i] & 0xFF];
    ready = true;
  }
  const uint8_t* p = (const uint8_t*)data;
  crc = ~crc;
  while (n >= 8) {
    uint32_t lo, hi;
    memcpy(&lo, p, 4); memcpy(&hi, p + 4, 4);
    lo ^= crc;
    crc = T[7][lo & 0xFF] ^ T[6][(lo >> 8) & 0xFF] ^ T[5][(lo >> 16) & 0xFF] ^ T[4][lo >> 24] ^
          T[3][hi & 0xFF] ^ T[2][(hi >> 8) & 0xFF] ^ T[1][(hi >> 16) & 0xFF] ^ T[0][hi >> 24];
    p += 8; n -= 8;
  }
  while (n--) crc = (crc >> 8) ^ T[0][(crc ^ *p++) & 0xFF];
  return ~crc;
}

extern "C" int64_t ctx_launch_count(void) { return (int64_t)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" int ctx_tc_available(void) {
  static int cached_dev = -1, cached = 0;                  // (called by every tensor-core entry point: one query per device)
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (dev == cached_dev) return cached;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cached = (major == 10 && tc_compiled()) ? 1 : 0;
  cached_dev = dev;
  return cached;
}

extern "C" int ctx_conv2d_fprop_ld(const float* x, int32_t C, const float* w, float* y, int32_t y_ld, int32_t Co,
                                   const ctx_conv_geom* g, const ctx_epilogue* epi, int32_t impl, void* stream) {
  int rc = check_geom(g, "conv2d_fprop");
  if (rc) return rc;
  CTX_REQUIRE(C > 0 && Co > 0, "conv2d_fprop: bad channel count");
  CTX_REQUIRE(y_ld >= Co, "conv2d_fprop: destination pixel stride %d is smaller than Co = %d", y_ld, Co);
  if (g->N == 0) return CTX_OK;
  CTX_REQUIRE(x && w && y, "conv2d_fprop: null pointer");
  Epi e = make_epi(epi);
  CTX_REQUIRE(y_ld == Co || (!e.act_ref && !e.accumulate), "conv2d_fprop: act_ref / accumulate need a dense output");
  const ImplChoice ic = impl_choice(impl);
  if (ic.tc) {
    int ok = tc_conv2d_fprop_eligible(C, Co, *g) && y_ld % 4 == 0;
    if (ok && ctx_tc_available()) {
      g_last_impl = ic.passes == 3 ? CTX_IMPL_TC3 : CTX_IMPL_TC;
      return tc_conv2d_fprop(x, C, w, y, y_ld, Co, *g, e, ic.passes, (cudaStream_t)stream);
    }
    if (ic.force) { set_error("conv2d_fprop: shape not eligible for the tcgen05 path"); return CTX_ERR_UNSUPPORTED; }
  }
  g_last_impl = CTX_IMPL_SIMT;
  return simt_conv2d_fprop(x, C, w, y, y_ld, Co, *g, e, (cudaStream_t)stream);
}

extern "C" int ctx_conv2d_fprop(const float* x, int32_t C, const float* w, float* y, int32_t Co,
                                const ctx_conv_geom* g, const ctx_epilogue* epi, int32_t impl, void* stream) {
  return ctx_conv2d_fprop_ld(x, C, w, y, Co, Co, g, epi, impl, stream);
}

extern "C" int ctx_conv2d_transpose(const float* x1, int32_t C1, const float* x2, int32_t C2, int32_t x2_batch,
                                    const float* w, float* y, int32_t Cn, const ctx_conv_geom* g,
                                    const ctx_epilogue* epi, int32_t impl, void* stream) {
  int rc = check_geom(g, "conv2d_transpose");
  if (rc) return rc;
  CTX_REQUIRE(C1 > 0 && Cn > 0 && C2 >= 0, "conv2d_transpose: bad channel count");
  if (g->N == 0) return CTX_OK;
  CTX_REQUIRE(x1 && w && y, "conv2d_transpose: null pointer");
  CTX_REQUIRE((C2 == 0) == (x2 == nullptr), "conv2d_transpose: x2 and C2 must both be set or both be empty");
  CTX_REQUIRE(C2 == 0 || x2_batch > 0, "conv2d_transpose: x2_batch must be positive");
  if (g->N == 0) return CTX_OK;
  Epi e = make_epi(epi);
  const ImplChoice ic = impl_choice(impl);
  if (ic.tc) {
    int ok = tc_conv2d_transpose_eligible(C1, C2, x2_batch, Cn, *g);
    if (ok && ctx_tc_available()) {
      g_last_impl = ic.passes == 3 ? CTX_IMPL_TC3 : CTX_IMPL_TC;
      return tc_conv2d_transpose(x1, C1, x2, C2, x2_batch, w, y, Cn, *g, e, ic.passes, (cudaStream_t)stream);
    }
    // few output channels (the last decoder layer, Cn = 3): one GEMM over (tap, channel) columns + scatter epilogue
    if (!ok && ctx_tc_available() && !e.act && !e.act_ref && !e.round &&
        tc_conv2d_transpose_scatter_eligible(C1, C2, x2_batch, Cn, *g)) {
      g_last_impl = ic.passes == 3 ? CTX_IMPL_TC3 : CTX_IMPL_TC;
      rc = bias_init(y, e.bias, (int64_t)g->N * g->H * g->W * Cn, Cn, e.accumulate, (cudaStream_t)stream);
      if (rc) return rc;
      return tc_conv2d_transpose_scatter(x1, C1, x2, C2, x2_batch, w, y, Cn, *g, ic.passes, (cudaStream_t)stream);
    }
    if (ic.force) { set_error("conv2d_transpose: shape not eligible for the tcgen05 path"); return CTX_ERR_UNSUPPORTED; }
  }
  g_last_impl = CTX_IMPL_SIMT;
  return simt_conv2d_transpose(x1, C1, x2, C2, x2_batch, w, y, Cn, *g, e, (cudaStream_t)stream);
}

// ---- convolutions on 16-bit operand planes (tc_h16.cu) ---------------------------------------------------------------------
static int p16_cfg(const ctx_p16* c, H16Cfg* out, const char* who) {
  CTX_REQUIRE(c != nullptr, "%s: null ctx_p16", who);
  CTX_REQUIRE(c->dtype == CTX_P16_F16 || c->dtype == CTX_P16_BF16, "%s: dtype must be CTX_P16_F16 or CTX_P16_BF16", who);
  CTX_REQUIRE(c->nprod == 1 || c->nprod == 3, "%s: nprod must be 1 or 3", who);
  CTX_REQUIRE(c->nprod == 3 || c->dtype == CTX_P16_F16, "%s: the single-product mode needs fp16 planes", who);
  out->f16 = c->dtype == CTX_P16_F16 ? 1 : 0;
  out->nprod = c->nprod;
  out->a_scale = c->a_scale;
  out->amax_bits = c->amax_bits;
  return CTX_OK;
}
static const ctx_p16 kBf16x3 = {CTX_P16_BF16, 3, nullptr, nullptr};

extern "C" int ctx_conv2d_fprop_p16_eligible(int32_t C, int32_t Co, const ctx_conv_geom* g) {
  if (!g || check_geom(g, "conv2d_fprop_p16")) return 0;
  return (ctx_tc_available() && h16_conv2d_fprop_eligible(C, Co, *g)) ? 1 : 0;
}

extern "C" int ctx_conv2d_fprop_p16(const void* x_planes, int32_t C, const void* w_planes, float* y, void* y_planes,
                                    int32_t y_planes_n, int32_t Co, const ctx_conv_geom* g, const ctx_epilogue* epi,
                                    const ctx_p16* cfg, void* stream) {
  int rc = check_geom(g, "conv2d_fprop_p16");
  if (rc) return rc;
  H16Cfg c;
  rc = p16_cfg(cfg, &c, "conv2d_fprop_p16");
  if (rc) return rc;
  if (g->N == 0) return CTX_OK;
  CTX_REQUIRE(x_planes && w_planes && y, "conv2d_fprop_p16: null pointer");
  if (!ctx_tc_available() || !h16_conv2d_fprop_eligible(C, Co, *g)) {
    set_error("conv2d_fprop_p16: shape not eligible (C, Co %% 64, exact 128-pixel tiling) or no sm_100 device");
    return CTX_ERR_UNSUPPORTED;
  }
  g_last_impl = c.nprod == 3 ? CTX_IMPL_P16X3 : CTX_IMPL_P16X1;
  return h16_conv2d_fprop(x_planes, C, w_planes, y, y_planes, y_planes_n, Co, *g, make_epi(epi), c, (cudaStream_t)stream);
}

extern "C" int ctx_conv2d_transpose_p16_eligible(int32_t C1, int32_t C2, int32_t x2_batch, int32_t Cn,
                                                 const ctx_conv_geom* g) {
  if (!g || check_geom(g, "conv2d_transpose_p16")) return 0;
  return (ctx_tc_available() && h16_conv2d_transpose_eligible(C1, C2, x2_batch, Cn, *g)) ? 1 : 0;
}

extern "C" int ctx_conv2d_transpose_p16(const void* x1_planes, int32_t C1, const void* x2_planes, int32_t C2, int32_t x2_batch,
                                        const void* w_planes, float* y, void* y_planes, int32_t y_planes_n, int32_t Cn,
                                        const ctx_conv_geom* g, const ctx_epilogue* epi, const ctx_p16* cfg, void* stream) {
  int rc = check_geom(g, "conv2d_transpose_p16");
  if (rc) return rc;
  H16Cfg c;
  rc = p16_cfg(cfg, &c, "conv2d_transpose_p16");
  if (rc) return rc;
  if (g->N == 0) return CTX_OK;
  CTX_REQUIRE(x1_planes && w_planes && y, "conv2d_transpose_p16: null pointer");
  CTX_REQUIRE((C2 == 0) == (x2_planes == nullptr), "conv2d_transpose_p16: x2 and C2 must both be set or both be empty");
  if (!ctx_tc_available() || !h16_conv2d_transpose_eligible(C1, C2, x2_batch, Cn, *g)) {
    set_error("conv2d_transpose_p16: shape not eligible (channels %% 64, exact 128-pixel phase tiling) or no sm_100 device");
    return CTX_ERR_UNSUPPORTED;
  }
  g_last_impl = c.nprod == 3 ? CTX_IMPL_P16X3 : CTX_IMPL_P16X1;
  return h16_conv2d_transpose(x1_planes, C1, x2_planes, C2, x2_batch, w_planes, y, y_planes, y_planes_n, Cn, *g, make_epi(epi),
                              c, (cudaStream_t)stream);
}

extern "C" int ctx_conv2d_transpose_merged_p16_eligible(int32_t C1, int32_t C2, int32_t x2_batch, int32_t Cn,
                                                        const ctx_conv_geom* g) {
  if (!g || check_geom(g, "conv2d_transpose_merged_p16")) return 0;
  return (ctx_tc_available() && h16_conv2d_transpose_merged_eligible(C1, C2, x2_batch, Cn, *g)) ? 1 : 0;
}

extern "C" int64_t ctx_conv2d_transpose_merged_rows(int32_t Cn, const ctx_conv_geom* g) {
  if (!g || check_geom(g, "conv2d_transpose_merged_rows")) return 0;
  return h16_conv2d_transpose_merged_rows(Cn, *g);
}

extern "C" int ctx_conv2d_transpose_merged_pack_p16(const float* w, int32_t Ck, int32_t Cn, const ctx_conv_geom* g, float scale,
                                                    void* wm_planes, int32_t nplanes, int32_t dtype, void* stream) {
  int rc = check_geom(g, "conv2d_transpose_merged_pack");
  if (rc) return rc;
  CTX_REQUIRE(w && wm_planes && Ck > 0 && Cn > 0, "conv2d_transpose_merged_pack: bad arguments");
  CTX_REQUIRE(dtype == CTX_P16_F16 || dtype == CTX_P16_BF16, "conv2d_transpose_merged_pack: bad dtype");
  return h16_conv2d_transpose_merged_pack(w, Ck, Cn, *g, scale, wm_planes, nplanes, dtype == CTX_P16_F16, (cudaStream_t)stream);
}

extern "C" int ctx_conv2d_transpose_merged_p16(const void* x1_planes, int32_t C1, const void* x2_planes, int32_t C2,
                                               int32_t x2_batch, const void* wm_planes, float* y, void* y_planes,
                                               int32_t y_planes_n, int32_t Cn, const ctx_conv_geom* g,
                                               const ctx_epilogue* epi, const ctx_p16* cfg, void* stream) {
  int rc = check_geom(g, "conv2d_transpose_merged_p16");
  if (rc) return rc;
  H16Cfg c;
  rc = p16_cfg(cfg, &c, "conv2d_transpose_merged_p16");
  if (rc) return rc;
  if (g->N == 0) return CTX_OK;
  CTX_REQUIRE(x1_planes && wm_planes && y, "conv2d_transpose_merged_p16: null pointer");
  CTX_REQUIRE((C2 == 0) == (x2_planes == nullptr), "conv2d_transpose_merged_p16: x2 and C2 must both be set or both be empty");
  if (!ctx_tc_available() || !h16_conv2d_transpose_merged_eligible(C1, C2, x2_batch, Cn, *g)) {
    set_error("conv2d_transpose_merged_p16: shape not eligible (stride 2, 4 * Cn == 256 or 512, channels %% 64, even output, "
              "input grid tiling into 128 pixels) or no sm_100 device");
    return CTX_ERR_UNSUPPORTED;
  }
  g_last_impl = c.nprod == 3 ? CTX_IMPL_P16X3 : CTX_IMPL_P16X1;
  return h16_conv2d_transpose_merged(x1_planes, C1, x2_planes, C2, x2_batch, wm_planes, y, y_planes, y_planes_n, Cn, *g,
                                     make_epi(epi), c, (cudaStream_t)stream);
}

extern "C" int ctx_conv2d_wgrad_p16_eligible(int32_t Ca, int32_t Cb1, int32_t Cb2, int32_t small2_batch,
                                             const ctx_conv_geom* g) {
  if (!g || check_geom(g, "conv2d_wgrad_p16")) return 0;
  return (ctx_tc_available() && h16_conv2d_wgrad_eligible(Ca, Cb1, Cb2, small2_batch, *g)) ? 1 : 0;
}

extern "C" int ctx_conv2d_wgrad_p16(const void* big_planes, int32_t Ca, const void* small1_planes, int32_t Cb1,
                                    const void* small2_planes, int32_t Cb2, int32_t small2_batch, float* dw,
                                    const ctx_conv_geom* g, const ctx_p16* cfg, void* stream) {
  int rc = check_geom(g, "conv2d_wgrad_p16");
  if (rc) return rc;
  H16Cfg c;
  rc = p16_cfg(cfg, &c, "conv2d_wgrad_p16");
  if (rc) return rc;
  CTX_REQUIRE(Ca > 0 && Cb1 > 0 && Cb2 >= 0, "conv2d_wgrad_p16: bad channel count");
  if (g->N == 0) return CTX_OK;
  CTX_REQUIRE(big_planes && small1_planes && dw, "conv2d_wgrad_p16: null pointer");
  CTX_REQUIRE((Cb2 == 0) == (small2_planes == nullptr), "conv2d_wgrad_p16: small2 and Cb2 must both be set or both be empty");
  CTX_REQUIRE(Cb2 == 0 || small2_batch > 0, "conv2d_wgrad_p16: small2_batch must be positive");
  if (!ctx_tc_available() || !h16_conv2d_wgrad_eligible(Ca, Cb1, Cb2, small2_batch, *g)) {
    set_error("conv2d_wgrad_p16: shape not eligible (channels %% 64, small grid tiling into 64 pixels) or no sm_100 device");
    return CTX_ERR_UNSUPPORTED;
  }
  g_last_impl = c.nprod == 3 ? CTX_IMPL_P16X3 : CTX_IMPL_P16X1;
  return h16_conv2d_wgrad(big_planes, Ca, small1_planes, Cb1, small2_planes, Cb2, small2_batch, dw, *g, c, (cudaStream_t)stream);
}

// the round-1 names: bf16 planes, three products, two result planes
extern "C" int ctx_conv2d_fprop_bf16x3_eligible(int32_t C, int32_t Co, const ctx_conv_geom* g) {
  return ctx_conv2d_fprop_p16_eligible(C, Co, g);
}
extern "C" int ctx_conv2d_fprop_bf16x3(const void* x_planes, int32_t C, const void* w_planes, float* y, void* y_planes,
                                       int32_t Co, const ctx_conv_geom* g, const ctx_epilogue* epi, void* stream) {
  return ctx_conv2d_fprop_p16(x_planes, C, w_planes, y, y_planes, 2, Co, g, epi, &kBf16x3, stream);
}
extern "C" int ctx_conv2d_transpose_bf16x3_eligible(int32_t C1, int32_t C2, int32_t x2_batch, int32_t Cn,
                                                    const ctx_conv_geom* g) {
  return ctx_conv2d_transpose_p16_eligible(C1, C2, x2_batch, Cn, g);
}
extern "C" int ctx_conv2d_transpose_bf16x3(const void* x1_planes, int32_t C1, const void* x2_planes, int32_t C2,
                                           int32_t x2_batch, const void* w_planes, float* y, void* y_planes, int32_t Cn,
                                           const ctx_conv_geom* g, const ctx_epilogue* epi, void* stream) {
  return ctx_conv2d_transpose_p16(x1_planes, C1, x2_planes, C2, x2_batch, w_planes, y, y_planes, 2, Cn, g, epi, &kBf16x3, stream);
}
extern "C" int ctx_conv2d_transpose_merged_bf16x3_eligible(int32_t C1, int32_t C2, int32_t x2_batch, int32_t Cn,
                                                           const ctx_conv_geom* g) {
  return ctx_conv2d_transpose_merged_p16_eligible(C1, C2, x2_batch, Cn, g);
}
extern "C" int ctx_conv2d_transpose_merged_pack(const float* w, int32_t Ck, int32_t Cn, const ctx_conv_geom* g,
                                                void* wm_planes, void* stream) {
  return ctx_conv2d_transpose_merged_pack_p16(w, Ck, Cn, g, 1.f, wm_planes, 2, CTX_P16_BF16, stream);
}
extern "C" int ctx_conv2d_transpose_merged_bf16x3(const void* x1_planes, int32_t C1, const void* x2_planes, int32_t C2,
                                                  int32_t x2_batch, const void* wm_planes, float* y, void* y_planes,
                                                  int32_t Cn, const ctx_conv_geom* g, const ctx_epilogue* epi,
                                                  void* stream) {
  return ctx_conv2d_transpose_merged_p16(x1_planes, C1, x2_planes, C2, x2_batch, wm_planes, y, y_planes, 2, Cn, g, epi,
                                         &kBf16x3, stream);
}

extern "C" int ctx_conv2d_wgrad(const float* big, int32_t Ca, const float* small1, int32_t Cb1, const float* small2,
                                int32_t Cb2, int32_t small2_batch, float* dw, const ctx_conv_geom* g, int32_t impl,
                                void* stream) {
  int rc = check_geom(g, "conv2d_wgrad");
  if (rc) return rc;
  CTX_REQUIRE(Ca > 0 && Cb1 > 0 && Cb2 >= 0, "conv2d_wgrad: bad channel count");
  if (g->N == 0) return CTX_OK;
  CTX_REQUIRE(big && small1 && dw, "conv2d_wgrad: null pointer");
  CTX_REQUIRE((Cb2 == 0) == (small2 == nullptr), "conv2d_wgrad: small2 and Cb2 must both be set or both be empty");
  CTX_REQUIRE(Cb2 == 0 || small2_batch > 0, "conv2d_wgrad: small2_batch must be positive");
  if (g->N == 0) return CTX_OK;
  const ImplChoice ic = impl_choice(impl);
  if (ic.tc) {
    int ok = tc_conv2d_wgrad_eligible(Ca, Cb1, Cb2, small2_batch, *g);
    if (ok && ctx_tc_available()) {
      g_last_impl = ic.passes == 3 ? CTX_IMPL_TC3 : CTX_IMPL_TC;
      return tc_conv2d_wgrad(big, Ca, small1, Cb1, small2, Cb2, small2_batch, dw, *g, ic.passes, (cudaStream_t)stream);
    }
    if (ic.force) { set_error("conv2d_wgrad: shape not eligible for the tcgen05 path"); return CTX_ERR_UNSUPPORTED; }
  }
  g_last_impl = CTX_IMPL_SIMT;
  return simt_conv2d_wgrad(big, Ca, small1, Cb1, small2, Cb2, small2_batch, dw, *g, (cudaStream_t)stream);
}

extern "C" int ctx_gemm(int32_t transA, int32_t transB, int32_t M, int32_t N, int32_t K, const float* A, int32_t lda,
                        const float* B, int32_t ldb, float* C, int32_t ldc, const ctx_epilogue* epi, int32_t impl,
                        void* stream) {
  CTX_REQUIRE(M >= 0 && N >= 0 && K >= 0, "gemm: negative dimension");
  CTX_REQUIRE(A && B && C, "gemm: null pointer");
  CTX_REQUIRE(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, "gemm: leading dimension too small");
  if (M == 0 || N == 0) return CTX_OK;
  Epi e = make_epi(epi);
  const ImplChoice ic = impl_choice(impl);
  if (ic.tc) {
    int ok = tc_gemm_eligible(transA, transB, M, N, K, A, lda, B, ldb, C, ldc);
    if (ok && ctx_tc_available()) {
      g_last_impl = ic.passes == 3 ? CTX_IMPL_TC3 : CTX_IMPL_TC;
      return tc_gemm(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, e, ic.passes, (cudaStream_t)stream);
    }
    if (ic.force) { set_error("gemm: shape not eligible for the tcgen05 path"); return CTX_ERR_UNSUPPORTED; }
  }
  g_last_impl = CTX_IMPL_SIMT;
  return simt_gemm(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, e, (cudaStream_t)stream);
}
