// The image-side layers: convolutions whose INPUT has a handful of channels (the 64x64x3 frames, and d loss / d pred of the
// 3-channel reconstruction), on tcgen05 with the patch matrix built in SHARED MEMORY — no im2col buffer in HBM.
//
//   reference layer (arm_shaping.py)            here
//   h0 = lrelu(conv2d(image, 5x5 s2, 3->64))    few_fprop, three products on fp16 (hi, lo): fp32-grade          (:1283)
//   d h0 filter / d loss                        few_wgrad: patches^T x dY                                         (autodiff)
//   d d_h4 input / d loss (128 <- 3)            few_fprop, one product: conv of dY[.,.,3] with the filter as HWIO  (autodiff of :1329)
//   d d_h4 filter / d loss                      few_wgrad: patches(dY)^T x [state ‖ skip]                          (autodiff of :1329)
//
// K layout of a patch row: k = r * 16 + (s * C + c) — one 16-element MMA K-step per filter ROW (kw * C <= 16 real values,
// zero padded), so a 5x5x3 patch is 80 halves instead of 75 and every K-step starts on a 32-byte boundary.  The input rows
// under a CTA's 128 output pixels arrive in shared memory by bulk copy (TMA, one tile ahead of the compute); each thread then
// assembles the patch row(s) of one output pixel — normalising uint8 frames (base.py:117-119), scaling a gradient input by
// the power of two its max |.| asks for, splitting into fp16 (hi, lo), zero padding — straight into the SWIZZLE_128B operand
// layout the tensor core reads (row r at r * 128 B, 16-byte chunk c at position c ^ (r & 7)).  That byte layout serves both GEMM forms: as a
// K-major A tile [128 pixels x K] for the convolution, and as an MN-major A tile [K = pixels][M = patch index] for the filter
// gradient, where the pixels are the reduction dimension and the result D[patch index][channel] stays in tensor memory across
// all pixel tiles of the CTA (split-K over CTAs, fp32 vector reductions into dw at the end).
// HBM traffic is the algorithmic minimum: the image once, the result once.
#include <cuda.h>
#include <cuda_fp16.h>
#include <string.h>
#include <algorithm>
#include "kernels.h"

namespace ctx {

constexpr int kFThreads = 256;
constexpr uint32_t kFAtom = 128 * 128;              // 128 rows x 128 B
// Filters are ~N(0, 0.02): the lo part of an fp16 (hi, lo) split would fall into fp16's subnormals (19 significant bits
// instead of 22), so the two-plane filter image holds 2^8 * w and the epilogue divides by 2^8.
constexpr float kFewWScale = 256.f;

struct FewParams {
  CUtensorMap mapB[2];                              // wgrad: small1 / small2 planes as {64 ch, 128 pixel rows} boxes
  CUtensorMap mapY, mapYp;                          // fprop: y as {32 fp32, 128 pixel rows} boxes; planes {64 halves, 128 rows, plane}
  const void* x;                                    // [N,H,W,C] fp32 or uint8
  int32_t x_u8;
  const uint32_t* x_amax;                           // bits of max |x| or null: x is scaled by 2^k (amax * 2^k in [2^14, 2^15))
  const float* r_scale;                             // device scalar or null: the result is divided by it
  const uint4* wimg;                                // fprop: packed filter image (few_pack)
  float* y;                                         // fprop: [N,P,Q,Co];  wgrad: dw [kh,kw,C,Cb1+Cb2]
  void* yp;
  int32_t yp_n;
  int64_t plane_stride;
  const float* bias;
  float slope;
  int32_t act;
  uint32_t* amax_bits;
  const uint16_t* s1;                               // wgrad: [rows, Cb1] fp16 plane
  const uint16_t* s2;                               // wgrad: [s2_batch * P * Q, Cb2] fp16 plane or null
  int32_t Cb1, Cb2, s2_batch;
  int32_t N, H, W, C, P, Q, Co, kh, kw, stride, pad_h, pad_w;
  int32_t BP, NB, tiles_per_img, num_tiles;
  int32_t RI;                                       // input rows of a tile's footprint, per image
  int32_t natoms, np;
  int32_t groups;                                   // fprop: 256-thread groups per CTA (1 | 2)
  uint32_t a_region, raw_bytes, st_elems, grp_bytes, RS;   // fprop: A / staging region; raw rows; staged halves per plane; per group
  int32_t tiles_per_cta;                            // wgrad: contiguous tile range per CTA
};

namespace few {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
// bounded: a protocol bug must trap (launch failure on the host), never hang the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 4000000000LL) __trap();
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
// SM100 shared-memory matrix descriptor (version 1), SWIZZLE_128B (see tc_h16.cu make_desc)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
__device__ __forceinline__ uint16_t f2h(float a) { return __half_as_ushort(__float2half_rn(fminf(fmaxf(a, -65504.f), 65504.f))); }
__device__ __forceinline__ float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float scale_for(uint32_t amax_bits) {      // = planes.cu p16_scale_for
  const int e = (int)((amax_bits >> 23) & 0xFFu);
  if (amax_bits == 0u || e == 0 || e == 255) return 1.f;
  int k = 14 - (e - 127);
  k = k > 126 ? 126 : (k < -126 ? -126 : k);
  return __uint_as_float((uint32_t)(k + 127) << 23);
}

// ---- raw input: the rows of the tile's footprint, one bulk copy (TMA, no tensor map) per image -------------------------------
// Image rows are contiguous in HBM, so the footprint of a tile is one contiguous run of bytes per image; rows above / below
// the image are simply not copied (build masks them).  The copy completes on an mbarrier and is issued one tile ahead, so
// HBM latency hides behind the previous tile's tensor-core and epilogue work.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void raw_issue(const FewParams& p, int tile, uint8_t* raw, uint32_t bar) {
  int n0, p0;
  if (p.NB > 1 || p.tiles_per_img == 1) { n0 = tile * p.NB; p0 = 0; }
  else { n0 = tile / p.tiles_per_img; p0 = (tile - n0 * p.tiles_per_img) * p.BP; }
  const int h0 = p0 * p.stride - p.pad_h;
  const uint32_t rowbytes = (uint32_t)(p.W * p.C) * (p.x_u8 ? 1u : 4u);
  const int hs = max(h0, 0), he = min(h0 + p.RI, p.H);
  const int nimg = min(p.NB, p.N - n0);
  const uint32_t per = he > hs ? (uint32_t)(he - hs) * rowbytes : 0u;
  mbar_arrive_expect_tx(bar, per * (uint32_t)max(nimg, 0));
  if (per == 0u) return;
  for (int img = 0; img < nimg; ++img)
    bulk_g2s(smem_u32(raw) + (uint32_t)(img * p.RI + hs - h0) * rowbytes,
             reinterpret_cast<const uint8_t*>(p.x) + ((int64_t)(n0 + img) * p.H + hs) * rowbytes, per, bar);
}

// ---- stage: raw rows -> fp16 (hi, lo) rows with the SAME zero padding, one conversion per INPUT element -----------------------
// Staged row = [pad_w * C zeros | W * C values | zeros], RS halves long; rows above / below the image are zeros.  `nthr`
// threads (a 256-thread group) cooperate, `tid` is the thread's index in the group.
__device__ __forceinline__ void stage_rows(const FewParams& p, int n0, int p0, float s, const uint8_t* raw, uint16_t* st_hi,
                                           uint16_t* st_lo, int tid, int nthr) {
  const int h0 = p0 * p.stride - p.pad_h;
  const int wc = p.W * p.C, lead = p.pad_w * p.C;
  const int rows = p.NB * p.RI;
  const uint32_t rowbytes = (uint32_t)wc * (p.x_u8 ? 1u : 4u);
  const int rq = p.RS >> 2;                                // RS is a multiple of 4: four halves (8 bytes) per thread and turn
  for (int i = tid; i < rows * rq; i += nthr) {
    const int row = i / rq, e4 = (i - row * rq) * 4;
    const int img = row / p.RI, ri = row - img * p.RI;
    const int h = h0 + ri;
    const bool ok = n0 + img < p.N && h >= 0 && h < p.H;
    const uint8_t* src = raw + (uint32_t)row * rowbytes;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int g = e4 + k - lead;
      float t = 0.f;
      if (ok && g >= 0 && g < wc) {
        if (p.x_u8) t = __fmul_rn(__fsub_rn(__fmul_rn((float)src[g], 1.0f / 255.0f), 0.5f), 2.0f);     // base.py:117-119
        else t = reinterpret_cast<const float*>(src)[g];
      }
      v[k] = t * s;
    }
    const uint32_t h01 = pack_h2(v[0], v[1]), h23 = pack_h2(v[2], v[3]);
    *reinterpret_cast<uint2*>(st_hi + (int64_t)row * p.RS + e4) = make_uint2(h01, h23);
    if (st_lo) {
      const uint32_t l01 = pack_h2(v[0] - h2f((uint16_t)(h01 & 0xFFFFu)), v[1] - h2f((uint16_t)(h01 >> 16)));
      const uint32_t l23 = pack_h2(v[2] - h2f((uint16_t)(h23 & 0xFFFFu)), v[3] - h2f((uint16_t)(h23 >> 16)));
      *reinterpret_cast<uint2*>(st_lo + (int64_t)row * p.RS + e4) = make_uint2(l01, l23);
    }
  }
}

// ---- build: thread (row, sub) copies filter rows sub, sub + 2, ... of output pixel `row` from the staged rows into the
// swizzled A tile(s) ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void build_rows(const FewParams& p, const uint16_t* st_hi, const uint16_t* st_lo, uint8_t* A_hi,
                                           uint8_t* A_lo, int tid) {
  const int row = tid & 127, sub = tid >> 7;
  const int per_img = p.BP * p.Q;
  const int img = row / per_img, rem = row - img * per_img;
  const int dp = rem / p.Q, dq = rem - dp * p.Q;
  const int kc = p.kw * p.C;
  const uint32_t sw = (uint32_t)(row & 7);
  for (int r = sub; r < p.kh; r += 2) {
    const int src = (img * p.RI + dp * p.stride + r) * p.RS + dq * p.stride * p.C;
    const uint32_t a = (uint32_t)r >> 2, c0 = ((uint32_t)r & 3u) * 2u;
    const uint32_t o0 = a * kFAtom + (uint32_t)row * 128u + ((c0 ^ sw) << 4), o1 = a * kFAtom + (uint32_t)row * 128u + (((c0 + 1u) ^ sw) << 4);
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
      const uint16_t* sp = pl ? st_lo : st_hi;
      uint8_t* A = pl ? A_lo : A_hi;
      if (!sp) break;
      uint32_t w[8];
      if ((src & 1) == 0) {                                // (staged rows start 4-byte aligned: RS is a multiple of 4)
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(sp + src);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t t = 2 * j < kc ? s32[j] : 0u;
          if (2 * j + 1 >= kc) t &= 0xFFFFu;
          w[j] = t;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t lo = 2 * j < kc ? (uint32_t)sp[src + 2 * j] : 0u;
          const uint32_t hi = 2 * j + 1 < kc ? (uint32_t)sp[src + 2 * j + 1] : 0u;
          w[j] = lo | (hi << 16);
        }
      }
      *reinterpret_cast<uint4*>(A + o0) = make_uint4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<uint4*>(A + o1) = make_uint4(w[4], w[5], w[6], w[7]);
    }
  }
}

__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory"); }

__device__ __forceinline__ void tile_origin(const FewParams& p, int tile, int* n0, int* p0) {
  if (p.NB > 1 || p.tiles_per_img == 1) { *n0 = tile * p.NB; *p0 = 0; }
  else { *n0 = tile / p.tiles_per_img; *p0 = (tile - *n0 * p.tiles_per_img) * p.BP; }
}
}  // namespace few

// ---- convolution: D[128 pixels x Co] = patches x W^T -----------------------------------------------------------------------------
// A CTA is TWO independent 256-thread groups that share the filter image: each group walks its own tiles (raw rows by bulk
// copy -> stage -> build -> MMAs -> epilogue) with its own operand tile, accumulator columns and barriers, so one group's
// tensor-core and store phases overlap the other's conversion phases (what two co-resident CTAs would do, minus a second
// copy of the filters, which would not fit).
constexpr int kFGroups = 2;                              // (p.groups = 1 when two operand tiles do not fit shared memory)

template <int kProd>
__global__ void __launch_bounds__(kFThreads * kFGroups) few_fprop_kernel(const __grid_constant__ FewParams p) {
  using namespace few;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int np = kProd == 3 ? 2 : 1;
  const int grp = threadIdx.x >> 8, tid = threadIdx.x & 255;
  const uint32_t a_plane = (uint32_t)p.natoms * kFAtom, w_plane = (uint32_t)p.natoms * (uint32_t)p.Co * 128u;
  uint8_t* Wt = sm;                                        // shared by both groups
  uint8_t* A = Wt + np * w_plane + (uint32_t)grp * p.grp_bytes;   // [A / epilogue staging | raw rows | staged hi | staged lo]
  uint8_t* raw = A + p.a_region;
  uint16_t* st_hi = reinterpret_cast<uint16_t*>(raw + p.raw_bytes);
  uint16_t* st_lo = np == 2 ? st_hi + p.st_elems : nullptr;
  __shared__ __align__(8) uint64_t bars_all[kFGroups][2];  // [0]: the tile's raw rows have landed; [1]: its MMAs are done
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[128];
  uint64_t* bars = bars_all[grp];
  const int warp = tid >> 5, lane = tid & 31;
  const uint32_t gcols = (uint32_t)(np * p.Co);            // accumulator columns of one group
  uint32_t ncols = 32;
  while (ncols < (uint32_t)p.groups * gcols) ncols <<= 1;
  const int first_tile = blockIdx.x * p.groups + grp, tile_step = gridDim.x * p.groups;
  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (first_tile < p.num_tiles) raw_issue(p, first_tile, raw, smem_u32(&bars[0]));
  }
  if (threadIdx.x >> 5 == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {   // the packed filter image is the shared-memory image: plain 128-bit copies
    const uint4* src = p.wimg;
    uint4* dst = reinterpret_cast<uint4*>(Wt);
    const int n16 = (int)(np * w_plane / 16);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
    if ((int)threadIdx.x < p.Co) s_bias[threadIdx.x] = p.bias ? __ldg(p.bias + threadIdx.x) : 0.f;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot + (uint32_t)grp * gcols;
  float s = 1.f;
  if (p.x_amax) s = scale_for(__ldg(p.x_amax));
  float inv = 1.f / s;
  if (p.r_scale) inv = 1.f / __ldg(p.r_scale);
  if constexpr (kProd == 3) inv *= 1.f / kFewWScale;     // the (hi, lo) filter image holds kFewWScale * w
  const uint32_t idesc = (1u << 4) | ((uint32_t)(p.Co >> 3) << 17) | ((128u >> 4) << 24);
  uint32_t amax = 0u;
  uint32_t it = 0;
  for (int tile = first_tile; tile < p.num_tiles; tile += tile_step, ++it) {
    int n0, p0;
    tile_origin(p, tile, &n0, &p0);
    mbar_wait(smem_u32(&bars[0]), it & 1u);
    stage_rows(p, n0, p0, s, raw, st_hi, st_lo, tid, kFThreads);
    if (tid == 0) tma_store_wait_read();                 // the previous tile's stores have drained the staging area = A
    group_sync(grp);
    // the raw rows are consumed: fetch the next tile's under this tile's build, MMAs and epilogue
    if (tid == 0 && tile + tile_step < p.num_tiles) raw_issue(p, tile + tile_step, raw, smem_u32(&bars[0]));
    build_rows(p, st_hi, st_lo, A, np == 2 ? A + a_plane : nullptr, tid);
    fence_async_smem();
    tc_fence_before();
    group_sync(grp);
    if (tid == 0) {
      tc_fence_after();
      const uint32_t sa = smem_u32(A), sw = smem_u32(Wt);
      for (int ks = 0; ks < p.kh; ++ks) {
        const uint32_t ao = (uint32_t)(ks >> 2) * kFAtom, wo = (uint32_t)(ks >> 2) * (uint32_t)p.Co * 128u;
        const uint64_t adv = (uint64_t)(ks & 3) * 2u;
        const uint64_t a_hi = make_desc(sa + ao, 16, 1024) + adv, b_hi = make_desc(sw + wo, 16, 1024) + adv;
        umma_f16(tmem_base, a_hi, b_hi, idesc, ks ? 1u : 0u);
        if constexpr (kProd == 3) {
          const uint64_t a_lo = make_desc(sa + a_plane + ao, 16, 1024) + adv, b_lo = make_desc(sw + w_plane + wo, 16, 1024) + adv;
          umma_f16(tmem_base + (uint32_t)p.Co, a_hi, b_lo, idesc, ks ? 1u : 0u);
          umma_f16(tmem_base + (uint32_t)p.Co, a_lo, b_hi, idesc, 1u);
        }
      }
      umma_commit(smem_u32(&bars[1]));
    }
    mbar_wait(smem_u32(&bars[1]), it & 1u);
    tc_fence_after();
    // ---- epilogue, 64 output channels at a time: tensor memory -> registers (row per thread) -> bias / lrelu / planes ->
    // shared memory in the 128-byte-swizzled box layout (conflict-free: chunk c of row r at position c ^ (r & 7)) -> TMA
    // tensor stores, which write whole 128-byte lines (a row-per-thread st.global touches 32 lines per instruction).  The
    // staging area is the A tile: its MMAs are complete.  Rows past the last image are clipped by the TMA unit.
    {
      const int q = warp & 3, half = warp >> 2;
      const int r = q * 32 + lane;
      const uint32_t sw = (uint32_t)(r & 7);
      uint8_t* st_f = A;                                   // [2 atoms][128 rows x 32 fp32]
      uint8_t* st_h = A + 2 * kFAtom;                      // [128 rows x 64 halves] hi, then lo
      const int row0 = tile * 128;
      for (int cg = 0; cg < p.Co; cg += 64) {
        const int c0 = cg + half * 32;                     // this warp's 32 columns = one fp32 atom
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        if constexpr (kProd == 3) {
          uint32_t v2[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.Co + c0), v2);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o;
          float* oo = reinterpret_cast<float*>(&o);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float t = __uint_as_float(v[j + e]) * inv + s_bias[c0 + j + e];
            if (p.act) t = fmaxf(t, p.slope * t);
            oo[e] = t;
            v[j + e] = __float_as_uint(t);
            amax = max(amax, __float_as_uint(t) & 0x7FFFFFFFu);
          }
          *reinterpret_cast<float4*>(st_f + (uint32_t)half * kFAtom + (uint32_t)r * 128u + ((((uint32_t)j >> 2) ^ sw) << 4)) = o;
        }
        if (p.yp) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float x0 = __uint_as_float(v[j + 2 * e]), x1 = __uint_as_float(v[j + 2 * e + 1]);
              hi[e] = pack_h2(x0, x1);
              lo[e] = pack_h2(x0 - h2f((uint16_t)(hi[e] & 0xFFFFu)), x1 - h2f((uint16_t)(hi[e] >> 16)));
            }
            const uint32_t o = (uint32_t)r * 128u + ((((uint32_t)half * 4u + ((uint32_t)j >> 3)) ^ sw) << 4);
            *reinterpret_cast<uint4*>(st_h + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (p.yp_n == 2) *reinterpret_cast<uint4*>(st_h + kFAtom + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        fence_async_smem();
        tc_fence_before();
        group_sync(grp);
        if (tid == 0) {
          tma_store_2d(&p.mapY, smem_u32(st_f), cg, row0);
          tma_store_2d(&p.mapY, smem_u32(st_f) + kFAtom, cg + 32, row0);
          if (p.yp) {
            tma_store_3d(&p.mapYp, smem_u32(st_h), cg, row0, 0);
            if (p.yp_n == 2) tma_store_3d(&p.mapYp, smem_u32(st_h) + kFAtom, cg, row0, 1);
          }
          tma_store_commit();
          if (cg + 64 < p.Co) tma_store_wait_read();       // the next column group reuses the staging area right away
        }
        // the accumulator is free again; the staging area = A tile once the stores have read it (waited for above, or —
        // last column group — under the next tile's stage pass)
        if (cg + 64 < p.Co) group_sync(grp);
      }
    }
  }
  if (p.amax_bits) {
    amax = __reduce_max_sync(0xffffffffu, amax);
    if (lane == 0 && amax) atomicMax(p.amax_bits, amax);
  }
  if (tid == 0) tma_store_wait_all();                      // shared memory must outlive the last stores
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x >> 5 == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(ncols) : "memory");
  }
}

// ---- filter gradient: D[patch index x Cb] += patches^T x [small1 ‖ small2], pixels = the reduction dimension ---------------------
// The MMAs of a tile are a few hundred cycles; staging and building its patch tile is the work.  So the CTA is small (one A and
// one B tile, ~70 KB: three CTAs share an SM and overlap each other's phases), the raw rows are fetched TWO tiles ahead, and B
// (the 16-bit plane rows of the tile, one TMA box per 64-channel atom) is requested as soon as the previous tile's MMAs retire.
__global__ void __launch_bounds__(kFThreads) few_wgrad_kernel(const __grid_constant__ FewParams p) {
  using namespace few;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int Cb = p.Cb1 + p.Cb2;
  const uint32_t a_bytes = 2u * kFAtom, b_bytes = (uint32_t)(Cb / 64) * kFAtom;     // A always spans M = 128 (two atoms)
  uint8_t* Abuf = sm;
  uint8_t* Bbuf = sm + a_bytes;
  uint8_t* rawb[2] = {Bbuf + b_bytes, Bbuf + b_bytes + p.raw_bytes};
  uint16_t* st_hi = reinterpret_cast<uint16_t*>(rawb[1] + p.raw_bytes);
  // [0], [1]: raw rows of buffer 0 / 1 landed; [2]: B landed; [3]: the tile's MMAs are done; [5]: all done
  __shared__ __align__(8) uint64_t bars[6];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t_begin = blockIdx.x * p.tiles_per_cta;
  const int t_end = min(p.num_tiles, t_begin + p.tiles_per_cta);
  if (t_begin >= t_end) return;                            // (uniform over the CTA)
  uint32_t ncols = 32;
  while (ncols < (uint32_t)Cb) ncols <<= 1;
  const int pq = p.P * p.Q;
  auto b_issue = [&](int tile) {                           // one thread
    int n0, p0;
    tile_origin(p, tile, &n0, &p0);
    const uint32_t bar = smem_u32(&bars[2]);
    mbar_arrive_expect_tx(bar, b_bytes);
    for (int a = 0; a < p.Cb1 / 64; ++a)
      tma_load_2d(smem_u32(Bbuf) + (uint32_t)a * kFAtom, &p.mapB[0], bar, a * 64, tile * 128);
    const int row2 = (n0 % p.s2_batch) * pq + p0 * p.Q;
    for (int a = 0; a < p.Cb2 / 64; ++a)
      tma_load_2d(smem_u32(Bbuf) + (uint32_t)(p.Cb1 / 64 + a) * kFAtom, &p.mapB[1], bar, a * 64, row2);
  };
  if (threadIdx.x == 0) {
    for (int i = 0; i < 6; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    raw_issue(p, t_begin, rawb[0], smem_u32(&bars[0]));
    if (t_begin + 1 < t_end) raw_issue(p, t_begin + 1, rawb[1], smem_u32(&bars[1]));
    b_issue(t_begin);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  float s = 1.f;
  if (p.x_amax) s = scale_for(__ldg(p.x_amax));
  float inv = 1.f / s;
  if (p.r_scale) inv = 1.f / __ldg(p.r_scale);
  // A and B MN-major (bits 15, 16), M = 128, N = Cb
  const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(Cb >> 3) << 17) | ((128u >> 4) << 24);
  for (int tile = t_begin, it = 0; tile < t_end; ++tile, ++it) {
    const int rb = it & 1;
    int n0, p0;
    tile_origin(p, tile, &n0, &p0);
    mbar_wait(smem_u32(&bars[rb]), (uint32_t)(it >> 1) & 1u);
    stage_rows(p, n0, p0, s, rawb[rb], st_hi, nullptr, threadIdx.x, kFThreads);
    if (it >= 1) mbar_wait(smem_u32(&bars[3]), (uint32_t)(it - 1) & 1u);         // the previous tile's MMAs released A and B
    __syncthreads();
    if (threadIdx.x == 0) {
      if (tile + 2 < t_end) raw_issue(p, tile + 2, rawb[rb], smem_u32(&bars[rb]));     // (this buffer's rows are consumed)
      if (it >= 1) b_issue(tile);
    }
    build_rows(p, st_hi, nullptr, Abuf, nullptr, threadIdx.x);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_wait(smem_u32(&bars[2]), (uint32_t)it & 1u);    // B of this tile has landed
      tc_fence_after();
      const uint32_t sa = smem_u32(Abuf), sb = smem_u32(Bbuf);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)                       // 16 pixels per K-step = 2048 B inside every atom
        umma_f16(tmem_base, make_desc(sa + ks * 2048, kFAtom, 1024), make_desc(sb + ks * 2048, kFAtom, 1024), idesc,
                 (it | ks) ? 1u : 0u);
      umma_commit(smem_u32(&bars[3]));
      if (tile + 1 == t_end) umma_commit(smem_u32(&bars[5]));
    }
  }
  mbar_wait(smem_u32(&bars[5]), 0);
  tc_fence_after();
  {
    const int q = warp & 3, half = warp >> 2;
    const int m = q * 32 + lane;                            // D row = patch index r * 16 + (s * C + c)
    const int kc = p.kw * p.C;
    const bool row_ok = (m >> 4) < p.kh && (m & 15) < kc;
    float* drow = p.y + ((int64_t)(m >> 4) * kc + (m & 15)) * Cb;
    for (int c0 = half * (Cb / 2); c0 < (half + 1) * (Cb / 2); c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      if (!row_ok) continue;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c0 + j), "f"(__uint_as_float(v[j]) * inv),
                     "f"(__uint_as_float(v[j + 1]) * inv), "f"(__uint_as_float(v[j + 2]) * inv),
                     "f"(__uint_as_float(v[j + 3]) * inv) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

// ---- filter [kh,kw,C,Co] fp32 -> the shared-memory image few_fprop reads: [plane][atom][Co rows x 128 B], swizzled -------------
__global__ void few_pack_kernel(const float* __restrict__ w, int kh, int kw, int C, int Co, int natoms, int nplanes,
                                uint4* __restrict__ img) {
  const int chunks = natoms * Co * 8;                      // 16-byte chunks per plane
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += gridDim.x * blockDim.x) {
    const int a = i / (Co * 8), rem = i - a * Co * 8;
    const int co = rem >> 3, c = rem & 7;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int kidx = a * 64 + c * 8 + e, r = kidx >> 4, j = kidx & 15;
      v[e] = (r < kh && j < kw * C) ? w[((int64_t)r * kw * C + j) * Co + co] * (nplanes == 2 ? kFewWScale : 1.f) : 0.f;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      hi[e] = few::pack_h2(v[2 * e], v[2 * e + 1]);
      lo[e] = few::pack_h2(v[2 * e] - few::h2f((uint16_t)(hi[e] & 0xFFFFu)), v[2 * e + 1] - few::h2f((uint16_t)(hi[e] >> 16)));
    }
    const int64_t o = ((int64_t)a * Co + co) * 8 + (c ^ (co & 7));
    img[o] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (nplanes == 2) img[(int64_t)natoms * Co * 8 + o] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ---- transposed convolution with a few OUTPUT channels (the reconstruction, arm_shaping.py:1329) -------------------------------
// y[n, oh, ow, co] = bias[co] + sum over (p, q, r, s) with oh = p * st - pad_h + r, ow = q * st - pad_w + s, and channels c of
// cat[n, p, q, c] * W[r, s, co, c].  As a GEMM the channel contraction comes first: G[pixel (p, q)][m = r * 16 + (s * Cn + co)]
// = cat[pixel][:] . W[r, s, co, :] — 128 (here 192) input pixels x KP = 16 kh columns x K = C1 + C2, a dense tensor-core shape
// (N = 80 for 5 x 5 x 3) — and the col2im sum over the <= ceil(kh / st) * ceil(kw / st) patch entries that land on an output pixel
// is a GATHER from G held in shared memory, so nothing is scattered with atomics and y is written once, coalesced.  A CTA tile
// owns BPo input rows (= BPo * st output rows) and also computes the halo rows above / below whose patches reach into its
// output rows (TMA zero-fills rows outside the image).  Operands: fp16 (hi, lo) planes of [state ‖ skip], one TMA box per
// 64-channel block; three products, cross terms in a second accumulator.  Optionally the reconstruction loss of the reference
// (l2_loss of out - target, arm_shaping.py:1352-1353) is evaluated in the same gather: loss sums, d loss / d y and its max.
struct FewTParams {
  CUtensorMap map[2];                               // state / skip planes: {64 ch, Q, P, images, plane}
  const uint4* wimg;                                // packed filter image [plane][64-channel block][KP rows x 128 B]
  float* y;
  const float* bias;
  float slope;
  int32_t act;
  int32_t N, H, W, P, Q, Cn, kh, kw, stride, pad_h, pad_w, x2_batch;
  int32_t nblk1, nblk;                              // 64-channel blocks of the state / of both sources
  int32_t RT, BPo, halo_top, tiles_per_img, num_tiles, KP, GS;   // GS: row stride of G in floats (odd)
  // fused l2 loss (all null: plain layer)
  const float* tgt;
  const float* resid;
  const float* weight;
  float* loss;
  float* dpred;
  uint32_t* dpred_amax;
  int32_t pass_images;
};

template <int ST>
__global__ void __launch_bounds__(kFThreads) few_transpose_kernel(const __grid_constant__ FewTParams p) {
  using namespace few;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int RTQ = p.RT * p.Q;
  const uint32_t a_plane = (uint32_t)RTQ * 128u, a_blk = 2u * a_plane;             // [block][plane][RTQ rows x 128 B]
  const uint32_t b_blk = (uint32_t)p.KP * 128u, b_plane = (uint32_t)p.nblk * b_blk;  // [plane][block][KP rows x 128 B]
  uint8_t* A = sm;
  uint8_t* Bw = A + p.nblk * a_blk;
  float* G = reinterpret_cast<float*>(Bw + 2 * b_plane);
  // [0]: the tile's operands have landed; [1], [2]: the MMAs of M tile 0 / 1 are done
  __shared__ __align__(8) uint64_t bars[3];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int MT = (RTQ + 127) / 128;
  const uint32_t dcols = 2u * (uint32_t)p.KP;              // D1 | D2 of one M tile
  uint32_t ncols = 32;
  while (ncols < (uint32_t)MT * dcols) ncols <<= 1;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    uint4* dst = reinterpret_cast<uint4*>(Bw);
    const int n16 = (int)(2 * b_plane / 16);
    for (int i = threadIdx.x; i < n16; i += kFThreads) dst[i] = __ldg(p.wimg + i);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(p.KP >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t tx_bytes = (uint32_t)p.nblk * a_blk;
  auto issue_loads = [&](int tile) {
    const int n = tile / p.tiles_per_img, p0 = (tile - n * p.tiles_per_img) * p.BPo;
    mbar_arrive_expect_tx(smem_u32(&bars[0]), tx_bytes);
    for (int blk = 0; blk < p.nblk; ++blk) {
      const bool second = blk >= p.nblk1;
      tma_load_5d(smem_u32(A + blk * a_blk), &p.map[second ? 1 : 0], smem_u32(&bars[0]), (second ? blk - p.nblk1 : blk) * 64, 0,
                  p0 - p.halo_top, second ? n % p.x2_batch : n, 0);
    }
  };
  if (threadIdx.x == 0 && (int)blockIdx.x < p.num_tiles) issue_loads(blockIdx.x);
  const float inv = 1.f / kFewWScale;
  uint32_t amax = 0u;
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
    const int n = tile / p.tiles_per_img, p0 = (tile - n * p.tiles_per_img) * p.BPo;
    mbar_wait(smem_u32(&bars[0]), it & 1u);
    if (threadIdx.x == 0) {
      tc_fence_after();
      for (int mt = 0; mt < MT; ++mt) {
        const uint32_t d1 = tmem_base + (uint32_t)mt * dcols, d2 = d1 + (uint32_t)p.KP;
        for (int blk = 0; blk < p.nblk; ++blk) {
          const uint32_t sa = smem_u32(A + blk * a_blk) + (uint32_t)mt * kFAtom, sb = smem_u32(Bw + blk * b_blk);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t a_hi = make_desc(sa, 16, 1024) + 2u * ks, a_lo = make_desc(sa + a_plane, 16, 1024) + 2u * ks;
            const uint64_t b_hi = make_desc(sb, 16, 1024) + 2u * ks, b_lo = make_desc(sb + b_plane, 16, 1024) + 2u * ks;
            const uint32_t first = (blk | ks) ? 1u : 0u;
            umma_f16(d1, a_hi, b_hi, idesc, first);
            umma_f16(d2, a_hi, b_lo, idesc, first);
            umma_f16(d2, a_lo, b_hi, idesc, 1u);
          }
        }
        umma_commit(smem_u32(&bars[1 + mt]));            // M tile 0 drains while M tile 1 multiplies
      }
    }
    // ---- tensor memory -> G (fp32 sums of the two accumulators, row stride GS: odd, so the row-per-thread writes spread
    // over the banks); warps 0..3 take M tile 0, warps 4..7 M tile 1
    {
      const int mt = warp >> 2, q = warp & 3;
      const int row = mt * 128 + q * 32 + lane;
      if (mt < MT) {
        mbar_wait(smem_u32(&bars[1 + mt]), it & 1u);
        tc_fence_after();
        for (int c0 = 0; c0 < p.KP; c0 += 16) {
          uint32_t v1[16], v2[16];
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)mt * dcols + (uint32_t)c0;
          tmem_ld16(ta, v1);
          tmem_ld16(ta + (uint32_t)p.KP, v2);
          if (row < RTQ) {
            float* gr = G + (int64_t)row * p.GS + c0;
#pragma unroll
            for (int j = 0; j < 16; ++j) gr[j] = __uint_as_float(v1[j]) + __uint_as_float(v2[j]);
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    // every MMA of the tile is done (both commits were waited for above): the operand tiles are free -> next tile's loads
    // run under the gather below
    if (threadIdx.x == 0 && tile + (int)gridDim.x < p.num_tiles) issue_loads(tile + gridDim.x);
    // ---- col2im gather + bias (+ lrelu) (+ the reconstruction loss and its gradient): one output pixel per thread and turn
    {
      const int out_rows = min(p.BPo * ST, p.H - p0 * ST);
      const int pbase = p0 - p.halo_top;
      const int g = p.pass_images > 0 ? n / p.pass_images : 0;
      const float wg = p.weight ? __ldg(p.weight + g) : 1.f;
      float ls = 0.f;
      // No division inside the loop: (row, column) of the thread's pixel advance by constants from turn to turn.  The patch
      // entries of one output pixel sit at constant strides in G: filter row r + ST <-> input row pp - 1, filter column
      // sx + ST <-> input column qq - 1 (shared-memory loads by 32-bit address; G is row-major [pixel][r * 16 + sx * Cn + co]).
      const uint32_t Gs = smem_u32(G);
      const int step_row = kFThreads / p.W, step_col = kFThreads - step_row * p.W;
      int ohl = (int)threadIdx.x / p.W, ow = (int)threadIdx.x - ohl * p.W;
      const int d_r = ST * 16 - p.Q * p.GS, d_s = ST * p.Cn - p.GS;          // G offset steps (floats) of r += ST, sx += ST
      const int total = out_rows * p.W;
      for (int pix = threadIdx.x; pix < total; pix += kFThreads) {
        const int oh = p0 * ST + ohl;
        float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        const unsigned uh = (unsigned)(oh + p.pad_h), uw = (unsigned)(ow + p.pad_w);
        const int r0 = (int)(uh % ST), s0 = (int)(uw % ST);
        int pp = (int)(uh / ST) - pbase;                               // input row of filter row r0 (rows outside hold zeros)
        const int q0 = (int)(uw / ST);
        int off_r = (pp * p.Q + q0) * p.GS + r0 * 16 + s0 * p.Cn;
        for (int r = r0; r < p.kh; r += ST, --pp, off_r += d_r) {
          if (pp < 0 || pp >= p.RT) continue;
          int qq = q0, off = off_r;
          for (int sx = s0; sx < p.kw; sx += ST, --qq, off += d_s) {
            if (qq < 0 || qq >= p.Q) continue;
            const uint32_t ga = Gs + 4u * (uint32_t)off;
            float g0, g1, g2;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(g0) : "r"(ga));
            if (p.Cn == 3) {
              asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(g1) : "r"(ga));
              asm volatile("ld.shared.f32 %0, [%1+8];" : "=f"(g2) : "r"(ga));
              acc[0] += g0; acc[1] += g1; acc[2] += g2;
            } else {
              acc[0] += g0;
#pragma unroll
              for (int co = 1; co < 5; ++co)
                if (co < p.Cn) {
                  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(g1) : "r"(ga + 4u * (uint32_t)co));
                  acc[co] += g1;
                }
            }
          }
        }
        const int64_t o = (((int64_t)n * p.H + oh) * p.W + ow) * p.Cn;
        const int64_t ot = (((int64_t)(n - g * p.pass_images) * p.H + oh) * p.W + ow) * p.Cn;
#pragma unroll
        for (int co = 0; co < 5; ++co) {
          if (co >= p.Cn) break;
          float t = acc[co] * inv + (p.bias ? __ldg(p.bias + co) : 0.f);
          if (p.act) t = fmaxf(t, p.slope * t);
          p.y[o + co] = t;
          if (p.tgt) {
            const float d = t + (p.resid ? __ldg(p.resid + ot + co) : 0.f) - __ldg(p.tgt + ot + co);
            ls += d * d;
            if (p.dpred) {
              const float dv = wg * d;
              p.dpred[o + co] = dv;
              amax = max(amax, __float_as_uint(dv) & 0x7FFFFFFFu);
            }
          }
        }
        ohl += step_row; ow += step_col;
        if (ow >= p.W) { ow -= p.W; ++ohl; }
      }
      if (p.tgt) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, o);
        if (lane == 0) atomicAdd(p.loss + g, 0.5f * ls);
      }
    }
    __syncthreads();                                       // G and the accumulators are free for the next tile
  }
  if (p.dpred_amax) {
    amax = __reduce_max_sync(0xffffffffu, amax);
    if (lane == 0 && amax) atomicMax(p.dpred_amax, amax);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

// filter [kh,kw,Cn,Ck] fp32 -> image [plane][block][KP rows x 128 B] (row m = r * 16 + s * Cn + co, K = 64 channels of the block)
__global__ void few_transpose_pack_kernel(const float* __restrict__ w, int kh, int kw, int Cn, int Ck, uint4* __restrict__ img) {
  const int KP = kh * 16, nblk = Ck / 64;
  const int chunks = nblk * KP * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += gridDim.x * blockDim.x) {
    const int blk = i / (KP * 8), rem = i - blk * KP * 8;
    const int m = rem >> 3, c = rem & 7;
    const int r = m >> 4, j = m & 15;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e)
      v[e] = j < kw * Cn ? w[((int64_t)r * kw * Cn + j) * Ck + blk * 64 + c * 8 + e] * kFewWScale : 0.f;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      hi[e] = few::pack_h2(v[2 * e], v[2 * e + 1]);
      lo[e] = few::pack_h2(v[2 * e] - few::h2f((uint16_t)(hi[e] & 0xFFFFu)), v[2 * e + 1] - few::h2f((uint16_t)(hi[e] >> 16)));
    }
    const int64_t o = ((int64_t)blk * KP + m) * 8 + (c ^ (m & 7));
    img[o] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    img[(int64_t)nblk * KP * 8 + o] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

static int few_transpose_plan(int C1, int C2, int Cn, const ctx_conv_geom& g, FewTParams* p) {
  if (C1 <= 0 || C1 % 64 != 0 || C2 % 64 != 0 || C2 < 0) return 0;
  if (Cn <= 0 || g.kw * Cn > 16 || g.kh > 8) return 0;
  if (g.stride != 1 && g.stride != 2) return 0;
  const int st = g.stride;
  const int halo_top = (g.kh - 1 - g.pad_h) / st;           // input rows above the owned ones whose patches reach them
  const int halo_bot = g.pad_h >= 1 ? (g.pad_h - 1) / st + 1 : 0;
  if (g.kh - 1 - g.pad_h < 0) return 0;
  int RT = 192 / g.Q;
  if (RT > g.P + halo_top + halo_bot) RT = g.P + halo_top + halo_bot;
  const int BPo = RT - halo_top - halo_bot;
  if (g.Q <= 0 || BPo < 1 || (RT * g.Q) % 8 != 0 || RT > 256 || g.Q > 256) return 0;
  const int nblk = (C1 + C2) / 64, KP = g.kh * 16;
  const size_t smem = (size_t)nblk * 2 * RT * g.Q * 128 + 2 * (size_t)nblk * KP * 128 + (size_t)RT * g.Q * (KP + 1) * 4;
  if (smem > 220 * 1024) return 0;
  if (((RT * g.Q + 127) / 128) * 2 * KP > 512) return 0;
  if (p) {
    p->RT = RT; p->BPo = BPo; p->halo_top = halo_top; p->KP = KP; p->GS = KP + 1;
    p->nblk1 = C1 / 64; p->nblk = nblk;
    p->tiles_per_img = (int)cdiv(g.P, BPo);
  }
  return 1;
}

// ---- host ---------------------------------------------------------------------------------------------------------------------
static bool few_tile_grid(int P, int Q, int* BP, int* NB, int* tiles_per_img) {
  if (Q <= 0 || P <= 0 || Q > 128 || 128 % Q != 0) return false;
  const int per_img = P * Q;
  if (per_img >= 128) {
    const int bp = 128 / Q;
    if (P % bp != 0) return false;
    *BP = bp; *NB = 1; *tiles_per_img = P / bp;
  } else {
    if (128 % per_img != 0) return false;
    *BP = P; *NB = 128 / per_img; *tiles_per_img = 1;
  }
  return true;
}

static int few_fill(FewParams& p, int C, const ctx_conv_geom& g, const char* who) {
  memset(&p, 0, sizeof(p));
  if (!few_tile_grid(g.P, g.Q, &p.BP, &p.NB, &p.tiles_per_img)) { set_error("%s: the output grid does not tile into 128 pixels", who); return CTX_ERR_UNSUPPORTED; }
  p.N = g.N; p.H = g.H; p.W = g.W; p.C = C; p.P = g.P; p.Q = g.Q;
  p.kh = g.kh; p.kw = g.kw; p.stride = g.stride; p.pad_h = g.pad_h; p.pad_w = g.pad_w;
  p.num_tiles = p.NB > 1 ? (int)cdiv(g.N, p.NB) : g.N * p.tiles_per_img;
  p.RI = (p.BP - 1) * g.stride + g.kh;
  p.natoms = (g.kh * 16 + 63) / 64;
  p.RS = (uint32_t)((((g.Q - 1) * g.stride + g.kw) * C + 3) & ~3);       // staged row: padded width x channels, in halves
  p.st_elems = (uint32_t)(((size_t)p.NB * p.RI * p.RS + 7) & ~(size_t)7);
  return CTX_OK;
}

static int few_eligible(int C, int Co, const ctx_conv_geom& g) {
  int bp, nb, tpi;
  if (C <= 0 || g.kw * C > 16 || g.kh > 8) return 0;
  if ((g.W * C) % 4 != 0) return 0;                      // image rows move as 16-byte-granular bulk copies (uint8: % 16)
  if (Co != 64 && Co != 128) return 0;
  if (g.stride != 1 && g.stride != 2) return 0;
  if (!few_tile_grid(g.P, g.Q, &bp, &nb, &tpi)) return 0;
  if ((int64_t)g.N * g.P * g.Q >= (1ll << 31)) return 0;
  return 1;
}

// (cudaDeviceGetAttribute / cudaFuncSetAttribute cost tens to hundreds of microseconds of HOST time: once per kernel and device)
static int few_smem_attr(const void* fn, size_t bytes, const char* who) {
  static int optin_dev[64] = {0};
  struct Seen { const void* fn; int dev; size_t bytes; };
  static Seen seen[32];
  static int nseen = 0;
  int dev = 0;
  CTX_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { set_error("%s: device index %d out of range", who, dev); return CTX_ERR_ARG; }
  if (optin_dev[dev] == 0) CTX_CUDA(cudaDeviceGetAttribute(&optin_dev[dev], cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if ((int64_t)bytes > optin_dev[dev] - 2048) { set_error("%s: the tile does not fit shared memory (%zu bytes)", who, bytes); return CTX_ERR_UNSUPPORTED; }
  for (int i = 0; i < nseen; ++i)
    if (seen[i].fn == fn && seen[i].dev == dev) {
      if (seen[i].bytes >= bytes) return CTX_OK;
      CTX_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
      seen[i].bytes = bytes;
      return CTX_OK;
    }
  CTX_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  if (nseen < 32) { seen[nseen].fn = fn; seen[nseen].dev = dev; seen[nseen].bytes = bytes; ++nseen; }
  return CTX_OK;
}

}  // namespace ctx

using namespace ctx;

extern "C" int ctx_conv2d_few_eligible(int32_t C, int32_t Co, const ctx_conv_geom* g) {
  if (!g || g->N < 0 || g->P <= 0 || g->Q <= 0) return 0;
  return (ctx_tc_available() && few_eligible(C, Co, *g)) ? 1 : 0;
}

extern "C" int64_t ctx_conv2d_few_pack_bytes(int32_t Co, const ctx_conv_geom* g, int32_t nplanes) {
  if (!g || g->kh <= 0 || g->kh > 8 || Co <= 0) return 0;
  return (int64_t)nplanes * ((g->kh * 16 + 63) / 64) * Co * 128;
}

extern "C" int ctx_conv2d_few_pack(const float* w, int32_t C, int32_t Co, const ctx_conv_geom* g, void* wimg, int32_t nplanes,
                                   void* stream) {
  CTX_REQUIRE(w && wimg && g, "conv2d_few_pack: null pointer");
  CTX_REQUIRE(nplanes == 1 || nplanes == 2, "conv2d_few_pack: nplanes must be 1 or 2");
  CTX_REQUIRE(C > 0 && g->kw * C <= 16 && g->kh <= 8 && Co > 0, "conv2d_few_pack: kw * C must be <= 16 and kh <= 8");
  CTX_REQUIRE(((uintptr_t)wimg & 15) == 0, "conv2d_few_pack: wimg must be 16-byte aligned");
  const int natoms = (g->kh * 16 + 63) / 64;
  const int chunks = natoms * Co * 8;
  few_pack_kernel<<<(chunks + 127) / 128, 128, 0, (cudaStream_t)stream>>>(w, g->kh, g->kw, C, Co, natoms, nplanes, (uint4*)wimg);
  CTX_LAUNCH_CHECK("conv2d_few_pack");
  return CTX_OK;
}

extern "C" int ctx_conv2d_few_fprop(const void* x, int32_t x_is_u8, int32_t C, const void* wimg, float* y, void* y_planes,
                                    int32_t y_planes_n, int32_t Co, const ctx_conv_geom* g, const ctx_epilogue* epi,
                                    int32_t nprod, const uint32_t* x_amax_bits, uint32_t* amax_bits, void* stream) {
  CTX_REQUIRE(g && x && wimg && y, "conv2d_few_fprop: null pointer");
  CTX_REQUIRE(nprod == 1 || nprod == 3, "conv2d_few_fprop: nprod must be 1 or 3");
  CTX_REQUIRE(!y_planes || y_planes_n == 1 || y_planes_n == 2, "conv2d_few_fprop: y_planes_n must be 1 or 2");
  auto al = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
  CTX_REQUIRE(al(wimg) && al(y) && al(y_planes), "conv2d_few_fprop: pointers must be 16-byte aligned");
  if (g->N == 0) return CTX_OK;
  if (!ctx_tc_available() || !few_eligible(C, Co, *g)) {
    set_error("conv2d_few_fprop: shape not eligible (kw * C <= 16, kh <= 8, Co 64 or 128, 128-pixel tiling) or no sm_100 device");
    return CTX_ERR_UNSUPPORTED;
  }
  Epi e = make_epi(epi);
  CTX_REQUIRE(!e.act_ref && !e.accumulate && !e.round, "conv2d_few_fprop: epilogue supports bias and lrelu only");
  FewParams p;
  int rc = few_fill(p, C, *g, "conv2d_few_fprop");
  if (rc) return rc;
  const int np = nprod == 3 ? 2 : 1;
  p.x = x; p.x_u8 = x_is_u8; p.x_amax = x_amax_bits; p.wimg = (const uint4*)wimg;
  p.y = y; p.yp = y_planes; p.yp_n = y_planes ? y_planes_n : 0; p.plane_stride = (int64_t)g->N * g->P * g->Q * Co;
  p.bias = e.bias; p.slope = e.slope; p.act = e.act; p.amax_bits = amax_bits;
  p.Co = Co; p.np = np;
  {
    const int64_t rows = (int64_t)g->N * g->P * g->Q;
    int64_t dims[2] = {Co, rows}, str[2] = {1, Co};
    int box[2] = {32, 128};
    rc = h16_make_map(&p.mapY, 2, y, 2, dims, str, box);
    if (rc) return rc;
    if (y_planes) {
      int64_t dims3[3] = {Co, rows, y_planes_n}, str3[3] = {1, Co, rows * Co};
      int box3[3] = {64, 128, 1};
      rc = h16_make_map(&p.mapYp, 1, y_planes, 3, dims3, str3, box3);
      if (rc) return rc;
    }
  }
  CTX_REQUIRE(((uintptr_t)x & 15) == 0, "conv2d_few_fprop: x must be 16-byte aligned");
  if (x_is_u8 && (g->W * C) % 16 != 0) { set_error("conv2d_few_fprop: uint8 rows must be multiples of 16 bytes"); return CTX_ERR_UNSUPPORTED; }
  const size_t raw_bytes = ((size_t)p.NB * p.RI * g->W * C * (x_is_u8 ? 1 : 4) + 127) & ~(size_t)127;
  // (the A region doubles as the epilogue's staging area: 2 fp32 atoms + hi + lo atoms of a 64-channel group)
  const size_t a_bytes = std::max((size_t)np * p.natoms * kFAtom, (size_t)(y_planes ? 4 : 2) * kFAtom);
  p.a_region = (uint32_t)a_bytes;
  p.raw_bytes = (uint32_t)raw_bytes;
  p.grp_bytes = (uint32_t)((a_bytes + raw_bytes + (size_t)np * p.st_elems * 2 + 1023) & ~(size_t)1023);
  p.groups = kFGroups;
  size_t smem = 1024 + (size_t)np * p.natoms * Co * 128 + (size_t)p.groups * p.grp_bytes;
  if (smem > 224 * 1024 || (size_t)p.groups * np * Co > 512) {
    p.groups = 1;
    smem = 1024 + (size_t)np * p.natoms * Co * 128 + p.grp_bytes;
  }
  const void* fn = nprod == 3 ? (const void*)few_fprop_kernel<3> : (const void*)few_fprop_kernel<1>;
  rc = few_smem_attr(fn, smem, "conv2d_few_fprop");
  if (rc) return rc;
  int grid = sm_count();                                   // one CTA (its groups walk tiles independently) per SM
  if (grid > (p.num_tiles + p.groups - 1) / p.groups) grid = (p.num_tiles + p.groups - 1) / p.groups;
  if (nprod == 3) few_fprop_kernel<3><<<grid, kFThreads * p.groups, smem, (cudaStream_t)stream>>>(p);
  else few_fprop_kernel<1><<<grid, kFThreads * p.groups, smem, (cudaStream_t)stream>>>(p);
  CTX_LAUNCH_CHECK("conv2d_few_fprop");
  return CTX_OK;
}

extern "C" int ctx_conv2d_few_wgrad(const void* x, int32_t x_is_u8, int32_t C, const void* small1_plane, int32_t Cb1,
                                    const void* small2_plane, int32_t Cb2, int32_t small2_batch, float* dw,
                                    const ctx_conv_geom* g, const uint32_t* x_amax_bits, const float* r_scale, void* stream) {
  CTX_REQUIRE(g && x && small1_plane && dw, "conv2d_few_wgrad: null pointer");
  CTX_REQUIRE((Cb2 == 0) == (small2_plane == nullptr), "conv2d_few_wgrad: small2 and Cb2 must both be set or both be empty");
  CTX_REQUIRE(Cb2 == 0 || small2_batch > 0, "conv2d_few_wgrad: small2_batch must be positive");
  auto al = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
  CTX_REQUIRE(al(small1_plane) && al(small2_plane) && al(dw), "conv2d_few_wgrad: pointers must be 16-byte aligned");
  if (g->N == 0) return CTX_OK;
  const int Cb = Cb1 + Cb2;
  if (!ctx_tc_available() || !few_eligible(C, Cb, *g) || Cb1 % 64 != 0 || Cb2 % 64 != 0) {
    set_error("conv2d_few_wgrad: shape not eligible (kw * C <= 16, kh <= 8, Cb1 + Cb2 64 or 128 in 64-channel sources, "
              "128-pixel tiling) or no sm_100 device");
    return CTX_ERR_UNSUPPORTED;
  }
  FewParams p;
  int rc = few_fill(p, C, *g, "conv2d_few_wgrad");
  if (rc) return rc;
  p.x = x; p.x_u8 = x_is_u8; p.x_amax = x_amax_bits; p.r_scale = r_scale;
  p.y = dw; p.s1 = (const uint16_t*)small1_plane; p.s2 = (const uint16_t*)small2_plane;
  p.Cb1 = Cb1; p.Cb2 = Cb2; p.s2_batch = small2_batch > 0 ? small2_batch : 1;
  p.Co = Cb; p.np = 1;
  CTX_REQUIRE(((uintptr_t)x & 15) == 0, "conv2d_few_wgrad: x must be 16-byte aligned");
  if (x_is_u8 && (g->W * C) % 16 != 0) { set_error("conv2d_few_wgrad: uint8 rows must be multiples of 16 bytes"); return CTX_ERR_UNSUPPORTED; }
  if (Cb2 > 0 && p.NB > 1 && small2_batch % p.NB != 0) {
    set_error("conv2d_few_wgrad: a tile of %d whole images needs small2_batch %% %d == 0", p.NB, p.NB);
    return CTX_ERR_UNSUPPORTED;
  }
  {
    const int64_t rows = (int64_t)g->N * g->P * g->Q;
    int64_t dims[2] = {Cb1, rows}, str[2] = {1, Cb1};
    int box[2] = {64, 128};
    rc = h16_make_map(&p.mapB[0], 1, small1_plane, 2, dims, str, box);
    if (rc) return rc;
  }
  if (Cb2 > 0) {
    const int64_t rows = (int64_t)small2_batch * g->P * g->Q;
    int64_t dims[2] = {Cb2, rows}, str[2] = {1, Cb2};
    int box[2] = {64, 128};
    rc = h16_make_map(&p.mapB[1], 1, small2_plane, 2, dims, str, box);
    if (rc) return rc;
  }
  const size_t raw_bytes = ((size_t)p.NB * p.RI * g->W * C * (x_is_u8 ? 1 : 4) + 127) & ~(size_t)127;
  p.raw_bytes = (uint32_t)raw_bytes;
  const size_t smem = 1024 + 2 * (size_t)kFAtom + (size_t)(Cb / 64) * kFAtom + 2 * raw_bytes + (size_t)p.st_elems * 2;
  rc = few_smem_attr((const void*)few_wgrad_kernel, smem, "conv2d_few_wgrad");
  if (rc) return rc;
  int occ = (int)(224 * 1024 / (smem + 1024));
  if (occ < 1) occ = 1;
  if (occ > 3) occ = 3;
  int grid = sm_count() * occ;
  if (grid > p.num_tiles) grid = p.num_tiles;
  p.tiles_per_cta = (int)cdiv(p.num_tiles, grid);
  grid = (int)cdiv(p.num_tiles, p.tiles_per_cta);
  few_wgrad_kernel<<<grid, kFThreads, smem, (cudaStream_t)stream>>>(p);
  CTX_LAUNCH_CHECK("conv2d_few_wgrad");
  return CTX_OK;
}

extern "C" int ctx_conv2d_transpose_few_eligible(int32_t C1, int32_t C2, int32_t x2_batch, int32_t Cn, const ctx_conv_geom* g) {
  if (!g || g->N < 0 || g->P <= 0 || g->Q <= 0) return 0;
  if (C2 > 0 && x2_batch <= 0) return 0;
  return (ctx_tc_available() && few_transpose_plan(C1, C2, Cn, *g, nullptr)) ? 1 : 0;
}

extern "C" int64_t ctx_conv2d_transpose_few_pack_bytes(int32_t Ck, const ctx_conv_geom* g) {
  if (!g || g->kh <= 0 || g->kh > 8 || Ck <= 0 || Ck % 64 != 0) return 0;
  return 2ll * (Ck / 64) * g->kh * 16 * 128;
}

extern "C" int ctx_conv2d_transpose_few_pack(const float* w, int32_t Ck, int32_t Cn, const ctx_conv_geom* g, void* wimg,
                                             void* stream) {
  CTX_REQUIRE(w && wimg && g, "conv2d_transpose_few_pack: null pointer");
  CTX_REQUIRE(Ck > 0 && Ck % 64 == 0 && Cn > 0 && g->kw * Cn <= 16 && g->kh <= 8,
              "conv2d_transpose_few_pack: Ck must be a multiple of 64, kw * Cn <= 16, kh <= 8");
  CTX_REQUIRE(((uintptr_t)wimg & 15) == 0, "conv2d_transpose_few_pack: wimg must be 16-byte aligned");
  const int chunks = (Ck / 64) * g->kh * 16 * 8;
  few_transpose_pack_kernel<<<(chunks + 127) / 128, 128, 0, (cudaStream_t)stream>>>(w, g->kh, g->kw, Cn, Ck, (uint4*)wimg);
  CTX_LAUNCH_CHECK("conv2d_transpose_few_pack");
  return CTX_OK;
}

extern "C" int ctx_conv2d_transpose_few(const void* x1_planes, int32_t C1, const void* x2_planes, int32_t C2, int32_t x2_batch,
                                        const void* wimg, float* y, int32_t Cn, const ctx_conv_geom* g,
                                        const ctx_epilogue* epi, const ctx_l2_fused* loss, void* stream) {
  CTX_REQUIRE(g && x1_planes && wimg && y, "conv2d_transpose_few: null pointer");
  CTX_REQUIRE((C2 == 0) == (x2_planes == nullptr), "conv2d_transpose_few: x2 and C2 must both be set or both be empty");
  CTX_REQUIRE(C2 == 0 || x2_batch > 0, "conv2d_transpose_few: x2_batch must be positive");
  auto al = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
  CTX_REQUIRE(al(x1_planes) && al(x2_planes) && al(wimg), "conv2d_transpose_few: pointers must be 16-byte aligned");
  if (g->N == 0) return CTX_OK;
  FewTParams p;
  memset(&p, 0, sizeof(p));
  if (!ctx_tc_available() || !few_transpose_plan(C1, C2, Cn, *g, &p)) {
    set_error("conv2d_transpose_few: shape not eligible (channels %% 64, kw * Cn <= 16, kh <= 8, tile fits shared memory) or no "
              "sm_100 device");
    return CTX_ERR_UNSUPPORTED;
  }
  Epi e = make_epi(epi);
  CTX_REQUIRE(!e.act_ref && !e.accumulate && !e.round, "conv2d_transpose_few: epilogue supports bias and lrelu only");
  p.wimg = (const uint4*)wimg; p.y = y; p.bias = e.bias; p.slope = e.slope; p.act = e.act;
  p.N = g->N; p.H = g->H; p.W = g->W; p.P = g->P; p.Q = g->Q; p.Cn = Cn; p.kh = g->kh; p.kw = g->kw;
  p.stride = g->stride; p.pad_h = g->pad_h; p.pad_w = g->pad_w; p.x2_batch = x2_batch > 0 ? x2_batch : 1;
  p.num_tiles = g->N * p.tiles_per_img;
  if (loss && loss->tgt) {
    CTX_REQUIRE(loss->loss && loss->pass_images > 0, "conv2d_transpose_few: the fused loss needs loss sums and pass_images");
    p.tgt = loss->tgt; p.resid = loss->resid; p.weight = loss->weight; p.loss = loss->loss; p.dpred = loss->dpred;
    p.dpred_amax = loss->dpred_amax; p.pass_images = loss->pass_images;
  }
  int rc;
  {
    const int64_t plane = (int64_t)g->N * g->P * g->Q * C1;
    int64_t dims[5] = {C1, g->Q, g->P, g->N, 2}, str[5] = {1, C1, (int64_t)g->Q * C1, (int64_t)g->P * g->Q * C1, plane};
    int box[5] = {64, g->Q, p.RT, 1, 2};
    rc = h16_make_map(&p.map[0], 1, x1_planes, 5, dims, str, box);
    if (rc) return rc;
  }
  if (C2 > 0) {
    const int64_t plane = (int64_t)x2_batch * g->P * g->Q * C2;
    int64_t dims[5] = {C2, g->Q, g->P, x2_batch, 2}, str[5] = {1, C2, (int64_t)g->Q * C2, (int64_t)g->P * g->Q * C2, plane};
    int box[5] = {64, g->Q, p.RT, 1, 2};
    rc = h16_make_map(&p.map[1], 1, x2_planes, 5, dims, str, box);
    if (rc) return rc;
  }
  const size_t smem = 1024 + (size_t)p.nblk * 2 * p.RT * g->Q * 128 + 2 * (size_t)p.nblk * p.KP * 128 +
                      (size_t)p.RT * g->Q * p.GS * 4;
  rc = few_smem_attr(g->stride == 2 ? (const void*)few_transpose_kernel<2> : (const void*)few_transpose_kernel<1>, smem,
                     "conv2d_transpose_few");
  if (rc) return rc;
  int grid = sm_count();
  if (grid > p.num_tiles) grid = p.num_tiles;
  if (g->stride == 2) few_transpose_kernel<2><<<grid, kFThreads, smem, (cudaStream_t)stream>>>(p);
  else few_transpose_kernel<1><<<grid, kFThreads, smem, (cudaStream_t)stream>>>(p);
  CTX_LAUNCH_CHECK("conv2d_transpose_few");
  return CTX_OK;
}
