// Forward convolutions from PRE-SPLIT bf16 planes on tcgen05 (kind::f16), fp32-grade:
//
//     a = hi + lo  (both bf16, |a - hi - lo| <= 2^-16 |a|),   D = A_hi*B_hi + (A_hi*B_lo + A_lo*B_hi),  fp32 accumulate in TMEM
//
// The 3-pass TF32 kernel (tc_gemm.cu) splits every fp32 operand tile inside the kernel, once per (tile, tap, consumer); here
// the producer of a tensor writes its (hi, lo) planes once — the epilogue below, ctx_split_bf16 for tensors made by other
// kernels, ctx_split_bf16_transpose for the weights — and the GEMM is a plain TMA -> shared memory -> MMA pipeline with no
// splitter warps.  One k-block = 64 channels (128-byte swizzle rows of bf16): ONE tensor-map box per operand fetches both
// planes (the plane is the outermost tensor-map dimension), 12 MMAs of K = 16 consume them.  Per product the tensor core
// spends 3 bf16 MMAs = 1.5 TF32-MMA times instead of 3, and shared memory serves half the operand bytes.
// Measured on the bench's layer shapes (tools/bf16x3_probe.cu, profiles/r1_bf16x3_probe.txt): 290-490 TFLOP/s algorithmic
// against 200-240 for the 3-pass TF32 tiles, max-norm error 5e-6..9e-6 per GEMM (K = 1600..6400).
//
// Kernel bf3_kernel<pair>: 320 threads — warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issue (one lane),
// warps 2..9 = epilogue (two warps per TMEM lane quarter): D1 + D2 -> bias -> lrelu -> fp32 store (+ bf16 (hi, lo) planes of
// the result for the next layer).  pair = cta_group::2: a cluster of two CTAs along M runs one 256 x BN MMA per step; each
// CTA loads its own 128 rows of A and HALF of the B tile (the probe: 1.45x over single CTAs at BN = 256, where the single
// CTA is bound by L2 -> SM operand traffic).
//
// Geometry is the conv-tap machinery of tc_gemm.cu restated for rank-5 bf16 maps {64 ch, Q, P, N, plane}: conv fprop
// (stride 1 / 2 through parity views) and transposed conv (stride phases merged into one launch, two sources = fused
// concat, the second source read at image n % batch).  Exact 128-pixel tilings only; everything else stays on tc_gemm.cu.
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>
#include "kernels.h"

namespace ctx {

constexpr int kBfMaxSeg = 64;
constexpr int kBfThreads = 320;
constexpr uint32_t kBfATile = 128 * 128;           // 128 pixels x 64 bf16

struct BfSeg {                                     // one K segment = one filter tap of one source tensor
  int16_t dh, dw;
  uint8_t mapA, wrap;
  uint16_t nblk;                                   // 64-channel blocks
  int32_t bk0, br0;                                // B: K coordinate of the first block; row offset added to the tile's n0
};

struct BfParams {
  CUtensorMap mapA[8];
  CUtensorMap mapB;
  BfSeg seg[kBfMaxSeg];
  int32_t BN, stages;
  int32_t nphase, ph_seg0[4], ph_nseg[4], ph_kblocks[4];
  int64_t ph_obase[4];
  int32_t NG, PG, QG, BP, NB, tiles_per_img, a_wrap, N;
  float* out;
  int64_t sN, sP, sQ;
  __nv_bfloat16* planes;                           // optional (hi, lo) copy of the output, indexed like `out`
  int64_t plane_stride;
  const float* bias;
  float slope;
  int32_t act;
  // phase-merged transposed conv: column j = phase * merge_cn + co, phase = ph * 2 + pw, goes to output pixel
  // (2p + ph, 2q + pw) = out + roff + (ph * merge_W + pw) * merge_cn + co (0 = plain column addressing)
  int32_t merge_cn, merge_W;
  int32_t* err_flag;
};

// ---- device helpers (raw PTX) ------------------------------------------------------------------------------------------
namespace bf {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
// Bounded waits: a protocol bug must trap (launch failure on the host), never hang the GPU.
template <bool kCluster>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int32_t* err_flag, int id) {
  auto probe = [&]() { return kCluster ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity); };
  if (probe()) return;
  const long long t0 = clock64();
  while (!probe()) {
    if (clock64() - t0 > 4000000000LL) {
      if (err_flag) atomicExch(err_flag, 0x2000 + id);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <bool kPair>
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if constexpr (kPair)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
template <bool kPair>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if constexpr (kPair)      // arrives on the barrier at this offset in BOTH CTAs of the pair
    asm volatile("{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
                 "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
                 ::"r"(bar) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// relaxed: the arrival only says "my stage has landed"; the data stays in this CTA's shared memory and is read by the
// pair's MMA datapath (see tc_gemm.cu mbar_arrive_remote for the measurement behind relaxed vs release)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
               "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(rank) : "memory");
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
// K-major SWIZZLE_128B operand tile: rows of 128 B (64 bf16 along K), 8-row groups 1024 B apart (SM100 descriptor v1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(16u >> 4) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
}  // namespace bf

// ---- the kernel -----------------------------------------------------------------------------------------------------------
template <bool kPair>
__global__ void __launch_bounds__(kBfThreads, 1) bf3_kernel(const __grid_constant__ BfParams p) {
  using namespace bf;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank = 0u;
  if constexpr (kPair) rank = cluster_ctarank();
  const uint32_t b_tile = (uint32_t)(kPair ? p.BN / 2 : p.BN) * 128u;     // this CTA's share of the B tile, per plane
  const uint32_t stage_bytes = 2u * (kBfATile + b_tile);                   // [A_hi | A_lo | B_hi | B_lo]
  const uint32_t b_off = 2u * kBfATile;
  __shared__ __align__(8) uint64_t bars[25];
  __shared__ uint32_t tmem_slot;
  const int S = p.stages;
  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[8 + s]); };
  auto peer_bar = [&](int s) { return smem_u32(&bars[16 + s]); };       // pair leader: the follower's stage has landed
  const uint32_t acc_bar = smem_u32(&bars[24]);

  const int z = blockIdx.z;                                              // stride phase of a transposed conv (else 0)
  const int seg0 = p.ph_seg0[z], nseg = p.ph_nseg[z], nkb = p.ph_kblocks[z];
  const int64_t obase = p.ph_obase[z];
  const int tile_m = blockIdx.x;
  int n0, p0;
  if (p.NB > 1 || p.tiles_per_img == 1) { n0 = tile_m * p.NB; p0 = 0; }
  else { n0 = tile_m / p.tiles_per_img; p0 = (tile_m - n0 * p.tiles_per_img) * p.BP; }
  const int ncol0 = blockIdx.y * p.BN;
  uint32_t ncols = 32;
  while (ncols < 2u * (uint32_t)p.BN) ncols <<= 1;                       // two accumulators: hi*hi | cross terms

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); mbar_init(peer_bar(s), 1); }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(ncols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(ncols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();          // both CTAs' barriers exist before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer ===================================================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int nb0 = ncol0 + (int)rank * (p.BN / 2);                    // pair: this CTA's half of the B tile
      for (int sg = 0; sg < nseg; ++sg) {
        const BfSeg s = p.seg[seg0 + sg];
        const int nimg = s.wrap ? (n0 % p.a_wrap) : n0;
        for (int cb = 0; cb < s.nblk; ++cb) {
          mbar_wait<false>(empty_bar(stage), phase ^ 1, p.err_flag, 1);
          const uint32_t sa = smem_base + stage * stage_bytes;
          mbar_arrive_expect_tx(full_bar(stage), stage_bytes);
          tma_load_5d(sa, &p.mapA[s.mapA], full_bar(stage), cb * 64, s.dw, p0 + s.dh, nimg, 0);
          tma_load_3d(sa + b_off, &p.mapB, full_bar(stage), s.bk0 + cb * 64, s.br0 + nb0, 0);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issue =======================================================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if (kPair && rank != 0) {
        for (int kb = 0; kb < nkb; ++kb) {              // follower: forward "my stage has landed" to the leader
          mbar_wait<false>(full_bar(stage), phase, p.err_flag, 4);
          mbar_arrive_remote(peer_bar(stage), 0u);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      } else {
        // kind::f16: D = f32 (1 << 4), A = B = bf16 (1 << 7, 1 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                               (((kPair ? 256u : 128u) >> 4) << 24);
        // two accumulators, as in the TF32 kernel: the tensor core truncates when it adds into a running fp32 sum, so the
        // small cross terms must not ride on the big one; D1 + D2 is formed once, in the epilogue
        const uint32_t d1 = tmem_base, d2 = tmem_base + (uint32_t)p.BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait<false>(full_bar(stage), phase, p.err_flag, 2);
          if constexpr (kPair) mbar_wait<true>(peer_bar(stage), phase, p.err_flag, 5);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + kBfATile);
          const uint64_t b_hi = make_desc(sa + b_off), b_lo = make_desc(sa + b_off + b_tile);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {              // K = 16 per MMA = 32 B = 2 sixteen-byte units along the swizzle row
            const uint32_t first = (kb | ks) ? 1u : 0u;
            umma_bf16<kPair>(d1, a_hi + 2u * ks, b_hi + 2u * ks, idesc, first);
            umma_bf16<kPair>(d2, a_hi + 2u * ks, b_lo + 2u * ks, idesc, first);
            umma_bf16<kPair>(d2, a_lo + 2u * ks, b_hi + 2u * ks, idesc, 1u);
          }
          umma_commit<kPair>(empty_bar(stage));         // frees the stage (in both CTAs of a pair) once the MMAs have read it
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        umma_commit<kPair>(acc_bar);
      }
    }
  } else {
    // =============================== epilogue: warps 2..9 ============================================================
    const int q = warp & 3;                              // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;                         // tile row
    const int half = (warp - 2) >> 2;                    // the two warps of a quarter split the columns
    const int c_begin = half ? p.BN / 2 : 0, c_end = half ? p.BN : p.BN / 2;
    const int per_img = p.BP * p.QG;
    const int dn = r / per_img, rem = r - dn * per_img;
    const int dp = rem / p.QG, dq = rem - dp * p.QG;
    const int gn = n0 + dn, gp = p0 + dp;
    const bool row_ok = gn < p.NG && gp < p.PG;
    const int64_t roff = obase + (int64_t)gn * p.sN + (int64_t)gp * p.sP + (int64_t)dq * p.sQ;
    mbar_wait<false>(acc_bar, 0, p.err_flag, 3);
    tc_fence_after();
    for (int c0 = c_begin; c0 < c_end; c0 += 32) {
      uint32_t v[32], v2[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.BN + c0), v2);
      const int col = ncol0 + c0;
      if (!row_ok || col >= p.N) continue;
      int co = col;                                      // output channel of the chunk's first column
      int64_t ooff = roff + col;
      if (p.merge_cn) {                                  // (warp-uniform: a 32-column chunk never straddles phases)
        const int phase = col / p.merge_cn;
        co = col - phase * p.merge_cn;
        ooff = roff + ((int64_t)(phase >> 1) * p.merge_W + (phase & 1)) * p.merge_cn + co;
      }
      float x[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float t = __uint_as_float(v[j]) + __uint_as_float(v2[j]);
        if (p.bias) t += __ldg(p.bias + co + j);
        if (p.act) t = fmaxf(t, p.slope * t);
        x[j] = t;
      }
      float* optr = p.out + ooff;
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(optr + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
      if (p.planes) {
        // (hi, lo) planes of the result: hi = bf16_rn(x), lo = bf16_rn(x - hi)  (x - hi is exact in fp32)
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * j]), h1 = __float2bfloat16_rn(x[2 * j + 1]);
          hi[j] = pack_bf16x2(h0, h1);
          lo[j] = pack_bf16x2(__float2bfloat16_rn(x[2 * j] - __bfloat162float(h0)),
                              __float2bfloat16_rn(x[2 * j + 1] - __bfloat162float(h1)));
        }
        __nv_bfloat16* ph = p.planes + ooff;
        __nv_bfloat16* pl = ph + p.plane_stride;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          *reinterpret_cast<uint4*>(ph + 2 * j) = make_uint4(hi[j], hi[j + 1], hi[j + 2], hi[j + 3]);
          *reinterpret_cast<uint4*>(pl + 2 * j) = make_uint4(lo[j], lo[j + 1], lo[j + 2], lo[j + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();          // the peer may still be reading this CTA's shared / tensor memory
  if (warp == 1) {
    tc_fence_after();
    if constexpr (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

// ---- host ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*BfEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static BfEncodeFn bf_get_encode() {
  static BfEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (BfEncodeFn)f;
    else
      cudaGetLastError();
  }
  return fn;
}

// bf16 tensor map, zero fill, 128-byte swizzle; dims / box innermost first; strides in ELEMENTS for dims 1..rank-1
static int bf_make_map(CUtensorMap* out, const void* base, int rank, const int64_t* dims, const int64_t* strides,
                       const int* box) {
  BfEncodeFn enc = bf_get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return CTX_ERR_CUDA; }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = (cuuint64_t)dims[i]; bx[i] = (cuuint32_t)box[i]; es[i] = 1;
    if (dims[i] <= 0 || box[i] <= 0 || box[i] > 256) { set_error("bf16 tensor map: bad dim/box %d: %lld/%d", i, (long long)dims[i], box[i]); return CTX_ERR_ARG; }
  }
  for (int i = 1; i < rank; ++i) {
    gs[i - 1] = (cuuint64_t)strides[i] * 2u;
    if (gs[i - 1] % 16 != 0) { set_error("bf16 tensor map: stride %d (%lld elements) is not 16-byte aligned", i, (long long)strides[i]); return CTX_ERR_ARG; }
  }
  if (((uintptr_t)base & 15) != 0) { set_error("bf16 tensor map: base pointer is not 16-byte aligned"); return CTX_ERR_ARG; }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, (void*)base, gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (bf16) failed with CUresult %d", (int)r); return CTX_ERR_CUDA; }
  return CTX_OK;
}

// rows = 128 pixels: BP rows x Q columns of one image (P % BP == 0), or NB whole images
static bool bf_tile_grid(int P, int Q, int* BP, int* NB, int* tiles_per_img) {
  const int rows = 128;
  if (Q <= 0 || P <= 0 || Q > rows || rows % Q != 0) return false;
  const int per_img = P * Q;
  if (per_img >= rows) {
    const int bp = rows / Q;
    if (P % bp != 0) return false;
    *BP = bp; *NB = 1; *tiles_per_img = P / bp;
  } else {
    if (rows % per_img != 0) return false;
    *BP = P; *NB = rows / per_img; *tiles_per_img = 1;
  }
  return true;
}

static int bf_pick_bn(int N) { return N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : 64); }

static int bf_launch(BfParams& p, int m_tiles, int phases, cudaStream_t stream, const char* name) {
  static int max_dyn_dev[64] = {0};
  int dev = 0;
  CTX_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { set_error("%s: device index %d out of range", name, dev); return CTX_ERR_ARG; }
  if (max_dyn_dev[dev] == 0) {
    cudaFuncAttributes fa[2];
    CTX_CUDA(cudaFuncGetAttributes(&fa[0], bf3_kernel<false>));
    CTX_CUDA(cudaFuncGetAttributes(&fa[1], bf3_kernel<true>));
    int optin = 0;
    CTX_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t stat = fa[0].sharedSizeBytes > fa[1].sharedSizeBytes ? fa[0].sharedSizeBytes : fa[1].sharedSizeBytes;
    const int dyn = optin - (int)stat;
    CTX_CUDA(cudaFuncSetAttribute(bf3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    CTX_CUDA(cudaFuncSetAttribute(bf3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    max_dyn_dev[dev] = dyn;
  }
  const bool pair = p.BN == 256;                     // 256-wide tiles run as CTA pairs (each CTA stages half of B)
  const uint32_t stage = 2u * (kBfATile + (uint32_t)(pair ? p.BN / 2 : p.BN) * 128u);
  int stages = (int)(((uint32_t)max_dyn_dev[dev] - 1024u) / stage);
  if (stages > 8) stages = 8;
  if (stages < 2) { set_error("%s: tile does not fit shared memory", name); return CTX_ERR_ARG; }
  p.stages = stages;
  p.err_flag = nullptr;
  const size_t smem = (size_t)stages * stage + 1024;
  const int n_tiles = (int)cdiv(p.N, p.BN);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(pair ? (m_tiles + 1) & ~1 : m_tiles), (unsigned)n_tiles, (unsigned)phases);
  cfg.blockDim = dim3(kBfThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cudaError_t e;
  if (pair) {
    // cluster of 2 along the M tiles; an odd last tile gets an all-out-of-range partner (TMA zero-fills, epilogue masks)
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, bf3_kernel<true>, p);
  } else {
    e = cudaLaunchKernelEx(&cfg, bf3_kernel<false>, p);
  }
  if (e != cudaSuccess) return cuda_fail(e, name);
  count_launch();
  return CTX_OK;
}

static void bf_fill_common(BfParams& p, const Epi& epi, float* y, __nv_bfloat16* yp, int64_t y_plane) {
  memset(&p, 0, sizeof(p));
  p.nphase = 1;
  p.a_wrap = 1;
  p.out = y;
  p.planes = yp;
  p.plane_stride = y_plane;
  p.bias = epi.bias; p.slope = epi.slope; p.act = epi.act;
}

static bool bf_aligned16(const void* ptr) { return ((uintptr_t)ptr & 15) == 0; }

// ---- conv fprop -----------------------------------------------------------------------------------------------------------
int bf3_conv2d_fprop_eligible(int C, int Co, const ctx_conv_geom& g) {
  int bp, nb, tpi;
  if (C % 64 != 0 || Co % 64 != 0) return 0;
  if (g.stride != 1 && g.stride != 2) return 0;
  if (g.kh * g.kw > kBfMaxSeg) return 0;
  if (!bf_tile_grid(g.P, g.Q, &bp, &nb, &tpi)) return 0;
  if (g.stride == 2 && (g.H < 2 || g.W < 2)) return 0;          // empty parity views stay on the TF32 kernel
  return 1;
}

// xp: planes [2][N,H,W,C]; wp: planes [2][kh*kw*Co][C] (per-tap K-major, ctx_split_bf16_transpose); y fp32 [N,P,Q,Co];
// yp: optional planes [2][N,P,Q,Co] of the result
int bf3_conv2d_fprop(const __nv_bfloat16* xp, int C, const __nv_bfloat16* wp, float* y, __nv_bfloat16* yp, int Co,
                     const ctx_conv_geom& g, const Epi& epi, cudaStream_t stream) {
  CTX_REQUIRE(bf_aligned16(xp) && bf_aligned16(wp) && bf_aligned16(y) && bf_aligned16(yp),
              "conv2d_fprop(bf16x3): pointers must be 16-byte aligned");
  CTX_REQUIRE(!epi.act_ref && !epi.accumulate && !epi.round, "conv2d_fprop(bf16x3): only bias / lrelu epilogues");
  BfParams p;
  bf_fill_common(p, epi, y, yp, (int64_t)g.N * g.P * g.Q * Co);
  int tpi;
  if (!bf_tile_grid(g.P, g.Q, &p.BP, &p.NB, &tpi)) { set_error("conv2d_fprop(bf16x3): grid does not tile"); return CTX_ERR_UNSUPPORTED; }
  p.tiles_per_img = tpi;
  const int m_tiles = p.NB > 1 ? (int)cdiv(g.N, p.NB) : g.N * tpi;
  p.BN = bf_pick_bn(Co);
  p.NG = g.N; p.PG = g.P; p.QG = g.Q; p.N = Co;
  const int64_t x_plane = (int64_t)g.N * g.H * g.W * C;
  if (g.stride == 1) {
    int64_t dims[5] = {C, g.W, g.H, g.N, 2}, str[5] = {1, C, (int64_t)g.W * C, (int64_t)g.H * g.W * C, x_plane};
    int box[5] = {64, g.Q, p.BP, p.NB, 2};
    int rc = bf_make_map(&p.mapA[0], xp, 5, dims, str, box);
    if (rc) return rc;
  } else {
    for (int hp = 0; hp < 2; ++hp)
      for (int wq = 0; wq < 2; ++wq) {
        const int Hh = (g.H - hp + 1) / 2, Wh = (g.W - wq + 1) / 2;
        int64_t dims[5] = {C, Wh, Hh, g.N, 2};
        int64_t str[5] = {1, 2 * (int64_t)C, 2 * (int64_t)g.W * C, (int64_t)g.H * g.W * C, x_plane};
        int box[5] = {64, g.Q, p.BP, p.NB, 2};
        int rc = bf_make_map(&p.mapA[hp * 2 + wq], xp + ((int64_t)hp * g.W + wq) * C, 5, dims, str, box);
        if (rc) return rc;
      }
  }
  {
    const int64_t rows = (int64_t)g.kh * g.kw * Co;
    int64_t dims[3] = {C, rows, 2}, str[3] = {1, C, rows * C};
    int box[3] = {64, p.BN == 256 ? 128 : p.BN, 2};
    int rc = bf_make_map(&p.mapB, wp, 3, dims, str, box);
    if (rc) return rc;
  }
  const int cblk = C / 64;
  int nseg = 0;
  for (int r = 0; r < g.kh; ++r)
    for (int s = 0; s < g.kw; ++s) {
      BfSeg& sg = p.seg[nseg++];
      const int rr = r - g.pad_h, ss = s - g.pad_w;
      if (g.stride == 1) { sg.mapA = 0; sg.dh = (int16_t)rr; sg.dw = (int16_t)ss; }
      else {
        const int hp = rr & 1, wq = ss & 1;                      // two's complement: correct for negatives
        sg.mapA = (uint8_t)(hp * 2 + wq); sg.dh = (int16_t)((rr - hp) / 2); sg.dw = (int16_t)((ss - wq) / 2);
      }
      sg.wrap = 0; sg.nblk = (uint16_t)cblk; sg.bk0 = 0; sg.br0 = (r * g.kw + s) * Co;
    }
  p.ph_seg0[0] = 0; p.ph_nseg[0] = nseg; p.ph_kblocks[0] = nseg * cblk; p.ph_obase[0] = 0;
  p.sN = (int64_t)g.P * g.Q * Co; p.sP = (int64_t)g.Q * Co; p.sQ = Co;
  return bf_launch(p, m_tiles, 1, stream, "bf3_conv2d_fprop");
}

// ---- transposed conv (stride phases merged into one launch; two sources = fused concat) ---------------------------------------
static void bf_phase_geom(const ctx_conv_geom& g, int ph, int pw, int* Hp, int* Wp, int* r0, int* s0, int* nr, int* ns,
                          int* dh0, int* dw0) {
  const int st = g.stride;
  *Hp = g.H > ph ? (g.H - ph + st - 1) / st : 0;
  *Wp = g.W > pw ? (g.W - pw + st - 1) / st : 0;
  *r0 = (ph + g.pad_h) % st; *s0 = (pw + g.pad_w) % st;
  *nr = *r0 < g.kh ? (g.kh - *r0 + st - 1) / st : 0;
  *ns = *s0 < g.kw ? (g.kw - *s0 + st - 1) / st : 0;
  *dh0 = (ph + g.pad_h - *r0) / st; *dw0 = (pw + g.pad_w - *s0) / st;
}

int bf3_conv2d_transpose_eligible(int C1, int C2, int x2_batch, int Cn, const ctx_conv_geom& g) {
  if (C1 % 64 != 0 || C2 % 64 != 0 || Cn % 64 != 0) return 0;
  if (g.stride != 1 && g.stride != 2) return 0;
  int Hp0 = -1, Wp0 = -1, segs = 0;
  for (int ph = 0; ph < g.stride; ++ph)
    for (int pw = 0; pw < g.stride; ++pw) {
      int Hp, Wp, r0, s0, nr, ns, dh0, dw0, bp, nb, tpi;
      bf_phase_geom(g, ph, pw, &Hp, &Wp, &r0, &s0, &nr, &ns, &dh0, &dw0);
      if (Hp == 0 || Wp == 0 || nr == 0 || ns == 0) return 0;
      if (Hp0 < 0) { Hp0 = Hp; Wp0 = Wp; }
      if (Hp != Hp0 || Wp != Wp0) return 0;                       // all phases in ONE launch: same pixel grid
      if (!bf_tile_grid(Hp, Wp, &bp, &nb, &tpi)) return 0;
      if (C2 > 0 && nb > 1 && x2_batch % nb != 0) return 0;
      segs += nr * ns * (C2 > 0 ? 2 : 1);
    }
  return segs <= kBfMaxSeg ? 1 : 0;
}

// x1p planes [2][N,P,Q,C1]; x2p planes [2][x2_batch,P,Q,C2] or null; wp planes [2][kh*kw*Cn][C1+C2] (the TF filter layout
// [kh,kw,Cn,Ck] as it is); y fp32 [N,H,W,Cn]; yp optional planes [2][N,H,W,Cn]
int bf3_conv2d_transpose(const __nv_bfloat16* x1p, int C1, const __nv_bfloat16* x2p, int C2, int x2_batch,
                         const __nv_bfloat16* wp, float* y, __nv_bfloat16* yp, int Cn, const ctx_conv_geom& g,
                         const Epi& epi, cudaStream_t stream) {
  CTX_REQUIRE(bf_aligned16(x1p) && bf_aligned16(x2p) && bf_aligned16(wp) && bf_aligned16(y) && bf_aligned16(yp),
              "conv2d_transpose(bf16x3): pointers must be 16-byte aligned");
  CTX_REQUIRE(!epi.act_ref && !epi.accumulate && !epi.round, "conv2d_transpose(bf16x3): only bias / lrelu epilogues");
  const int st = g.stride, Ck = C1 + C2;
  BfParams p;
  bf_fill_common(p, epi, y, yp, (int64_t)g.N * g.H * g.W * Cn);
  struct Ph { int ph, pw, Hp, Wp, r0, s0, nr, ns, dh0, dw0, taps; };
  Ph phs[4];
  int nph = 0;
  for (int ph = 0; ph < st; ++ph)
    for (int pw = 0; pw < st; ++pw) {
      Ph q;
      q.ph = ph; q.pw = pw;
      bf_phase_geom(g, ph, pw, &q.Hp, &q.Wp, &q.r0, &q.s0, &q.nr, &q.ns, &q.dh0, &q.dw0);
      q.taps = q.nr * q.ns;
      phs[nph++] = q;
    }
  for (int i = 1; i < nph; ++i)                         // heaviest phase first
    for (int j = i; j > 0 && phs[j].taps > phs[j - 1].taps; --j) { Ph t = phs[j]; phs[j] = phs[j - 1]; phs[j - 1] = t; }
  int tpi;
  if (!bf_tile_grid(phs[0].Hp, phs[0].Wp, &p.BP, &p.NB, &tpi)) { set_error("conv2d_transpose(bf16x3): grid does not tile"); return CTX_ERR_UNSUPPORTED; }
  p.tiles_per_img = tpi;
  const int m_tiles = p.NB > 1 ? (int)cdiv(g.N, p.NB) : g.N * tpi;
  p.BN = bf_pick_bn(Cn);
  p.NG = g.N; p.PG = phs[0].Hp; p.QG = phs[0].Wp; p.N = Cn;
  p.a_wrap = x2_batch > 0 ? x2_batch : 1;
  {
    const int64_t plane = (int64_t)g.N * g.P * g.Q * C1;
    int64_t dims[5] = {C1, g.Q, g.P, g.N, 2}, str[5] = {1, C1, (int64_t)g.Q * C1, (int64_t)g.P * g.Q * C1, plane};
    int box[5] = {64, phs[0].Wp, p.BP, p.NB, 2};
    int rc = bf_make_map(&p.mapA[0], x1p, 5, dims, str, box);
    if (rc) return rc;
  }
  if (C2 > 0) {
    const int64_t plane = (int64_t)x2_batch * g.P * g.Q * C2;
    int64_t dims[5] = {C2, g.Q, g.P, x2_batch, 2}, str[5] = {1, C2, (int64_t)g.Q * C2, (int64_t)g.P * g.Q * C2, plane};
    int box[5] = {64, phs[0].Wp, p.BP, p.NB, 2};
    int rc = bf_make_map(&p.mapA[1], x2p, 5, dims, str, box);
    if (rc) return rc;
  }
  {
    const int64_t rows = (int64_t)g.kh * g.kw * Cn;
    int64_t dims[3] = {Ck, rows, 2}, str[3] = {1, Ck, rows * Ck};
    int box[3] = {64, p.BN == 256 ? 128 : p.BN, 2};
    int rc = bf_make_map(&p.mapB, wp, 3, dims, str, box);
    if (rc) return rc;
  }
  int seg0 = 0;
  for (int i = 0; i < nph; ++i) {
    const Ph& q = phs[i];
    int nseg = 0, kblocks = 0;
    for (int a = 0; a < q.nr; ++a)
      for (int b = 0; b < q.ns; ++b) {
        const int tap = (q.r0 + st * a) * g.kw + (q.s0 + st * b);
        for (int src = 0; src < (C2 > 0 ? 2 : 1); ++src) {
          if (seg0 + nseg >= kBfMaxSeg) { set_error("conv2d_transpose(bf16x3): too many taps"); return CTX_ERR_UNSUPPORTED; }
          BfSeg& sg = p.seg[seg0 + nseg++];
          sg.mapA = (uint8_t)src; sg.wrap = (uint8_t)src;
          sg.dh = (int16_t)(q.dh0 - a); sg.dw = (int16_t)(q.dw0 - b);
          sg.nblk = (uint16_t)((src ? C2 : C1) / 64);
          sg.bk0 = src ? C1 : 0; sg.br0 = tap * Cn;
          kblocks += sg.nblk;
        }
      }
    p.ph_seg0[i] = seg0; p.ph_nseg[i] = nseg; p.ph_kblocks[i] = kblocks;
    p.ph_obase[i] = ((int64_t)q.ph * g.W + q.pw) * Cn;
    seg0 += nseg;
  }
  p.nphase = nph;
  p.sN = (int64_t)g.H * g.W * Cn; p.sP = (int64_t)st * g.W * Cn; p.sQ = (int64_t)st * Cn;
  return bf_launch(p, m_tiles, nph, stream, "bf3_conv2d_transpose");
}

// ---- transposed conv, stride 2, few output channels (4 * Cn == 256): the four stride phases merged along N ------------------
// A 128 x 64 tile moves 48 KB of operands per k-block for 1/4 of the MACs of a 256-wide pair tile, and the per-phase
// formulation re-reads every input pixel once per filter tap (25x for 5x5).  Here ONE GEMM covers all four phases:
// rows = input pixels (p, q), columns = (phase, co), K = (distinct input offset (dh, dw), source, channel).  Phase
// (ph, pw) at offset (dh, dw) uses exactly one tap (r, s) or none; the packed filter matrix Wm[offset][phase * Cn + co][c]
// holds that tap's weights or zeros (ctx_conv2d_transpose_merged_pack).  The input is fetched once per distinct offset
// (9 for 5x5 instead of 25), every tile is a 256-wide CTA pair, and the zero blocks cost MMA time only (36 k-block
// columns issued for 25 needed).
struct MergedPlan {
  int nh, nw;                 // distinct row / column offsets
  int dh[3], dw[3];
  int tap_r[2][3], tap_s[2][3];   // [phase bit][offset index] -> filter row / column, or -1
};

static bool merged_plan(const ctx_conv_geom& g, MergedPlan* mp) {
  if (g.stride != 2 || g.H != 2 * g.P || g.W != 2 * g.Q) return false;
  // one axis at a time: phase bit b, taps r = r0 + 2a at input offset d = dh0 - a
  auto axis = [&](int k, int pad, int* n, int* d, int (*tap)[3]) {
    *n = 0;
    for (int b = 0; b < 2; ++b) for (int j = 0; j < 3; ++j) tap[b][j] = -1;
    for (int b = 0; b < 2; ++b) {
      const int r0 = (b + pad) % 2, d0 = (b + pad - r0) / 2;
      const int nr = r0 < k ? (k - r0 + 1) / 2 : 0;
      if (nr == 0) return false;
      for (int a = 0; a < nr; ++a) {
        const int off = d0 - a;
        int j = 0;
        while (j < *n && d[j] != off) ++j;
        if (j == *n) { if (*n == 3) return false; d[(*n)++] = off; }
        tap[b][j] = r0 + 2 * a;
      }
    }
    return true;
  };
  return axis(g.kh, g.pad_h, &mp->nh, mp->dh, mp->tap_r) && axis(g.kw, g.pad_w, &mp->nw, mp->dw, mp->tap_s);
}

int bf3_conv2d_transpose_merged_eligible(int C1, int C2, int x2_batch, int Cn, const ctx_conv_geom& g) {
  MergedPlan mp;
  int bp, nb, tpi;
  if (4 * Cn != 256 || C1 % 64 != 0 || C2 % 64 != 0) return 0;
  if (!merged_plan(g, &mp)) return 0;
  if (!bf_tile_grid(g.P, g.Q, &bp, &nb, &tpi)) return 0;
  if (C2 > 0 && nb > 1 && x2_batch % nb != 0) return 0;
  return mp.nh * mp.nw * (C2 > 0 ? 2 : 1) <= kBfMaxSeg ? 1 : 0;
}

__global__ void merged_pack_kernel(const float* __restrict__ w, int kw, int Cn, int Ck, MergedPlan mp,
                                   __nv_bfloat16* __restrict__ planes) {
  // planes [2][noff * 4 * Cn][Ck]; one thread per (row, 4 channels)
  const int64_t rows = (int64_t)mp.nh * mp.nw * 4 * Cn, total = rows * (Ck / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / (Ck / 4);
    const int c4 = (int)(i - row * (Ck / 4)) * 4;
    const int off = (int)(row / (4 * Cn)), rem = (int)(row - (int64_t)off * 4 * Cn);
    const int phase = rem / Cn, co = rem - phase * Cn;
    const int r = mp.tap_r[phase >> 1][off / mp.nw], s_ = mp.tap_s[phase & 1][off % mp.nw];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= 0 && s_ >= 0) v = *reinterpret_cast<const float4*>(w + (((int64_t)r * kw + s_) * Cn + co) * Ck + c4);
    const float x[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16* hi = planes + row * Ck + c4;
    __nv_bfloat16* lo = hi + rows * Ck;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat16 h = __float2bfloat16_rn(x[e]);
      hi[e] = h;
      lo[e] = __float2bfloat16_rn(x[e] - __bfloat162float(h));
    }
  }
}

int64_t bf3_conv2d_transpose_merged_rows(int Cn, const ctx_conv_geom& g) {
  MergedPlan mp;
  if (!merged_plan(g, &mp)) return 0;
  return (int64_t)mp.nh * mp.nw * 4 * Cn;
}

// w [kh,kw,Cn,Ck] fp32 -> packed planes [2][rows][Ck], rows = bf3_conv2d_transpose_merged_rows
int bf3_conv2d_transpose_merged_pack(const float* w, int Ck, int Cn, const ctx_conv_geom& g, __nv_bfloat16* planes,
                                     cudaStream_t stream) {
  MergedPlan mp;
  CTX_REQUIRE(merged_plan(g, &mp) && Ck % 4 == 0, "conv2d_transpose_merged_pack: geometry not eligible");
  CTX_REQUIRE(bf_aligned16(w) && bf_aligned16(planes), "conv2d_transpose_merged_pack: pointers must be 16-byte aligned");
  const int64_t total = (int64_t)mp.nh * mp.nw * 4 * Cn * (Ck / 4);
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  merged_pack_kernel<<<blocks, 256, 0, stream>>>(w, g.kw, Cn, Ck, mp, planes);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "merged_pack");
  count_launch();
  return CTX_OK;
}

int bf3_conv2d_transpose_merged(const __nv_bfloat16* x1p, int C1, const __nv_bfloat16* x2p, int C2, int x2_batch,
                                const __nv_bfloat16* wmp, float* y, __nv_bfloat16* yp, int Cn, const ctx_conv_geom& g,
                                const Epi& epi, cudaStream_t stream) {
  CTX_REQUIRE(bf_aligned16(x1p) && bf_aligned16(x2p) && bf_aligned16(wmp) && bf_aligned16(y) && bf_aligned16(yp),
              "conv2d_transpose_merged(bf16x3): pointers must be 16-byte aligned");
  CTX_REQUIRE(!epi.act_ref && !epi.accumulate && !epi.round, "conv2d_transpose_merged(bf16x3): only bias / lrelu epilogues");
  MergedPlan mp;
  if (!merged_plan(g, &mp)) { set_error("conv2d_transpose_merged(bf16x3): geometry not eligible"); return CTX_ERR_UNSUPPORTED; }
  const int Ck = C1 + C2;
  BfParams p;
  bf_fill_common(p, epi, y, yp, (int64_t)g.N * g.H * g.W * Cn);
  int tpi;
  if (!bf_tile_grid(g.P, g.Q, &p.BP, &p.NB, &tpi)) { set_error("conv2d_transpose_merged(bf16x3): grid does not tile"); return CTX_ERR_UNSUPPORTED; }
  p.tiles_per_img = tpi;
  const int m_tiles = p.NB > 1 ? (int)cdiv(g.N, p.NB) : g.N * tpi;
  p.BN = 256;
  p.NG = g.N; p.PG = g.P; p.QG = g.Q; p.N = 4 * Cn;
  p.a_wrap = x2_batch > 0 ? x2_batch : 1;
  p.merge_cn = Cn; p.merge_W = g.W;
  {
    const int64_t plane = (int64_t)g.N * g.P * g.Q * C1;
    int64_t dims[5] = {C1, g.Q, g.P, g.N, 2}, str[5] = {1, C1, (int64_t)g.Q * C1, (int64_t)g.P * g.Q * C1, plane};
    int box[5] = {64, g.Q, p.BP, p.NB, 2};
    int rc = bf_make_map(&p.mapA[0], x1p, 5, dims, str, box);
    if (rc) return rc;
  }
  if (C2 > 0) {
    const int64_t plane = (int64_t)x2_batch * g.P * g.Q * C2;
    int64_t dims[5] = {C2, g.Q, g.P, x2_batch, 2}, str[5] = {1, C2, (int64_t)g.Q * C2, (int64_t)g.P * g.Q * C2, plane};
    int box[5] = {64, g.Q, p.BP, p.NB, 2};
    int rc = bf_make_map(&p.mapA[1], x2p, 5, dims, str, box);
    if (rc) return rc;
  }
  {
    const int64_t rows = (int64_t)mp.nh * mp.nw * 4 * Cn;
    int64_t dims[3] = {Ck, rows, 2}, str[3] = {1, Ck, rows * Ck};
    int box[3] = {64, 128, 2};
    int rc = bf_make_map(&p.mapB, wmp, 3, dims, str, box);
    if (rc) return rc;
  }
  int nseg = 0, kblocks = 0;
  for (int a = 0; a < mp.nh; ++a)
    for (int b = 0; b < mp.nw; ++b)
      for (int src = 0; src < (C2 > 0 ? 2 : 1); ++src) {
        BfSeg& sg = p.seg[nseg++];
        sg.mapA = (uint8_t)src; sg.wrap = (uint8_t)src;
        sg.dh = (int16_t)mp.dh[a]; sg.dw = (int16_t)mp.dw[b];
        sg.nblk = (uint16_t)((src ? C2 : C1) / 64);
        sg.bk0 = src ? C1 : 0; sg.br0 = (a * mp.nw + b) * 4 * Cn;
        kblocks += sg.nblk;
      }
  p.ph_seg0[0] = 0; p.ph_nseg[0] = nseg; p.ph_kblocks[0] = kblocks; p.ph_obase[0] = 0;
  p.nphase = 1;
  p.sN = (int64_t)g.H * g.W * Cn; p.sP = 2 * (int64_t)g.W * Cn; p.sQ = 2 * (int64_t)Cn;
  return bf_launch(p, m_tiles, 1, stream, "bf3_conv2d_transpose_merged");
}

}  // namespace ctx
