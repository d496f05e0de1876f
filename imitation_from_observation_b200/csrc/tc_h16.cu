// Implicit-GEMM convolutions on 16-BIT OPERAND PLANES (fp16 or bf16) with tcgen05 kind::f16, fp32 accumulation in TMEM.
//
// A "planes" tensor is [planes][shape] of 16-bit values: plane 0 = hi = rn16(s * x), plane 1 = lo = rn16(s * x - hi)
// (s = an optional power-of-two scale kept in a device word, 1 for activations and weights).  Two arithmetic modes:
//
//   nprod = 3   D = A_hi*B_hi + (A_hi*B_lo + A_lo*B_hi)        fp32-grade: hi + lo carries 22 (fp16) / 16 (bf16) significant
//               bits; the cross terms go to a second accumulator (the tensor core truncates when it adds into a running
//               fp32 sum, so small terms must not ride on the big one) and D1 + D2 is formed once, in the epilogue.
//               This is the forward pass (every lrelu mask is decided there).
//   nprod = 1   D = A_hi*B_hi with fp16 planes: fp16 has the TF32 significand (1 + 10 bits), rounded to NEAREST by the
//               producer — single-pass-TF32 accuracy without the hardware's operand truncation — at twice the tensor-core
//               rate and HALF the operand bytes.  The backward GEMMs were bound by L2 -> SM operand traffic with fp32
//               operands (profiles/r1_ncu_full_summary.txt: 1.85 GB in 192 us = 9.6 TB/s for d_h2's data gradient), so
//               halving the bytes is what makes them faster.  Gradients are scaled into fp16's range per tensor by the
//               producer (ctx_grad_planes); the epilogue here multiplies the result by 1 / scale.
//
// Three GEMM forms through ONE warp-specialised kernel h16_kernel<pair, nprod> (320 threads: warp 0 = TMA producer, warp 1
// = TMEM allocator + MMA issue, warps 2..9 = epilogue, two warps per TMEM lane quarter):
//   * conv fprop (stride 1 / 2 through parity views) — also the data gradient of a transposed conv;
//   * transposed conv, stride phases merged into one launch, two sources = fused concat (arm_shaping.py:1323-1329) — also
//     the data gradient of a conv; and the phase-merged form for few output channels (columns = (phase, channel));
//   * conv filter gradient: pixels are the K dimension, so both operands are MN-MAJOR tiles (64-channel atoms of
//     [64 pixels x 128 B], SWIZZLE_128B: LBO = atom stride, SBO = 1024 B between 8-pixel groups), split-K over pixel
//     blocks, fp32 vector reductions into dw.
// K-major operand tiles are one TMA box per operand and stage: {64 ch, Q, P, N, plane} pixel boxes taken at the filter
// tap's offset (out-of-range coordinates are zero-filled by the TMA unit = TF SAME padding) and {64 k, rows, plane} weight
// rows.  256-wide N tiles run as CTA pairs (cta_group::2, cluster of 2 along M): one 256 x 256 MMA per step for two SMs,
// each CTA stages its own 128 rows of A and HALF of the B tile.  nprod = 1 tiles need 256 TMEM columns and 32 KB per
// stage, so two CTAs (or pairs) are co-resident per SM and overlap each other's prologue / epilogue.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <string.h>
#include <stdlib.h>
#include "kernels.h"

namespace ctx {

constexpr int kHMaxSeg = 64;
constexpr int kHThreads = 320;
constexpr uint32_t kHATile = 128 * 128;            // 128 rows x 64 elements x 2 B per plane (either major)

struct HSeg {                                      // K-major: one filter tap of one source tensor; MN-major: one filter tap
  int16_t dh, dw;
  uint8_t mapA, wrap;
  uint16_t nblk;                                   // 64-channel blocks (K-major)
  int32_t bk0, br0;                                // B (K-major): K coordinate of the first block; row offset added to n0
};

struct HParams {
  CUtensorMap mapA[8];                             // K-major: up to 4 parity views / 2 sources; MN-major: [view + 4 * plane]
  CUtensorMap mapB[2];                             // K-major: [0]; MN-major: small1 / small2
  CUtensorMap mapY[4], mapYP[4];                   // result (fp32) / its planes per stride phase: TMA-store epilogue
  int32_t tma_out, epi_bufs;                       // 1: epilogue through shared memory + TMA stores; staging buffers (1 | 2)
  HSeg seg[kHMaxSeg];
  int32_t mn;                                      // 0 = K-major operands (conv / transposed conv), 1 = MN-major (filter gradient)
  int32_t BN, stages, f16;
  int32_t nphase, ph_seg0[4], ph_nseg[4], ph_kblocks[4];
  int64_t ph_obase[4];
  int32_t NG, PG, QG, BP, NB, tiles_per_img, a_wrap, N;
  // MN-major (filter gradient): M = taps * Ca rows; K blocks are 64-pixel boxes of the small grid
  int32_t M, Ca, a_box, b_box, Cb1, kPG, kQG, kBP, kNB, k_tiles_per_img, total_kblocks, kblocks_per_split;
  // output
  float* out;
  int64_t sN, sP, sQ;
  void* planes;                                    // optional 16-bit planes of the result, indexed like `out`
  int64_t plane_stride;
  int32_t out_np;                                  // planes written: 1 | 2
  const float* bias;
  const float* act_ref;
  float slope;
  int32_t act, accumulate, atomic_out;
  const float* a_scale;                            // device scalar s: operand planes hold s * x, the result is divided by s
  uint32_t* amax_bits;                             // device word: atomicMax of the bit patterns of |result|
  int32_t merge_cn, merge_W;                       // phase-merged transposed conv (see the host function)
  int32_t* err_flag;
};

// ---- device helpers (raw PTX) ------------------------------------------------------------------------------------------
namespace h16 {
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
      if (err_flag) atomicExch(err_flag, 0x3000 + id);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <bool kPair>
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
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
// pair's MMA datapath (tc_gemm.cu mbar_arrive_remote has the measurement behind relaxed vs release)
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// SM100 shared-memory matrix descriptor (version 1), SWIZZLE_128B.  K-major: rows of 128 B (64 elements along K), 8-row
// groups SBO = 1024 B apart, LBO unused.  MN-major: atoms of [k rows][128 B = 64 elements along M/N]; the next 8 k-rows are
// SBO = 1024 B further, the next 64-element M/N atom LBO bytes further; one K = 16 MMA spans two k-groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// two fp32 -> one 32-bit word of 16-bit values (low half = first), round to nearest even, saturating for fp16
__device__ __forceinline__ uint32_t pack16(float a, float b, bool f16) {
  if (f16) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
  }
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
}
__device__ __forceinline__ float unpack16_lo(uint32_t w, bool f16) {
  return f16 ? __half2float(__ushort_as_half((unsigned short)(w & 0xFFFFu))) : __uint_as_float(w << 16);
}
__device__ __forceinline__ float unpack16_hi(uint32_t w, bool f16) {
  return f16 ? __half2float(__ushort_as_half((unsigned short)(w >> 16))) : __uint_as_float(w & 0xFFFF0000u);
}
}  // namespace h16

// ---- the kernel -----------------------------------------------------------------------------------------------------------
template <bool kPair, int kProd>
__global__ void __launch_bounds__(kHThreads, kProd == 1 ? 2 : 1) h16_kernel(const __grid_constant__ HParams p) {
  using namespace h16;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank = 0u;
  if constexpr (kPair) rank = cluster_ctarank();
  constexpr uint32_t np = kProd == 3 ? 2u : 1u;                          // operand planes read
  const int bn_cta = kPair ? p.BN / 2 : p.BN;
  const uint32_t b_tile = (uint32_t)bn_cta * 128u;                       // this CTA's share of the B tile, per plane
  const uint32_t stage_bytes = np * (kHATile + b_tile);                  // [A planes | B planes]
  const uint32_t b_off = np * kHATile;
  __shared__ __align__(8) uint64_t bars[25];
  __shared__ uint32_t tmem_slot;
  const int S = p.stages;
  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[8 + s]); };
  auto peer_bar = [&](int s) { return smem_u32(&bars[16 + s]); };       // pair leader: the follower's stage has landed
  const uint32_t acc_bar = smem_u32(&bars[24]);

  const int z = blockIdx.z;                      // K-major: stride phase of a transposed conv; MN-major: K split
  int seg0 = 0, nseg = 0, nkb, kb_begin = 0;
  int64_t obase = 0;
  if (!p.mn) {
    seg0 = p.ph_seg0[z]; nseg = p.ph_nseg[z]; nkb = p.ph_kblocks[z]; obase = p.ph_obase[z];
  } else {
    kb_begin = z * p.kblocks_per_split;
    const int kb_end = min(p.total_kblocks, kb_begin + p.kblocks_per_split);
    nkb = max(0, kb_end - kb_begin);
  }
  const int tile_m = blockIdx.x;
  int n0, p0;
  if (p.NB > 1 || p.tiles_per_img == 1) { n0 = tile_m * p.NB; p0 = 0; }
  else { n0 = tile_m / p.tiles_per_img; p0 = (tile_m - n0 * p.tiles_per_img) * p.BP; }
  const int ncol0 = blockIdx.y * p.BN;
  uint32_t ncols = 32;
  while (ncols < (kProd == 3 ? 2u : 1u) * (uint32_t)p.BN) ncols <<= 1;

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

  if (nkb > 0) {
    if (warp == 0) {
      // =============================== TMA producer ===================================================================
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const int nb0 = ncol0 + (int)rank * bn_cta;                        // pair: this CTA's half of the B tile
        if (!p.mn) {
          // K order: source, 64-channel block, filter tap.  The taps of one channel block read the SAME input bytes at
          // shifted pixel offsets, so issuing them back to back turns 8 of 9 (24 of 25) reads of every input byte into L2
          // hits; tap-major order re-read the tile's input from DRAM once per tap (ncu: 8.1 GB for 0.34 GB of operands on
          // Inception2's h0, profiles/r2_ncu_inception2_3x3.txt).
          for (int src = 0; src < 2; ++src) {
            int nblk = 0;
            for (int sg = 0; sg < nseg; ++sg)
              if (p.seg[seg0 + sg].wrap == src) { nblk = p.seg[seg0 + sg].nblk; break; }
            const int nimg = src ? (n0 % p.a_wrap) : n0;
            for (int cb = 0; cb < nblk; ++cb)
              for (int sg = 0; sg < nseg; ++sg) {
                const HSeg s = p.seg[seg0 + sg];
                if (s.wrap != src) continue;
                mbar_wait<false>(empty_bar(stage), phase ^ 1, p.err_flag, 1);
                const uint32_t sa = smem_base + stage * stage_bytes;
                mbar_arrive_expect_tx(full_bar(stage), stage_bytes);
                tma_load_5d(sa, &p.mapA[s.mapA], full_bar(stage), cb * 64, s.dw, p0 + s.dh, nimg, 0);
                tma_load_3d(sa + b_off, &p.mapB[0], full_bar(stage), s.bk0 + cb * 64, s.br0 + nb0, 0);
                if (++stage == S) { stage = 0; phase ^= 1; }
              }
          }
        } else {
          // filter gradient: the 128 tile rows are two 64-channel atoms (tap, channel block); K blocks are 64-pixel boxes
          const int m0 = tile_m * 128;
          int natomA = 0;
          for (int j = 0; j < 2; ++j) natomA += (m0 + 64 * j < p.M) ? 1 : 0;
          const uint32_t tx = np * ((uint32_t)natomA * 8192u + b_tile);
          for (int kb = kb_begin; kb < kb_begin + nkb; ++kb) {
            int kn0, kp0;
            if (p.kNB > 1 || p.k_tiles_per_img == 1) { kn0 = kb * p.kNB; kp0 = 0; }
            else { kn0 = kb / p.k_tiles_per_img; kp0 = (kb - kn0 * p.k_tiles_per_img) * p.kBP; }
            mbar_wait<false>(empty_bar(stage), phase ^ 1, p.err_flag, 2);
            const uint32_t sa = smem_base + stage * stage_bytes;
            const uint32_t sb = sa + b_off;
            mbar_arrive_expect_tx(full_bar(stage), tx);
            for (int j = 0; j < 2; j += p.a_box) {
              const int m = m0 + 64 * j;
              if (m >= p.M) break;
              const int tap = m / p.Ca, a0 = m - tap * p.Ca;
              const HSeg s = p.seg[tap];
#pragma unroll
              for (uint32_t pl = 0; pl < np; ++pl)
                tma_load_5d(sa + pl * kHATile + j * 8192, &p.mapA[s.mapA + 4 * pl], full_bar(stage), 0, s.dw, kp0 + s.dh, kn0,
                            a0 >> 6);
            }
            const int prow = (kn0 * p.kPG + kp0) * p.kQG;                    // first pixel row of this K block
            const int prow2 = ((kn0 % p.a_wrap) * p.kPG + kp0) * p.kQG;
            for (int a = 0; a < bn_cta / 64; a += p.b_box) {
              const int col = nb0 + 64 * a;
              const bool first = col < p.Cb1;
#pragma unroll
              for (uint32_t pl = 0; pl < np; ++pl)
                tma_load_4d(sb + pl * b_tile + a * 8192, &p.mapB[first ? 0 : 1], full_bar(stage), 0, first ? prow : prow2,
                            (first ? col : col - p.Cb1) >> 6, (int)pl);
            }
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
          // kind::f16: D = f32 (1 << 4); A / B format at bits 7 / 10 (0 = f16, 1 = bf16); bits 15 / 16 = A / B MN-major;
          // N >> 3 at bit 17, M >> 4 at bit 24
          const uint32_t fmt = p.f16 ? 0u : 1u, mnb = p.mn ? 1u : 0u;
          const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (mnb << 15) | (mnb << 16) |
                                 ((uint32_t)(p.BN >> 3) << 17) | (((kPair ? 256u : 128u) >> 4) << 24);
          const uint32_t d1 = tmem_base, d2 = tmem_base + (uint32_t)p.BN;
          const uint32_t lbo = p.mn ? 8192u : 16u, sbo = 1024u;
          const uint64_t kstep = p.mn ? 128u : 2u;        // 16-byte units per K = 16 step: 2048 B (two 8-row groups) | 32 B
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait<false>(full_bar(stage), phase, p.err_flag, 2);
            if constexpr (kPair) mbar_wait<true>(peer_bar(stage), phase, p.err_flag, 5);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * stage_bytes;
            const uint64_t a_hi = make_desc(sa, lbo, sbo), b_hi = make_desc(sa + b_off, lbo, sbo);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t first = (kb | ks) ? 1u : 0u;
              umma_f16<kPair>(d1, a_hi + kstep * ks, b_hi + kstep * ks, idesc, first);
              if constexpr (kProd == 3) {
                const uint64_t a_lo = make_desc(sa + kHATile, lbo, sbo), b_lo = make_desc(sa + b_off + b_tile, lbo, sbo);
                umma_f16<kPair>(d2, a_hi + kstep * ks, b_lo + kstep * ks, idesc, first);
                umma_f16<kPair>(d2, a_lo + kstep * ks, b_hi + kstep * ks, idesc, 1u);
              }
            }
            umma_commit<kPair>(empty_bar(stage));         // frees the stage (in both CTAs of a pair) once the MMAs have read it
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
          umma_commit<kPair>(acc_bar);
        }
      }
    }
  }
  if (warp >= 2 && (nkb > 0 || !p.atomic_out)) {
    // =============================== epilogue: warps 2..9 ==============================================================
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
    const bool f16 = p.f16 != 0;
    float inv_scale = 1.f;
    if (p.a_scale) inv_scale = 1.f / __ldg(p.a_scale);   // (a power of two: exact)
    uint32_t amax = 0u;
    // lrelu' mask of a 32-column chunk as 32 sign bits of the saved activation.  The epilogue warps are idle during the main
    // loop, so the first chunk's bits are fetched BEFORE waiting for the accumulator and each further chunk's right after the
    // previous one left the registers: the row-per-thread global loads never sit between tensor memory and the stores.
    auto ref_mask = [&](int cg) -> uint32_t {
      if (!p.act_ref) return 0xFFFFFFFFu;
      const int col = ncol0 + cg + half * 32;
      int64_t ooff = roff + col;
      if (p.merge_cn) {
        const int mi = col / p.merge_cn;
        ooff = roff + ((int64_t)(mi >> 1) * p.merge_W + (mi & 1)) * p.merge_cn + (col - mi * p.merge_cn);
      }
      if (!(row_ok && col < p.N)) return 0xFFFFFFFFu;
      uint32_t m = 0u;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p.act_ref + ooff + j));
        m |= (a.x >= 0.f ? 1u : 0u) << j | (a.y >= 0.f ? 1u : 0u) << (j + 1) | (a.z >= 0.f ? 1u : 0u) << (j + 2) |
             (a.w >= 0.f ? 1u : 0u) << (j + 3);
      }
      return m;
    };
    uint32_t mask = p.tma_out ? ref_mask(0) : 0u;
    // accumulate (+=): the chunk's previous values, fetched the same way — eight independent 128-bit loads in flight per thread
    // before the accumulator wait / under the previous chunk's stores, instead of one load -> add -> next load chain.
    float4 prev[8];
    auto load_prev = [&](int cg) {
      const int col = ncol0 + cg + half * 32;
      int64_t ooff = roff + col;
      if (p.merge_cn) {
        const int mi = col / p.merge_cn;
        ooff = roff + ((int64_t)(mi >> 1) * p.merge_W + (mi & 1)) * p.merge_cn + (col - mi * p.merge_cn);
      }
      const bool ok = row_ok && col < p.N;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        prev[j] = ok ? *reinterpret_cast<const float4*>(p.out + ooff + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    if (p.tma_out && p.accumulate) load_prev(0);
    if (nkb > 0) {
      mbar_wait<false>(acc_bar, 0, p.err_flag, 3);
      tc_fence_after();
    }
    if (p.tma_out) {
      // ---- epilogue through shared memory: 64 columns at a time, every epilogue thread writes its row's 32 columns (and
      // their planes) into a 128-byte-swizzled box image in the (now idle) operand stages, and one thread hands the boxes to
      // the TMA unit — whole 128-byte lines instead of 32 partial lines per st.global instruction.  Two staging buffers:
      // the stores of one column group drain while the next one is computed.
      const uint32_t sw = (uint32_t)(r & 7);
      const uint32_t need = (2u + (uint32_t)p.out_np) * kHATile;
      const bool elected = warp == 2 && lane == 0;
      int gi = 0;
      for (int cg = 0; cg < p.BN; cg += 64, ++gi) {
        const int c0 = cg + half * 32;
        uint32_t v[32];
        if (nkb > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
          if constexpr (kProd == 3) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t v2[16];
              tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.BN + c0 + 16 * h), v2);
#pragma unroll
              for (int j = 0; j < 16; ++j)
                v[16 * h + j] = __float_as_uint(__uint_as_float(v[16 * h + j]) + __uint_as_float(v2[j]));
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        const int col = ncol0 + c0;
        int co = col, mi = z;
        int64_t ooff = roff + col;
        if (p.merge_cn) {
          mi = col / p.merge_cn;
          co = col - mi * p.merge_cn;
          ooff = roff + ((int64_t)(mi >> 1) * p.merge_W + (mi & 1)) * p.merge_cn + co;
        }
        const bool live = row_ok && col < p.N;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float* pp = reinterpret_cast<const float*>(&prev[j >> 2]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float t = __uint_as_float(v[j + e]) * inv_scale;
            if (p.bias) t += __ldg(p.bias + co + j + e);
            if (p.act) t = fmaxf(t, p.slope * t);
            if (p.act_ref) t *= ((mask >> (j + e)) & 1u) ? 1.f : p.slope;
            if (p.accumulate) t += pp[e];
            v[j + e] = __float_as_uint(t);
            if (live) amax = max(amax, __float_as_uint(t) & 0x7FFFFFFFu);
          }
        }
        // the buffer this group goes into was last read by the stores of group gi - epi_bufs
        const uint32_t sbuf = smem_base + (uint32_t)(gi % p.epi_bufs) * need;
        if (elected && gi >= p.epi_bufs) {
          if (p.epi_bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        {
          const uint32_t fo = sbuf + (uint32_t)half * kHATile + (uint32_t)r * 128u;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(fo + ((((uint32_t)j >> 2) ^ sw) << 4)), "r"(v[j]),
                         "r"(v[j + 1]), "r"(v[j + 2]), "r"(v[j + 3]) : "memory");
          if (p.planes) {
            const uint32_t po = sbuf + 2u * kHATile + (uint32_t)r * 128u;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float x0 = __uint_as_float(v[j + 2 * e]), x1 = __uint_as_float(v[j + 2 * e + 1]);
                hi[e] = pack16(x0, x1, f16);
                lo[e] = pack16(x0 - unpack16_lo(hi[e], f16), x1 - unpack16_hi(hi[e], f16), f16);
              }
              const uint32_t ch = ((((uint32_t)half * 4u + ((uint32_t)j >> 3)) ^ sw) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(po + ch), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
              if (p.out_np == 2)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(po + kHATile + ch), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (cg + 64 < p.BN) {                               // (the chunk's registers are free: next chunk's sign bits / sums)
          mask = ref_mask(cg + 64);
          if (p.accumulate) load_prev(cg + 64);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (elected && ncol0 + cg < p.N) {
          // (a 64-column group never straddles stride phases of the merged form: merge_cn is a multiple of 64)
          const int colg = ncol0 + cg;
          int mg = z, cc = colg;
          if (p.merge_cn) { mg = colg / p.merge_cn; cc = colg - mg * p.merge_cn; }
          tma_store_4d(&p.mapY[mg], sbuf, cc, 0, p0, n0);
          tma_store_4d(&p.mapY[mg], sbuf + kHATile, cc + 32, 0, p0, n0);
          if (p.planes) {
            tma_store_5d(&p.mapYP[mg], sbuf + 2u * kHATile, cc, 0, p0, n0, 0);
            if (p.out_np == 2) tma_store_5d(&p.mapYP[mg], sbuf + 3u * kHATile, cc, 0, p0, n0, 1);
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (elected) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // shared memory outlives the stores
    } else
    for (int c0 = c_begin; c0 < c_end; c0 += 32) {
      uint32_t v[32];
      if (nkb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        if constexpr (kProd == 3) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t v2[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.BN + c0 + 16 * h), v2);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              v[16 * h + j] = __float_as_uint(__uint_as_float(v[16 * h + j]) + __uint_as_float(v2[j]));
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      const int col = ncol0 + c0;
      if (!row_ok || col >= p.N) continue;
      int co = col;                                      // output channel of the chunk's first column
      int64_t ooff = roff + col;
      if (p.merge_cn) {                                  // (warp-uniform: a 32-column chunk never straddles phases)
        const int phase = col / p.merge_cn;
        co = col - phase * p.merge_cn;
        ooff = roff + ((int64_t)(phase >> 1) * p.merge_W + (phase & 1)) * p.merge_cn + co;
      }
      float* optr = p.out + ooff;
      if (p.atomic_out) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(optr + j),
                       "f"(__uint_as_float(v[j]) * inv_scale), "f"(__uint_as_float(v[j + 1]) * inv_scale),
                       "f"(__uint_as_float(v[j + 2]) * inv_scale), "f"(__uint_as_float(v[j + 3]) * inv_scale) : "memory");
        continue;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 ref = make_float4(0.f, 0.f, 0.f, 0.f), prev = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.act_ref) ref = __ldg(reinterpret_cast<const float4*>(p.act_ref + ooff + j));
        if (p.accumulate) prev = *reinterpret_cast<const float4*>(optr + j);
        const float* rr = reinterpret_cast<const float*>(&ref);
        const float* pp = reinterpret_cast<const float*>(&prev);
        float4 o;
        float* oo = reinterpret_cast<float*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float t = __uint_as_float(v[j + e]) * inv_scale;
          if (p.bias) t += __ldg(p.bias + co + j + e);
          if (p.act) t = fmaxf(t, p.slope * t);
          if (p.act_ref) t *= (rr[e] >= 0.f ? 1.f : p.slope);
          if (p.accumulate) t += pp[e];
          oo[e] = t;
          v[j + e] = __float_as_uint(t);
          amax = max(amax, __float_as_uint(t) & 0x7FFFFFFFu);
        }
        *reinterpret_cast<float4*>(optr + j) = o;
      }
      if (p.planes) {
        // 16-bit planes of the result: hi = rn16(x), lo = rn16(x - hi)  (x - hi is exact in fp32)
        uint8_t* ph = reinterpret_cast<uint8_t*>(p.planes) + 2 * ooff;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float x0 = __uint_as_float(v[j + 2 * e]), x1 = __uint_as_float(v[j + 2 * e + 1]);
            hi[e] = pack16(x0, x1, f16);
            lo[e] = pack16(x0 - unpack16_lo(hi[e], f16), x1 - unpack16_hi(hi[e], f16), f16);
          }
          *reinterpret_cast<uint4*>(ph + 2 * j) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          if (p.out_np == 2) *reinterpret_cast<uint4*>(ph + 2 * p.plane_stride + 2 * j) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
    }
    if (p.amax_bits) {
      amax = __reduce_max_sync(0xffffffffu, amax);
      if (lane == 0 && amax) atomicMax(p.amax_bits, amax);
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
typedef CUresult (*HEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static HEncodeFn h_get_encode() {
  static HEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (HEncodeFn)f;
    else
      cudaGetLastError();
  }
  return fn;
}

// 16-bit tensor map, zero fill, 128-byte swizzle; dims / box innermost first; strides in ELEMENTS for dims 1..rank-1
int h16_make_map(CUtensorMap* out, int f16, const void* base, int rank, const int64_t* dims, const int64_t* strides,
                 const int* box) {
  HEncodeFn enc = h_get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return CTX_ERR_CUDA; }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = (cuuint64_t)dims[i]; bx[i] = (cuuint32_t)box[i]; es[i] = 1;
    if (dims[i] <= 0 || box[i] <= 0 || box[i] > 256) { set_error("16-bit tensor map: bad dim/box %d: %lld/%d", i, (long long)dims[i], box[i]); return CTX_ERR_ARG; }
  }
  for (int i = 1; i < rank; ++i) {
    gs[i - 1] = (cuuint64_t)strides[i] * (f16 == 2 ? 4u : 2u);
    if (gs[i - 1] % 16 != 0) { set_error("16-bit tensor map: stride %d (%lld elements) is not 16-byte aligned", i, (long long)strides[i]); return CTX_ERR_ARG; }
  }
  if (((uintptr_t)base & 15) != 0) { set_error("16-bit tensor map: base pointer is not 16-byte aligned"); return CTX_ERR_ARG; }
  const CUtensorMapDataType dt = f16 == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                          : (f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  CUresult r = enc(out, dt, (cuuint32_t)rank, (void*)base,
                   gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (16-bit) failed with CUresult %d", (int)r); return CTX_ERR_CUDA; }
  return CTX_OK;
}

// rows = `rows` pixels: BP rows x Q columns of one image (P % BP == 0), or NB whole images
static bool h_tile_grid(int P, int Q, int rows, int* BP, int* NB, int* tiles_per_img) {
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

static int h_pick_bn(int N) { return N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : 64); }

// CTA pairs (cta_group::2, two SMs on one 256-row tile, each staging half of B): every 256-wide tile.  CTX_H16_PAIR128=1 also
// pairs the 128-wide THREE-product tiles (a single-CTA 128 x 128 three-product tile moves 64 KB per k-block for 768 MMA cycles,
// a pair 48 KB per CTA) — measured a wash (h1's forward 169 vs 160 us, d_h2 311 vs 316 us, the step 4.80 ms either way: those
// layers are not held back by operand traffic alone), so it stays off.
static bool h_pair(int BN, int m_tiles, int nprod) {
  static int knob = -1;
  if (knob < 0) { const char* e = getenv("CTX_H16_PAIR128"); knob = (e && e[0] == '1') ? 1 : 0; }
  return m_tiles >= 2 && (BN == 256 || (BN == 128 && nprod == 3 && knob));
}

static int h_nprod_ok(const H16Cfg& cfg, const char* name) {
  CTX_REQUIRE(cfg.nprod == 1 || cfg.nprod == 3, "%s: nprod must be 1 or 3", name);
  CTX_REQUIRE(cfg.nprod == 3 || cfg.f16, "%s: single-product mode needs fp16 planes (bf16 keeps 8 significant bits)", name);
  return CTX_OK;
}

static int h_launch(HParams& p, int m_tiles, int zdim, const H16Cfg& cfg, cudaStream_t stream, const char* name) {
  static int max_dyn_dev[64] = {0};
  int dev = 0;
  CTX_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { set_error("%s: device index %d out of range", name, dev); return CTX_ERR_ARG; }
  if (max_dyn_dev[dev] == 0) {
    cudaFuncAttributes fa[4];
    CTX_CUDA(cudaFuncGetAttributes(&fa[0], h16_kernel<false, 1>));
    CTX_CUDA(cudaFuncGetAttributes(&fa[1], h16_kernel<true, 1>));
    CTX_CUDA(cudaFuncGetAttributes(&fa[2], h16_kernel<false, 3>));
    CTX_CUDA(cudaFuncGetAttributes(&fa[3], h16_kernel<true, 3>));
    int optin = 0;
    CTX_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    size_t stat = 0;
    for (int i = 0; i < 4; ++i) stat = fa[i].sharedSizeBytes > stat ? fa[i].sharedSizeBytes : stat;
    const int dyn = optin - (int)stat;
    CTX_CUDA(cudaFuncSetAttribute(h16_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    CTX_CUDA(cudaFuncSetAttribute(h16_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    CTX_CUDA(cudaFuncSetAttribute(h16_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    CTX_CUDA(cudaFuncSetAttribute(h16_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    max_dyn_dev[dev] = dyn;
  }
  const bool pair = p.mn ? (p.BN == 256 && m_tiles >= 2) : h_pair(p.BN, m_tiles, cfg.nprod);   // (each CTA stages half of B)
  const uint32_t np = cfg.nprod == 3 ? 2u : 1u;
  const uint32_t stage = np * (kHATile + (uint32_t)(pair ? p.BN / 2 : p.BN) * 128u);
  // single-product tiles: two CTAs (pairs) per SM overlap each other's prologue / epilogue (256 TMEM columns each)
  const int occ = cfg.nprod == 1 ? 2 : 1;
  const uint32_t budget = occ == 1 ? (uint32_t)max_dyn_dev[dev] : (uint32_t)(max_dyn_dev[dev] - 2048) / 2u;
  int stages = (int)((budget - 1024u) / stage);
  if (stages > 8) stages = 8;
  if (stages < 2) { set_error("%s: tile does not fit shared memory", name); return CTX_ERR_ARG; }
  p.stages = stages;
  if (p.tma_out) {                                   // the epilogue stages 64-column groups in the idle operand stages
    const uint32_t need = (2u + (uint32_t)p.out_np) * kHATile;
    const uint32_t have = (uint32_t)stages * stage;
    p.epi_bufs = have >= 2 * need ? 2 : 1;
    if (have < need || p.atomic_out || p.mn) p.tma_out = 0;
  }
  p.f16 = cfg.f16 ? 1 : 0;
  p.a_scale = cfg.a_scale;
  p.amax_bits = cfg.amax_bits;
  p.err_flag = nullptr;
  const size_t smem = (size_t)stages * stage + 1024;
  const int n_tiles = (int)cdiv(p.N, p.BN);
  if (n_tiles > 65535 || zdim > 65535) { set_error("%s: grid too large", name); return CTX_ERR_ARG; }
  cudaLaunchConfig_t cfgl;
  memset(&cfgl, 0, sizeof(cfgl));
  cfgl.gridDim = dim3((unsigned)(pair ? (m_tiles + 1) & ~1 : m_tiles), (unsigned)n_tiles, (unsigned)zdim);
  cfgl.blockDim = dim3(kHThreads, 1, 1);
  cfgl.dynamicSmemBytes = smem;
  cfgl.stream = stream;
  cudaLaunchAttribute attr[1];
  cudaError_t e;
  if (pair) {
    // cluster of 2 along the M tiles; an odd last tile gets an all-out-of-range partner (TMA zero-fills, epilogue masks)
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfgl.attrs = attr; cfgl.numAttrs = 1;
    e = cfg.nprod == 3 ? cudaLaunchKernelEx(&cfgl, h16_kernel<true, 3>, p) : cudaLaunchKernelEx(&cfgl, h16_kernel<true, 1>, p);
  } else {
    e = cfg.nprod == 3 ? cudaLaunchKernelEx(&cfgl, h16_kernel<false, 3>, p) : cudaLaunchKernelEx(&cfgl, h16_kernel<false, 1>, p);
  }
  if (e != cudaSuccess) return cuda_fail(e, name);
  count_launch();
  return CTX_OK;
}

static void h_fill_common(HParams& p, const Epi& epi, float* y, void* yp, int yp_n, int64_t y_plane) {
  memset(&p, 0, sizeof(p));
  p.nphase = 1;
  p.a_wrap = 1;
  p.Cb1 = 1 << 30;
  p.out = y;
  p.planes = yp;
  p.out_np = yp ? yp_n : 0;
  p.plane_stride = y_plane;
  p.bias = epi.bias; p.slope = epi.slope; p.act = epi.act; p.act_ref = epi.act_ref; p.accumulate = epi.accumulate;
}

static bool h_aligned16(const void* ptr) { return ((uintptr_t)ptr & 15) == 0; }

// Output boxes of the TMA-store epilogue for one stride phase: the tile's pixels (n, p, q) of a [Cn, Qg, Pg, N] view of y with
// element strides (1, sQ, sP, sN) starting `base` elements into y; the planes get the same view plus a plane dimension.
static int h_out_maps(HParams& p, int idx, float* y, void* yp, int yp_n, int f16, int64_t base, int Cn, int Qg, int Pg, int N,
                      int64_t sQ, int64_t sP, int64_t sN, int64_t plane_stride) {
  int64_t dims[5] = {Cn, Qg, Pg, N, yp_n > 0 ? yp_n : 1}, str[5] = {1, sQ, sP, sN, plane_stride};
  int box[5] = {32, Qg, p.BP, p.NB, 1};
  int rc = h16_make_map(&p.mapY[idx], 2, y + base, 4, dims, str, box);
  if (rc) return rc;
  if (yp) {
    box[0] = 64;
    rc = h16_make_map(&p.mapYP[idx], f16, (uint16_t*)yp + base, 5, dims, str, box);
    if (rc) return rc;
  }
  return CTX_OK;
}
static bool h_tma_out_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("CTX_TMA_STORE"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

// ---- conv fprop -----------------------------------------------------------------------------------------------------------
int h16_conv2d_fprop_eligible(int C, int Co, const ctx_conv_geom& g) {
  int bp, nb, tpi;
  if (C % 64 != 0 || Co % 64 != 0) return 0;
  if (g.stride != 1 && g.stride != 2) return 0;
  if (g.kh * g.kw > kHMaxSeg) return 0;
  if (!h_tile_grid(g.P, g.Q, 128, &bp, &nb, &tpi)) return 0;
  if (g.stride == 2 && (g.H < 2 || g.W < 2)) return 0;          // empty parity views stay on the TF32 kernel
  return 1;
}

// xp: planes [.][N,H,W,C]; wp: planes [.][kh*kw*Co][C] (per-tap K-major, ctx_split_planes_transpose); y fp32 [N,P,Q,Co];
// yp: optional planes [yp_n][N,P,Q,Co] of the result
int h16_conv2d_fprop(const void* xp, int C, const void* wp, float* y, void* yp, int yp_n, int Co, const ctx_conv_geom& g,
                     const Epi& epi, const H16Cfg& cfg, cudaStream_t stream) {
  CTX_REQUIRE(h_aligned16(xp) && h_aligned16(wp) && h_aligned16(y) && h_aligned16(yp),
              "conv2d_fprop(p16): pointers must be 16-byte aligned");
  CTX_REQUIRE(!epi.round, "conv2d_fprop(p16): no TF32 rounding in this path");
  CTX_REQUIRE(!yp || yp_n == 1 || yp_n == 2, "conv2d_fprop(p16): y_planes_n must be 1 or 2");
  int rc = h_nprod_ok(cfg, "conv2d_fprop(p16)");
  if (rc) return rc;
  HParams p;
  h_fill_common(p, epi, y, yp, yp_n, (int64_t)g.N * g.P * g.Q * Co);
  int tpi;
  if (!h_tile_grid(g.P, g.Q, 128, &p.BP, &p.NB, &tpi)) { set_error("conv2d_fprop(p16): grid does not tile"); return CTX_ERR_UNSUPPORTED; }
  p.tiles_per_img = tpi;
  const int m_tiles = p.NB > 1 ? (int)cdiv(g.N, p.NB) : g.N * tpi;
  p.BN = h_pick_bn(Co);
  p.NG = g.N; p.PG = g.P; p.QG = g.Q; p.N = Co;
  const int np = cfg.nprod == 3 ? 2 : 1;
  const bool pair = h_pair(p.BN, m_tiles, cfg.nprod);
  const int64_t x_plane = (int64_t)g.N * g.H * g.W * C;
  if (g.stride == 1) {
    int64_t dims[5] = {C, g.W, g.H, g.N, np}, str[5] = {1, C, (int64_t)g.W * C, (int64_t)g.H * g.W * C, x_plane};
    int box[5] = {64, g.Q, p.BP, p.NB, np};
    rc = h16_make_map(&p.mapA[0], cfg.f16, xp, 5, dims, str, box);
    if (rc) return rc;
  } else {
    for (int hp = 0; hp < 2; ++hp)
      for (int wq = 0; wq < 2; ++wq) {
        const int Hh = (g.H - hp + 1) / 2, Wh = (g.W - wq + 1) / 2;
        int64_t dims[5] = {C, Wh, Hh, g.N, np};
        int64_t str[5] = {1, 2 * (int64_t)C, 2 * (int64_t)g.W * C, (int64_t)g.H * g.W * C, x_plane};
        int box[5] = {64, g.Q, p.BP, p.NB, np};
        rc = h16_make_map(&p.mapA[hp * 2 + wq], cfg.f16, (const uint16_t*)xp + ((int64_t)hp * g.W + wq) * C, 5, dims, str, box);
        if (rc) return rc;
      }
  }
  {
    const int64_t rows = (int64_t)g.kh * g.kw * Co;
    int64_t dims[3] = {C, rows, np}, str[3] = {1, C, rows * C};
    int box[3] = {64, pair ? p.BN / 2 : p.BN, np};
    rc = h16_make_map(&p.mapB[0], cfg.f16, wp, 3, dims, str, box);
    if (rc) return rc;
  }
  const int cblk = C / 64;
  int nseg = 0;
  for (int r = 0; r < g.kh; ++r)
    for (int s = 0; s < g.kw; ++s) {
      HSeg& sg = p.seg[nseg++];
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
  if (h_tma_out_enabled()) {
    rc = h_out_maps(p, 0, y, yp, yp_n, cfg.f16, 0, Co, g.Q, g.P, g.N, p.sQ, p.sP, p.sN, p.plane_stride);
    if (rc) return rc;
    p.tma_out = 1;
  }
  return h_launch(p, m_tiles, 1, cfg, stream, "h16_conv2d_fprop");
}

// ---- transposed conv (stride phases merged into one launch; two sources = fused concat) ---------------------------------------
static void h_phase_geom(const ctx_conv_geom& g, int ph, int pw, int* Hp, int* Wp, int* r0, int* s0, int* nr, int* ns,
                         int* dh0, int* dw0) {
  const int st = g.stride;
  *Hp = g.H > ph ? (g.H - ph + st - 1) / st : 0;
  *Wp = g.W > pw ? (g.W - pw + st - 1) / st : 0;
  *r0 = (ph + g.pad_h) % st; *s0 = (pw + g.pad_w) % st;
  *nr = *r0 < g.kh ? (g.kh - *r0 + st - 1) / st : 0;
  *ns = *s0 < g.kw ? (g.kw - *s0 + st - 1) / st : 0;
  *dh0 = (ph + g.pad_h - *r0) / st; *dw0 = (pw + g.pad_w - *s0) / st;
}

int h16_conv2d_transpose_eligible(int C1, int C2, int x2_batch, int Cn, const ctx_conv_geom& g) {
  if (C1 % 64 != 0 || C2 % 64 != 0 || Cn % 64 != 0) return 0;
  if (g.stride != 1 && g.stride != 2) return 0;
  int Hp0 = -1, Wp0 = -1, segs = 0;
  for (int ph = 0; ph < g.stride; ++ph)
    for (int pw = 0; pw < g.stride; ++pw) {
      int Hp, Wp, r0, s0, nr, ns, dh0, dw0, bp, nb, tpi;
      h_phase_geom(g, ph, pw, &Hp, &Wp, &r0, &s0, &nr, &ns, &dh0, &dw0);
      if (Hp == 0 || Wp == 0 || nr == 0 || ns == 0) return 0;
      if (Hp0 < 0) { Hp0 = Hp; Wp0 = Wp; }
      if (Hp != Hp0 || Wp != Wp0) return 0;                       // all phases in ONE launch: same pixel grid
      if (!h_tile_grid(Hp, Wp, 128, &bp, &nb, &tpi)) return 0;
      if (C2 > 0 && nb > 1 && x2_batch % nb != 0) return 0;
      segs += nr * ns * (C2 > 0 ? 2 : 1);
    }
  return segs <= kHMaxSeg ? 1 : 0;
}

// x1p planes [.][N,P,Q,C1]; x2p planes [.][x2_batch,P,Q,C2] or null; wp planes [.][kh*kw*Cn][C1+C2] (the TF filter layout
// [kh,kw,Cn,Ck] as it is); y fp32 [N,H,W,Cn]; yp optional planes [yp_n][N,H,W,Cn]
int h16_conv2d_transpose(const void* x1p, int C1, const void* x2p, int C2, int x2_batch, const void* wp, float* y, void* yp,
                         int yp_n, int Cn, const ctx_conv_geom& g, const Epi& epi, const H16Cfg& cfg, cudaStream_t stream) {
  CTX_REQUIRE(h_aligned16(x1p) && h_aligned16(x2p) && h_aligned16(wp) && h_aligned16(y) && h_aligned16(yp),
              "conv2d_transpose(p16): pointers must be 16-byte aligned");
  CTX_REQUIRE(!epi.round, "conv2d_transpose(p16): no TF32 rounding in this path");
  CTX_REQUIRE(!yp || yp_n == 1 || yp_n == 2, "conv2d_transpose(p16): y_planes_n must be 1 or 2");
  int rc = h_nprod_ok(cfg, "conv2d_transpose(p16)");
  if (rc) return rc;
  const int st = g.stride, Ck = C1 + C2;
  const int np = cfg.nprod == 3 ? 2 : 1;
  HParams p;
  h_fill_common(p, epi, y, yp, yp_n, (int64_t)g.N * g.H * g.W * Cn);
  struct Ph { int ph, pw, Hp, Wp, r0, s0, nr, ns, dh0, dw0, taps; };
  Ph phs[4];
  int nph = 0;
  for (int ph = 0; ph < st; ++ph)
    for (int pw = 0; pw < st; ++pw) {
      Ph q;
      q.ph = ph; q.pw = pw;
      h_phase_geom(g, ph, pw, &q.Hp, &q.Wp, &q.r0, &q.s0, &q.nr, &q.ns, &q.dh0, &q.dw0);
      q.taps = q.nr * q.ns;
      phs[nph++] = q;
    }
  for (int i = 1; i < nph; ++i)                         // heaviest phase first
    for (int j = i; j > 0 && phs[j].taps > phs[j - 1].taps; --j) { Ph t = phs[j]; phs[j] = phs[j - 1]; phs[j - 1] = t; }
  int tpi;
  if (!h_tile_grid(phs[0].Hp, phs[0].Wp, 128, &p.BP, &p.NB, &tpi)) { set_error("conv2d_transpose(p16): grid does not tile"); return CTX_ERR_UNSUPPORTED; }
  p.tiles_per_img = tpi;
  const int m_tiles = p.NB > 1 ? (int)cdiv(g.N, p.NB) : g.N * tpi;
  p.BN = h_pick_bn(Cn);
  const bool pair = h_pair(p.BN, m_tiles, cfg.nprod);
  p.NG = g.N; p.PG = phs[0].Hp; p.QG = phs[0].Wp; p.N = Cn;
  p.a_wrap = x2_batch > 0 ? x2_batch : 1;
  {
    const int64_t plane = (int64_t)g.N * g.P * g.Q * C1;
    int64_t dims[5] = {C1, g.Q, g.P, g.N, np}, str[5] = {1, C1, (int64_t)g.Q * C1, (int64_t)g.P * g.Q * C1, plane};
    int box[5] = {64, phs[0].Wp, p.BP, p.NB, np};
    rc = h16_make_map(&p.mapA[0], cfg.f16, x1p, 5, dims, str, box);
    if (rc) return rc;
  }
  if (C2 > 0) {
    const int64_t plane = (int64_t)x2_batch * g.P * g.Q * C2;
    int64_t dims[5] = {C2, g.Q, g.P, x2_batch, np}, str[5] = {1, C2, (int64_t)g.Q * C2, (int64_t)g.P * g.Q * C2, plane};
    int box[5] = {64, phs[0].Wp, p.BP, p.NB, np};
    rc = h16_make_map(&p.mapA[1], cfg.f16, x2p, 5, dims, str, box);
    if (rc) return rc;
  }
  {
    const int64_t rows = (int64_t)g.kh * g.kw * Cn;
    int64_t dims[3] = {Ck, rows, np}, str[3] = {1, Ck, rows * Ck};
    int box[3] = {64, pair ? p.BN / 2 : p.BN, np};
    rc = h16_make_map(&p.mapB[0], cfg.f16, wp, 3, dims, str, box);
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
          if (seg0 + nseg >= kHMaxSeg) { set_error("conv2d_transpose(p16): too many taps"); return CTX_ERR_UNSUPPORTED; }
          HSeg& sg = p.seg[seg0 + nseg++];
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
  if (h_tma_out_enabled()) {
    for (int i = 0; i < nph; ++i) {
      rc = h_out_maps(p, i, y, yp, yp_n, cfg.f16, p.ph_obase[i], Cn, phs[i].Wp, phs[i].Hp, g.N, p.sQ, p.sP, p.sN, p.plane_stride);
      if (rc) return rc;
    }
    p.tma_out = 1;
  }
  return h_launch(p, m_tiles, nph, cfg, stream, "h16_conv2d_transpose");
}

// ---- transposed conv, stride 2, few output channels (4 * Cn == 256 or 512): the four stride phases merged along N -------------
// A 128 x 64 tile moves 48 KB of operands per k-block for 1/4 of the MACs of a 256-wide pair tile, and the per-phase
// formulation re-reads every input pixel once per filter tap (25x for 5x5).  Here ONE GEMM covers all four phases:
// rows = input pixels (p, q), columns = (phase, co), K = (distinct input offset (dh, dw), source, channel).  Phase
// (ph, pw) at offset (dh, dw) uses exactly one tap (r, s) or none; the packed filter matrix Wm[offset][phase * Cn + co][c]
// holds that tap's weights or zeros (ctx_conv2d_transpose_merged_pack).  The input is fetched once per distinct offset
// (9 for 5x5 instead of 25), every tile is a 256-wide CTA pair, and the zero blocks cost MMA time only (36 k-block
// columns issued for 25 needed).
struct HMergedPlan {
  int nh, nw;                 // distinct row / column offsets
  int dh[3], dw[3];
  int tap_r[2][3], tap_s[2][3];   // [phase bit][offset index] -> filter row / column, or -1
};

static bool h_merged_plan(const ctx_conv_geom& g, HMergedPlan* mp) {
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

int h16_conv2d_transpose_merged_eligible(int C1, int C2, int x2_batch, int Cn, const ctx_conv_geom& g) {
  HMergedPlan mp;
  int bp, nb, tpi;
  if ((4 * Cn != 256 && 4 * Cn != 512) || C1 % 64 != 0 || C2 % 64 != 0) return 0;
  if (!h_merged_plan(g, &mp)) return 0;
  if (!h_tile_grid(g.P, g.Q, 128, &bp, &nb, &tpi)) return 0;
  if (C2 > 0 && nb > 1 && x2_batch % nb != 0) return 0;
  return mp.nh * mp.nw * (C2 > 0 ? 2 : 1) <= kHMaxSeg ? 1 : 0;
}

__global__ void h16_merged_pack_kernel(const float* __restrict__ w, int kw, int Cn, int Ck, HMergedPlan mp, void* planes,
                                       int nplanes, int f16, float scale) {
  // planes [nplanes][noff * 4 * Cn][Ck]; one thread per (row, 4 channels)
  const int64_t rows = (int64_t)mp.nh * mp.nw * 4 * Cn, total = rows * (Ck / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / (Ck / 4);
    const int c4 = (int)(i - row * (Ck / 4)) * 4;
    const int off = (int)(row / (4 * Cn)), rem = (int)(row - (int64_t)off * 4 * Cn);
    const int phase = rem / Cn, co = rem - phase * Cn;
    const int r = mp.tap_r[phase >> 1][off / mp.nw], s_ = mp.tap_s[phase & 1][off % mp.nw];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= 0 && s_ >= 0) v = *reinterpret_cast<const float4*>(w + (((int64_t)r * kw + s_) * Cn + co) * Ck + c4);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    const uint32_t h0 = h16::pack16(v.x, v.y, f16 != 0), h1 = h16::pack16(v.z, v.w, f16 != 0);
    uint16_t* hi = reinterpret_cast<uint16_t*>(planes) + row * Ck + c4;
    *reinterpret_cast<uint2*>(hi) = make_uint2(h0, h1);
    if (nplanes == 2) {
      const uint32_t l0 = h16::pack16(v.x - h16::unpack16_lo(h0, f16 != 0), v.y - h16::unpack16_hi(h0, f16 != 0), f16 != 0);
      const uint32_t l1 = h16::pack16(v.z - h16::unpack16_lo(h1, f16 != 0), v.w - h16::unpack16_hi(h1, f16 != 0), f16 != 0);
      *reinterpret_cast<uint2*>(hi + rows * Ck) = make_uint2(l0, l1);
    }
  }
}

int64_t h16_conv2d_transpose_merged_rows(int Cn, const ctx_conv_geom& g) {
  HMergedPlan mp;
  if (!h_merged_plan(g, &mp)) return 0;
  return (int64_t)mp.nh * mp.nw * 4 * Cn;
}

// w [kh,kw,Cn,Ck] fp32 -> packed planes [nplanes][rows][Ck], rows = h16_conv2d_transpose_merged_rows
int h16_conv2d_transpose_merged_pack(const float* w, int Ck, int Cn, const ctx_conv_geom& g, float scale, void* planes,
                                     int nplanes, int f16, cudaStream_t stream) {
  HMergedPlan mp;
  CTX_REQUIRE(h_merged_plan(g, &mp) && Ck % 4 == 0, "conv2d_transpose_merged_pack: geometry not eligible");
  CTX_REQUIRE(h_aligned16(w) && h_aligned16(planes), "conv2d_transpose_merged_pack: pointers must be 16-byte aligned");
  CTX_REQUIRE(nplanes == 1 || nplanes == 2, "conv2d_transpose_merged_pack: nplanes must be 1 or 2");
  const int64_t total = (int64_t)mp.nh * mp.nw * 4 * Cn * (Ck / 4);
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  h16_merged_pack_kernel<<<blocks, 256, 0, stream>>>(w, g.kw, Cn, Ck, mp, planes, nplanes, f16, scale);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "merged_pack");
  count_launch();
  return CTX_OK;
}

int h16_conv2d_transpose_merged(const void* x1p, int C1, const void* x2p, int C2, int x2_batch, const void* wmp, float* y,
                                void* yp, int yp_n, int Cn, const ctx_conv_geom& g, const Epi& epi, const H16Cfg& cfg,
                                cudaStream_t stream) {
  CTX_REQUIRE(h_aligned16(x1p) && h_aligned16(x2p) && h_aligned16(wmp) && h_aligned16(y) && h_aligned16(yp),
              "conv2d_transpose_merged(p16): pointers must be 16-byte aligned");
  CTX_REQUIRE(!epi.round, "conv2d_transpose_merged(p16): no TF32 rounding in this path");
  CTX_REQUIRE(!yp || yp_n == 1 || yp_n == 2, "conv2d_transpose_merged(p16): y_planes_n must be 1 or 2");
  int rc = h_nprod_ok(cfg, "conv2d_transpose_merged(p16)");
  if (rc) return rc;
  HMergedPlan mp;
  if (!h_merged_plan(g, &mp)) { set_error("conv2d_transpose_merged(p16): geometry not eligible"); return CTX_ERR_UNSUPPORTED; }
  const int Ck = C1 + C2;
  const int np = cfg.nprod == 3 ? 2 : 1;
  HParams p;
  h_fill_common(p, epi, y, yp, yp_n, (int64_t)g.N * g.H * g.W * Cn);
  int tpi;
  if (!h_tile_grid(g.P, g.Q, 128, &p.BP, &p.NB, &tpi)) { set_error("conv2d_transpose_merged(p16): grid does not tile"); return CTX_ERR_UNSUPPORTED; }
  p.tiles_per_img = tpi;
  const int m_tiles = p.NB > 1 ? (int)cdiv(g.N, p.NB) : g.N * tpi;
  p.BN = 256;
  const bool pair = m_tiles >= 2;
  p.NG = g.N; p.PG = g.P; p.QG = g.Q; p.N = 4 * Cn;
  p.a_wrap = x2_batch > 0 ? x2_batch : 1;
  p.merge_cn = Cn; p.merge_W = g.W;
  {
    const int64_t plane = (int64_t)g.N * g.P * g.Q * C1;
    int64_t dims[5] = {C1, g.Q, g.P, g.N, np}, str[5] = {1, C1, (int64_t)g.Q * C1, (int64_t)g.P * g.Q * C1, plane};
    int box[5] = {64, g.Q, p.BP, p.NB, np};
    rc = h16_make_map(&p.mapA[0], cfg.f16, x1p, 5, dims, str, box);
    if (rc) return rc;
  }
  if (C2 > 0) {
    const int64_t plane = (int64_t)x2_batch * g.P * g.Q * C2;
    int64_t dims[5] = {C2, g.Q, g.P, x2_batch, np}, str[5] = {1, C2, (int64_t)g.Q * C2, (int64_t)g.P * g.Q * C2, plane};
    int box[5] = {64, g.Q, p.BP, p.NB, np};
    rc = h16_make_map(&p.mapA[1], cfg.f16, x2p, 5, dims, str, box);
    if (rc) return rc;
  }
  {
    const int64_t rows = (int64_t)mp.nh * mp.nw * 4 * Cn;
    int64_t dims[3] = {Ck, rows, np}, str[3] = {1, Ck, rows * Ck};
    int box[3] = {64, pair ? 128 : 256, np};
    rc = h16_make_map(&p.mapB[0], cfg.f16, wmp, 3, dims, str, box);
    if (rc) return rc;
  }
  int nseg = 0, kblocks = 0;
  for (int a = 0; a < mp.nh; ++a)
    for (int b = 0; b < mp.nw; ++b)
      for (int src = 0; src < (C2 > 0 ? 2 : 1); ++src) {
        HSeg& sg = p.seg[nseg++];
        sg.mapA = (uint8_t)src; sg.wrap = (uint8_t)src;
        sg.dh = (int16_t)mp.dh[a]; sg.dw = (int16_t)mp.dw[b];
        sg.nblk = (uint16_t)((src ? C2 : C1) / 64);
        sg.bk0 = src ? C1 : 0; sg.br0 = (a * mp.nw + b) * 4 * Cn;
        kblocks += sg.nblk;
      }
  p.ph_seg0[0] = 0; p.ph_nseg[0] = nseg; p.ph_kblocks[0] = kblocks; p.ph_obase[0] = 0;
  p.nphase = 1;
  p.sN = (int64_t)g.H * g.W * Cn; p.sP = 2 * (int64_t)g.W * Cn; p.sQ = 2 * (int64_t)Cn;
  if (h_tma_out_enabled()) {
    for (int ph = 0; ph < 4; ++ph) {
      rc = h_out_maps(p, ph, y, yp, yp_n, cfg.f16, ((int64_t)(ph >> 1) * g.W + (ph & 1)) * Cn, Cn, g.Q, g.P, g.N, p.sQ, p.sP,
                      p.sN, p.plane_stride);
      if (rc) return rc;
    }
    p.tma_out = 1;
  }
  return h_launch(p, m_tiles, 1, cfg, stream, "h16_conv2d_transpose_merged");
}

// ---- conv filter gradient (MN-major operands, split-K) --------------------------------------------------------------------------
int h16_conv2d_wgrad_eligible(int Ca, int Cb1, int Cb2, int s2_batch, const ctx_conv_geom& g) {
  int bp, nb, tpi;
  if (Ca % 64 != 0 || Cb1 % 64 != 0 || Cb2 % 64 != 0) return 0;
  if (g.stride != 1 && g.stride != 2) return 0;
  if (g.kh * g.kw > kHMaxSeg) return 0;
  if (!h_tile_grid(g.P, g.Q, 64, &bp, &nb, &tpi)) return 0;
  if (Cb2 > 0 && nb > 1 && s2_batch % nb != 0) return 0;
  if (g.stride == 2 && (g.H < 2 || g.W < 2)) return 0;
  if ((int64_t)g.N * g.P * g.Q >= (1ll << 31)) return 0;
  return 1;
}

// Split-K factor: `occ` CTAs per SM are resident, so the launch runs in waves; minimise waves x (k-blocks per CTA + fixed cost)
static int h_pick_splits(int tiles, int kblocks, int min_per_split, int occ) {
  const int sms = sm_count() * occ;
  int maxs = kblocks / min_per_split;
  if (maxs < 1) maxs = 1;
  if (maxs > 4 * sms) maxs = 4 * sms;
  const int64_t ovh = 6;
  int best = 1;
  int64_t best_cost = -1;
  for (int s = 1; s <= maxs; ++s) {
    const int64_t per = cdiv(kblocks, s);
    const int64_t real = cdiv(kblocks, per);
    const int64_t waves = cdiv((int64_t)tiles * real, sms);
    const int64_t cost = waves * (per + ovh);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
  }
  return best;
}

// big planes [.][N,H,W,Ca] (cfg.a_scale applies to it: it is the gradient operand of a conv filter gradient and the dY of a
// transposed-conv one), s1 planes [.][N,P,Q,Cb1], s2 planes [.][s2_batch,P,Q,Cb2] or null; dw [kh,kw,Ca,Cb1+Cb2] fp32 (+=)
int h16_conv2d_wgrad(const void* big, int Ca, const void* s1, int Cb1, const void* s2, int Cb2, int s2_batch, float* dw,
                     const ctx_conv_geom& g, const H16Cfg& cfg, cudaStream_t stream) {
  CTX_REQUIRE(h_aligned16(big) && h_aligned16(s1) && h_aligned16(dw) && h_aligned16(s2),
              "conv2d_wgrad(p16): pointers must be 16-byte aligned");
  int rc = h_nprod_ok(cfg, "conv2d_wgrad(p16)");
  if (rc) return rc;
  const int np = cfg.nprod == 3 ? 2 : 1;
  HParams p;
  Epi none = make_epi(nullptr);
  h_fill_common(p, none, dw, nullptr, 0, 0);
  int tpi;
  if (!h_tile_grid(g.P, g.Q, 64, &p.kBP, &p.kNB, &tpi)) { set_error("conv2d_wgrad(p16): grid does not tile"); return CTX_ERR_UNSUPPORTED; }
  p.k_tiles_per_img = tpi;
  p.mn = 1;
  const int N = Cb1 + Cb2;
  p.BN = h_pick_bn(N);
  p.kPG = g.P; p.kQG = g.Q;
  p.Ca = Ca; p.M = g.kh * g.kw * Ca; p.N = N;
  p.Cb1 = Cb2 > 0 ? Cb1 : (1 << 30);
  p.a_wrap = s2_batch > 0 ? s2_batch : 1;
  const int m_tiles = (int)cdiv(p.M, 128);
  const bool pair = p.BN == 256 && m_tiles >= 2;
  const int bn_cta = pair ? p.BN / 2 : p.BN;
  // atoms fetched per TMA instruction: A — both 64-channel atoms of a tile when they share a filter tap; B — as many
  // consecutive atoms as always come from ONE source
  p.a_box = (Ca % 128 == 0) ? 2 : 1;
  p.b_box = bn_cta / 64;
  while (p.b_box > 1 && Cb2 > 0 && (Cb1 % (64 * p.b_box) != 0)) p.b_box /= 2;
  // A: per parity view and plane, rank-5 {64 ch, W', H', N, Ca/64}
  const int64_t big_plane = (int64_t)g.N * g.H * g.W * Ca;
  for (int pl = 0; pl < np; ++pl) {
    const uint16_t* base0 = (const uint16_t*)big + pl * big_plane;
    if (g.stride == 1) {
      int64_t dims[5] = {64, g.W, g.H, g.N, Ca / 64}, str[5] = {1, Ca, (int64_t)g.W * Ca, (int64_t)g.H * g.W * Ca, 64};
      int box[5] = {64, g.Q, p.kBP, p.kNB, p.a_box};
      rc = h16_make_map(&p.mapA[4 * pl], cfg.f16, base0, 5, dims, str, box);
      if (rc) return rc;
    } else {
      for (int hp = 0; hp < 2; ++hp)
        for (int wq = 0; wq < 2; ++wq) {
          const int Hh = (g.H - hp + 1) / 2, Wh = (g.W - wq + 1) / 2;
          int64_t dims[5] = {64, Wh, Hh, g.N, Ca / 64};
          int64_t str[5] = {1, 2 * (int64_t)Ca, 2 * (int64_t)g.W * Ca, (int64_t)g.H * g.W * Ca, 64};
          int box[5] = {64, g.Q, p.kBP, p.kNB, p.a_box};
          rc = h16_make_map(&p.mapA[4 * pl + hp * 2 + wq], cfg.f16, base0 + ((int64_t)hp * g.W + wq) * Ca, 5, dims, str, box);
          if (rc) return rc;
        }
    }
  }
  // B: [pixels][Cb] matrices, rank-4 {64 ch, pixels, Cb/64, plane}
  {
    const int64_t rows = (int64_t)g.N * g.P * g.Q;
    int64_t dims[4] = {64, rows, Cb1 / 64, np}, str[4] = {1, Cb1, 64, rows * Cb1};
    int box[4] = {64, 64, p.b_box, 1};
    rc = h16_make_map(&p.mapB[0], cfg.f16, s1, 4, dims, str, box);
    if (rc) return rc;
  }
  if (Cb2 > 0) {
    const int64_t rows = (int64_t)s2_batch * g.P * g.Q;
    int64_t dims[4] = {64, rows, Cb2 / 64, np}, str[4] = {1, Cb2, 64, rows * Cb2};
    int box[4] = {64, 64, p.b_box, 1};
    rc = h16_make_map(&p.mapB[1], cfg.f16, s2, 4, dims, str, box);
    if (rc) return rc;
  }
  for (int r = 0; r < g.kh; ++r)
    for (int s = 0; s < g.kw; ++s) {
      HSeg& sg = p.seg[r * g.kw + s];
      const int rr = r - g.pad_h, ss = s - g.pad_w;
      if (g.stride == 1) { sg.mapA = 0; sg.dh = (int16_t)rr; sg.dw = (int16_t)ss; }
      else {
        const int hp = rr & 1, wq = ss & 1;
        sg.mapA = (uint8_t)(hp * 2 + wq); sg.dh = (int16_t)((rr - hp) / 2); sg.dw = (int16_t)((ss - wq) / 2);
      }
    }
  p.total_kblocks = p.kNB > 1 ? (int)cdiv(g.N, p.kNB) : g.N * tpi;
  const int n_tiles = (int)cdiv(N, p.BN);
  const int splits = h_pick_splits((pair ? (m_tiles + 1) / 2 * 2 : m_tiles) * n_tiles, p.total_kblocks, 8, cfg.nprod == 1 ? 2 : 1);
  p.kblocks_per_split = (int)cdiv(p.total_kblocks, splits);
  const int real_splits = (int)cdiv(p.total_kblocks, p.kblocks_per_split);
  p.atomic_out = 1;
  // D rows are dw rows, addressed linearly: a 128-"image" tile of 1x1 grids
  p.NG = p.M; p.PG = 1; p.QG = 1; p.BP = 1; p.NB = 128; p.tiles_per_img = 1;
  p.sN = N; p.sP = 0; p.sQ = 0;
  return h_launch(p, m_tiles, real_splits, cfg, stream, "h16_conv2d_wgrad");
}

}  // namespace ctx
