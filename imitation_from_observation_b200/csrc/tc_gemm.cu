// tcgen05 / TMEM / TMA implicit-GEMM kernels for sm_100a: conv fprop, transposed conv (per stride phase), conv
// filter gradient and dense GEMM, all through ONE warp-specialised kernel.
//
//   D[128 x BN] (+)= A[128 x K] * B[BN x K]^T        fp32 accumulate in TMEM, operands TF32
//
// * Operands are staged in shared memory by TMA (cp.async.bulk.tensor, 128-byte swizzle) — im2col is never
//   materialised: an A tile of a convolution is ONE 4-D TMA box {32 channels, BQ cols, BP rows, NB images} taken
//   at the filter tap's (dh, dw) offset; out-of-range coordinates are zero-filled by the TMA unit, which is
//   exactly TF SAME padding.  Stride-2 convolutions read through four "parity views" of the input
//   (base + (hp*W+wp)*C, strides doubled), so every tap is again a dense box.
// * The channel concatenation of a decoder level ([state ‖ skip], arm_shaping.py:1323-1329) is two K segments
//   with two tensor maps — it never exists in memory.
// * Any P x Q pixel grid: exact 128-pixel tiles when the grid allows, else partial boxes BQ x BP x NB <= 128 pixels
//   (column tiles for wide maps) whose unused MMA rows are masked by the epilogue; a ragged last 32-channel block is
//   zero-filled by the TMA unit.  A transposed conv with few output channels (the 3-channel image) is ONE GEMM over
//   (tap, channel) columns whose epilogue scatter-adds (col2im).
// * Filter gradients use MN-major operands (pixels are the K dimension): the same pixel boxes, four 32-channel
//   atoms per 128-row tile fetched with ONE rank-5 box (one TMA instruction costs the TMA unit a few hundred cycles
//   whatever its size); split-K over pixel blocks sized per wave, reduced with vector fp32 atomics.
// * Precision: single pass = plain TF32 (the hardware truncates fp32 operands: producers round to
//   nearest, weights come from a rounded shadow copy).  3-pass = the fp32-faithful split
//   a*b ~= hi(a)*hi(b) + hi(a)*lo(b) + lo(a)*hi(b): eight "splitter" warps turn each landed fp32 tile into (hi, lo);
//   A's (hi, lo) go to TENSOR MEMORY (tcgen05.st, 4-deep ring) and are consumed by TS-form MMAs, B's lo goes to a
//   small shared-memory ring; three tcgen05.mma per K step into two accumulators (the small cross terms must not
//   ride on the big sum: the tensor core adds into its fp32 accumulator with truncation).  ~2^-21 per product.
//
// Three instantiations tc_igemm_kernel<pair, 3-pass>:
//   <0,0> single pass, two CTAs per SM (they overlap each other's prologue / epilogue / pipeline-fill bubbles);
//   <0,1> 3-pass, one CTA per SM (two for narrow or very short tiles);
//   <1,1> 3-pass CTA PAIR (cta_group::2, cluster of 2 along M): one 256 x 128 MMA per step for two SMs; each CTA
//         stages and splits its own 128 rows of A and HALF of the B tile; the follower forwards "my tiles are split"
//         to the leader with a relaxed remote mbarrier arrive; commits are multicast to both CTAs.
//
// Warp roles (320 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + warp-convergent MMA issue
// (elect.sync-predicated tcgen05.mma, vote-based barrier waits), warps 2..9 = splitters during the main loop, then the
// epilogue (two warps per TMEM lane quarter: TMEM -> registers -> fused bias / lrelu / lrelu' / accumulate / TF32
// rounding -> 128-bit stores, vector atomics for split-K, or scatter-add).
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <string>
#include <string.h>
#include <stdlib.h>
#include "kernels.h"

namespace ctx {

// ------------------------------------------------------------------------------------------------------------
// parameters
// ------------------------------------------------------------------------------------------------------------
enum { A_PIX = 0, A_MN = 1 };     // A operand: K-major pixel boxes | MN-major (4 atoms of 32 channels)
enum { B_K = 0, B_MN = 1 };       // B operand: K-major 2-D rows | MN-major 2-D (BN/32 atoms)

constexpr int kMaxSeg = 64;
constexpr int kMaxMapsA = 8;
constexpr int kThreads = 320;     // warp 0 TMA, warp 1 MMA, warps 2..9 splitters + epilogue
constexpr int kWorkers = 256;     // splitter / epilogue threads (two warps per tensor-memory lane quarter)

struct Seg {                       // one K segment = one filter tap of one source tensor
  int16_t dh, dw;                  // pixel offset of the A box
  uint8_t mapA;                    // which tensor map
  uint8_t wrap;                    // source 2: image index taken modulo a_wrap
  uint16_t nblk;                   // number of 32-wide K blocks in this segment
  int32_t bk0;                     // B: K coordinate of the segment's first block
  int32_t br0;                     // B (K-major only): row offset added to the tile's n0
};

struct TcParams {
  CUtensorMap mapA[kMaxMapsA];
  CUtensorMap mapB[2];
  Seg seg[kMaxSeg];
  int32_t nseg;
  int32_t a_mode, b_mode;
  int32_t passes;                  // 1 | 3
  int32_t BN;                      // 64 | 128 | 256
  int32_t raw_slots, lo_slots;     // ring depths: fp32 tiles landed by TMA | their lo parts (3-pass only)
  int32_t pair;                    // cta_group::2: two CTAs (a cluster of 2 along the M tiles) run ONE 256 x BN MMA; each
                                   // CTA stages its own 128 rows of A and HALF of the B tile (see the kernel)
  int32_t a_tmem;                  // 3-pass, K-major A: the splitters put A (hi, lo) into TENSOR MEMORY, the MMAs read
                                   // A from there — shared memory then only serves B (see the kernel header)
  int32_t total_kblocks;           // whole K loop (all splits)
  int32_t kblocks_per_split;
  // stride phases of a transposed conv merged into one launch (blockIdx.z = phase; no split-K then)
  int32_t nphase;
  int32_t ph_seg0[4], ph_nseg[4], ph_kblocks[4];
  int64_t ph_obase[4];
  // row grid of the A/D tile: rows r -> (dn, dp, dq)
  int32_t NG, PG, QG;              // grid extents (images, rows, cols); A_MN: NG = M, PG = QG = 1
  int32_t BP, NB;                  // tile = NB images x BP rows x QG cols  (== 128 rows), or 32 for A_MN k-blocks
  int32_t tiles_per_img;           // row tiles per image (PG / BP, rounded up) when NB == 1, else 1
  // partial tiles (conv fprop on grids that do not tile into 128 pixels): the A box is BQ x BP x NB <= 128 pixels
  // (a_rows), q_tiles column tiles per image row band; MMA rows past a_rows are dead and masked by the epilogue.
  // 0 = exact tiling (BQ = QG, one column tile, 128 rows).
  int32_t BQ, q_tiles, a_rows;
  int32_t a_wrap;                  // batch of the wrapped source
  // A_MN: K blocks are 32-pixel boxes of this pixel grid (images x kPG x kQG), kNB images x kBP rows each
  int32_t kPG, kQG, kBP, kNB, k_tiles_per_img;
  // A_MN (wgrad): per-atom tap decode
  int32_t Ca;                      // channels per tap on the M axis
  int32_t M;                       // valid rows of D
  int32_t N;                       // valid cols of D
  // B_MN second source (wgrad over [small1 ‖ small2])
  int32_t Cb1;                     // columns < Cb1 come from mapB[0], the rest from mapB[1] (pixel rows wrapped)
  // MN-major operands: 32-column atoms fetched per TMA instruction.  One instruction costs the TMA unit a few hundred
  // cycles whatever its size, so the atoms of a tile are fetched together through a rank+1 tensor map whose
  // outermost dimension walks the atoms (b_atoms0 / b_atoms1 for the two B sources; 0 = one instruction per atom).
  int32_t a_atoms, b_atoms0, b_atoms1;
  // output
  float* out;
  int64_t sN, sP, sQ, obase;       // element strides of D rows over (n, p, q)
  int32_t atomic_out;
  // split-K inside a thread-block CLUSTER (1 x 1 x cluster_k CTAs, one K range each): the partial accumulators of ranks
  // 1.. travel through distributed shared memory into rank 0's idle operand ring, rank 0 adds them to its own and runs
  // the fused epilogue — one launch and a deterministic sum where split-K over the grid needs a zero fill, atomics and
  // a second epilogue pass.  0 / 1 = off.
  int32_t cluster_k;
  // scatter epilogue (transposed conv with few output channels computed as ONE pixel x (tap, channel) GEMM):
  // column j = tap*sc_Cn + co of row (n, p, q) is added to out[n, p*sc_st + r - sc_padh, q*sc_st + s - sc_padw, co]
  int32_t scatter, sc_kw, sc_Cn, sc_st, sc_padh, sc_padw, sc_H, sc_W;
  Epi epi;
  int32_t* err_flag;
};

// ------------------------------------------------------------------------------------------------------------
// device helpers (raw PTX; names follow the PTX ISA)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// Bounded wait: a protocol bug must not hang the GPU (a hung box is a lost lease) — after ~2 s it records the
// barrier id and traps, which surfaces as a launch failure on the host.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int32_t* err_flag, int id) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      if (err_flag) atomicExch(err_flag, 0x1000 + id);
      __trap();
    }
  }
}
// Warp-uniform wait (all 32 lanes call it): the loop condition is a vote, so control flow stays convergent and
// the compiler can keep the caller's loop state in uniform registers.  Bounded like mbar_wait.
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, int32_t* err_flag, int id) {
  uint32_t spins = 0;
  while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
    if (++spins > 2000000u) {
      if (err_flag) atomicExch(err_flag, 0x1000 + id);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from tensor memory (rows = lanes, K along 32-bit columns), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 columns: thread t of the warp writes row (lane_base + t), columns col .. col+31
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Warp-convergent forms: every lane executes the statement, `lead` (non-zero in exactly one lane) predicates the
// tcgen05 instruction inside the asm.  The surrounding loop then stays convergent, so the compiler keeps descriptors
// and addresses in uniform registers instead of shuffling them through a per-lane waterfall for every MMA.
__device__ __forceinline__ void umma_tf32_p(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate, uint32_t lead) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(lead)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_ts_p(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate, uint32_t lead) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(lead)
      : "memory");
}
__device__ __forceinline__ void umma_commit_p(uint32_t bar, uint32_t lead) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(bar), "r"(lead)
      : "memory");
}
// ---- cta_group::2 (CTA pair) variants ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster.  Relaxed: the arrival only
// carries "my tiles are split" — the tiles themselves stay in this CTA's tensor / shared memory, are read by this
// SM's own MMA datapath, and were already ordered by the splitters' proxy fence + local barrier; a cluster-scope
// release here puts a full memory barrier per k-block on the follower's critical path (measured: the pair ran 20 % slower).
// 16 bytes into the shared memory of CTA `rank` of this cluster, at the address `addr` has in this CTA's window
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t rank, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
               "st.shared::cluster.v4.b32 [ra], {%2, %3, %4, %5};\n\t}" ::"r"(addr), "r"(rank), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(rank)
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait_warp_cluster(uint32_t bar, uint32_t parity, int32_t* err_flag, int id) {
  uint32_t spins = 0;
  while (!__all_sync(0xffffffffu, mbar_try_wait_cluster(bar, parity))) {
    if (++spins > 2000000u) {
      if (err_flag) atomicExch(err_flag, 0x1000 + id);
      __trap();
    }
  }
}
__device__ __forceinline__ void umma2_tf32_ts_p(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "@q tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit of pair MMAs: arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma2_commit_p(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t.reg .b16 m;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "mov.b16 m, 3;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
      ::"r"(bar)
      : "memory");
}

// descriptor words: lo = address[17:4] | LBO << 16 ; hi = SBO | version 1 << 14 | layout type << 29
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout_type << 29);
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 columns of fp32: thread t of the warp gets row (lane_base + t), columns col .. col+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (SM100 "version 1").
//   K-major : SWIZZLE_128B (16-byte chunks): rows of 128 B (32 tf32 along K), 8-row groups SBO = 1024 B apart;
//             LBO unused.
//   MN-major: 32-bit operands only exist as SWIZZLE_128B_BASE32B (32-byte chunks; TMA's SWIZZLE_128B_ATOM_32B):
//             atoms of [4 k-rows x 128 B (32 elements along M/N)]; the next 4 k-rows are SBO = 512 B further,
//             the next 32-element MN atom LBO bytes further.  One K=8 MMA therefore spans two k-groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61; // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}

// ------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------
constexpr uint32_t kATileBytes = 128 * 128;          // 128 rows (or 4 atoms x 32 rows) x 128 B

// TMEM columns to allocate: one accumulator (two in 3-pass mode) of BN columns, rounded up to a power of two >= 32
// (+ a_tmem stages x (A_hi, A_lo) x 32 columns when the A operand lives in tensor memory; a_tmem = lo ring depth)
__host__ __device__ __forceinline__ uint32_t tmem_cols(int BN, int passes, int a_tmem) {
  const uint32_t need = (uint32_t)BN * (passes == 3 ? 2u : 1u) + (a_tmem ? 64u * (uint32_t)a_tmem : 0u);
  uint32_t c = 32u;
  while (c < need) c <<= 1;
  return c;
}

// lo part of NC 16-byte chunks spaced kWorkers chunks apart: all loads first, then mask / subtract / store
template <int NC>
__device__ __forceinline__ void split_chunks(const float4* __restrict__ s4, float4* __restrict__ d4) {
  float4 v[NC];
#pragma unroll
  for (int u = 0; u < NC; ++u) v[u] = s4[u * kWorkers];
#pragma unroll
  for (int u = 0; u < NC; ++u) {
    float4 l;
    l.x = v[u].x - __uint_as_float(__float_as_uint(v[u].x) & 0xFFFFE000u);
    l.y = v[u].y - __uint_as_float(__float_as_uint(v[u].y) & 0xFFFFE000u);
    l.z = v[u].z - __uint_as_float(__float_as_uint(v[u].z) & 0xFFFFE000u);
    l.w = v[u].w - __uint_as_float(__float_as_uint(v[u].w) & 0xFFFFE000u);
    d4[u * kWorkers] = l;
  }
}

// kPair = true is the cta_group::2 build: a kernel that contains pair instructions can only be launched as a cluster
// of even size, so the CTA-pair variant is a separate instantiation.
template <bool kPair, bool kX3>
__global__ void __launch_bounds__(kThreads, 2) tc_igemm_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128-byte swizzle atoms
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // CTA pair: this CTA stages (and splits) only its half of the B tile
  constexpr bool pair = kPair;
  uint32_t cta_rank = 0u;
  if constexpr (kPair) cta_rank = cluster_ctarank();
  const uint32_t b_tile_bytes = (uint32_t)(pair ? p.BN / 2 : p.BN) * 128u;
  const uint32_t raw_bytes = kATileBytes + b_tile_bytes;            // fp32 tiles as landed by TMA
  // smem: [raw ring: R x (A | B)] [lo ring: L x (A_lo | B_lo)].  The raw ring is deep (it hides the TMA latency),
  // the lo ring only bridges splitter -> tensor core.
  const int R = p.raw_slots, L = p.lo_slots;
  const uint32_t lo_base = smem_base + (uint32_t)R * raw_bytes;
  // a lo slot holds (A_lo | B_lo), or only B_lo when A lives in tensor memory
  const uint32_t lo_bytes = p.a_tmem ? b_tile_bytes : raw_bytes;
  const uint32_t lo_b_off = p.a_tmem ? 0u : kATileBytes;

  // full[R], ready[L], raw_empty[R], lo_empty[L], acc, peer_ready[L] (pair leader: the other CTA's tiles are split)
  __shared__ __align__(8) uint64_t bars[8 + 4 + 8 + 4 + 1 + 4];
  __shared__ uint32_t tmem_slot;
  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto ready_bar = [&](int s) { return smem_u32(&bars[8 + s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[12 + s]); };
  auto lo_empty_bar = [&](int s) { return smem_u32(&bars[20 + s]); };
  const uint32_t acc_bar = smem_u32(&bars[24]);
  auto peer_ready_bar = [&](int s) { return smem_u32(&bars[25 + s]); };

  // this CTA's K range: split-K over blockIdx.z, or (merged transposed-conv phases) the whole K of phase z
  int kb_begin, kb_end, seg0 = 0, nseg = p.nseg;
  int64_t obase = p.obase;
  if (p.nphase > 1) {
    const int z = blockIdx.z;
    seg0 = p.ph_seg0[z]; nseg = p.ph_nseg[z]; obase = p.ph_obase[z];
    kb_begin = 0; kb_end = p.ph_kblocks[z];
  } else {
    kb_begin = blockIdx.z * p.kblocks_per_split;
    kb_end = min(p.total_kblocks, kb_begin + p.kblocks_per_split);
  }
  const int nkb = kb_end - kb_begin;
  // tile origin on the row grid
  const int tile_m = blockIdx.x;
  int n0, p0, q0 = 0;
  const int BQ = p.BQ > 0 ? p.BQ : p.QG, a_rows = p.a_rows > 0 ? p.a_rows : 128;
  if (p.q_tiles > 1) {
    const int per = p.tiles_per_img * p.q_tiles;
    n0 = tile_m / per;
    const int rem_t = tile_m - n0 * per, pt = rem_t / p.q_tiles;
    p0 = pt * p.BP; q0 = (rem_t - pt * p.q_tiles) * BQ;
  } else if (p.NB > 1 || p.tiles_per_img == 1) { n0 = tile_m * p.NB; p0 = 0; }
  else { n0 = tile_m / p.tiles_per_img; p0 = (tile_m - n0 * p.tiles_per_img) * p.BP; }
  const int ncol0 = blockIdx.y * p.BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < R; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < L; ++s) {
      mbar_init(ready_bar(s), kWorkers);
      mbar_init(lo_empty_bar(s), 1);
      mbar_init(peer_ready_bar(s), 1);
    }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    // 3-pass mode keeps two accumulators: D1 = hi*hi, D2 = the two cross terms.  The tensor core truncates when
    // it adds into a running fp32 sum (measured: error grows linearly with the number of accumulations), so the
    // small terms must not ride on the big sum; D1 + D2 is formed once, in the epilogue, with a rounded fp32 add.
    const uint32_t ncols = tmem_cols(p.BN, kX3 ? 3 : 1, p.a_tmem);
    if constexpr (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                   "r"(ncols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                   "r"(ncols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();         // both CTAs' barriers exist before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (nkb > 0) {
    if (warp == 0) {
      // =============================== TMA producer (one lane) =================================================
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        if (p.a_mode == A_PIX) {
          // enumerate (segment, channel block) pairs; skip to kb_begin
          int kb = 0;
          for (int sg = 0; sg < nseg && kb < kb_end; ++sg) {
            const Seg s = p.seg[seg0 + sg];
            if (kb + s.nblk <= kb_begin) { kb += s.nblk; continue; }
            for (int cb = 0; cb < s.nblk && kb < kb_end; ++cb, ++kb) {
              if (kb < kb_begin) continue;
              mbar_wait(empty_bar(stage), phase ^ 1, p.err_flag, 1);
              const uint32_t sa = smem_base + stage * raw_bytes;
              const uint32_t sb = sa + kATileBytes;
              mbar_arrive_expect_tx(full_bar(stage), (uint32_t)a_rows * 128u + b_tile_bytes);
              const int nimg = s.wrap ? (n0 % p.a_wrap) : n0;
              tma_load_4d(sa, &p.mapA[s.mapA], full_bar(stage), cb * 32, q0 + s.dw, p0 + s.dh, nimg);
              const int nb0 = ncol0 + (int)cta_rank * (p.BN / 2);      // pair: this CTA's half of the B tile
              if (p.b_mode == B_K) {
                tma_load_2d(sb, &p.mapB[0], full_bar(stage), s.bk0 + cb * 32, s.br0 + nb0);
              } else if (p.b_atoms0 > 0) {
                tma_load_3d(sb, &p.mapB[0], full_bar(stage), 0, s.bk0 + cb * 32, nb0 >> 5);
              } else {
                for (int a = 0; a < p.BN / 32; ++a)
                  tma_load_2d(sb + a * 4096, &p.mapB[0], full_bar(stage), ncol0 + a * 32, s.bk0 + cb * 32);
              }
              if (++stage == R) { stage = 0; phase ^= 1; }
            }
          }
        } else {
          // A_MN: K blocks are 32-pixel boxes; the 128 tile rows are 4 atoms (tap, 32 channels)
          const int m0 = tile_m * 128;
          for (int kb = kb_begin; kb < kb_end; ++kb) {
            int kn0, kp0;
            if (p.kNB > 1 || p.k_tiles_per_img == 1) { kn0 = kb * p.kNB; kp0 = 0; }
            else { kn0 = kb / p.k_tiles_per_img; kp0 = (kb - kn0 * p.k_tiles_per_img) * p.kBP; }
            mbar_wait(empty_bar(stage), phase ^ 1, p.err_flag, 2);
            const uint32_t sa = smem_base + stage * raw_bytes;
            const uint32_t sb = sa + kATileBytes;
            int natomA = 0;
            for (int j = 0; j < 4; ++j) natomA += (m0 + 32 * j < p.M) ? 1 : 0;
            mbar_arrive_expect_tx(full_bar(stage), (uint32_t)natomA * 4096u + b_tile_bytes);
            const int astep = p.a_atoms > 1 ? p.a_atoms : 1;
            for (int j = 0; j < 4; j += astep) {
              const int m = m0 + 32 * j;
              if (m >= p.M) break;
              const int tap = m / p.Ca, a0 = m - tap * p.Ca;
              const Seg s = p.seg[tap];
              if (p.a_atoms > 1)
                tma_load_5d(sa + j * 4096, &p.mapA[s.mapA], full_bar(stage), 0, s.dw, kp0 + s.dh, kn0, a0 >> 5);
              else
                tma_load_4d(sa + j * 4096, &p.mapA[s.mapA], full_bar(stage), a0, s.dw, kp0 + s.dh, kn0);
            }
            const int prow = (kn0 * p.kPG + kp0) * p.kQG;               // first pixel row of this K block
            const int prow2 = ((kn0 % p.a_wrap) * p.kPG + kp0) * p.kQG;
            if (p.b_atoms0 > 0) {
              if (ncol0 < p.Cb1 || p.N <= p.Cb1) {
                tma_load_3d(sb, &p.mapB[0], full_bar(stage), 0, prow, ncol0 >> 5);
                if (p.b_atoms1 > 0 && ncol0 + 32 * p.b_atoms0 < ncol0 + p.BN && p.N > p.Cb1)
                  tma_load_3d(sb + p.b_atoms0 * 4096, &p.mapB[1], full_bar(stage), 0, prow2, 0);
              } else {
                tma_load_3d(sb, &p.mapB[1], full_bar(stage), 0, prow2, (ncol0 - p.Cb1) >> 5);
              }
            } else {
              for (int a = 0; a < p.BN / 32; ++a) {
                const int n = ncol0 + a * 32;
                if (n < p.Cb1 || p.N <= p.Cb1) {       // (columns past N are zero-filled by the TMA unit)
                  tma_load_2d(sb + a * 4096, &p.mapB[0], full_bar(stage), n, prow);
                } else {
                  tma_load_2d(sb + a * 4096, &p.mapB[1], full_bar(stage), n - p.Cb1, prow2);
                }
              }
            }
            if (++stage == R) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      // =============================== MMA issuer (whole warp, one lane issues) ==================================
      if (kPair && cta_rank != 0) {
        // pair follower: the leader CTA issues the 256-row MMAs; this warp only forwards "my tiles are split"
        int lst = 0;
        uint32_t lphase = 0;
        for (int i = 0; i < nkb; ++i) {
          mbar_wait_warp(ready_bar(lst), lphase, p.err_flag, 7);
          if constexpr (kPair) { if (lane == 0) mbar_arrive_remote(peer_ready_bar(lst), 0u); }
          __syncwarp();
          if (++lst == L) { lst = 0; lphase ^= 1; }
        }
      } else {
        const uint32_t lead = lane == 0 ? 1u : 0u;
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((p.a_mode == A_MN ? 1u : 0u) << 15) |
                               ((p.b_mode == B_MN ? 1u : 0u) << 16) | ((uint32_t)(p.BN >> 3) << 17) |
                               (((pair ? 256u : 128u) >> 4) << 24);
        const bool amn = p.a_mode == A_MN, bmn = p.b_mode == B_MN;
        const uint32_t a_lbo = amn ? 4096u : 16u, a_step = amn ? 64u : 2u;     // steps in 16-byte units
        const uint32_t b_lbo = bmn ? 4096u : 16u, b_step = bmn ? 64u : 2u;
        const uint32_t a_hi = desc_hi(amn ? 512u : 1024u, amn ? 1u : 2u);
        const uint32_t b_hi = desc_hi(bmn ? 512u : 1024u, bmn ? 1u : 2u);
        const uint32_t d1 = tmem_base, d2 = tmem_base + (uint32_t)p.BN;
        constexpr bool x3 = kX3;
        const bool atm = kX3 && p.a_tmem != 0;
        // descriptor low words of ring slot 0; they advance by one slot per k-block (loop-carried, warp-uniform)
        const uint32_t raw16 = raw_bytes >> 4, lo16 = lo_bytes >> 4;
        const uint32_t a_base = desc_lo(smem_base, a_lbo), b_base = desc_lo(smem_base + kATileBytes, b_lbo);
        const uint32_t al_base = desc_lo(lo_base, a_lbo), bl_base = desc_lo(lo_base + lo_b_off, b_lbo);
        const uint32_t ta_base = tmem_base + 2u * (uint32_t)p.BN;
        uint32_t a0 = a_base, b0 = b_base, al0 = al_base, bl0 = bl_base, ta_hi = ta_base;
        int stage = 0, lst = 0;
        uint32_t phase = 0, lphase = 0;
        for (int i = 0; i < nkb; ++i) {
          if (x3) mbar_wait_warp(ready_bar(lst), lphase, p.err_flag, 3);
          else mbar_wait_warp(full_bar(stage), phase, p.err_flag, 3);
          if constexpr (kPair) mbar_wait_warp(peer_ready_bar(lst), lphase, p.err_flag, 8);
          tc_fence_after();
          const uint32_t first = i ? 1u : 0u;
          if constexpr (kPair) {
            // one 256 x BN MMA per step for the CTA pair: A from each CTA's tensor memory, B halves from each
            // CTA's shared memory (same offsets in both CTAs)
            const uint32_t ta_lo = ta_hi + 32u;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t acc = ks ? 1u : first;
              umma2_tf32_ts_p(d1, ta_hi + 8u * ks, b0 + ks * b_step, b_hi, idesc, acc);
              umma2_tf32_ts_p(d2, ta_hi + 8u * ks, bl0 + ks * b_step, b_hi, idesc, acc);
              umma2_tf32_ts_p(d2, ta_lo + 8u * ks, b0 + ks * b_step, b_hi, idesc, 1u);
            }
          } else if (atm) {
            // A (hi, lo) of this k-block sit in tensor memory columns [2*BN + 64*lst, +64): 8 columns per K = 8 step
            const uint32_t ta_lo = ta_hi + 32u;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t acc = ks ? 1u : first;
              umma_tf32_ts_p(d1, ta_hi + 8u * ks, b0 + ks * b_step, b_hi, idesc, acc, lead);
              umma_tf32_ts_p(d2, ta_hi + 8u * ks, bl0 + ks * b_step, b_hi, idesc, acc, lead);
              umma_tf32_ts_p(d2, ta_lo + 8u * ks, b0 + ks * b_step, b_hi, idesc, 1u, lead);
            }
          } else if (x3) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t acc = ks ? 1u : first;
              umma_tf32_p(d1, a0 + ks * a_step, a_hi, b0 + ks * b_step, b_hi, idesc, acc, lead);
              umma_tf32_p(d2, a0 + ks * a_step, a_hi, bl0 + ks * b_step, b_hi, idesc, acc, lead);
              umma_tf32_p(d2, al0 + ks * a_step, a_hi, b0 + ks * b_step, b_hi, idesc, 1u, lead);
            }
          } else {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_tf32_p(d1, a0 + ks * a_step, a_hi, b0 + ks * b_step, b_hi, idesc, ks ? 1u : first, lead);
          }
          if constexpr (kPair) umma2_commit_p(empty_bar(stage));     // (both CTAs' slots)
          else umma_commit_p(empty_bar(stage), lead);     // frees the raw slot once these MMAs have read it
          a0 += raw16; b0 += raw16;
          if (++stage == R) { stage = 0; phase ^= 1; a0 = a_base; b0 = b_base; }
          if (x3) {
            if constexpr (kPair) umma2_commit_p(lo_empty_bar(lst));
            else umma_commit_p(lo_empty_bar(lst), lead);  // ... and the lo slot
            al0 += lo16; bl0 += lo16; ta_hi += 64u;
            if (++lst == L) { lst = 0; lphase ^= 1; al0 = al_base; bl0 = bl_base; ta_hi = ta_base; }
          }
        }
        if constexpr (kPair) umma2_commit_p(acc_bar);
        else umma_commit_p(acc_bar, lead);                 // accumulator complete
      }
    } else if constexpr (kX3) {
      // =============================== splitters: fp32 tile -> (hi, lo) =======================================
      const int t = threadIdx.x - 64;              // 0..255
      const int arow = (warp & 3) * 32 + lane;     // tensor-memory lane quarter this warp may access = tile row
      const int ahalf = (warp - 2) >> 2;           // the two warps of a quarter take K columns 0..15 / 16..31
      // generic pointer to the 1024-byte aligned operand ring (plain loads / stores: the compiler schedules them
      // freely between the barrier waits, which are compiler memory barriers)
      uint8_t* const ring = smem_raw + (smem_base - smem_u32(smem_raw));
      const int a_chunks = p.a_tmem ? 0 : (int)(kATileBytes >> 4) / kWorkers;     // 16-byte chunks per thread
      const int b_chunks = (int)(b_tile_bytes >> 4) / kWorkers;                   // = BN / 32 (exact: BN % 32 == 0
      const int tail = (int)(b_tile_bytes >> 4) - b_chunks * kWorkers;            //   except BN = 80 -> tail chunks)
      int stage = 0, lst = 0;
      uint32_t phase = 0, lphase = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(full_bar(stage), phase, p.err_flag, 4);
        mbar_wait(lo_empty_bar(lst), lphase ^ 1, p.err_flag, 6);
        const uint8_t* src = ring + (size_t)stage * raw_bytes;
        uint8_t* dst = ring + (size_t)R * raw_bytes + (size_t)lst * lo_bytes;
        if (p.a_tmem) {
          // A: this thread owns half of tile row `arow` (64 B = 16 K-elements; 16-byte chunks XOR-swizzled by
          // row & 7); hi (= the raw word, the tensor core ignores the low 13 bits) and lo go to tensor memory
          tc_fence_after();
          uint32_t x[16];
          const uint4* row = reinterpret_cast<const uint4*>(src + (size_t)arow * 128);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 q4 = row[(4 * ahalf + c) ^ (arow & 7)];
            x[4 * c] = q4.x; x[4 * c + 1] = q4.y; x[4 * c + 2] = q4.z; x[4 * c + 3] = q4.w;
          }
          const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 2u * (uint32_t)p.BN +
                              64u * (uint32_t)lst + 16u * (uint32_t)ahalf;
          tmem_st16(ta, x);
#pragma unroll
          for (int e = 0; e < 16; ++e)
            x[e] = __float_as_uint(__uint_as_float(x[e]) - __uint_as_float(x[e] & 0xFFFFE000u));
          tmem_st16(ta + 32u, x);
          src += kATileBytes;                              // the shared-memory pass below only covers B
        }
        // the tensor core reads an fp32 word as TF32 by ignoring its low 13 bits (verified on B200:
        // tools/tc_check.py "trunc"), so the raw tile already IS the hi part; lo = x - trunc(x) is exact.
        // Elementwise, hence layout-agnostic: thread t takes chunks t, t+256, ... in branch-free batches.
        {
          const float4* s4 = reinterpret_cast<const float4*>(src) + t;
          float4* d4 = reinterpret_cast<float4*>(dst) + t;
          int left = a_chunks + b_chunks;
          for (; left >= 8; left -= 8, s4 += 8 * kWorkers, d4 += 8 * kWorkers) split_chunks<8>(s4, d4);
          if (left >= 4) { split_chunks<4>(s4, d4); s4 += 4 * kWorkers; d4 += 4 * kWorkers; left -= 4; }
          if (left >= 2) { split_chunks<2>(s4, d4); s4 += 2 * kWorkers; d4 += 2 * kWorkers; left -= 2; }
          if (left >= 1) { split_chunks<1>(s4, d4); s4 += kWorkers; d4 += kWorkers; }
          if (t < tail) split_chunks<1>(s4, d4);
        }
        if (p.a_tmem) { tmem_st_wait(); tc_fence_before(); }
        fence_async_smem();                        // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(ready_bar(lst));
        if (++stage == R) { stage = 0; phase ^= 1; }
        if (++lst == L) { lst = 0; lphase ^= 1; }
      }
    }
  }

  // ================================= cluster split-K: partial sums -> rank 0 ====================================
  uint32_t krank = 0u;
  if (p.cluster_k > 1) {
    krank = cluster_ctarank();
    if (nkb > 0) {
      mbar_wait(acc_bar, 0, p.err_flag, 9);        // (every thread: the accumulator and so the operand ring are final)
      tc_fence_after();
    }
    __syncwarp();
    cluster_sync_all();                            // every CTA's MMAs have read their rings: rank 0's is free to receive
    if (krank != 0 && warp >= 2) {
      // partial (rank - 1) as [column quad][row] float4: consecutive lanes write consecutive 16-byte words
      const int q = warp & 3, r = q * 32 + lane;
      const int nch = (p.BN + 31) >> 5, ch_half = (nch + 1) >> 1;
      const int c_begin = ((warp - 2) >> 2) ? ch_half * 32 : 0, c_end = ((warp - 2) >> 2) ? p.BN : min(p.BN, ch_half * 32);
      const uint32_t dst = smem_base + (krank - 1u) * (uint32_t)p.BN * 512u + (uint32_t)r * 16u;
      for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        uint32_t v[32];
        if (nkb > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
          if (kX3) {
            uint32_t v2[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.BN + c0), v2);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          st_cluster_v4(dst + (uint32_t)((c0 + j) >> 2) * 2048u, 0u, v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
    __syncwarp();
    cluster_sync_all();                            // (release / acquire: the partials are visible to rank 0)
  }

  // ================================= epilogue: warps 2..9 ======================================================
  // warp w reads TMEM lane quarter w % 4 (hardware rule); the two warps of a quarter split the 32-column chunks
  if (warp >= 2 && krank == 0u) {
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;                   // tile row
    const int nch = (p.BN + 31) >> 5, ch_half = (nch + 1) >> 1;
    const int c_begin = ((warp - 2) >> 2) ? ch_half * 32 : 0, c_end = ((warp - 2) >> 2) ? p.BN : min(p.BN, ch_half * 32);
    if (nkb > 0) {
      mbar_wait(acc_bar, 0, p.err_flag, 5);
      tc_fence_after();
    }
    // row -> (n, p, q) -> output offset
    const int per_img = p.BP * BQ;
    const int dn = r / per_img, rem = r - dn * per_img;
    const int dp = rem / BQ, dq = rem - dp * BQ;
    const int gn = n0 + dn, gp = p0 + dp, gq = q0 + dq;
    const bool row_ok = r < a_rows && gn < p.NG && gp < p.PG && gq < p.QG && (nkb > 0 || !p.atomic_out);
    const int64_t roff = obase + (int64_t)gn * p.sN + (int64_t)gp * p.sP + (int64_t)gq * p.sQ;
    for (int c0 = c_begin; c0 < c_end; c0 += 32) {
      uint32_t v[32];
      if (nkb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        if (kX3) {
          uint32_t v2[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.BN + c0), v2);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (p.cluster_k > 1) {                       // + the other ranks' partial sums, in rank order (deterministic)
        for (int s = 0; s < p.cluster_k - 1; ++s) {
          const uint32_t src = smem_base + (uint32_t)s * (uint32_t)p.BN * 512u + (uint32_t)r * 16u;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 t;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                         : "r"(src + (uint32_t)((c0 + j) >> 2) * 2048u));
            v[j] = __float_as_uint(__uint_as_float(v[j]) + t.x);
            v[j + 1] = __float_as_uint(__uint_as_float(v[j + 1]) + t.y);
            v[j + 2] = __float_as_uint(__uint_as_float(v[j + 2]) + t.z);
            v[j + 3] = __float_as_uint(__uint_as_float(v[j + 3]) + t.w);
          }
        }
      }
      if (!row_ok) continue;
      const int col = ncol0 + c0;
      if (col >= p.N) continue;
      float* optr = p.out + roff + col;
      if (p.scatter) {
        // (tap, channel) of the first column; all lanes walk the same columns, so this is warp-uniform
        int tap = col / p.sc_Cn, co = col - tap * p.sc_Cn;
        int rr = tap / p.sc_kw, ss = tap - rr * p.sc_kw;
        const int hb = gp * p.sc_st - p.sc_padh, wb = gq * p.sc_st - p.sc_padw;
        float* obase_n = p.out + (int64_t)gn * p.sc_H * p.sc_W * p.sc_Cn;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (col + j < p.N) {
            const int h = hb + rr, w = wb + ss;
            if ((unsigned)h < (unsigned)p.sc_H && (unsigned)w < (unsigned)p.sc_W)
              asm volatile("red.global.add.f32 [%0], %1;" ::"l"(obase_n + ((int64_t)h * p.sc_W + w) * p.sc_Cn + co),
                           "f"(__uint_as_float(v[j]))
                           : "memory");
          }
          if (++co == p.sc_Cn) { co = 0; if (++ss == p.sc_kw) { ss = 0; ++rr; } }
        }
        continue;
      }
      if (p.atomic_out) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (col + j + 3 < p.N) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(optr + j), "f"(__uint_as_float(v[j])),
                         "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                         : "memory");
          } else {
            for (int e = 0; e < 4; ++e)
              if (col + j + e < p.N) atomicAdd(optr + j + e, __uint_as_float(v[j + e]));
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (col + j + 3 < p.N) {
            float4 o;
            float* oo = reinterpret_cast<float*>(&o);
            float4 prev = make_float4(0.f, 0.f, 0.f, 0.f), ref = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.epi.accumulate) prev = *reinterpret_cast<const float4*>(optr + j);
            if (p.epi.act_ref) ref = __ldg(reinterpret_cast<const float4*>(p.epi.act_ref + roff + col + j));
            const float* pp = reinterpret_cast<const float*>(&prev);
            const float* rr = reinterpret_cast<const float*>(&ref);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x = __uint_as_float(v[j + e]);
              if (p.epi.bias) x += __ldg(p.epi.bias + col + j + e);
              if (p.epi.act) x = fmaxf(x, p.epi.slope * x);
              if (p.epi.act_ref) x *= (rr[e] >= 0.f ? 1.f : p.epi.slope);
              if (p.epi.accumulate) x += pp[e];
              if (p.epi.round) x = round_tf32(x);
              oo[e] = x;
            }
            *reinterpret_cast<float4*>(optr + j) = o;
          } else {
            for (int e = 0; e < 4; ++e)
              if (col + j + e < p.N)
                optr[j + e] = p.epi.apply(__uint_as_float(v[j + e]), col + j + e, roff + col + j + e, p.out);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();         // the peer may still be reading this CTA's tensor / shared memory
  if (warp == 1) {
    const uint32_t ncols = tmem_cols(p.BN, kX3 ? 3 : 1, p.a_tmem);
    if constexpr (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------
// host: tensor maps
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
    else
      cudaGetLastError();
  }
  return fn;
}

// fp32 tensor map, rank <= 5, zero fill; 128-byte swizzle of 16-byte chunks (K-major operand tiles) or of
// 32-byte chunks (mn_major: the only layout tcgen05 accepts for MN-major 32-bit operands).  dims/box innermost first; strides in ELEMENTS for
// dims 1..rank-1.
static int make_map(CUtensorMap* out, const float* base, int rank, const int64_t* dims, const int64_t* strides,
                    const int* box, bool mn_major = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return CTX_ERR_CUDA; }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = (cuuint64_t)dims[i];
    bx[i] = (cuuint32_t)box[i];
    es[i] = 1;
    if (dims[i] <= 0 || box[i] <= 0 || box[i] > 256) { set_error("tensor map: bad dim/box %d: %lld/%d", i, (long long)dims[i], box[i]); return CTX_ERR_ARG; }
  }
  for (int i = 1; i < rank; ++i) {
    gs[i - 1] = (cuuint64_t)strides[i] * 4u;
    if (gs[i - 1] % 16 != 0) { set_error("tensor map: stride %d (%lld elements) is not 16-byte aligned", i, (long long)strides[i]); return CTX_ERR_ARG; }
  }
  if (((uintptr_t)base & 15) != 0) { set_error("tensor map: base pointer is not 16-byte aligned"); return CTX_ERR_ARG; }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)base, gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return CTX_ERR_CUDA; }
  return CTX_OK;
}

bool tc_compiled() { return true; }

// Tile a (P x Q) pixel grid into boxes of exactly `rows` pixels: full-width rows, either BP rows of one image
// or NB whole images.
static bool tile_grid(int P, int Q, int rows, int* BP, int* NB, int* tiles_per_img) {
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

// Partial tiling of a P x Q pixel grid for conv fprop when it does not tile exactly into 128 pixels (InceptionV3's
// 17x17, 35x35, 71..149-wide maps): the A box is BQ x BP x NB <= 128 pixels.
static void plan_partial(int P, int Q, int* BQ, int* BP, int* NB, int* tiles_p, int* tiles_q) {
  if (Q <= 128 && P * Q <= 128) {                       // whole images per tile
    *BQ = Q; *BP = P; *NB = 128 / (P * Q); if (*NB < 1) *NB = 1; *tiles_p = 1; *tiles_q = 1;
    return;
  }
  // search the column split that wastes the fewest MMA rows: utilisation = P*Q / (tiles * 128)
  int best_tq = 0, best_tiles = 0, best_bq = 0, best_bp = 0, best_tp = 0;
  for (int tq = 1; tq <= 8; ++tq) {
    const int bq = (Q + tq - 1) / tq;
    if (bq > 128) continue;
    int bp = 128 / bq;
    if (bp > P) bp = P;
    const int tp = (P + bp - 1) / bp;
    bp = (P + tp - 1) / tp;                             // even out the row bands
    const int tiles = tp * tq;
    if (best_tq == 0 || tiles < best_tiles) { best_tq = tq; best_tiles = tiles; best_bq = bq; best_bp = bp; best_tp = tp; }
  }
  *BQ = best_bq; *BP = best_bp; *NB = 1; *tiles_p = best_tp; *tiles_q = best_tq;
}

// N tile: 128 for the 3-pass mode (with a 5-deep raw ring it is the best-balanced 1-CTA shape: tensor 768, splitter
// ~512, TMA ~42 B/cycle per k-block), 256 for single-pass; halved while the grid would not fill the SMs.
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}
// experiment knobs (environment, read once): CTX_TC_BN3 = N tile of the 3-pass mode, CTX_TC_OCC = CTAs per SM the
// shared-memory ring is sized for
static int knob_bn3() { static int v = env_int("CTX_TC_BN3", 128); return v; }
static int knob_lo_depth() { static int v = env_int("CTX_TC_LO_DEPTH", 4); return v < 1 ? 1 : (v > 4 ? 4 : v); }
static int knob_pair() { static int v = env_int("CTX_TC_PAIR", 1); return v; }
static int knob_atmem() { static int v = env_int("CTX_TC_ATMEM", 1); return v; }
static int knob_occ() { static int v = env_int("CTX_TC_OCC", 0); return v; }
// Cluster split-K for the skinny dense GEMMs is OFF by default: measured (profiles/r2_layer_gemm_cluster_k.txt) it loses to
// the wide-tile grid split — 84 vs 67 us for h4_lin's forward, 81 vs 62 us for d_h0_lin's data gradient: four times the
// CTAs at BN = 64 re-read A four times as often, and the two cluster barriers + the DSMEM hop cost more than the zero
// fill + second pass they replace.  CTX_TC_CLUSTER_K=1 turns it on (bit-reproducible sums).
static int knob_cluster_k() { static int v = env_int("CTX_TC_CLUSTER_K", 0); return v; }

// CTAs per SM the operand ring is sized for.  Two co-resident CTAs overlap each other's prologue / epilogue /
// pipeline-fill bubbles and double the TMA instructions in flight: measured +30..70 % for single-pass tiles
// (profiles/r1_layer_bench_*.txt).  The 3-pass mode is bound by shared-memory bandwidth at BN = 128 and needs its
// two-slot lo ring, so it keeps one CTA per SM except for narrow (BN <= 64) or very short (<= 8 k-blocks) tiles.
static int tc_occupancy(int passes, int BN, int kblocks) {
  const int k = knob_occ();
  if (k == 1 || k == 2) return k;
  if (passes != 3) return 2;
  if (BN > 128) return 1;                               // (2 x BN accumulator columns leave no room for a second CTA)
  return (BN <= 64 || kblocks <= 8) ? 2 : 1;
}

static int pick_bn(int N, int64_t m_tiles, int passes) {
  // 128 < N <= 192 (InceptionV3's 192-channel layers): one exact 192-wide tile instead of 128 + a half-empty 128
  // (fits the 3-pass TMEM budget: 2 x 192 accumulator columns + 2 stages of A)
  if (N > 128 && N <= 192 && N % 32 == 0 && m_tiles >= sm_count()) return 192;
  if (N <= 32) return 32;                               // (InceptionV3's 32-channel stem convs)
  int bn = passes == 3 ? knob_bn3() : 256;
  const int sms = sm_count();
  while (bn > 64 && (N <= bn / 2 || m_tiles * cdiv(N, bn) < sms)) bn /= 2;
  return bn;
}

// CTA pairs (cta_group::2) for the 3-pass K-major tiles: the pair runs one 256 x 128 MMA per step, each CTA staging and
// splitting only half of the B tile — the 1-CTA 3-pass tile is bound by shared-memory bandwidth, and B is the
// operand both CTAs would otherwise each stage, split and feed in full.  Decided before the B tensor map is built
// (its box is half as tall).
static bool want_pair(int passes, int a_mode, int BN, int N, int m_tiles, int kblocks) {
  return knob_pair() && knob_atmem() && passes == 3 && a_mode == A_PIX && BN == 128 && N % 128 == 0 && m_tiles >= 2 &&
         tc_occupancy(passes, BN, kblocks) == 1;
}

static int launch_tc(TcParams& p, int m_tiles, int splits, cudaStream_t stream, const char* name) {
  const int sms = sm_count();
  if (sms <= 0) return cuda_fail(cudaErrorNoDevice, name);
  const uint32_t bn_cta = (uint32_t)(p.pair ? p.BN / 2 : p.BN);          // B rows staged by one CTA
  const uint32_t raw = kATileBytes + bn_cta * 128u;
  // the opt-in shared-memory limit is a per-device function attribute: set it once for every device this process uses
  static int max_dyn_dev[64] = {0};
  int dev = 0;
  CTX_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { set_error("%s: device ordinal %d out of range", name, dev); return CTX_ERR_ARG; }
  if (max_dyn_dev[dev] == 0) {
    cudaFuncAttributes fa[3];
    CTX_CUDA(cudaFuncGetAttributes(&fa[0], tc_igemm_kernel<false, false>));
    CTX_CUDA(cudaFuncGetAttributes(&fa[1], tc_igemm_kernel<false, true>));
    CTX_CUDA(cudaFuncGetAttributes(&fa[2], tc_igemm_kernel<true, true>));
    int optin = 0;
    CTX_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    size_t stat = 0;
    for (int i = 0; i < 3; ++i) stat = fa[i].sharedSizeBytes > stat ? fa[i].sharedSizeBytes : stat;
    const int dyn = optin - (int)stat;
    CTX_CUDA(cudaFuncSetAttribute(tc_igemm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    CTX_CUDA(cudaFuncSetAttribute(tc_igemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    CTX_CUDA(cudaFuncSetAttribute(tc_igemm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
    max_dyn_dev[dev] = dyn;
  }
  const int max_dyn = max_dyn_dev[dev];
  const int occ = tc_occupancy(p.passes, p.BN, p.nphase > 1 ? p.ph_kblocks[0] : p.kblocks_per_split);
  const uint32_t budget = occ == 1 ? (uint32_t)max_dyn : (uint32_t)(max_dyn - 2048) / 2u;
  // 3-pass with a K-major A operand: A (hi, lo) goes through tensor memory when the columns are there
  // (2 accumulators + 2 stages x 64 columns; 512 columns per SM shared by the resident CTAs)
  const int tmem_budget = occ == 2 ? 256 : 512;
  const bool atm = knob_atmem() && p.passes == 3 && p.a_mode == A_PIX && !p.scatter && 2 * p.BN + 128 <= tmem_budget;
  if (p.pair && (!atm || occ != 1)) { set_error("%s: internal: CTA pair without tensor-memory A", name); return CTX_ERR_ARG; }
  const uint32_t lo_b = p.passes == 3 ? (atm ? bn_cta * 128u : raw) : 0u;
  if (atm) {
    // lo ring depth = tensor-memory A stages: as many as the columns allow (<= 4); a CTA pair needs the depth to
    // hide the follower -> leader "tiles are split" hop
    int depth = (tmem_budget - 2 * p.BN) / 64;
    if (depth > knob_lo_depth()) depth = knob_lo_depth();
    p.lo_slots = depth;
  } else {
    p.lo_slots = p.passes == 3 ? ((budget - 1024u) >= 2 * raw + 2 * lo_b ? 2 : 1) : 0;
  }
  p.a_tmem = atm ? p.lo_slots : 0;
  const int slots = (int)((budget - 1024u - (uint32_t)p.lo_slots * lo_b) / raw);
  p.raw_slots = slots;
  if (p.raw_slots > 8) p.raw_slots = 8;
  if (p.raw_slots < 2) { set_error("%s: tile does not fit shared memory", name); return CTX_ERR_ARG; }
  size_t smem = (size_t)p.raw_slots * raw + (size_t)p.lo_slots * lo_b + 1024;
  if (p.cluster_k > 1) {
    // rank 0 receives cluster_k - 1 partial tiles of 128 x BN fp32 in its (then idle) operand ring: make the allocation
    // large enough for them, within this occupancy's share of the SM
    const size_t need = (size_t)(p.cluster_k - 1) * p.BN * 512 + 1024;
    if (need > (size_t)budget) { set_error("%s: internal: cluster split-K partials do not fit shared memory", name); return CTX_ERR_ARG; }
    if (need > smem) smem = need;
  }
  p.err_flag = nullptr;
  const int n_tiles = (int)cdiv(p.N, p.BN);
  if (n_tiles > 65535 || splits > 65535) { set_error("%s: grid too large", name); return CTX_ERR_ARG; }
  if (p.pair) {
    // cluster of 2 CTAs along the M tiles (an odd last tile gets an all-out-of-range partner: TMA zero-fills it and
    // the epilogue masks its rows)
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)((m_tiles + 1) & ~1), (unsigned)n_tiles, (unsigned)splits);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, tc_igemm_kernel<true, true>, p);
    if (e != cudaSuccess) {
      int nclusters = -1;
      cudaError_t e2 = cudaOccupancyMaxActiveClusters(&nclusters, tc_igemm_kernel<true, true>, &cfg);
      cudaGetLastError();
      set_error("CUDA error in %s (CTA pair launch): %s; grid %u x %u x %u, smem %zu, max active clusters %d (query: %s)",
                name, cudaGetErrorName(e), cfg.gridDim.x, cfg.gridDim.y, cfg.gridDim.z, smem, nclusters,
                cudaGetErrorName(e2));
      return CTX_ERR_CUDA;
    }
    count_launch();
    return CTX_OK;
  }
  if (p.cluster_k > 1) {
    if (splits != p.cluster_k || p.atomic_out || p.nphase > 1 || p.scatter || p.BN % 32 != 0 ||
        (size_t)(p.cluster_k - 1) * p.BN * 512 > smem - 1024) {
      set_error("%s: internal: cluster split-K misconfigured", name);
      return CTX_ERR_ARG;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)m_tiles, (unsigned)n_tiles, (unsigned)splits);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = (unsigned)p.cluster_k;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = p.passes == 3 ? cudaLaunchKernelEx(&cfg, tc_igemm_kernel<false, true>, p)
                                  : cudaLaunchKernelEx(&cfg, tc_igemm_kernel<false, false>, p);
    if (e != cudaSuccess) return cuda_fail(e, name);
    count_launch();
    return CTX_OK;
  }
  dim3 grid((unsigned)m_tiles, (unsigned)n_tiles, (unsigned)splits);
  if (p.passes == 3) tc_igemm_kernel<false, true><<<grid, kThreads, smem, stream>>>(p);
  else tc_igemm_kernel<false, false><<<grid, kThreads, smem, stream>>>(p);
  CTX_LAUNCH_CHECK(name);
  return CTX_OK;
}

static void fill_common(TcParams& p, int passes, const Epi& epi) {
  memset(&p, 0, sizeof(p));
  p.passes = passes;
  p.nphase = 1;
  p.epi = epi;
  p.a_wrap = 1;
  p.Cb1 = 1 << 30;
}

static bool aligned16(const void* ptr) { return ((uintptr_t)ptr & 15) == 0; }

// ---- conv fprop ------------------------------------------------------------------------------------------------
int tc_conv2d_fprop_eligible(int C, int Co, const ctx_conv_geom& g) {
  // channels: 16-byte TMA strides (C, Co % 4); a ragged last 32-channel block is zero-filled by the TMA unit (A: channel
  // coordinates past C; B: the matching rows are multiplied by those zeros, rows past kh*kw*C are zero-filled too)
  if (C % 4 != 0 || Co % 4 != 0 || C < 32 || Co < 32) return 0;
  if (g.stride != 1 && g.stride != 2) return 0;
  if (g.kh * g.kw > kMaxSeg) return 0;
  return 1;                                          // any P x Q grid: exact 128-pixel tiles or partial tiles
}

// Tensor maps of a (possibly stride-2) NHWC tensor read at conv-tap offsets.  For stride 2, map[hp*2+wp] is the
// parity view x[n, 2*h2+hp, 2*w2+wp, c]; a tap r (relative to the output pixel: r' = r - pad) reads parity
// hp = r' & 1 at row offset (r' - hp) / 2.
// atoms > 1 (MN-major only, C % 32 == 0): rank-5 maps whose outermost dimension walks the 32-channel atoms, so one
// TMA instruction fetches `atoms` consecutive atoms (coordinates {0, w, h, n, channel / 32}).
static int make_tap_maps(CUtensorMap* maps, const float* x, int C, int N, int H, int W, int stride, int box_q,
                         int box_p, int box_n, bool mn_major, int atoms = 1) {
  const int rank = atoms > 1 ? 5 : 4;
  const int c0 = atoms > 1 ? 32 : C;
  if (stride == 1) {
    int64_t dims[5] = {c0, W, H, N, C / 32}, str[5] = {1, C, (int64_t)W * C, (int64_t)H * W * C, 32};
    int box[5] = {32, box_q, box_p, box_n, atoms};
    return make_map(&maps[0], x, rank, dims, str, box, mn_major);
  }
  for (int hp = 0; hp < 2; ++hp)
    for (int wp = 0; wp < 2; ++wp) {
      int Hh = (H - hp + 1) / 2, Wh = (W - wp + 1) / 2;
      const float* base = x + ((int64_t)hp * W + wp) * C;
      if (Hh <= 0 || Wh <= 0) {
        // empty parity view (a 1-pixel-high / wide input): keep a VALID descriptor over one pixel of x; tap_seg()
        // pushes every access through it out of range, so the TMA unit zero-fills (the taps contribute nothing)
        Hh = Hh <= 0 ? 1 : Hh; Wh = Wh <= 0 ? 1 : Wh; base = x;
      }
      int64_t dims[5] = {c0, Wh, Hh, N, C / 32};
      int64_t str[5] = {1, 2 * (int64_t)C, 2 * (int64_t)W * C, (int64_t)H * W * C, 32};
      int box[5] = {32, box_q, box_p, box_n, atoms};
      int rc = make_map(&maps[hp * 2 + wp], base, rank, dims, str, box, mn_major);
      if (rc) return rc;
    }
  return CTX_OK;
}

// MN-major B operand [K rows, N columns] (row stride ld): all BN/32 atoms of a tile in ONE instruction through a
// rank-3 map {32, K, N/32} (coordinates {0, k, column / 32}); needs N % 32 == 0.
static int make_bmn_map(CUtensorMap* map, const float* B, int64_t K, int N, int64_t ld, int atoms) {
  int64_t dims[3] = {32, K, N / 32}, str[3] = {1, ld, 32};
  int box[3] = {32, 32, atoms};
  return make_map(map, B, 3, dims, str, box, true);
}

static void tap_seg(Seg& s, int r, int sx, const ctx_conv_geom& g) {
  const int rr = r - g.pad_h, ss = sx - g.pad_w;
  if (g.stride == 1) { s.mapA = 0; s.dh = (int16_t)rr; s.dw = (int16_t)ss; return; }
  const int hp = rr & 1, wp = ss & 1;                 // two's complement: correct for negatives
  s.mapA = (uint8_t)(hp * 2 + wp);
  s.dh = (int16_t)((rr - hp) / 2);
  s.dw = (int16_t)((ss - wp) / 2);
  if ((g.H - hp + 1) / 2 <= 0 || (g.W - wp + 1) / 2 <= 0) s.dh = 16000;   // empty parity view: always out of range
}

int tc_conv2d_fprop(const float* x, int C, const float* w, float* y, int y_ld, int Co, const ctx_conv_geom& g,
                    const Epi& epi, int passes, cudaStream_t stream) {
  TcParams p;
  fill_common(p, passes, epi);
  int tpi, box_q = g.Q, m_tiles;
  CTX_REQUIRE(aligned16(x) && aligned16(w) && aligned16(y), "conv2d_fprop(tc): pointers must be 16-byte aligned");
  if (tile_grid(g.P, g.Q, 128, &p.BP, &p.NB, &tpi)) {
    p.tiles_per_img = tpi;
    m_tiles = p.NB > 1 ? (int)cdiv(g.N, p.NB) : g.N * tpi;
  } else {
    int tq;
    plan_partial(g.P, g.Q, &p.BQ, &p.BP, &p.NB, &tpi, &tq);
    p.tiles_per_img = tpi; p.q_tiles = tq; p.a_rows = p.BQ * p.BP * p.NB;
    box_q = p.BQ;
    m_tiles = p.NB > 1 ? (int)cdiv(g.N, p.NB) : g.N * tpi * tq;
    if (tq == 1 && p.NB == 1 && tpi > 1) p.q_tiles = 0;   // plain row bands: same decode as exact tiling
  }
  p.a_mode = A_PIX; p.b_mode = B_MN;
  p.BN = pick_bn(Co, m_tiles, passes);
  p.NG = g.N; p.PG = g.P; p.QG = g.Q;
  p.M = g.N * g.P * g.Q; p.N = Co;
  int rc = make_tap_maps(p.mapA, x, C, g.N, g.H, g.W, g.stride, box_q, p.BP, p.NB, false);
  if (rc) return rc;
  const int cblk = (int)cdiv(C, 32);
  p.pair = (Co % 32 == 0 && want_pair(passes, A_PIX, p.BN, Co, m_tiles, g.kh * g.kw * cblk)) ? 1 : 0;
  if (Co % 32 == 0) {
    p.b_atoms0 = (p.pair ? p.BN / 2 : p.BN) / 32;      // all atoms of the tile in one instruction
    rc = make_bmn_map(&p.mapB[0], w, (int64_t)g.kh * g.kw * C, Co, Co, p.b_atoms0);
  } else {
    p.b_atoms0 = 0;                                    // ragged last atom: one (zero-filled) box per atom
    int64_t dims[2] = {Co, (int64_t)g.kh * g.kw * C}, str[2] = {1, Co};
    int box[2] = {32, 32};
    rc = make_map(&p.mapB[0], w, 2, dims, str, box, true);
  }
  if (rc) return rc;
  p.nseg = g.kh * g.kw;
  for (int r = 0; r < g.kh; ++r)
    for (int s = 0; s < g.kw; ++s) {
      Seg& sg = p.seg[r * g.kw + s];
      tap_seg(sg, r, s, g);
      sg.wrap = 0; sg.nblk = (uint16_t)cblk; sg.bk0 = (r * g.kw + s) * C; sg.br0 = 0;
    }
  p.total_kblocks = p.nseg * cblk;
  p.kblocks_per_split = p.total_kblocks;
  p.out = y;
  p.sN = (int64_t)g.P * g.Q * y_ld; p.sP = (int64_t)g.Q * y_ld; p.sQ = y_ld; p.obase = 0;   // y_ld >= Co: wider destination
  p.atomic_out = 0;
  return launch_tc(p, m_tiles, 1, stream, "tc_conv2d_fprop");
}

// ---- transposed conv (one launch per stride phase) ---------------------------------------------------------------
static void phase_geom(const ctx_conv_geom& g, int ph, int pw, int* Hp, int* Wp, int* r0, int* s0, int* nr, int* ns,
                       int* dh0, int* dw0) {
  const int st = g.stride;
  *Hp = g.H > ph ? (g.H - ph + st - 1) / st : 0;
  *Wp = g.W > pw ? (g.W - pw + st - 1) / st : 0;
  *r0 = (ph + g.pad_h) % st; *s0 = (pw + g.pad_w) % st;
  *nr = *r0 < g.kh ? (g.kh - *r0 + st - 1) / st : 0;
  *ns = *s0 < g.kw ? (g.kw - *s0 + st - 1) / st : 0;
  *dh0 = (ph + g.pad_h - *r0) / st; *dw0 = (pw + g.pad_w - *s0) / st;
}

int tc_conv2d_transpose_eligible(int C1, int C2, int x2_batch, int Cn, const ctx_conv_geom& g) {
  if (C1 % 32 != 0 || C2 % 32 != 0 || Cn % 4 != 0 || Cn < 32) return 0;
  if (g.stride != 1 && g.stride != 2) return 0;
  for (int ph = 0; ph < g.stride; ++ph)
    for (int pw = 0; pw < g.stride; ++pw) {
      int Hp, Wp, r0, s0, nr, ns, dh0, dw0, bp, nb, tpi;
      phase_geom(g, ph, pw, &Hp, &Wp, &r0, &s0, &nr, &ns, &dh0, &dw0);
      if (Hp == 0 || Wp == 0) continue;
      if (!tile_grid(Hp, Wp, 128, &bp, &nb, &tpi)) return 0;
      if (C2 > 0 && nb > 1 && x2_batch % nb != 0) return 0;
      if (nr * ns * 2 > kMaxSeg) return 0;
    }
  return 1;
}

int tc_conv2d_transpose(const float* x1, int C1, const float* x2, int C2, int x2_batch, const float* w, float* y,
                        int Cn, const ctx_conv_geom& g, const Epi& epi, int passes, cudaStream_t stream) {
  CTX_REQUIRE(aligned16(x1) && aligned16(w) && aligned16(y) && (x2 == nullptr || aligned16(x2)),
              "conv2d_transpose(tc): pointers must be 16-byte aligned");
  const int st = g.stride, Ck = C1 + C2;
  // Collect the stride phases.  Phases with identical tile geometry (always the case for even sizes) are merged
  // into ONE launch, blockIdx.z = phase, heaviest phase first (it has the most filter taps).
  struct Ph { int ph, pw, Hp, Wp, r0, s0, nr, ns, dh0, dw0, BP, NB, tpi, taps; };
  Ph phs[4];
  int nph = 0;
  for (int ph = 0; ph < st; ++ph)
    for (int pw = 0; pw < st; ++pw) {
      Ph q;
      q.ph = ph; q.pw = pw;
      phase_geom(g, ph, pw, &q.Hp, &q.Wp, &q.r0, &q.s0, &q.nr, &q.ns, &q.dh0, &q.dw0);
      if (q.Hp == 0 || q.Wp == 0) continue;
      if (!tile_grid(q.Hp, q.Wp, 128, &q.BP, &q.NB, &q.tpi)) return CTX_ERR_UNSUPPORTED;
      q.taps = q.nr * q.ns;
      phs[nph++] = q;
    }
  for (int i = 1; i < nph; ++i)                         // insertion sort, descending taps
    for (int j = i; j > 0 && phs[j].taps > phs[j - 1].taps; --j) { Ph t = phs[j]; phs[j] = phs[j - 1]; phs[j - 1] = t; }
  bool same = nph > 1;
  int total_segs = 0;
  for (int i = 0; i < nph; ++i) {
    same = same && phs[i].Hp == phs[0].Hp && phs[i].Wp == phs[0].Wp;
    total_segs += phs[i].taps * (C2 > 0 ? 2 : 1);
  }
  same = same && total_segs <= kMaxSeg;

  auto fill_phase_segs = [&](TcParams& p, const Ph& q, int seg0, int* kblocks) {
    int nseg = 0;
    *kblocks = 0;
    for (int i = 0; i < q.nr; ++i)
      for (int j = 0; j < q.ns; ++j) {
        const int tap = (q.r0 + st * i) * g.kw + (q.s0 + st * j);
        for (int src = 0; src < (C2 > 0 ? 2 : 1); ++src) {
          Seg& sg = p.seg[seg0 + nseg++];
          sg.mapA = (uint8_t)src; sg.wrap = (uint8_t)src;
          sg.dh = (int16_t)(q.dh0 - i); sg.dw = (int16_t)(q.dw0 - j);
          sg.nblk = (uint16_t)((src ? C2 : C1) / 32);
          sg.bk0 = src ? C1 : 0; sg.br0 = tap * Cn;
          *kblocks += sg.nblk;
        }
      }
    return nseg;
  };

  for (int first = 0; first < nph; first += (same ? nph : 1)) {
    const Ph& q0 = phs[first];
    const int count = same ? nph : 1;
    TcParams p;
    fill_common(p, passes, epi);
    p.BP = q0.BP; p.NB = q0.NB; p.tiles_per_img = q0.tpi;
    p.a_mode = A_PIX; p.b_mode = B_K;
    const int m_tiles = p.NB > 1 ? (int)cdiv(g.N, p.NB) : g.N * q0.tpi;
    p.BN = pick_bn(Cn, (int64_t)m_tiles * count, passes);
    p.pair = want_pair(passes, A_PIX, p.BN, Cn, m_tiles, q0.taps * (Ck / 32)) ? 1 : 0;
    p.NG = g.N; p.PG = q0.Hp; p.QG = q0.Wp;
    p.M = g.N * q0.Hp * q0.Wp; p.N = Cn;
    p.a_wrap = x2_batch > 0 ? x2_batch : 1;
    {
      int64_t dims[4] = {C1, g.Q, g.P, g.N}, str[4] = {1, C1, (int64_t)g.Q * C1, (int64_t)g.P * g.Q * C1};
      int box[4] = {32, q0.Wp, p.BP, p.NB};
      int rc = make_map(&p.mapA[0], x1, 4, dims, str, box);
      if (rc) return rc;
    }
    if (C2 > 0) {
      int64_t dims[4] = {C2, g.Q, g.P, x2_batch}, str[4] = {1, C2, (int64_t)g.Q * C2, (int64_t)g.P * g.Q * C2};
      int box[4] = {32, q0.Wp, p.BP, p.NB};
      int rc = make_map(&p.mapA[1], x2, 4, dims, str, box);
      if (rc) return rc;
    }
    {
      int64_t dims[2] = {Ck, (int64_t)g.kh * g.kw * Cn}, str[2] = {1, Ck};
      int box[2] = {32, p.pair ? p.BN / 2 : p.BN};
      int rc = make_map(&p.mapB[0], w, 2, dims, str, box);
      if (rc) return rc;
    }
    int seg0 = 0;
    for (int i = 0; i < count; ++i) {
      const Ph& q = phs[first + i];
      int kblocks = 0;
      const int nseg = fill_phase_segs(p, q, seg0, &kblocks);
      p.ph_seg0[i] = seg0; p.ph_nseg[i] = nseg; p.ph_kblocks[i] = kblocks;
      p.ph_obase[i] = ((int64_t)q.ph * g.W + q.pw) * Cn;
      seg0 += nseg;
    }
    p.nphase = count;
    p.nseg = p.ph_nseg[0];
    p.total_kblocks = p.ph_kblocks[0];
    p.kblocks_per_split = p.ph_kblocks[0];
    p.obase = p.ph_obase[0];
    p.out = y;
    p.sN = (int64_t)g.H * g.W * Cn; p.sP = (int64_t)st * g.W * Cn; p.sQ = (int64_t)st * Cn;
    p.atomic_out = 0;
    int rc = launch_tc(p, m_tiles, count, stream, "tc_conv2d_transpose");
    if (rc) return rc;
  }
  return CTX_OK;
}

// ---- transposed conv with few output channels: ONE pixel x (tap, channel) GEMM + scatter-add epilogue ---------------
// y[n, p*st + r - pad_h, q*st + s - pad_w, co] += sum_c cat[n,p,q,c] * w[r,s,co,c]: the filter [kh,kw,Cn,Ck] IS a
// K-major [kh*kw*Cn, Ck] matrix, so the whole layer is a single GEMM with N = kh*kw*Cn columns whose epilogue
// scatters each column to its output pixel (col2im) with fp32 reductions.  The input is read once (the per-phase
// formulation re-reads it once per tap, which costs 25x the traffic when Cn = 3 cannot amortise it).
// y must already hold the bias (or zeros / the value to accumulate into).
int tc_conv2d_transpose_scatter_eligible(int C1, int C2, int x2_batch, int Cn, const ctx_conv_geom& g) {
  int bp, nb, tpi;
  if (C1 % 32 != 0 || C2 % 32 != 0) return 0;
  if (g.kh * g.kw * Cn > 256) return 0;
  if (!tile_grid(g.P, g.Q, 128, &bp, &nb, &tpi)) return 0;
  if (C2 > 0 && nb > 1 && x2_batch % nb != 0) return 0;
  return 1;
}

int tc_conv2d_transpose_scatter(const float* x1, int C1, const float* x2, int C2, int x2_batch, const float* w,
                                float* y, int Cn, const ctx_conv_geom& g, int passes, cudaStream_t stream) {
  CTX_REQUIRE(aligned16(x1) && aligned16(w) && (x2 == nullptr || aligned16(x2)),
              "conv2d_transpose(tc scatter): pointers must be 16-byte aligned");
  TcParams p;
  fill_common(p, passes, make_epi(nullptr));
  int tpi;
  if (!tile_grid(g.P, g.Q, 128, &p.BP, &p.NB, &tpi)) return CTX_ERR_UNSUPPORTED;
  p.tiles_per_img = tpi;
  p.a_mode = A_PIX; p.b_mode = B_K;
  const int Ck = C1 + C2, N = g.kh * g.kw * Cn;
  p.BN = (int)cdiv(N, 16) * 16;
  p.NG = g.N; p.PG = g.P; p.QG = g.Q;
  p.M = g.N * g.P * g.Q; p.N = N;
  p.a_wrap = x2_batch > 0 ? x2_batch : 1;
  const int m_tiles = p.NB > 1 ? (int)cdiv(g.N, p.NB) : g.N * tpi;
  {
    int64_t dims[4] = {C1, g.Q, g.P, g.N}, str[4] = {1, C1, (int64_t)g.Q * C1, (int64_t)g.P * g.Q * C1};
    int box[4] = {32, g.Q, p.BP, p.NB};
    int rc = make_map(&p.mapA[0], x1, 4, dims, str, box);
    if (rc) return rc;
  }
  if (C2 > 0) {
    int64_t dims[4] = {C2, g.Q, g.P, x2_batch}, str[4] = {1, C2, (int64_t)g.Q * C2, (int64_t)g.P * g.Q * C2};
    int box[4] = {32, g.Q, p.BP, p.NB};
    int rc = make_map(&p.mapA[1], x2, 4, dims, str, box);
    if (rc) return rc;
  }
  {
    int64_t dims[2] = {Ck, N}, str[2] = {1, Ck};       // rows past N are zero-filled by the TMA unit
    int box[2] = {32, p.BN};
    int rc = make_map(&p.mapB[0], w, 2, dims, str, box);
    if (rc) return rc;
  }
  p.nseg = C2 > 0 ? 2 : 1;
  for (int src = 0; src < p.nseg; ++src) {
    Seg& sg = p.seg[src];
    sg.mapA = (uint8_t)src; sg.wrap = (uint8_t)src; sg.dh = 0; sg.dw = 0;
    sg.nblk = (uint16_t)((src ? C2 : C1) / 32); sg.bk0 = src ? C1 : 0; sg.br0 = 0;
  }
  p.total_kblocks = Ck / 32;
  p.kblocks_per_split = p.total_kblocks;
  p.out = y;
  p.sN = 0; p.sP = 0; p.sQ = 0; p.obase = 0;
  p.atomic_out = 0;
  p.scatter = 1; p.sc_kw = g.kw; p.sc_Cn = Cn; p.sc_st = g.stride; p.sc_padh = g.pad_h; p.sc_padw = g.pad_w;
  p.sc_H = g.H; p.sc_W = g.W;
  return launch_tc(p, m_tiles, 1, stream, "tc_conv2d_transpose_scatter");
}

// ---- conv filter gradient ------------------------------------------------------------------------------------------
int tc_conv2d_wgrad_eligible(int Ca, int Cb1, int Cb2, int s2_batch, const ctx_conv_geom& g) {
  int bp, nb, tpi;
  if (Ca % 32 != 0 || Cb1 % 32 != 0 || Cb2 % 32 != 0) return 0;
  if (g.stride != 1 && g.stride != 2) return 0;
  if (g.kh * g.kw > kMaxSeg) return 0;
  if (!tile_grid(g.P, g.Q, 32, &bp, &nb, &tpi)) return 0;
  if (Cb2 > 0 && nb > 1 && s2_batch % nb != 0) return 0;
  return 1;
}

// Split-K factor: one or two CTAs per SM are resident (tc_occupancy), so the launch runs in waves of that many CTAs.  Choose the split that minimises  waves x (k-blocks per CTA + fixed cost per CTA), the
// fixed cost (prologue, TMEM allocation, atomic epilogue) being worth about `ovh` k-blocks.
static int pick_splits(int tiles, int kblocks, int min_per_split, int passes, int BN) {
  const int sms = sm_count() * tc_occupancy(passes, BN, kblocks);
  int maxs = kblocks / min_per_split;
  if (maxs < 1) maxs = 1;
  if (maxs > 4 * sms) maxs = 4 * sms;
  const int64_t ovh = 10;
  int best = 1;
  int64_t best_cost = -1;
  for (int s = 1; s <= maxs; ++s) {
    const int64_t per = cdiv(kblocks, s);
    const int64_t real = cdiv(kblocks, per);                 // splits that actually get work
    const int64_t waves = cdiv((int64_t)tiles * real, sms);
    const int64_t cost = waves * (per + ovh);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
  }
  return best;
}

int tc_conv2d_wgrad(const float* big, int Ca, const float* s1, int Cb1, const float* s2, int Cb2, int s2_batch,
                    float* dw, const ctx_conv_geom& g, int passes, cudaStream_t stream) {
  CTX_REQUIRE(aligned16(big) && aligned16(s1) && aligned16(dw) && (s2 == nullptr || aligned16(s2)),
              "conv2d_wgrad(tc): pointers must be 16-byte aligned");
  TcParams p;
  fill_common(p, passes, make_epi(nullptr));
  int tpi;
  if (!tile_grid(g.P, g.Q, 32, &p.kBP, &p.kNB, &tpi)) return CTX_ERR_UNSUPPORTED;
  p.k_tiles_per_img = tpi;
  p.a_mode = A_MN; p.b_mode = B_MN;
  const int N = Cb1 + Cb2;
  p.BN = pick_bn(N, 1 << 20, passes);           // split-K fills the machine, so take the widest tile
  p.kPG = g.P; p.kQG = g.Q;
  p.Ca = Ca; p.M = g.kh * g.kw * Ca; p.N = N;
  p.Cb1 = Cb1; p.a_wrap = s2_batch > 0 ? s2_batch : 1;
  // atoms of an M tile that share a filter tap are consecutive 32-channel blocks: fetch them together
  p.a_atoms = (Ca % 128 == 0) ? 4 : (Ca % 64 == 0) ? 2 : 1;
  int rc = make_tap_maps(p.mapA, big, Ca, g.N, g.H, g.W, g.stride, g.Q, p.kBP, p.kNB, true, p.a_atoms);
  if (rc) return rc;
  // B tile atoms: all from one source (tile inside small1 or inside small2), or the single tile [small1 ‖ small2]
  if (Cb2 == 0 || (Cb1 % p.BN == 0 && Cb2 % p.BN == 0)) { p.b_atoms0 = p.BN / 32; p.b_atoms1 = Cb2 > 0 ? p.BN / 32 : 0; }
  else if (Cb1 + Cb2 <= p.BN) { p.b_atoms0 = Cb1 / 32; p.b_atoms1 = (p.BN - Cb1) / 32; }
  else { p.b_atoms0 = p.b_atoms1 = 0; }
  if (p.b_atoms0 > 0) {
    rc = make_bmn_map(&p.mapB[0], s1, (int64_t)g.N * g.P * g.Q, Cb1, Cb1, p.b_atoms0);
    if (rc) return rc;
    if (Cb2 > 0) {
      rc = make_bmn_map(&p.mapB[1], s2, (int64_t)s2_batch * g.P * g.Q, Cb2, Cb2, p.b_atoms1);
      if (rc) return rc;
    }
  } else {
    {
      int64_t dims[2] = {Cb1, (int64_t)g.N * g.P * g.Q}, str[2] = {1, Cb1};
      int box[2] = {32, 32};
      rc = make_map(&p.mapB[0], s1, 2, dims, str, box, true);
      if (rc) return rc;
    }
    if (Cb2 > 0) {
      int64_t dims[2] = {Cb2, (int64_t)s2_batch * g.P * g.Q}, str[2] = {1, Cb2};
      int box[2] = {32, 32};
      rc = make_map(&p.mapB[1], s2, 2, dims, str, box, true);
      if (rc) return rc;
    }
  }
  p.nseg = g.kh * g.kw;
  for (int r = 0; r < g.kh; ++r)
    for (int s = 0; s < g.kw; ++s) tap_seg(p.seg[r * g.kw + s], r, s, g);
  p.total_kblocks = p.kNB > 1 ? (int)cdiv(g.N, p.kNB) : g.N * tpi;
  // epilogue row grid: rows are dw rows
  const int m_tiles = (int)cdiv(p.M, 128);
  const int n_tiles = (int)cdiv(N, p.BN);
  const int splits = pick_splits(m_tiles * n_tiles, p.total_kblocks, 8, passes, p.BN);
  p.kblocks_per_split = (int)cdiv(p.total_kblocks, splits);
  const int real_splits = (int)cdiv(p.total_kblocks, p.kblocks_per_split);
  p.out = dw;
  p.atomic_out = 1;
  // D rows are addressed linearly: a 128-"image" tile of 1x1 grids
  p.NG = p.M; p.PG = 1; p.QG = 1; p.BP = 1; p.NB = 128; p.tiles_per_img = 1;
  p.sN = N; p.sP = 0; p.sQ = 0; p.obase = 0;
  return launch_tc(p, m_tiles, real_splits, stream, "tc_conv2d_wgrad");
}

// ---- dense GEMM ------------------------------------------------------------------------------------------------------
int tc_gemm_eligible(int tA, int tB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                     const float* C, int ldc) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (!aligned16(A) || !aligned16(B) || !aligned16(C)) return 0;
  // TMA needs 16-byte strides; M and K are free (out-of-range rows / columns are zero-filled by the TMA unit)
  if (lda % 4 != 0 || ldb % 4 != 0 || ldc % 4 != 0 || N % 4 != 0) return 0;
  if (N < 32) return 0;
  if (tA && tB) return 0;                          // MN-major A with K-major B is not wired up (unused by the graph)
  if (!tB && (N % 32 != 0)) return 0;
  return 1;
}

// Deferred epilogue of a split-K GEMM: C holds the raw sums, apply bias / lrelu / lrelu' mask / TF32 rounding in place.
__global__ void gemm_epilogue_kernel(float* C, int M, int N, int ldc, Epi epi) {
  const int64_t total = (int64_t)M * (N >> 2);
  const int nq = N >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / nq;
    const int n = (int)(i - m * nq) << 2;
    const int64_t off = m * ldc + n;
    float4 v = *reinterpret_cast<float4*>(C + off);
    v.x = epi.apply(v.x, n, off, C); v.y = epi.apply(v.y, n + 1, off + 1, C);
    v.z = epi.apply(v.z, n + 2, off + 2, C); v.w = epi.apply(v.w, n + 3, off + 3, C);
    *reinterpret_cast<float4*>(C + off) = v;
  }
}

int tc_gemm(int tA, int tB, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
            const Epi& epi, int passes, cudaStream_t stream) {
  // Skinny GEMMs (the `linear` layers: M = batch <= a few hundred rows, K up to 8192) leave most SMs idle and run a
  // long serial K loop.  When the fused epilogue is not a plain accumulation, split K anyway: zero C, accumulate
  // the raw partial sums with the split-K path, then apply the epilogue in a second, tiny pass over C.
  // Few tiles and a long K, any fused epilogue: split K over a thread-block cluster (partials meet in rank 0's shared memory).
  int cluster_k = 1;
  if (!epi.accumulate && knob_cluster_k()) {
    const int bn = pick_bn(N, cdiv(M, 128), passes);
    const int64_t tiles = cdiv(M, 128) * cdiv(N, bn);
    const int kblocks = (int)cdiv(K, 32);
    for (int sk = 8; sk >= 2; sk >>= 1) {
      const int occ = tc_occupancy(passes, bn, (int)cdiv(kblocks, sk));
      const int64_t slots = (int64_t)sm_count() * occ;
      // bytes rank 0 can receive: this occupancy's share of the SM's shared memory (launch_tc sizes the allocation)
      const int64_t ring = (occ == 1 ? 200 : 108) * 1024;
      if (tiles * sk > slots) continue;                                    // (the whole cluster grid in one wave)
      if (kblocks / sk < 4 || bn % 32 != 0) continue;
      if ((int64_t)(sk - 1) * bn * 512 > ring) continue;
      if ((int64_t)(sk - 1) * cdiv(kblocks, sk) >= kblocks) continue;      // every rank gets work
      cluster_k = sk;
      break;
    }
  }
  if (!epi.accumulate && cluster_k == 1) {
    const int bn = pick_bn(N, 1 << 20, passes);
    const int64_t tiles = cdiv(M, 128) * cdiv(N, bn);
    const int kblocks = (int)cdiv(K, 32);
    if (tiles * 2 <= sm_count() * tc_occupancy(passes, bn, kblocks) && kblocks >= 64 && !(epi.act_ref && ldc != N)) {
      CTX_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, stream));
      Epi acc = make_epi(nullptr);
      acc.accumulate = 1;
      int rc = tc_gemm(tA, tB, M, N, K, A, lda, B, ldb, C, ldc, acc, passes, stream);
      if (rc) return rc;
      if (epi.bias || epi.act || epi.act_ref || epi.round) {
        const int64_t work = (int64_t)M * (N >> 2);
        int blocks = (int)cdiv(work, 256);
        if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
        gemm_epilogue_kernel<<<blocks, 256, 0, stream>>>(C, M, N, ldc, epi);
        CTX_LAUNCH_CHECK("gemm_epilogue");
      }
      return CTX_OK;
    }
  }
  TcParams p;
  fill_common(p, passes, epi);
  const bool plain_accumulate = epi.accumulate && !epi.bias && !epi.act && !epi.act_ref && !epi.round;
  const int m_tiles = (int)cdiv(M, 128);
  // split-K (plain accumulation only) fills the machine by itself; otherwise narrow the N tile on small grids
  p.BN = pick_bn(N, plain_accumulate ? (1 << 20) : m_tiles, passes);
  p.M = M; p.N = N;
  int rc;
  if (!tA) {
    // A [M,K] row-major: K-major rows; tile = 128 consecutive rows ("images"), one 1x1 pixel each
    p.a_mode = A_PIX;
    int64_t dims[4] = {K, 1, 1, M}, str[4] = {1, lda, lda, lda};
    int box[4] = {32, 1, 1, 128};
    rc = make_map(&p.mapA[0], A, 4, dims, str, box);
    if (rc) return rc;
    p.nseg = 1;
    Seg& sg = p.seg[0];
    sg.mapA = 0; sg.wrap = 0; sg.dh = 0; sg.dw = 0; sg.nblk = (uint16_t)cdiv(K, 32); sg.bk0 = 0; sg.br0 = 0;
    CTX_REQUIRE(cdiv(K, 32) < 65536, "gemm(tc): K too large");
    p.total_kblocks = (int)cdiv(K, 32);
  } else {
    // A stored [K,M]: MN-major; K blocks are 32 consecutive rows of the stored matrix
    p.a_mode = A_MN;
    p.Ca = M; p.kPG = 1; p.kQG = 1; p.kBP = 1; p.kNB = 32; p.k_tiles_per_img = 1;
    int64_t dims[4] = {M, 1, 1, K}, str[4] = {1, lda, lda, lda};
    int box[4] = {32, 1, 1, 32};
    rc = make_map(&p.mapA[0], A, 4, dims, str, box, true);
    if (rc) return rc;
    p.nseg = 1;
    Seg& sg = p.seg[0];
    sg.mapA = 0; sg.wrap = 0; sg.dh = 0; sg.dw = 0; sg.nblk = 0; sg.bk0 = 0; sg.br0 = 0;
    p.total_kblocks = (int)cdiv(K, 32);
  }
  if (!tB) {
    p.b_mode = B_MN;                                 // B [K,N] row-major (N % 32 == 0 is an eligibility condition)
    p.b_atoms0 = p.BN / 32;
    rc = make_bmn_map(&p.mapB[0], B, K, N, ldb, p.b_atoms0);
  } else {
    p.b_mode = B_K;                                  // B stored [N,K]
    int64_t dims[2] = {K, N}, str[2] = {1, ldb};
    int box[2] = {32, p.BN};
    rc = make_map(&p.mapB[0], B, 2, dims, str, box);
  }
  if (rc) return rc;
  const int n_tiles = (int)cdiv(N, p.BN);
  int splits = 1;
  if (plain_accumulate) splits = pick_splits(m_tiles * n_tiles, p.total_kblocks, 4, passes, p.BN);
  if (cluster_k > 1) splits = cluster_k;
  p.kblocks_per_split = (int)cdiv(p.total_kblocks, splits);
  splits = (int)cdiv(p.total_kblocks, p.kblocks_per_split);
  if (cluster_k > 1) {
    if (splits == cluster_k && p.BN % 32 == 0) p.cluster_k = cluster_k;
    else { splits = 1; p.kblocks_per_split = p.total_kblocks; }             // (cannot happen: checked above)
  }
  // plain += always goes out as vector reductions (fire and forget): the read-modify-write epilogue is a chain of dependent
  // row-per-thread loads (ncu, hz_lin's filter gradient: 37 us for 8 k-blocks, every FADD waiting on its own load)
  p.atomic_out = plain_accumulate ? 1 : 0;
  if (p.atomic_out) p.epi.accumulate = 0;
  p.out = C;
  // D (and a K-major A) rows are matrix rows: a 128-"image" tile of 1x1 grids
  p.NG = M; p.PG = 1; p.QG = 1; p.BP = 1; p.NB = 128; p.tiles_per_img = 1;
  p.sN = ldc; p.sP = 0; p.sQ = 0; p.obase = 0;
  return launch_tc(p, m_tiles, splits, stream, "tc_gemm");
}

}  // namespace ctx
