// Probe for the round-2 backward path: GEMMs on 16-bit operand planes (fp16 or bf16) with tcgen05 kind::f16, fp32
// accumulation in TMEM.  Questions it answers on a B200 before the library kernel is written:
//   (1) shared-memory descriptors of MN-MAJOR 16-bit operands (the filter gradient: pixels are the K dimension and
//       channels are contiguous in NHWC) — LBO / SBO / K-advance are command-line overridable so several hypotheses
//       can be checked in one GPU call;
//   (2) throughput of a single-product fp16 pipeline (fp16 has the TF32 significand: 1 + 10 bits) against the
//       three-product (hi, lo) scheme, single CTAs and CTA pairs, ring depth, split-K.
// Not part of the library.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o h16_probe h16_probe.cu -lcuda
//   ./h16_probe M N K BN stages nprod mn f16 pair splits reps [lbo sbo kstep]
//
// K-major: A [planes][M][K], B [planes][N][K].  MN-major (mn = 1): A [planes][K][M], B [planes][K][N].
// nprod = 1: D = A0*B0.  nprod = 3: D = A0*B0 + (A0*B1 + A1*B0) with the cross terms in a second accumulator.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Params {
  CUtensorMap mapA, mapB;
  float* D;
  int M, N, K, BN, stages, nprod, mn, f16, kb_per_split;
  uint32_t lbo, sbo, kstep;          // MN-major descriptor fields (bytes)
  int* err;
};

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
template <bool kCluster>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int id) {
  auto probe = [&]() { return kCluster ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity); };
  if (probe()) return;
  const long long t0 = clock64();
  while (!probe()) {
    if (clock64() - t0 > 2000000000LL) { if (err) atomicExch(err, 0x1000 + id); __trap(); }
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
  if constexpr (kPair)
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
// SM100 shared-memory descriptor, SWIZZLE_128B: start address, leading / stride byte offsets (>> 4), version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

constexpr int kThreads = 192;                       // warp 0: TMA, warp 1: MMA, warps 2..5: epilogue
constexpr uint32_t kATile = 128 * 128;              // 128 rows x 64 elements x 2 B (either major)

template <bool kPair>
__global__ void __launch_bounds__(kThreads, 1) probe_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if constexpr (kPair) rank = cluster_ctarank();
  const int planes = p.nprod == 3 ? 2 : 1;
  const int bn_cta = kPair ? p.BN / 2 : p.BN;
  const uint32_t b_tile = (uint32_t)bn_cta * 128u;
  const uint32_t stage_bytes = (uint32_t)planes * (kATile + b_tile);
  const uint32_t b_off = (uint32_t)planes * kATile;            // [A0 | A1 | B0 | B1]
  __shared__ __align__(8) uint64_t bars[25];
  __shared__ uint32_t tmem_slot;
  const int S = p.stages;
  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[8 + s]); };
  auto peer_bar = [&](int s) { return smem_u32(&bars[16 + s]); };
  const uint32_t acc_bar = smem_u32(&bars[24]);
  const int nkb_total = p.K / 64;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int kb1 = min(nkb_total, kb0 + p.kb_per_split);
  const int nkb = kb1 - kb0;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * p.BN;
  uint32_t ncols = 32;
  while (ncols < (uint32_t)(p.BN * (p.nprod == 3 ? 2 : 1))) ncols <<= 1;

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
  if constexpr (kPair) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const int nb0 = n0 + (int)rank * bn_cta;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait<false>(empty_bar(stage), phase ^ 1, p.err, 1);
        const uint32_t sa = smem_base + stage * stage_bytes;
        mbar_arrive_expect_tx(full_bar(stage), stage_bytes);
        if (!p.mn) {
          tma_load_3d(sa, &p.mapA, full_bar(stage), kb * 64, m0, 0);              // box {64 k, 128 rows, planes}
          tma_load_3d(sa + b_off, &p.mapB, full_bar(stage), kb * 64, nb0, 0);     // box {64 k, bn_cta rows, planes}
        } else {
          tma_load_4d(sa, &p.mapA, full_bar(stage), 0, kb * 64, m0 >> 6, 0);           // box {64 m, 64 k, 2 atoms, planes}
          tma_load_4d(sa + b_off, &p.mapB, full_bar(stage), 0, kb * 64, nb0 >> 6, 0);  // box {64 n, 64 k, bn_cta/64 atoms, planes}
        }
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      if (kPair && rank != 0) {
        for (int kb = 0; kb < nkb; ++kb) {                // follower: forward "my stage has landed" to the leader
          mbar_wait<false>(full_bar(stage), phase, p.err, 4);
          mbar_arrive_remote(peer_bar(stage), 0u);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      } else {
        // kind::f16: D = f32 (1 << 4); A, B format at bits 7 / 10: 0 = f16, 1 = bf16; bit 15 / 16 = A / B MN-major;
        // N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t fmt = p.f16 ? 0u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((p.mn ? 1u : 0u) << 15) | ((p.mn ? 1u : 0u) << 16) |
                               ((uint32_t)(p.BN >> 3) << 17) | (((kPair ? 256u : 128u) >> 4) << 24);
        const uint32_t d1 = tmem_base, d2 = tmem_base + (uint32_t)p.BN;
        const uint32_t lbo = p.mn ? p.lbo : 16u, sbo = p.mn ? p.sbo : 1024u;
        const uint64_t kstep = (uint64_t)((p.mn ? p.kstep : 32u) >> 4);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait<false>(full_bar(stage), phase, p.err, 2);
          if constexpr (kPair) mbar_wait<true>(peer_bar(stage), phase, p.err, 5);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint64_t a0 = make_desc(sa, lbo, sbo), a1 = make_desc(sa + kATile, lbo, sbo);
          const uint64_t b0 = make_desc(sa + b_off, lbo, sbo), b1 = make_desc(sa + b_off + b_tile, lbo, sbo);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {                // K = 16 per MMA
            const uint32_t first = (kb | ks) ? 1u : 0u;
            umma_f16<kPair>(d1, a0 + kstep * ks, b0 + kstep * ks, idesc, first);
            if (p.nprod == 3) {
              umma_f16<kPair>(d2, a0 + kstep * ks, b1 + kstep * ks, idesc, first);
              umma_f16<kPair>(d2, a1 + kstep * ks, b0 + kstep * ks, idesc, 1u);
            }
          }
          umma_commit<kPair>(empty_bar(stage));
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        umma_commit<kPair>(acc_bar);
      }
    }
  } else {
    const int q = warp & 3, r = q * 32 + lane;
    if (nkb > 0) {
      mbar_wait<false>(acc_bar, 0, p.err, 3);
      tc_fence_after();
      float* orow = p.D + (size_t)(m0 + r) * p.N + n0;
      const bool atomic = gridDim.z > 1;
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        if (p.nprod == 3) {
          uint32_t v2[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.BN + c0), v2);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
        }
        if (m0 + r < p.M) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (atomic)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + c0 + j), "f"(__uint_as_float(v[j])),
                           "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3])) : "memory");
            else
              *reinterpret_cast<float4*>(orow + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                      __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

// ---- host ------------------------------------------------------------------------------------------------------------------
static uint16_t bf16_rn(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7FFFu + ((u >> 16) & 1u); return (uint16_t)(u >> 16); }
static float bf16_f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
static uint16_t f16_rn(float f) { __half h = __float2half_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }
static float f16_f(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void encode(EncodeTiledFn enc, CUtensorMap* m, int f16, void* base, int rank, const cuuint64_t* gd, const cuuint64_t* gs,
                   const cuuint32_t* bx) {
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
}

static float frand(uint64_t& s) {                   // ~N(0,1)
  float a = 0;
  for (int i = 0; i < 4; ++i) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; a += (float)((s >> 40) & 0xFFFFFF) / 16777216.0f; }
  return (a - 2.0f) * 1.7320508f;
}

int main(int argc, char** argv) {
  int M = argc > 1 ? atoi(argv[1]) : 32768, N = argc > 2 ? atoi(argv[2]) : 512, K = argc > 3 ? atoi(argv[3]) : 3200;
  int BN = argc > 4 ? atoi(argv[4]) : 256, stages = argc > 5 ? atoi(argv[5]) : 6, nprod = argc > 6 ? atoi(argv[6]) : 1;
  int mn = argc > 7 ? atoi(argv[7]) : 0, f16 = argc > 8 ? atoi(argv[8]) : 1, pair = argc > 9 ? atoi(argv[9]) : 1;
  int splits = argc > 10 ? atoi(argv[10]) : 1, reps = argc > 11 ? atoi(argv[11]) : 20;
  uint32_t lbo = argc > 12 ? (uint32_t)atoi(argv[12]) : 8192u, sbo = argc > 13 ? (uint32_t)atoi(argv[13]) : 1024u;
  uint32_t kstep = argc > 14 ? (uint32_t)atoi(argv[14]) : 2048u;
  const int planes = nprod == 3 ? 2 : 1;
  if (M % 128 || N % BN || K % 64 || BN % 64 || BN > 256 || stages < 1 || stages > 8 || BN * planes > 512 || (nprod != 1 && nprod != 3)) {
    printf("bad shape\n"); return 1;
  }
  const int Mc = M < 256 ? M : 256;                 // rows verified against fp64 on the host
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaFree(0));
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)f;

  // host data: rows of A repeat with period Mc (keeps generation cheap); B fully random
  std::vector<float> A((size_t)Mc * K), B((size_t)N * K);
  uint64_t seed = 12345;
  for (auto& x : A) x = fabsf(frand(seed)) * 0.5f + 0.1f * frand(seed);
  for (auto& x : B) x = 0.02f * frand(seed);
  auto rn = [&](float v) { return f16 ? f16_rn(v) : bf16_rn(v); };
  auto fl = [&](uint16_t h) { return f16 ? f16_f(h) : bf16_f(h); };
  std::vector<uint16_t> Ap((size_t)planes * M * K), Bp((size_t)planes * N * K);
  std::vector<float> A0((size_t)Mc * K), A1((size_t)Mc * K), B0((size_t)N * K), B1((size_t)N * K);
  for (int m = 0; m < Mc; ++m)
    for (int k = 0; k < K; ++k) {
      const float a = A[(size_t)m * K + k];
      const uint16_t hi = rn(a), lo = rn(a - fl(hi));
      A0[(size_t)m * K + k] = fl(hi); A1[(size_t)m * K + k] = fl(lo);
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const float b = B[(size_t)n * K + k];
      const uint16_t hi = rn(b), lo = rn(b - fl(hi));
      B0[(size_t)n * K + k] = fl(hi); B1[(size_t)n * K + k] = fl(lo);
    }
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      const size_t src = (size_t)(m % Mc) * K + k;
      const size_t dst = mn ? (size_t)k * M + m : (size_t)m * K + k;
      Ap[dst] = rn(A0[src]);
      if (planes == 2) Ap[(size_t)M * K + dst] = rn(A1[src]);
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const size_t src = (size_t)n * K + k;
      const size_t dst = mn ? (size_t)k * N + n : src;
      Bp[dst] = rn(B0[src]);
      if (planes == 2) Bp[(size_t)N * K + dst] = rn(B1[src]);
    }
  uint16_t *dA, *dB; float* dD; int* derr;
  CK(cudaMalloc(&dA, Ap.size() * 2)); CK(cudaMalloc(&dB, Bp.size() * 2)); CK(cudaMalloc(&dD, (size_t)M * N * 4));
  CK(cudaMalloc(&derr, 4)); CK(cudaMemset(derr, 0, 4));
  CK(cudaMemcpy(dA, Ap.data(), Ap.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bp.data(), Bp.size() * 2, cudaMemcpyHostToDevice));

  Params p;
  memset(&p, 0, sizeof(p));
  const int bn_cta = pair ? BN / 2 : BN;
  if (!mn) {
    cuuint64_t gdA[3] = {(cuuint64_t)K, (cuuint64_t)M, (cuuint64_t)planes}, gsA[2] = {(cuuint64_t)K * 2, (cuuint64_t)M * K * 2};
    cuuint32_t bxA[3] = {64, 128, (cuuint32_t)planes};
    encode(enc, &p.mapA, f16, dA, 3, gdA, gsA, bxA);
    cuuint64_t gdB[3] = {(cuuint64_t)K, (cuuint64_t)N, (cuuint64_t)planes}, gsB[2] = {(cuuint64_t)K * 2, (cuuint64_t)N * K * 2};
    cuuint32_t bxB[3] = {64, (cuuint32_t)bn_cta, (cuuint32_t)planes};
    encode(enc, &p.mapB, f16, dB, 3, gdB, gsB, bxB);
  } else {
    // stored [K][M]: dims {64 m, K, M/64 atoms, planes}
    cuuint64_t gdA[4] = {64, (cuuint64_t)K, (cuuint64_t)M / 64, (cuuint64_t)planes};
    cuuint64_t gsA[3] = {(cuuint64_t)M * 2, 128, (cuuint64_t)M * K * 2};
    cuuint32_t bxA[4] = {64, 64, 2, (cuuint32_t)planes};
    encode(enc, &p.mapA, f16, dA, 4, gdA, gsA, bxA);
    cuuint64_t gdB[4] = {64, (cuuint64_t)K, (cuuint64_t)N / 64, (cuuint64_t)planes};
    cuuint64_t gsB[3] = {(cuuint64_t)N * 2, 128, (cuuint64_t)N * K * 2};
    cuuint32_t bxB[4] = {64, 64, (cuuint32_t)bn_cta / 64, (cuuint32_t)planes};
    encode(enc, &p.mapB, f16, dB, 4, gdB, gsB, bxB);
  }
  p.D = dD; p.M = M; p.N = N; p.K = K; p.BN = BN; p.stages = stages; p.nprod = nprod; p.mn = mn; p.f16 = f16; p.err = derr;
  p.lbo = lbo; p.sbo = sbo; p.kstep = kstep;
  const int nkb = K / 64;
  p.kb_per_split = (nkb + splits - 1) / splits;
  splits = (nkb + p.kb_per_split - 1) / p.kb_per_split;
  const size_t smem = (size_t)stages * planes * (kATile + bn_cta * 128) + 1024;
  if (smem > 232448 || (pair && (M / 128) % 2)) { printf("does not fit\n"); return 1; }
  CK(cudaFuncSetAttribute(probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(M / 128, N / BN, splits);
  auto launch = [&]() {
    if (splits > 1) CK(cudaMemsetAsync(dD, 0, (size_t)M * N * 4));
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = dim3(kThreads, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute attr[1];
    if (pair) {
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, probe_kernel<true>, p));
    } else {
      CK(cudaLaunchKernelEx(&cfg, probe_kernel<false>, p));
    }
  };
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  const double us = ms * 1e3 / reps, tf = 2.0 * M * N * (double)K / (us * 1e-6) / 1e12;

  std::vector<float> D((size_t)Mc * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double max_ref = 0, max_err = 0, sum2 = 0, sum2e = 0;
  const int rows_chk = Mc < 128 ? Mc : 128;
  for (int m = 0; m < rows_chk; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) {
        const size_t ia = (size_t)m * K + k, ib = (size_t)n * K + k;
        ref += (double)A0[ia] * B0[ib];
        if (nprod == 3) ref += (double)A0[ia] * B1[ib] + (double)A1[ia] * B0[ib];
      }
      const double e = fabs((double)D[(size_t)m * N + n] - ref);
      if (fabs(ref) > max_ref) max_ref = fabs(ref);
      if (e > max_err) max_err = e;
      sum2 += ref * ref; sum2e += e * e;
    }
  int herr = 0; CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
  printf("M=%d N=%d K=%d BN=%d stages=%d nprod=%d mn=%d %s pair=%d splits=%d lbo=%u sbo=%u kstep=%u | %.1f us  %.1f TF/s algorithmic "
         "(%.1f TF/s of MMAs) | vs exact products of the planes: max-norm err %.2e  rel-L2 %.2e | err_flag %d %s\n",
         M, N, K, BN, stages, nprod, mn, f16 ? "f16" : "bf16", pair, splits, lbo, sbo, kstep, us, tf, tf * nprod,
         max_err / max_ref, sqrt(sum2e / sum2), herr, sqrt(sum2e / sum2) < 1e-4 ? "OK" : "WRONG");
  return 0;
}
