// SIMT (CUDA-core, exact fp32) implicit-GEMM kernels: conv fprop, transposed conv, conv wgrad, dense GEMM.
//
// One register-tiled kernel, parameterised by a "problem" functor that says how element (m,k) of A, (k,n) of B
// and the output offset of row m are found in memory.  This is the exact-fp32 path (CTX_IMPL_SIMT) and the
// path for shapes the tcgen05 kernels do not take (channel counts that are not multiples of 32: the Cin=3
// first conv, the Cout=3 last transposed conv, all of ContextAEReal).  im2col is never materialised.
#include "common.cuh"

namespace ctx {

// ---- problem functors -----------------------------------------------------------------------------------

// Dense GEMM  C = opA(A) opB(B)
template <bool TA, bool TB>
struct GemmProb {
  static constexpr bool A_KCONTIG = !TA;
  static constexpr bool B_NCONTIG = !TB;
  static constexpr bool ATOMIC_OUT = false;
  int M, N, K;
  const float* A; const float* B; float* C;
  int lda, ldb, ldc;
  Epi epi;
  struct RowA { int64_t off; bool ok; };
  struct ColB { int64_t off; bool ok; };
  struct KInfo { int k; bool ok; };
  __device__ __forceinline__ RowA rowA(int m) const { return {TA ? (int64_t)m : (int64_t)m * lda, m < M}; }
  __device__ __forceinline__ ColB colB(int n) const { return {TB ? (int64_t)n * ldb : (int64_t)n, n < N}; }
  __device__ __forceinline__ KInfo decodeK(int k, int kend) const { return {k, k < kend}; }
  __device__ __forceinline__ float fetchA(const RowA& r, const KInfo& k) const {
    return (r.ok && k.ok) ? __ldg(A + r.off + (TA ? (int64_t)k.k * lda : (int64_t)k.k)) : 0.f;
  }
  __device__ __forceinline__ float fetchB(const KInfo& k, const ColB& c) const {
    return (c.ok && k.ok) ? __ldg(B + c.off + (TB ? (int64_t)k.k : (int64_t)k.k * ldb)) : 0.f;
  }
  __device__ __forceinline__ int64_t rowOut(int m) const { return (int64_t)m * ldc; }
  __device__ __forceinline__ float* out() const { return C; }
};

// Strided SAME conv, NHWC x HWIO
struct FpropProb {
  static constexpr bool A_KCONTIG = true;
  static constexpr bool B_NCONTIG = true;
  static constexpr bool ATOMIC_OUT = false;
  int M, N, K;                       // M = imgs*P*Q, N = Co, K = kh*kw*C
  const float* x; const float* w; float* y;
  int C, H, W, st, padh, padw, kw;
  int ldy;                           // channels per pixel of the destination tensor (>= N: write into a wider tensor)
  FastDiv divPQ, divQ, divC, divKw;
  Epi epi;
  struct RowA { int64_t img; int h0, w0; bool ok; };
  struct ColB { int n; bool ok; };
  struct KInfo { int r, s, c, k; bool ok; };
  __device__ __forceinline__ RowA rowA(int m) const {
    int n, rem, p, q;
    divPQ.divmod(m, n, rem);
    divQ.divmod(rem, p, q);
    return {(int64_t)n * H * W * C, p * st - padh, q * st - padw, m < M};
  }
  __device__ __forceinline__ ColB colB(int n) const { return {n, n < N}; }
  __device__ __forceinline__ KInfo decodeK(int k, int kend) const {
    int tap, c, r, s;
    divC.divmod(k, tap, c);
    divKw.divmod(tap, r, s);
    return {r, s, c, k, k < kend};
  }
  __device__ __forceinline__ float fetchA(const RowA& r, const KInfo& k) const {
    int h = r.h0 + k.r, ww = r.w0 + k.s;
    bool ok = r.ok && k.ok && (unsigned)h < (unsigned)H && (unsigned)ww < (unsigned)W;
    return ok ? __ldg(x + r.img + ((int64_t)h * W + ww) * C + k.c) : 0.f;
  }
  __device__ __forceinline__ float fetchB(const KInfo& k, const ColB& c) const {
    return (c.ok && k.ok) ? __ldg(w + (int64_t)k.k * N + c.n) : 0.f;
  }
  __device__ __forceinline__ int64_t rowOut(int m) const { return (int64_t)m * ldy; }
  __device__ __forceinline__ float* out() const { return y; }
};

// One stride-phase of a SAME transposed conv with a two-source (concatenated) input.
struct TransProb {
  static constexpr bool A_KCONTIG = true;
  static constexpr bool B_NCONTIG = false;
  static constexpr bool ATOMIC_OUT = false;
  int M, N, K;                       // M = imgs*Hp*Wp, N = Cn, K = nr*ns*Ck
  const float* x1; const float* x2; const float* w; float* y;
  int C1, C2, Ck, x2_batch;
  int P, Q, H, W, st, kw;
  int ph, pw, r0, s0, ns, dh0, dw0;
  FastDiv divHpWp, divWp, divCk, divNs, divX2;
  Epi epi;
  struct RowA { int n, n2, hq, wq; bool ok; };
  struct ColB { int n; bool ok; };
  struct KInfo { int i, j, c, tap; bool ok; };
  __device__ __forceinline__ void decodeM(int m, int& n, int& hp, int& wp) const {
    int rem;
    divHpWp.divmod(m, n, rem);
    divWp.divmod(rem, hp, wp);
  }
  __device__ __forceinline__ RowA rowA(int m) const {
    int n, hp, wp;
    decodeM(m, n, hp, wp);
    int q2, n2;
    divX2.divmod(n, q2, n2);
    return {n, n2, hp + dh0, wp + dw0, m < M};
  }
  __device__ __forceinline__ ColB colB(int n) const { return {n, n < N}; }
  __device__ __forceinline__ KInfo decodeK(int k, int kend) const {
    int t, c, i, j;
    divCk.divmod(k, t, c);
    divNs.divmod(t, i, j);
    return {i, j, c, (r0 + st * i) * kw + (s0 + st * j), k < kend};
  }
  __device__ __forceinline__ float fetchA(const RowA& r, const KInfo& k) const {
    int p = r.hq - k.i, q = r.wq - k.j;
    bool ok = r.ok && k.ok && (unsigned)p < (unsigned)P && (unsigned)q < (unsigned)Q;
    if (!ok) return 0.f;
    if (k.c < C1) return __ldg(x1 + (((int64_t)r.n * P + p) * Q + q) * C1 + k.c);
    return __ldg(x2 + (((int64_t)r.n2 * P + p) * Q + q) * C2 + (k.c - C1));
  }
  __device__ __forceinline__ float fetchB(const KInfo& k, const ColB& c) const {
    return (c.ok && k.ok) ? __ldg(w + ((int64_t)k.tap * N + c.n) * Ck + k.c) : 0.f;
  }
  __device__ __forceinline__ int64_t rowOut(int m) const {
    int n, hp, wp;
    decodeM(m, n, hp, wp);
    return (((int64_t)n * H + (ph + st * hp)) * W + (pw + st * wp)) * N;
  }
  __device__ __forceinline__ float* out() const { return y; }
};

// Filter gradient:  dw[(tap,a), b] += sum_{pixels} big[shifted pixel, a] * small[pixel, b]
struct WgradProb {
  static constexpr bool A_KCONTIG = false;
  static constexpr bool B_NCONTIG = true;
  static constexpr bool ATOMIC_OUT = true;
  int M, N, K;                       // M = kh*kw*Ca, N = Cb1+Cb2, K = imgs*P*Q
  const float* big; const float* s1; const float* s2; float* dw;
  int Ca, Cb1, Cb2, s2_batch;
  int P, Q, H, W, st, padh, padw, kw;
  FastDiv divCa, divKw, divPQ, divQ, divS2;
  Epi epi;
  struct RowA { int a, dh, dw; bool ok; };
  struct ColB { int n; bool ok; };
  struct KInfo { int64_t base; int hs, ws; int64_t pix1, pix2; bool ok; };
  __device__ __forceinline__ RowA rowA(int m) const {
    int tap, a, r, s;
    divCa.divmod(m, tap, a);
    divKw.divmod(tap, r, s);
    return {a, r - padh, s - padw, m < M};
  }
  __device__ __forceinline__ ColB colB(int n) const { return {n, n < N}; }
  __device__ __forceinline__ KInfo decodeK(int k, int kend) const {
    int n, rem, p, q, q2, n2;
    divPQ.divmod(k, n, rem);
    divQ.divmod(rem, p, q);
    divS2.divmod(n, q2, n2);
    return {(int64_t)n * H * W, p * st, q * st, ((int64_t)n * P + p) * Q + q, ((int64_t)n2 * P + p) * Q + q,
            k < kend};
  }
  __device__ __forceinline__ float fetchA(const RowA& r, const KInfo& k) const {
    int h = k.hs + r.dh, ww = k.ws + r.dw;
    bool ok = r.ok && k.ok && (unsigned)h < (unsigned)H && (unsigned)ww < (unsigned)W;
    return ok ? __ldg(big + (k.base + (int64_t)h * W + ww) * Ca + r.a) : 0.f;
  }
  __device__ __forceinline__ float fetchB(const KInfo& k, const ColB& c) const {
    if (!(c.ok && k.ok)) return 0.f;
    if (c.n < Cb1) return __ldg(s1 + k.pix1 * Cb1 + c.n);
    return __ldg(s2 + k.pix2 * Cb2 + (c.n - Cb1));
  }
  __device__ __forceinline__ int64_t rowOut(int m) const { return (int64_t)m * N; }
  __device__ __forceinline__ float* out() const { return dw; }
};

// ---- the kernel --------------------------------------------------------------------------------------------
template <int T, int B>
__device__ __forceinline__ constexpr int frag_index(int t, int i) {
  // T >= 4: groups of 4 consecutive elements, groups strided across the tile (conflict-free float4 smem reads)
  return (T >= 4) ? (i / 4) * (B / (T / 4)) + t * 4 + (i % 4) : t * T + i;
}

template <class Prob, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256) simt_gemm_kernel(const Prob p, int k_per_split) {
  constexpr int BK = 8, NT = 256;
  static_assert((BM / TM) * (BN / TN) == NT, "thread layout");
  constexpr int EA = (BM * BK + NT - 1) / NT;
  constexpr int EB = (BN * BK + NT - 1) / NT;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int t = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(p.K, kbeg + k_per_split);

  int a_x[EA], a_k[EA], b_x[EB], b_k[EB];
  bool a_v[EA], b_v[EB];
  typename Prob::RowA ra[EA];
  typename Prob::ColB cb[EB];
#pragma unroll
  for (int i = 0; i < EA; ++i) {
    int e = t + NT * i;
    a_v[i] = e < BM * BK;
    if (Prob::A_KCONTIG) { a_k[i] = e % BK; a_x[i] = e / BK; } else { a_x[i] = e % BM; a_k[i] = e / BM; }
    ra[i] = p.rowA(m0 + a_x[i]);
    ra[i].ok = ra[i].ok && a_v[i];
  }
#pragma unroll
  for (int i = 0; i < EB; ++i) {
    int e = t + NT * i;
    b_v[i] = e < BN * BK;
    if (Prob::B_NCONTIG) { b_x[i] = e % BN; b_k[i] = e / BN; } else { b_k[i] = e % BK; b_x[i] = e / BK; }
    cb[i] = p.colB(n0 + b_x[i]);
    cb[i].ok = cb[i].ok && b_v[i];
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float pa[EA], pb[EB];
  auto gload = [&](int k0) {
    if (Prob::A_KCONTIG) {
      typename Prob::KInfo ki = p.decodeK(k0 + a_k[0], kend);
#pragma unroll
      for (int i = 0; i < EA; ++i) pa[i] = p.fetchA(ra[i], ki);
    } else {
#pragma unroll
      for (int i = 0; i < EA; ++i) pa[i] = p.fetchA(ra[i], p.decodeK(k0 + a_k[i], kend));
    }
    if (!Prob::B_NCONTIG) {
      typename Prob::KInfo ki = p.decodeK(k0 + b_k[0], kend);
#pragma unroll
      for (int i = 0; i < EB; ++i) pb[i] = p.fetchB(ki, cb[i]);
    } else {
#pragma unroll
      for (int i = 0; i < EB; ++i) pb[i] = p.fetchB(p.decodeK(k0 + b_k[i], kend), cb[i]);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < EA; ++i)
      if (a_v[i]) As[buf][a_k[i]][a_x[i]] = pa[i];
#pragma unroll
    for (int i = 0; i < EB; ++i)
      if (b_v[i]) Bs[buf][b_k[i]][b_x[i]] = pb[i];
  };

  const int ty = t / (BN / TN), tx = t % (BN / TN);
  const int nk = (kend - kbeg + BK - 1) / BK;
  if (nk > 0) { gload(kbeg); sstore(0); }
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) gload(kbeg + (kb + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
      if (TM >= 4) {
#pragma unroll
        for (int g = 0; g < TM / 4; ++g) {
          float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][frag_index<TM, BM>(ty, g * 4)]);
          a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[buf][kk][frag_index<TM, BM>(ty, i)];
      }
      if (TN >= 4) {
#pragma unroll
        for (int g = 0; g < TN / 4; ++g) {
          float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][frag_index<TN, BN>(tx, g * 4)]);
          b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[buf][kk][frag_index<TN, BN>(tx, j)];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kb + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
  }

  float* out = p.out();
  const bool atomic = Prob::ATOMIC_OUT || gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + frag_index<TM, BM>(ty, i);
    if (m >= p.M) continue;
    const int64_t ro = p.rowOut(m);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + frag_index<TN, BN>(tx, j);
      if (n >= p.N) continue;
      const int64_t idx = ro + n;
      if (atomic) atomicAdd(out + idx, acc[i][j]);
      else out[idx] = p.epi.apply(acc[i][j], n, idx, out);
    }
  }
}

// ---- host dispatch ---------------------------------------------------------------------------------------------
template <class Prob>
static int launch_simt(const Prob& p, int max_splits, cudaStream_t stream, const char* name) {
  if (p.M <= 0 || p.N <= 0) return CTX_OK;
  const int sms = sm_count();
  if (sms <= 0) return cuda_fail(cudaErrorNoDevice, name);
  auto run = [&](auto kern, int BM, int BN) -> int {
    int64_t gx = cdiv(p.M, BM), gy = cdiv(p.N, BN);
    int splits = 1;
    if (max_splits > 1) {
      int64_t tiles = gx * gy;
      int64_t want = cdiv(2 * (int64_t)sms, tiles);
      int64_t kblocks = cdiv(p.K, 8 * 16);                 // at least 16 k-steps per split
      splits = (int)(want < 1 ? 1 : (want > kblocks ? (kblocks < 1 ? 1 : kblocks) : want));
      if (splits > max_splits) splits = max_splits;
    }
    int kper = (int)(cdiv(cdiv(p.K, splits), 8) * 8);
    if (kper <= 0) kper = 8;
    splits = (int)cdiv(p.K, kper);
    if (splits < 1) splits = 1;
    if (gy > 65535 || splits > 65535) { set_error("%s: grid too large", name); return CTX_ERR_ARG; }
    dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)splits);
    kern<<<grid, 256, 0, stream>>>(p, kper);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, name);
    count_launch();
    return CTX_OK;
  };
  if (p.N <= 16) return run(simt_gemm_kernel<Prob, 128, 16, 8, 1>, 128, 16);
  if (p.M >= 1024 && p.N >= 96) return run(simt_gemm_kernel<Prob, 128, 128, 8, 8>, 128, 128);
  return run(simt_gemm_kernel<Prob, 64, 64, 4, 4>, 64, 64);
}

static bool fits_i32(int64_t v) { return v >= 0 && v < (int64_t)2147483647; }

int simt_conv2d_fprop(const float* x, int C, const float* w, float* y, int y_ld, int Co, const ctx_conv_geom& g,
                      const Epi& epi, cudaStream_t stream) {
  FpropProb p;
  int64_t M = (int64_t)g.N * g.P * g.Q, K = (int64_t)g.kh * g.kw * C;
  CTX_REQUIRE(fits_i32(M) && fits_i32(K) && fits_i32((int64_t)g.N * g.H * g.W * C) && fits_i32(M * y_ld),
              "conv2d_fprop: tensor too large for 32-bit indexing");
  p.M = (int)M; p.N = Co; p.K = (int)K; p.ldy = y_ld;
  p.x = x; p.w = w; p.y = y;
  p.C = C; p.H = g.H; p.W = g.W; p.st = g.stride; p.padh = g.pad_h; p.padw = g.pad_w; p.kw = g.kw;
  p.divPQ = FastDiv(g.P * g.Q); p.divQ = FastDiv(g.Q); p.divC = FastDiv(C); p.divKw = FastDiv(g.kw);
  p.epi = epi;
  return launch_simt(p, 1, stream, "simt_conv2d_fprop");
}

int simt_conv2d_transpose(const float* x1, int C1, const float* x2, int C2, int x2_batch, const float* w,
                          float* y, int Cn, const ctx_conv_geom& g, const Epi& epi, cudaStream_t stream) {
  const int st = g.stride;
  CTX_REQUIRE(fits_i32((int64_t)g.N * g.H * g.W * Cn) && fits_i32((int64_t)g.N * g.P * g.Q * (C1 + C2)),
              "conv2d_transpose: tensor too large for 32-bit indexing");
  for (int ph = 0; ph < st; ++ph) {
    for (int pw = 0; pw < st; ++pw) {
      const int Hp = g.H > ph ? (g.H - ph + st - 1) / st : 0;
      const int Wp = g.W > pw ? (g.W - pw + st - 1) / st : 0;
      if (Hp == 0 || Wp == 0) continue;
      TransProb p;
      p.r0 = (ph + g.pad_h) % st; p.s0 = (pw + g.pad_w) % st;
      const int nr = p.r0 < g.kh ? (g.kh - p.r0 + st - 1) / st : 0;
      const int ns = p.s0 < g.kw ? (g.kw - p.s0 + st - 1) / st : 0;
      p.ns = ns > 0 ? ns : 1;
      p.dh0 = (ph + g.pad_h - p.r0) / st; p.dw0 = (pw + g.pad_w - p.s0) / st;
      p.M = g.N * Hp * Wp; p.N = Cn; p.K = nr * ns * (C1 + C2);
      p.x1 = x1; p.x2 = x2; p.w = w; p.y = y;
      p.C1 = C1; p.C2 = C2; p.Ck = C1 + C2; p.x2_batch = x2_batch > 0 ? x2_batch : 1;
      p.P = g.P; p.Q = g.Q; p.H = g.H; p.W = g.W; p.st = st; p.kw = g.kw; p.ph = ph; p.pw = pw;
      p.divHpWp = FastDiv(Hp * Wp); p.divWp = FastDiv(Wp); p.divCk = FastDiv(p.Ck); p.divNs = FastDiv(p.ns);
      p.divX2 = FastDiv(p.x2_batch);
      p.epi = epi;
      int rc = launch_simt(p, 1, stream, "simt_conv2d_transpose");
      if (rc != CTX_OK) return rc;
    }
  }
  return CTX_OK;
}

int simt_conv2d_wgrad(const float* big, int Ca, const float* s1, int Cb1, const float* s2, int Cb2,
                      int s2_batch, float* dw, const ctx_conv_geom& g, cudaStream_t stream) {
  WgradProb p;
  int64_t K = (int64_t)g.N * g.P * g.Q;
  CTX_REQUIRE(fits_i32(K) && fits_i32((int64_t)g.N * g.H * g.W * Ca) && fits_i32(K * (Cb1 + Cb2)),
              "conv2d_wgrad: tensor too large for 32-bit indexing");
  p.M = g.kh * g.kw * Ca; p.N = Cb1 + Cb2; p.K = (int)K;
  p.big = big; p.s1 = s1; p.s2 = s2; p.dw = dw;
  p.Ca = Ca; p.Cb1 = Cb1; p.Cb2 = Cb2; p.s2_batch = s2_batch > 0 ? s2_batch : 1;
  p.P = g.P; p.Q = g.Q; p.H = g.H; p.W = g.W; p.st = g.stride; p.padh = g.pad_h; p.padw = g.pad_w; p.kw = g.kw;
  p.divCa = FastDiv(Ca); p.divKw = FastDiv(g.kw); p.divPQ = FastDiv(g.P * g.Q); p.divQ = FastDiv(g.Q);
  p.divS2 = FastDiv(p.s2_batch);
  p.epi = make_epi(nullptr);
  return launch_simt(p, 1024, stream, "simt_conv2d_wgrad");
}

int simt_gemm(int tA, int tB, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
              int ldc, const Epi& epi, cudaStream_t stream) {
  auto fill = [&](auto& p) {
    p.M = M; p.N = N; p.K = K; p.A = A; p.B = B; p.C = C; p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.epi = epi;
  };
  if (!tA && !tB) { GemmProb<false, false> p; fill(p); return launch_simt(p, 1, stream, "simt_gemm_nn"); }
  if (!tA && tB)  { GemmProb<false, true> p;  fill(p); return launch_simt(p, 1, stream, "simt_gemm_nt"); }
  if (tA && !tB)  { GemmProb<true, false> p;  fill(p); return launch_simt(p, 1, stream, "simt_gemm_tn"); }
  GemmProb<true, true> p; fill(p); return launch_simt(p, 1, stream, "simt_gemm_tt");
}

}  // namespace ctx
