// Dense per-rank matvec y = op(A) x  (the pylops.MatrixMult block applied by
// MPIBlockDiag / MPIVStack: pylops_mpi/basicoperators/BlockDiag.py:127-129,
// 139-141; VStack.py:129-131,144-145; and the single-RHS tile product of
// MPIMatrixMult, MatrixMult.py:366-370, 670).
//
// Single right-hand side => 0.5..1 flop per byte of A => HBM-bound; tensor cores
// do not apply.  Algorithmic bytes = m*n*sizeof(A) (+ vectors).
//  * op = N : one warp per row, 16-byte loads along the row, 4 in flight per lane,
//            x served from L1/L2, warp-shuffle reduction.
//  * op = T/H: one lane per 16-byte column vector, warps stride over the rows of a
//            row chunk, CTA-level smem fold, chunk partials folded in chunk order
//            by the last CTA of each column tile (deterministic, no atomics on y).
#include <cstdlib>
#include "common.cuh"

namespace {

// ---- element traits ---------------------------------------------------------
struct cf32 { float re, im; };
struct cf64 { double re, im; };

template <typename TA> struct ElemTraits;
template <> struct ElemTraits<float> {
  using X = float; using Acc = float; static constexpr int V = 4;
  __device__ static __forceinline__ Acc zero() { return 0.f; }
  __device__ static __forceinline__ void fma_(Acc& acc, float a, float x, bool) { acc = fmaf(a, x, acc); }
  __device__ static __forceinline__ Acc add(Acc a, Acc b) { return a + b; }
};
template <> struct ElemTraits<double> {
  using X = double; using Acc = double; static constexpr int V = 2;
  __device__ static __forceinline__ Acc zero() { return 0.0; }
  __device__ static __forceinline__ void fma_(Acc& acc, double a, double x, bool) { acc = fma(a, x, acc); }
  __device__ static __forceinline__ Acc add(Acc a, Acc b) { return a + b; }
};
template <> struct ElemTraits<cf32> {
  using X = cf32; using Acc = cf32; static constexpr int V = 2;
  __device__ static __forceinline__ Acc zero() { return {0.f, 0.f}; }
  __device__ static __forceinline__ void fma_(Acc& acc, cf32 a, cf32 x, bool conj) {
    float ai = conj ? -a.im : a.im;
    acc.re = fmaf(a.re, x.re, fmaf(-ai, x.im, acc.re));
    acc.im = fmaf(a.re, x.im, fmaf(ai, x.re, acc.im));
  }
  __device__ static __forceinline__ Acc add(Acc a, Acc b) { return {a.re + b.re, a.im + b.im}; }
};
template <> struct ElemTraits<cf64> {
  using X = cf64; using Acc = cf64; static constexpr int V = 1;
  __device__ static __forceinline__ Acc zero() { return {0.0, 0.0}; }
  __device__ static __forceinline__ void fma_(Acc& acc, cf64 a, cf64 x, bool conj) {
    double ai = conj ? -a.im : a.im;
    acc.re = fma(a.re, x.re, fma(-ai, x.im, acc.re));
    acc.im = fma(a.re, x.im, fma(ai, x.re, acc.im));
  }
  __device__ static __forceinline__ Acc add(Acc a, Acc b) { return {a.re + b.re, a.im + b.im}; }
};
template <> struct ElemTraits<__nv_bfloat16> {
  using X = float; using Acc = float; static constexpr int V = 8;
  __device__ static __forceinline__ Acc zero() { return 0.f; }
  __device__ static __forceinline__ void fma_(Acc& acc, __nv_bfloat16 a, float x, bool) {
    acc = fmaf(__bfloat162float(a), x, acc);
  }
  __device__ static __forceinline__ Acc add(Acc a, Acc b) { return a + b; }
};

template <typename A> __device__ __forceinline__ A shfl_xor_t(A v, int o);
template <> __device__ __forceinline__ float shfl_xor_t(float v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
template <> __device__ __forceinline__ double shfl_xor_t(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
template <> __device__ __forceinline__ cf32 shfl_xor_t(cf32 v, int o) {
  return {__shfl_xor_sync(0xffffffffu, v.re, o), __shfl_xor_sync(0xffffffffu, v.im, o)};
}
template <> __device__ __forceinline__ cf64 shfl_xor_t(cf64 v, int o) {
  return {__shfl_xor_sync(0xffffffffu, v.re, o), __shfl_xor_sync(0xffffffffu, v.im, o)};
}

template <typename TA>
struct AVec {  // 16 bytes of A
  TA v[ElemTraits<TA>::V];
};
template <typename TA>
__device__ __forceinline__ AVec<TA> load_a(const TA* p) {
  uint4 r = ldg_stream16(p);
  AVec<TA> o;
  *reinterpret_cast<uint4*>(&o) = r;
  return o;
}

// cached scalar loads for every element type
__device__ __forceinline__ float ldg_t(const float* p) { return __ldg(p); }
__device__ __forceinline__ double ldg_t(const double* p) { return __ldg(p); }
__device__ __forceinline__ cf32 ldg_t(const cf32* p) {
  float2 t = __ldg(reinterpret_cast<const float2*>(p));
  return {t.x, t.y};
}
__device__ __forceinline__ cf64 ldg_t(const cf64* p) {
  double2 t = __ldg(reinterpret_cast<const double2*>(p));
  return {t.x, t.y};
}
__device__ __forceinline__ float ldcg_t(const float* p) { return __ldcg(p); }
__device__ __forceinline__ double ldcg_t(const double* p) { return __ldcg(p); }
__device__ __forceinline__ cf32 ldcg_t(const cf32* p) {
  float2 t = __ldcg(reinterpret_cast<const float2*>(p));
  return {t.x, t.y};
}
__device__ __forceinline__ cf64 ldcg_t(const cf64* p) {
  double2 t = __ldcg(reinterpret_cast<const double2*>(p));
  return {t.x, t.y};
}

template <typename TA>
struct XVec {  // the x elements matching one 16-byte vector of A
  typename ElemTraits<TA>::X v[ElemTraits<TA>::V];
};
template <typename TA>
__device__ __forceinline__ XVec<TA> load_x(const typename ElemTraits<TA>::X* p) {
  XVec<TA> o;
  constexpr int NQ = sizeof(XVec<TA>) / 16;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&o);
#pragma unroll
  for (int i = 0; i < NQ; ++i) d[i] = __ldg(q + i);
  return o;
}

// -------------------------------------------------------------------------
// op = N : y_i = sum_j A_ij x_j
// -------------------------------------------------------------------------
constexpr int GN_WARPS = 8;
constexpr int GN_UNROLL = 4;

template <typename TA, bool VEC>
__global__ void __launch_bounds__(GN_WARPS * 32)
gemv_n_kernel(const TA* __restrict__ A, size_t lda, size_t m, size_t n,
              const typename ElemTraits<TA>::X* __restrict__ x,
              typename ElemTraits<TA>::X* __restrict__ y) {
  using Tr = ElemTraits<TA>;
  using Acc = typename Tr::Acc;
  constexpr int V = Tr::V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t row = (size_t)blockIdx.x * GN_WARPS + warp;
  if (row >= m) return;
  const TA* a = A + row * lda;
  Acc acc[GN_UNROLL];
#pragma unroll
  for (int u = 0; u < GN_UNROLL; ++u) acc[u] = Tr::zero();
  size_t j = 0;
  if (VEC) {
    const size_t nvec = n / V;
    size_t v = lane;
    for (; v + (GN_UNROLL - 1) * 32 < nvec; v += GN_UNROLL * 32) {
      AVec<TA> av[GN_UNROLL];
#pragma unroll
      for (int u = 0; u < GN_UNROLL; ++u) av[u] = load_a(a + (v + u * 32) * V);
#pragma unroll
      for (int u = 0; u < GN_UNROLL; ++u) {
        XVec<TA> xv = load_x<TA>(x + (v + u * 32) * V);
#pragma unroll
        for (int e = 0; e < V; ++e) Tr::fma_(acc[u], av[u].v[e], xv.v[e], false);
      }
    }
    for (; v < nvec; v += 32) {
      AVec<TA> av = load_a(a + v * V);
      XVec<TA> xv = load_x<TA>(x + v * V);
#pragma unroll
      for (int e = 0; e < V; ++e) Tr::fma_(acc[0], av.v[e], xv.v[e], false);
    }
    j = nvec * V;
  }
  for (size_t jj = j + lane; jj < n; jj += 32) Tr::fma_(acc[0], a[jj], x[jj], false);
  Acc s = Tr::add(Tr::add(acc[0], acc[1]), Tr::add(acc[2], acc[3]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s = Tr::add(s, shfl_xor_t(s, o));
  if (lane == 0) y[row] = s;
}


// Long rows (>= 32 KB): with one warp per row a warp streams its whole row alone, so the kernel ends with a long
// tail in which the last few hundred warps run latency-bound (measured on 32768-column bf16 panels: a fixed ~40 us
// on top of bytes / bandwidth -- 0.94 of the HBM peak at 32768 rows, 0.84 at 16384, ~0.65 at 8192).  Here the 8 warps
// of a CTA sweep R rows TOGETHER, each warp taking every 8th 512-byte piece of all R rows: 2 R independent 16-byte
// loads per lane in flight, every x vector loaded once per R rows, work per CTA R rows instead of 8 -- a shorter,
// steeper tail.  Partial sums meet in shared memory and are added in warp order (deterministic).
constexpr int GS_WARPS = 8;
template <typename TA, int R>
__global__ void __launch_bounds__(GS_WARPS * 32)
gemv_n_split_kernel(const TA* __restrict__ A, size_t lda, size_t m, size_t n,
                    const typename ElemTraits<TA>::X* __restrict__ x,
                    typename ElemTraits<TA>::X* __restrict__ y) {
  using Tr = ElemTraits<TA>;
  using Acc = typename Tr::Acc;
  constexpr int V = Tr::V;
  constexpr size_t STEP = GS_WARPS * 32;
  __shared__ Acc part[GS_WARPS][R];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t row0 = (size_t)blockIdx.x * R;
  const TA* a[R];
  Acc acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const size_t row = row0 + r < m ? row0 + r : m - 1;     // rows past the end re-read the last row; never stored
    a[r] = A + row * lda;
    acc[r] = Tr::zero();
  }
  const size_t nvec = n / V;
  size_t v = (size_t)warp * 32 + lane;
  for (; v + STEP < nvec; v += 2 * STEP) {
    AVec<TA> a0[R], a1[R];
#pragma unroll
    for (int r = 0; r < R; ++r) a0[r] = load_a(a[r] + v * V);
#pragma unroll
    for (int r = 0; r < R; ++r) a1[r] = load_a(a[r] + (v + STEP) * V);
    const XVec<TA> x0 = load_x<TA>(x + v * V), x1 = load_x<TA>(x + (v + STEP) * V);
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int e = 0; e < V; ++e) Tr::fma_(acc[r], a0[r].v[e], x0.v[e], false);
#pragma unroll
      for (int e = 0; e < V; ++e) Tr::fma_(acc[r], a1[r].v[e], x1.v[e], false);
    }
  }
  for (; v < nvec; v += STEP) {
    const XVec<TA> x0 = load_x<TA>(x + v * V);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const AVec<TA> a0 = load_a(a[r] + v * V);
#pragma unroll
      for (int e = 0; e < V; ++e) Tr::fma_(acc[r], a0.v[e], x0.v[e], false);
    }
  }
  if (warp == 0)
    for (size_t jj = nvec * V + lane; jj < n; jj += 32) {
#pragma unroll
      for (int r = 0; r < R; ++r) Tr::fma_(acc[r], a[r][jj], x[jj], false);
    }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    Acc s = acc[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = Tr::add(s, shfl_xor_t(s, o));
    if (lane == 0) part[warp][r] = s;
  }
  __syncthreads();
  if (threadIdx.x < R && row0 + threadIdx.x < m) {
    Acc s = part[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < GS_WARPS; ++w) s = Tr::add(s, part[w][threadIdx.x]);
    y[row0 + threadIdx.x] = s;
  }
}

inline bool gemv_split_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("B2_GEMV_SPLIT");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

// -------------------------------------------------------------------------
// op = T / H : y_j = sum_i op(A_ij) x_i
// grid = (column tiles, row chunks); CTA = 8 warps; lane <-> 16-byte column vector
// -------------------------------------------------------------------------
constexpr int GT_WARPS = 8;
constexpr int GT_ROWS = 128;   // rows per chunk

template <typename TA, bool VEC>
__global__ void __launch_bounds__(GT_WARPS * 32)
gemv_t_kernel(const TA* __restrict__ A, size_t lda, size_t m, size_t n,
              const typename ElemTraits<TA>::X* __restrict__ x,
              typename ElemTraits<TA>::X* __restrict__ y,
              typename ElemTraits<TA>::Acc* __restrict__ partials,
              unsigned int* __restrict__ tickets, bool conj) {
  using Tr = ElemTraits<TA>;
  using Acc = typename Tr::Acc;
  constexpr int V = VEC ? Tr::V : 1;
  constexpr int TILE = 32 * V;   // columns per CTA
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t col0 = (size_t)blockIdx.x * TILE + (size_t)lane * V;
  const size_t r0 = (size_t)blockIdx.y * GT_ROWS;
  const size_t r1 = (r0 + GT_ROWS < m) ? r0 + GT_ROWS : m;
  Acc acc[V];
#pragma unroll
  for (int e = 0; e < V; ++e) acc[e] = Tr::zero();
  if (VEC) {
    if (col0 + V <= n) {
      size_t i = r0 + warp;
      // 4 rows in flight per warp
      for (; i + 3 * GT_WARPS < r1; i += 4 * GT_WARPS) {
        AVec<TA> av[4];
        typename Tr::X xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) av[u] = load_a(A + (i + u * GT_WARPS) * lda + col0);
#pragma unroll
        for (int u = 0; u < 4; ++u) xv[u] = ldg_t(x + i + u * GT_WARPS);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int e = 0; e < V; ++e) Tr::fma_(acc[e], av[u].v[e], xv[u], conj);
      }
      for (; i < r1; i += GT_WARPS) {
        AVec<TA> av = load_a(A + i * lda + col0);
        typename Tr::X xv = ldg_t(x + i);
#pragma unroll
        for (int e = 0; e < V; ++e) Tr::fma_(acc[e], av.v[e], xv, conj);
      }
    } else {
      for (size_t i = r0 + warp; i < r1; i += GT_WARPS) {
        typename Tr::X xv = x[i];
#pragma unroll
        for (int e = 0; e < V; ++e)
          if (col0 + e < n) Tr::fma_(acc[e], A[i * lda + col0 + e], xv, conj);
      }
    }
  } else {
    if (col0 < n)
      for (size_t i = r0 + warp; i < r1; i += GT_WARPS) Tr::fma_(acc[0], A[i * lda + col0], x[i], conj);
  }
  // fold the 8 warps through shared memory (fixed order)
  __shared__ Acc smem[GT_WARPS][32 * (VEC ? Tr::V : 1)];
  __shared__ bool is_last;
#pragma unroll
  for (int e = 0; e < V; ++e) smem[warp][lane * V + e] = acc[e];
  __syncthreads();
  const int nchunks = gridDim.y;
  for (int c = threadIdx.x; c < TILE; c += GT_WARPS * 32) {
    Acc s = smem[0][c];
#pragma unroll
    for (int w = 1; w < GT_WARPS; ++w) s = Tr::add(s, smem[w][c]);
    const size_t col = (size_t)blockIdx.x * TILE + c;
    if (col < n) {
      if (nchunks == 1) y[col] = s;
      else partials[(size_t)blockIdx.y * n + col] = s;
    }
  }
  if (nchunks == 1) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(&tickets[blockIdx.x], 1u);
    is_last = (t == (unsigned)nchunks - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    for (int c = threadIdx.x; c < TILE; c += GT_WARPS * 32) {
      const size_t col = (size_t)blockIdx.x * TILE + c;
      if (col < n) {
        Acc s = ldcg_t(&partials[col]);
        for (int k = 1; k < nchunks; ++k) s = Tr::add(s, ldcg_t(&partials[(size_t)k * n + col]));
        y[col] = s;
      }
    }
    if (threadIdx.x == 0) tickets[blockIdx.x] = 0u;
  }
}


template <typename TA>
int launch_gemv(b2_ctx* ctx, const void* A, size_t lda, size_t m, size_t n, const void* x,
                void* y, int op, cudaStream_t st) {
  using Tr = ElemTraits<TA>;
  using X = typename Tr::X;
  using Acc = typename Tr::Acc;
  constexpr int V = Tr::V;
  const bool vec = b2_aligned16(A) && ((lda * sizeof(TA)) % 16 == 0);
  if (op == B2_OP_N) {
    if (m == 0) return B2_OK;
    unsigned grid = (unsigned)((m + GN_WARPS - 1) / GN_WARPS);
    if (vec && b2_aligned16(x) && n * sizeof(TA) >= 32768 && gemv_split_enabled()) {
      constexpr int R = 4;
      gemv_n_split_kernel<TA, R><<<(unsigned)((m + R - 1) / R), GS_WARPS * 32, 0, st>>>((const TA*)A, lda, m, n, (const X*)x,
                                                                                        (X*)y);
      B2_LAUNCH_CHECK();
      return B2_OK;
    }
    if (vec && b2_aligned16(x) && n >= (size_t)V)
      gemv_n_kernel<TA, true><<<grid, GN_WARPS * 32, 0, st>>>((const TA*)A, lda, m, n, (const X*)x, (X*)y);
    else
      gemv_n_kernel<TA, false><<<grid, GN_WARPS * 32, 0, st>>>((const TA*)A, lda, m, n, (const X*)x, (X*)y);
    B2_LAUNCH_CHECK();
    return B2_OK;
  }
  if (n == 0) return B2_OK;
  const bool conj = (op == B2_OP_H);
  const int tile = 32 * (vec ? V : 1);
  const size_t ntiles = (n + tile - 1) / tile;
  size_t nchunks = (m + GT_ROWS - 1) / GT_ROWS;
  if (nchunks < 1) nchunks = 1;
  if (ntiles > (size_t)B2_TICKETS) return B2_ERR_WORKSPACE;
  if (nchunks > 65535) return B2_ERR_ARG;
  const size_t need = nchunks * n * sizeof(Acc);
  if (nchunks > 1 && need > ctx->gemv_partials_bytes) {
    // grow scratch (stream-ordered would be nicer; this happens once per shape class)
    B2_CUDA(cudaStreamSynchronize(st));
    if (ctx->gemv_partials) B2_CUDA(cudaFree(ctx->gemv_partials));
    ctx->gemv_partials = nullptr;
    ctx->gemv_partials_bytes = 0;
    B2_CUDA(cudaMalloc((void**)&ctx->gemv_partials, need));
    ctx->gemv_partials_bytes = need;
  }
  dim3 grid((unsigned)ntiles, (unsigned)nchunks);
  // tickets for gemv live after the first 64 slots (slot 0 is the reduction ticket)
  unsigned int* tk = ctx->tickets + 64;
  if (ntiles + 64 > (size_t)B2_TICKETS) return B2_ERR_WORKSPACE;
  if (vec)
    gemv_t_kernel<TA, true><<<grid, GT_WARPS * 32, 0, st>>>((const TA*)A, lda, m, n, (const X*)x, (X*)y, (Acc*)ctx->gemv_partials, tk, conj);
  else
    gemv_t_kernel<TA, false><<<grid, GT_WARPS * 32, 0, st>>>((const TA*)A, lda, m, n, (const X*)x, (X*)y, (Acc*)ctx->gemv_partials, tk, conj);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

}  // namespace

extern "C" int b2_gemv(b2_ctx* ctx, const void* A, size_t lda, size_t m, size_t n, const void* x,
                       void* y, int op, int dtype_a, int dtype_xy, void* stream) {
  if (!ctx) return B2_ERR_ARG;
  if (op != B2_OP_N && op != B2_OP_T && op != B2_OP_H) return B2_ERR_ARG;
  if (m == 0 && n == 0) return B2_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // empty contraction -> zero output
  const size_t out_len = (op == B2_OP_N) ? m : n, in_len = (op == B2_OP_N) ? n : m;
  if (in_len == 0) {
    if (out_len) B2_CUDA(cudaMemsetAsync(y, 0, out_len * b2_dtype_size(dtype_xy), st));
    return B2_OK;
  }
  if (!A || !x || !y) return B2_ERR_ARG;
  if (lda < n) return B2_ERR_ARG;
  if (dtype_a == B2_BF16) {
    if (dtype_xy != B2_F32) return B2_ERR_DTYPE;
    return launch_gemv<__nv_bfloat16>(ctx, A, lda, m, n, x, y, op, st);
  }
  if (dtype_a != dtype_xy) return B2_ERR_DTYPE;
  switch (dtype_a) {
    case B2_F32: return launch_gemv<float>(ctx, A, lda, m, n, x, y, op, st);
    case B2_F64: return launch_gemv<double>(ctx, A, lda, m, n, x, y, op, st);
    case B2_C64: return launch_gemv<cf32>(ctx, A, lda, m, n, x, y, op, st);
    case B2_C128: return launch_gemv<cf64>(ctx, A, lda, m, n, x, y, op, st);
    default: return B2_ERR_DTYPE;
  }
}
