// Generic (f32 / f64 / c64 / c128) dense tile product on the SIMT pipes:
//   C[b] (+)= op(A[b]) B[b],  A[b] is m x k (op=N) or k x m (op=T/H), B[b] k x n, C[b] m x n,
// all row-major, optional batch strides.  Serves
//   * the multi-column tile product of MPIMatrixMult for the dtypes the tensor
//     cores do not cover (pylops_mpi/basicoperators/MatrixMult.py:366-370, 409-413,
//     663-670, 742-763; parity cases of tests/test_matrixmult.py), and
//   * the per-slice product of MPIFredholm1 (signalprocessing/Fredholm1.py:119-129,
//     147-167), one slice per blockIdx.z.
// 64x64 CTA tile, 16-deep K slices through shared memory, 4x4 register tile per
// thread.  The bf16 tensor-core path lives in gemm_tc.cu.
#include <string.h>
#include "common.cuh"

namespace {

struct c32 { float re, im; };
struct c64 { double re, im; };

template <typename T> struct Num;
template <> struct Num<float> {
  __device__ static __forceinline__ float zero() { return 0.f; }
  __device__ static __forceinline__ void fma_(float& c, float a, float b) { c = fmaf(a, b, c); }
  __device__ static __forceinline__ float conj(float a) { return a; }
  __device__ static __forceinline__ float add(float a, float b) { return a + b; }
};
template <> struct Num<double> {
  __device__ static __forceinline__ double zero() { return 0.0; }
  __device__ static __forceinline__ void fma_(double& c, double a, double b) { c = fma(a, b, c); }
  __device__ static __forceinline__ double conj(double a) { return a; }
  __device__ static __forceinline__ double add(double a, double b) { return a + b; }
};
template <> struct Num<c32> {
  __device__ static __forceinline__ c32 zero() { return {0.f, 0.f}; }
  __device__ static __forceinline__ void fma_(c32& c, c32 a, c32 b) {
    c.re = fmaf(a.re, b.re, fmaf(-a.im, b.im, c.re));
    c.im = fmaf(a.re, b.im, fmaf(a.im, b.re, c.im));
  }
  __device__ static __forceinline__ c32 conj(c32 a) { return {a.re, -a.im}; }
  __device__ static __forceinline__ c32 add(c32 a, c32 b) { return {a.re + b.re, a.im + b.im}; }
};
template <> struct Num<c64> {
  __device__ static __forceinline__ c64 zero() { return {0.0, 0.0}; }
  __device__ static __forceinline__ void fma_(c64& c, c64 a, c64 b) {
    c.re = fma(a.re, b.re, fma(-a.im, b.im, c.re));
    c.im = fma(a.re, b.im, fma(a.im, b.re, c.im));
  }
  __device__ static __forceinline__ c64 conj(c64 a) { return {a.re, -a.im}; }
  __device__ static __forceinline__ c64 add(c64 a, c64 b) { return {a.re + b.re, a.im + b.im}; }
};

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

// extra destinations of the SAME logical output buffer in peer GPUs' memory (IPC-mapped, NVLink):
// the epilogue stores every result element locally and to each peer -> the all-gather of
// MPIFredholm1 (Fredholm1.py:131-132) happens inside the product kernel, tile by tile.
struct PeerDst {
  void* p[8];
  int n;
};

template <typename T, bool TRANS_A, bool FUSED = false>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const T* __restrict__ A, size_t lda, size_t sA, const T* __restrict__ B,
                 size_t ldb, size_t sB, T* __restrict__ C, size_t ldc, size_t sC, size_t m,
                 size_t n, size_t k, bool conj_a, bool accumulate, PeerDst peers = PeerDst{}) {
  using N_ = Num<T>;
  __shared__ T As[BK][BM + 1];
  __shared__ T Bs[BK][BN + 1];
  A += (size_t)blockIdx.z * sA;
  B += (size_t)blockIdx.z * sB;
  C += (size_t)blockIdx.z * sC;
  const size_t m0 = (size_t)blockIdx.y * BM, n0 = (size_t)blockIdx.x * BN;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  T acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = N_::zero();

  for (size_t k0 = 0; k0 < k; k0 += BK) {
    // A tile -> As[kk][i]
    for (int e = threadIdx.x; e < BM * BK; e += 256) {
      int i, kk;
      if (TRANS_A) { i = e % BM; kk = e / BM; }   // contiguous along m
      else { kk = e % BK; i = e / BK; }            // contiguous along k
      const size_t gi = m0 + i, gk = k0 + kk;
      T v = N_::zero();
      if (gi < m && gk < k) {
        v = TRANS_A ? A[gk * lda + gi] : A[gi * lda + gk];
        if (conj_a) v = N_::conj(v);
      }
      As[kk][i] = v;
    }
    for (int e = threadIdx.x; e < BN * BK; e += 256) {
      const int j = e % BN, kk = e / BN;
      const size_t gj = n0 + j, gk = k0 + kk;
      Bs[kk][j] = (gj < n && gk < k) ? B[gk * ldb + gj] : N_::zero();
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      T a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) N_::fma_(acc[i][j], a[i], b[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const size_t gi = m0 + ty * TM + i;
    if (gi >= m) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const size_t gj = n0 + tx + 16 * j;
      if (gj >= n) continue;
      T* c = C + gi * ldc + gj;
      const T v = accumulate ? N_::add(*c, acc[i][j]) : acc[i][j];
      *c = v;
      if (FUSED) {
        const size_t off = (size_t)blockIdx.z * sC + gi * ldc + gj;
        for (int d = 0; d < peers.n; ++d) reinterpret_cast<T*>(peers.p[d])[off] = v;   // P2P store over NVLink
      }
    }
  }
}

template <typename T>
int launch_gemm(const void* A, size_t lda, size_t sA, const void* B, size_t ldb, size_t sB,
                void* C, size_t ldc, size_t sC, size_t m, size_t n, size_t k, size_t batch,
                int op_a, bool accumulate, cudaStream_t st) {
  if (m == 0 || n == 0 || batch == 0) return B2_OK;
  if (batch > 65535) return B2_ERR_ARG;
  dim3 grid((unsigned)((n + BN - 1) / BN), (unsigned)((m + BM - 1) / BM), (unsigned)batch);
  if (grid.y > 65535u) return B2_ERR_ARG;
  const bool conj = (op_a == B2_OP_H);
  if (op_a == B2_OP_N)
    gemm_simt_kernel<T, false><<<grid, 256, 0, st>>>((const T*)A, lda, sA, (const T*)B, ldb, sB, (T*)C, ldc, sC, m, n, k, false, accumulate);
  else
    gemm_simt_kernel<T, true><<<grid, 256, 0, st>>>((const T*)A, lda, sA, (const T*)B, ldb, sB, (T*)C, ldc, sC, m, n, k, conj, accumulate);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

int dispatch(const void* A, size_t lda, size_t sA, const void* B, size_t ldb, size_t sB, void* C,
             size_t ldc, size_t sC, size_t m, size_t n, size_t k, size_t batch, int op_a,
             bool accumulate, int dtype, cudaStream_t st) {
  switch (dtype) {
    case B2_F32: return launch_gemm<float>(A, lda, sA, B, ldb, sB, C, ldc, sC, m, n, k, batch, op_a, accumulate, st);
    case B2_F64: return launch_gemm<double>(A, lda, sA, B, ldb, sB, C, ldc, sC, m, n, k, batch, op_a, accumulate, st);
    case B2_C64: return launch_gemm<c32>(A, lda, sA, B, ldb, sB, C, ldc, sC, m, n, k, batch, op_a, accumulate, st);
    case B2_C128: return launch_gemm<c64>(A, lda, sA, B, ldb, sB, C, ldc, sC, m, n, k, batch, op_a, accumulate, st);
    default: return B2_ERR_DTYPE;
  }
}

}  // namespace

template <typename T>
static int launch_fused(const void* G, const void* x, void* y, void* const* peers, int npeers, size_t nsl, size_t nx,
                 size_t ny, size_t nz, int adjoint, cudaStream_t st) {
  const size_t m = adjoint ? ny : nx, k = adjoint ? nx : ny;
  PeerDst pd;
  pd.n = npeers;
  for (int d = 0; d < 8; ++d) pd.p[d] = d < npeers ? peers[d] : nullptr;
  dim3 grid((unsigned)((nz + BN - 1) / BN), (unsigned)((m + BM - 1) / BM), (unsigned)nsl);
  if (adjoint)
    gemm_simt_kernel<T, true, true><<<grid, 256, 0, st>>>((const T*)G, ny, nx * ny, (const T*)x, nz, k * nz, (T*)y, nz,
                                                          m * nz, m, nz, k, true, false, pd);
  else
    gemm_simt_kernel<T, false, true><<<grid, 256, 0, st>>>((const T*)G, ny, nx * ny, (const T*)x, nz, k * nz, (T*)y, nz,
                                                           m * nz, m, nz, k, false, false, pd);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

// y[s] = op(G[s]) x[s] written to the local output AND to the same offsets of `npeers` peer buffers
// (fused product + all-gather over NVLink peer memory).  peers_host[d] must already point at the
// position in peer d's buffer that corresponds to y.
extern "C" int b2_batched_gemm_allgather(b2_ctx* ctx, const void* G, const void* x, void* y,
                                         void* const* peers_host, int npeers, size_t nsl, size_t nx, size_t ny,
                                         size_t nz, int adjoint, int dtype, void* stream) {
  if (!ctx || npeers < 0 || npeers > 8) return B2_ERR_ARG;
  if (nsl == 0) return B2_OK;
  if (!G || !x || !y || (npeers && !peers_host)) return B2_ERR_ARG;
  if (nsl > 65535) return B2_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case B2_F32: return launch_fused<float>(G, x, y, peers_host, npeers, nsl, nx, ny, nz, adjoint, st);
    case B2_F64: return launch_fused<double>(G, x, y, peers_host, npeers, nsl, nx, ny, nz, adjoint, st);
    case B2_C64: return launch_fused<c32>(G, x, y, peers_host, npeers, nsl, nx, ny, nz, adjoint, st);
    case B2_C128: return launch_fused<c64>(G, x, y, peers_host, npeers, nsl, nx, ny, nz, adjoint, st);
    default: return B2_ERR_DTYPE;
  }
}

// symmetric (peer-mappable) buffers: plain cudaMalloc + CUDA IPC handles exchanged by the caller
extern "C" int b2_symm_alloc(size_t bytes, void** out) {
  if (!out) return B2_ERR_ARG;
  B2_CUDA(cudaMalloc(out, bytes));
  return B2_OK;
}
extern "C" int b2_symm_free(void* p) {
  if (p) B2_CUDA(cudaFree(p));
  return B2_OK;
}
extern "C" int b2_ipc_get_handle(void* p, void* handle64_host) {
  if (!p || !handle64_host) return B2_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  B2_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(handle64_host, &h, sizeof h);
  return B2_OK;
}
extern "C" int b2_ipc_open_handle(const void* handle64_host, void** out) {
  if (!handle64_host || !out) return B2_ERR_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64_host, sizeof h);
  B2_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return B2_OK;
}
extern "C" int b2_ipc_close_handle(void* p) {
  if (p) B2_CUDA(cudaIpcCloseMemHandle(p));
  return B2_OK;
}

extern "C" int b2_gemm(b2_ctx* ctx, const void* A, size_t lda, const void* B, size_t ldb, void* C,
                       size_t ldc, size_t m, size_t n, size_t k, int op_a, int accumulate,
                       int dtype, void* stream) {
  if (!ctx) return B2_ERR_ARG;
  if (op_a != B2_OP_N && op_a != B2_OP_T && op_a != B2_OP_H) return B2_ERR_ARG;
  if (m && n && (!C)) return B2_ERR_ARG;
  if (m && n && k && (!A || !B)) return B2_ERR_ARG;
  return dispatch(A, lda, 0, B, ldb, 0, C, ldc, 0, m, n, k, 1, op_a, accumulate != 0, dtype,
                  (cudaStream_t)stream);
}

extern "C" int b2_batched_gemm(b2_ctx* ctx, const void* G, const void* x, void* y, size_t nsl,
                               size_t nx, size_t ny, size_t nz, int adjoint, int dtype,
                               void* stream) {
  if (!ctx) return B2_ERR_ARG;
  if (nsl == 0) return B2_OK;
  if (!G || !x || !y) return B2_ERR_ARG;
  // forward: y[s] (nx x nz) = G[s] (nx x ny) x[s] (ny x nz)
  // adjoint: y[s] (ny x nz) = G[s]^H (ny x nx) x[s] (nx x nz)
  const size_t m = adjoint ? ny : nx, k = adjoint ? nx : ny;
  size_t done = 0;
  while (done < nsl) {   // blockIdx.z limit
    size_t b = nsl - done < 65535 ? nsl - done : 65535;
    const size_t es = b2_dtype_size(dtype);
    int rc = dispatch((const char*)G + done * nx * ny * es, ny, nx * ny,
                      (const char*)x + done * k * nz * es, nz, k * nz,
                      (char*)y + done * m * nz * es, nz, m * nz, m, nz, k, b,
                      adjoint ? B2_OP_H : B2_OP_N, false, dtype, (cudaStream_t)stream);
    if (rc) return rc;
    done += b;
  }
  return B2_OK;
}
