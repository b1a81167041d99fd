// Fused proximal-gradient update behind ISTA / FISTA ("next" row: pylops_mpi/optimization/cls_sparsity.py:270-343,
// 578-662).  The reference makes 8+ passes over the model per iteration (x.copy, x + grad, threshold, x - xold, two
// norms, FISTA's z update); this is ONE pass:
//     u = base + alpha * g ;  v = thresh(u) ;  xnew = v ;  znew = v + c (v - xold)
//     sums[0] = sum |v - xold|^2 ,  sums[1] = sum |v|        (local partials; the caller all-reduces them)
// HBM-bound: algorithmic bytes per element = (2..3 reads + 1..2 writes) * sizeof(T).
// Thresholds restate third-party pylops.optimization.cls_sparsity._soft/_hard/_halfthreshold (pylops 2.x).
#include <math.h>
#include "common.cuh"

namespace {

constexpr int SP_THREADS = 256;

struct SpParams {
  const void* base;
  const void* g;
  const void* xold;
  void* xnew;
  void* znew;
  double alpha, thresh, c;
  double hard_cut, half_cut;
  size_t n_real;
  int kind;
};

__device__ __forceinline__ double thr_real(double u, const SpParams& p) {
  const double a = fabs(u);
  switch (p.kind) {
    case B2_THRESH_SOFT: return copysign(fmax(a - p.thresh, 0.0), u) * (a > 0.0 ? 1.0 : 0.0);
    case B2_THRESH_HARD: return a <= p.hard_cut ? 0.0 : u;
    case B2_THRESH_HALF: {
      if (a <= p.half_cut) return 0.0;
      double arg = (p.thresh / 8.0) * pow(a / 3.0, -1.5);
      arg = fmin(fmax(arg, -1.0), 1.0);
      const double phi = 2.0 / 3.0 * acos(arg);
      return 2.0 / 3.0 * u * (1.0 + cos(2.0 * 3.14159265358979323846 / 3.0 - phi));
    }
    default: return u;
  }
}
__device__ __forceinline__ void thr_cx(double& ur, double& ui, const SpParams& p) {
  const double a = hypot(ur, ui);
  double s = 1.0;
  if (p.kind == B2_THRESH_SOFT) s = a > 0.0 ? fmax(a - p.thresh, 0.0) / a : 0.0;
  else if (p.kind == B2_THRESH_HARD) s = a <= p.hard_cut ? 0.0 : 1.0;
  ur *= s;
  ui *= s;
}

// one "item" = 1 real scalar or 1 complex pair, all arithmetic in double
template <typename T, bool CX>
__device__ __forceinline__ void item(const SpParams& p, const T* b, const T* g, const T* xo, T* xn, T* zn,
                                     bool has_g, bool has_xo, bool has_zn, double* acc) {
  if (!CX) {
    double u = (double)b[0];
    if (has_g) u = fma(p.alpha, (double)g[0], u);
    const double v = thr_real(u, p);
    const double d = has_xo ? v - (double)xo[0] : 0.0;
    xn[0] = (T)v;
    if (has_zn) zn[0] = (T)fma(p.c, d, v);
    acc[0] = fma(d, d, acc[0]);
    acc[1] += fabs(v);
  } else {
    double ur = b[0], ui = b[1];
    if (has_g) { ur = fma(p.alpha, (double)g[0], ur); ui = fma(p.alpha, (double)g[1], ui); }
    thr_cx(ur, ui, p);
    const double dr = has_xo ? ur - (double)xo[0] : 0.0, di = has_xo ? ui - (double)xo[1] : 0.0;
    xn[0] = (T)ur; xn[1] = (T)ui;
    if (has_zn) { zn[0] = (T)fma(p.c, dr, ur); zn[1] = (T)fma(p.c, di, ui); }
    acc[0] += dr * dr + di * di;
    acc[1] += hypot(ur, ui);
  }
}

template <typename T, bool CX, bool VEC>
__global__ void __launch_bounds__(SP_THREADS)
sparse_update_kernel(const __grid_constant__ SpParams p, double* __restrict__ partials,
                     unsigned int* __restrict__ ticket, double* __restrict__ out) {
  constexpr int V = Vec16<T>::N;
  constexpr int STEP = CX ? 2 : 1;
  const T* base = (const T*)p.base;
  const T* g = (const T*)p.g;
  const T* xo = (const T*)p.xold;
  T* xn = (T*)p.xnew;
  T* zn = (T*)p.znew;
  const bool has_g = g != nullptr, has_xo = xo != nullptr, has_zn = zn != nullptr;
  double acc[2] = {0.0, 0.0};
  const size_t stride = (size_t)gridDim.x * SP_THREADS;
  size_t i = (size_t)blockIdx.x * SP_THREADS + threadIdx.x;
  if (VEC) {
    const size_t nvec = p.n_real / V;
    for (; i < nvec; i += stride) {
      // coherent loads: xnew / znew may alias base / xold (in-place update)
      Vec16<T> vb = load_vec_coherent(base + i * V), vg = vb, vo = vb, vx, vz;
      if (has_g) vg = load_vec_coherent(g + i * V);
      if (has_xo) vo = (xo == base) ? vb : load_vec_coherent(xo + i * V);
#pragma unroll
      for (int k = 0; k < V; k += STEP)
        item<T, CX>(p, vb.v + k, vg.v + k, vo.v + k, vx.v + k, vz.v + k, has_g, has_xo, has_zn, acc);
      store_vec(xn + i * V, vx);
      if (has_zn) store_vec(zn + i * V, vz);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      for (size_t t = nvec * V; t + STEP - 1 < p.n_real; t += STEP)
        item<T, CX>(p, base + t, g + t, xo + t, xn + t, zn + t, has_g, has_xo, has_zn, acc);
    }
  } else {
    const size_t nit = p.n_real / STEP;
    for (; i < nit; i += stride) {
      const size_t t = i * STEP;
      item<T, CX>(p, base + t, g + t, xo + t, xn + t, zn + t, has_g, has_xo, has_zn, acc);
    }
  }
  __shared__ double smem[2][SP_THREADS / 32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    double v = warp_sum(acc[k]);
    if (lane == 0) smem[k][warp] = v;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      double v = lane < SP_THREADS / 32 ? smem[k][lane] : 0.0;
      v = warp_sum(v);
      if (lane == 0) partials[(size_t)blockIdx.x * 2 + k] = v;
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && warp == 0) {
    __threadfence();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      double v = 0.0;
      for (unsigned int b = lane; b < gridDim.x; b += 32) v += __ldcg(&partials[(size_t)b * 2 + k]);
      v = warp_sum(v);
      if (lane == 0 && out) out[k] = v;
    }
    if (lane == 0) *ticket = 0u;
  }
}

template <typename T, bool CX>
int launch(b2_ctx* ctx, const SpParams& p, double* sums, cudaStream_t st) {
  constexpr int V = Vec16<T>::N;
  const bool vec = b2_aligned16(p.base) && b2_aligned16(p.g) && b2_aligned16(p.xold) && b2_aligned16(p.xnew) &&
                   b2_aligned16(p.znew) && p.n_real >= (size_t)V;
  size_t items = vec ? p.n_real / V : p.n_real;
  size_t need = (items + SP_THREADS - 1) / SP_THREADS;
  size_t cap = (size_t)ctx->sm_count * 8;
  if (cap > (size_t)B2_RED_MAX_BLOCKS) cap = B2_RED_MAX_BLOCKS;
  int grid = (int)(need < 1 ? 1 : (need < cap ? need : cap));
  if (vec)
    sparse_update_kernel<T, CX, true><<<grid, SP_THREADS, 0, st>>>(p, ctx->red_partials, ctx->tickets, sums);
  else
    sparse_update_kernel<T, CX, false><<<grid, SP_THREADS, 0, st>>>(p, ctx->red_partials, ctx->tickets, sums);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

}  // namespace

extern "C" int b2_sparse_update(b2_ctx* ctx, const void* base, const void* g, double alpha, const void* xold,
                                double thresh, int kind, void* xnew, void* znew, double c, double* sums_dev,
                                size_t n, int dtype, void* stream) {
  if (!ctx) return B2_ERR_ARG;
  if (kind < B2_THRESH_NONE || kind > B2_THRESH_HALF || thresh < 0.0) return B2_ERR_ARG;
  const bool cx = dtype == B2_C64 || dtype == B2_C128;
  if (cx && kind == B2_THRESH_HALF) return B2_ERR_UNSUPPORTED;
  if (b2_dtype_size(dtype) == 0 || dtype == B2_BF16 || dtype == B2_I64) return B2_ERR_DTYPE;
  if (n == 0) {  // a rank may own no model elements; its partial sums are still all-reduced
    if (sums_dev) B2_CUDA(cudaMemsetAsync(sums_dev, 0, 2 * sizeof(double), (cudaStream_t)stream));
    return B2_OK;
  }
  if (!base || !xnew || (znew && !xold)) return B2_ERR_ARG;
  SpParams p;
  p.base = base; p.g = g; p.xold = xold; p.xnew = xnew; p.znew = znew;
  p.alpha = alpha; p.thresh = thresh; p.c = c; p.kind = kind;
  p.hard_cut = sqrt(2.0 * thresh);
  p.half_cut = (cbrt(54.0) / 4.0) * pow(thresh, 2.0 / 3.0);
  p.n_real = cx ? 2 * n : n;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case B2_F32: return launch<float, false>(ctx, p, sums_dev, st);
    case B2_F64: return launch<double, false>(ctx, p, sums_dev, st);
    case B2_C64: return launch<float, true>(ctx, p, sums_dev, st);
    case B2_C128: return launch<double, true>(ctx, p, sums_dev, st);
    default: return B2_ERR_DTYPE;
  }
}
