// MPIFirstDerivative per-rank apply (reference:
// pylops_mpi/basicoperators/FirstDerivative.py:129-319).
//
// The operator is the banded matrix D (forward) or D^T (adjoint) acting along
// axis 0 of a [nrows_global x ncols] array; this rank owns rows
// [row0, row0+nrows_local) and receives up to 2 neighbour rows each side
// (the add_ghost_cells payload).  Every output row i is
//     y[i,:] = (sum_{k=-2..2} tap_k(i) * x[i+k,:]) * (1/sampling)
// with constant taps for interior rows and per-row taps for the first / last
// four GLOBAL rows (edge handling).  Adjoint taps are generated as the exact
// transpose of the forward taps, tap^T_k(i) = tap_{-k}(i+k), so <Dx,y> = <x,D^T y>
// holds by construction for every N, kind, order and edge flag.
//
// Fast path (HBM-bound, target >= 60 % of the HBM roofline): one thread owns a
// 16-byte column vector and a short chunk of rows held in a register window;
// every y element is written once with a streaming store.  Measured tuning
// (profiles/r01_stencil_tuning.md): SHORT chunks win -- 4 rows per CTA keeps the
// concurrently running CTAs on a narrow band of whole rows (DRAM-page and L2
// friendly; the 2R overlap rows are L2 hits, DRAM traffic stays ~algorithmic),
// 64-row chunks (first design) reached 0.85 of the copy peak, 4-row chunks 1.04.
// Algorithmic bytes: 2*sizeof(T) per element.
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int R = 2;           // max stencil radius
constexpr int NT = 2 * R + 1;  // taps per row
constexpr int SPECIAL = 4;     // rows at each global edge with their own taps

struct StencilParams {
  double interior[NT];
  double top[SPECIAL][NT];     // taps of global rows 0..3
  double bot[SPECIAL][NT];     // taps of global rows N-1, N-2, N-3, N-4
  double scale;                // 1/sampling
  long long nloc, ncols, row0, nglob;
  int n_lo, n_hi;
  long long batch_stride;      // elements between consecutive [nloc x ncols] problems (local axis-k derivatives)
  int nbatch;
};

// ---- halo rows over NVLink PEER MEMORY inside the stencil kernel (round 2) -----------------------------------
// Every rank owns a mailbox ("box") in IPC-mapped memory: flags + 2 parities x 2 sides of halo rows.  The first
// column-tile CTAs of the kernel push this rank's boundary rows straight into the neighbours' boxes (16-byte P2P
// stores) and publish a system-scope flag; only the CTAs that own the first / last row chunk wait for the
// neighbour's flag, and they are scheduled LAST, so the exchange hides behind the interior rows: ONE launch per
// apply, no NCCL call, no side stream (the reference does 2-4 add_ghost_cells exchanges per apply,
// FirstDerivative.py:221-247, 276-319).  Sequence number and tickets live in device memory (graph-capturable).
constexpr size_t HALO_HDR = 256;
struct HaloBox {
  unsigned long long flag[2][2];   // [parity][side]: side 0 = rows from rank-1, side 1 = rows from rank+1
};
struct HaloPeer {
  char* mine;        // this rank's box (local mapping)
  char* prev;        // rank-1's box (peer mapping) or nullptr
  char* next;        // rank+1's box or nullptr
  size_t cap;        // bytes per (parity, side) slot
  unsigned long long* seq;   // device: number of completed exchanges
  unsigned int* tickets;     // device: [0] push ticket, [1] edge ticket, [2] push-done marker
  int send_lo, send_hi;      // rows this rank sends to rank-1 / rank+1
};
__device__ __forceinline__ char* halo_slot(char* box, size_t cap, int par, int side) {
  return box + HALO_HDR + ((size_t)par * 2 + side) * cap;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// forward taps of global row i (offsets -2..2), before the 1/sampling scale
// second-derivative taps (MPISecondDerivative, basicoperators/SecondDerivative.py:125-257)
void fwd_taps2(long long i, long long N, int kind, int edge, double t[NT]) {
  if (kind == B2_FD_FORWARD) {            // y[i] = x[i] - 2 x[i+1] + x[i+2], i <= N-3      (:128-133)
    if (i <= N - 3) { t[R] = 1.0; t[R + 1] = -2.0; t[R + 2] = 1.0; }
  } else if (kind == B2_FD_BACKWARD) {    // y[i] = x[i-2] - 2 x[i-1] + x[i], i >= 2       (:160-165)
    if (i >= 2) { t[R - 2] = 1.0; t[R - 1] = -2.0; t[R] = 1.0; }
  } else {                                // centered                                      (:193-208)
    if (i >= 1 && i <= N - 2) { t[R - 1] = 1.0; t[R] = -2.0; t[R + 1] = 1.0; }
    else if (edge && N >= 3) {
      if (i == 0) { t[R] = 1.0; t[R + 1] = -2.0; t[R + 2] = 1.0; }
      if (i == N - 1) { t[R - 2] = 1.0; t[R - 1] = -2.0; t[R] = 1.0; }
    }
  }
}

void fwd_taps(long long i, long long N, int deriv, int kind, int order, int edge, double t[NT]) {
  for (int k = 0; k < NT; ++k) t[k] = 0.0;
  if (i < 0 || i >= N) return;
  if (deriv == 2) { fwd_taps2(i, N, kind, edge, t); return; }
  if (kind == B2_FD_FORWARD) {
    if (i <= N - 2) { t[R] = -1.0; t[R + 1] = 1.0; }
  } else if (kind == B2_FD_BACKWARD) {
    if (i >= 1) { t[R - 1] = -1.0; t[R] = 1.0; }
  } else if (order == 3) {
    if (i >= 1 && i <= N - 2) { t[R - 1] = -0.5; t[R + 1] = 0.5; }
    else if (edge && N >= 2) {
      if (i == 0) { t[R] += -1.0; t[R + 1] += 1.0; }
      if (i == N - 1) { t[R - 1] += -1.0; t[R] += 1.0; }
    }
  } else {  // centered, order 5
    if (i >= 2 && i <= N - 3) {
      t[R - 2] = 1.0 / 12.0; t[R - 1] = -2.0 / 3.0; t[R + 1] = 2.0 / 3.0; t[R + 2] = -1.0 / 12.0;
    } else if (edge) {
      // FirstDerivative.py:263-272: rank-0 writes y[0], y[1]; last rank writes y[-1], y[-2]
      // (later assignments overwrite earlier ones when N is tiny)
      double a[NT] = {0, 0, 0, 0, 0};
      bool set = false;
      if (i == 0 && N >= 2) { a[R] = -1.0; a[R + 1] = 1.0; set = true; }
      if (i == 1 && N >= 3) { for (int k = 0; k < NT; ++k) a[k] = 0; a[R - 1] = -0.5; a[R + 1] = 0.5; set = true; }
      if (i == N - 1 && N >= 2) { for (int k = 0; k < NT; ++k) a[k] = 0; a[R - 1] = -1.0; a[R] = 1.0; set = true; }
      if (i == N - 2 && N >= 3) { for (int k = 0; k < NT; ++k) a[k] = 0; a[R - 1] = -0.5; a[R + 1] = 0.5; set = true; }
      if (set) for (int k = 0; k < NT; ++k) t[k] = a[k];
    }
  }
}

void row_taps(long long i, long long N, int deriv, int kind, int order, int edge, int adjoint, double t[NT]) {
  if (!adjoint) { fwd_taps(i, N, deriv, kind, order, edge, t); return; }
  for (int k = -R; k <= R; ++k) {
    double f[NT];
    fwd_taps(i + k, N, deriv, kind, order, edge, f);   // zero outside [0, N)
    t[k + R] = f[-k + R];
  }
}

template <typename T>
__device__ __forceinline__ const T* row_ptr(const StencilParams& p, const T* x, const T* lo,
                                            const T* hi, long long r) {
  // r is a LOCAL row index in [-n_lo, nloc + n_hi); anything else -> nullptr
  if (r >= 0 && r < p.nloc) return x + r * p.ncols;
  if (r < 0) return (r >= -(long long)p.n_lo) ? lo + (p.n_lo + r) * p.ncols : nullptr;
  long long h = r - p.nloc;
  return (h < p.n_hi) ? hi + h * p.ncols : nullptr;
}

// row r of the extended block as a 16-byte vector; halo rows (written by a PEER GPU during this kernel in the
// peer-memory mode) go through the coherent load path, local rows through the read-only one
template <typename T>
__device__ __forceinline__ bool load_row(const StencilParams& p, const T* x, const T* lo, const T* hi, long long r,
                                         size_t coff, Vec16<T>& out) {
  if (r >= 0 && r < p.nloc) { out = load_vec(x + r * p.ncols + coff); return true; }
  const T* rp = row_ptr(p, x, lo, hi, r);
  if (!rp) return false;
  out = load_vec_coherent(rp + coff);
  return true;
}

__device__ __forceinline__ const double* special_taps(const StencilParams& p, long long gi) {
  if (gi < SPECIAL) return p.top[gi];
  if (gi >= p.nglob - SPECIAL) return p.bot[p.nglob - 1 - gi];
  return nullptr;
}

// -------------------------------------------------------------------------
// fast path: 16-byte column vectors, rolling window down a chunk of rows.
// MASK bit (k+R) set <=> interior tap k is non-zero (compile-time skip).
// -------------------------------------------------------------------------
// tuning knobs (template parameters): ST_COLS threads along columns, ST_ROWS rows per chunk
// (per thread), ST_U rows loaded per step.  Variant 0 is the default; B2_STENCIL_VARIANT selects
// another one at run time (used by profiles/tune_stencil.py).
template <typename T, int MASK, int ST_ROWS, int ST_U, int ST_COLS, bool PEER = false>
__global__ void __launch_bounds__(ST_COLS)
stencil_vec_kernel(const T* __restrict__ x, T* __restrict__ y, const T* __restrict__ lo,
                   const T* __restrict__ hi, const __grid_constant__ StencilParams p,
                   const HaloPeer hp = HaloPeer{}) {
  constexpr int V = Vec16<T>::N;
  const long long ncv = p.ncols / V;
  // 1-D grid, column tile fastest: concurrently running CTAs cover whole rows
  const long long n_ct = (ncv + ST_COLS - 1) / ST_COLS;
  const long long ct = (long long)blockIdx.x % n_ct;
  long long rc = (long long)blockIdx.x / n_ct;
  x += (size_t)blockIdx.y * (size_t)p.batch_stride;     // batched local problems (blockIdx.y = 0 otherwise)
  y += (size_t)blockIdx.y * (size_t)p.batch_stride;
  const long long cv = ct * ST_COLS + threadIdx.x;
  [[maybe_unused]] bool edge_cta = false;
  [[maybe_unused]] unsigned long long seq = 0;
  [[maybe_unused]] long long n_edge = 0;
  if constexpr (PEER) {
    // chunk order: interior chunks first, the chunks next to the neighbours (0, n_rc-2, n_rc-1) LAST
    const long long n_rc = (p.nloc + ST_ROWS - 1) / ST_ROWS, j = rc;
    n_edge = n_rc < 3 ? n_rc : 3;
    if (j < n_rc - n_edge) rc = j + 1;
    else {
      const long long e = j - (n_rc - n_edge);
      rc = (e == 0) ? 0 : n_rc - n_edge + e;
      edge_cta = true;
    }
    seq = *reinterpret_cast<volatile unsigned long long*>(hp.seq) + 1ull;
    const int par = (int)(seq & 1ull);
    if (j == 0) {
      // push my boundary rows into the neighbours' boxes (this CTA's column tile)
      if (cv < ncv) {
        const size_t cb = (size_t)cv * 16;
        if (hp.prev)
          for (int rr = 0; rr < hp.send_lo; ++rr)
            stg_stream16(halo_slot(hp.prev, hp.cap, par, 1) + (size_t)rr * p.ncols * sizeof(T) + cb,
                         ldg_stream16(x + (size_t)rr * p.ncols + (size_t)cv * V));
        if (hp.next)
          for (int rr = 0; rr < hp.send_hi; ++rr)
            stg_stream16(halo_slot(hp.next, hp.cap, par, 0) + (size_t)rr * p.ncols * sizeof(T) + cb,
                         ldg_stream16(x + (size_t)(p.nloc - hp.send_hi + rr) * p.ncols + (size_t)cv * V));
      }
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&hp.tickets[0], 1u);
        if (t == (unsigned int)n_ct - 1u) {      // every column tile of this rank is on its way: publish
          __threadfence_system();
          if (hp.prev) st_release_sys_u64(&reinterpret_cast<HaloBox*>(hp.prev)->flag[par][1], seq);
          if (hp.next) st_release_sys_u64(&reinterpret_cast<HaloBox*>(hp.next)->flag[par][0], seq);
          *reinterpret_cast<volatile unsigned int*>(&hp.tickets[2]) = 1u;
        }
      }
    }
    lo = reinterpret_cast<const T*>(halo_slot(hp.mine, hp.cap, par, 0));
    hi = reinterpret_cast<const T*>(halo_slot(hp.mine, hp.cap, par, 1));
  }
  const long long r0 = rc * ST_ROWS;
  const long long r1 = (r0 + ST_ROWS < p.nloc) ? r0 + ST_ROWS : p.nloc;
  if constexpr (PEER) {
    const int par = (int)(seq & 1ull);
    const bool needs_lo = hp.prev && p.n_lo > 0 && r0 - R < 0;
    const bool needs_hi = hp.next && p.n_hi > 0 && r1 - 1 + R >= p.nloc;
    if (needs_lo || needs_hi) {
      if (threadIdx.x == 0) {
        const HaloBox* me = reinterpret_cast<const HaloBox*>(hp.mine);
        if (needs_lo) while (ld_acquire_sys_u64(&me->flag[par][0]) < seq) { }
        if (needs_hi) while (ld_acquire_sys_u64(&me->flag[par][1]) < seq) { }
      }
      __syncthreads();
    }
  }
  if (PEER ? false : (cv >= ncv)) return;
  const long long g0 = p.row0 + r0, g1 = p.row0 + r1;  // global rows [g0, g1)
  const bool has_special = (g0 < SPECIAL) || (g1 > p.nglob - SPECIAL);
  const size_t coff = (size_t)cv * V;
  const T scale = (T)p.scale;
  const bool do_scale = (p.scale != 1.0);

  if (PEER && cv >= ncv) {
    // inactive column lanes of a peer-mode CTA still take part in the barriers below
  } else if (!has_special) {
    T c[NT];
#pragma unroll
    for (int k = 0; k < NT; ++k) c[k] = (T)p.interior[k];
    // window w[j] holds local row (r - R + j) for the output row r being produced
    Vec16<T> w[NT + ST_U - 1];
#pragma unroll
    for (int j = 0; j < 2 * R; ++j) {
      if (!load_row(p, x, lo, hi, r0 - R + j, coff, w[j])) {
#pragma unroll
        for (int e = 0; e < V; ++e) w[j].v[e] = (T)0;
      }
    }
    for (long long r = r0; r < r1; r += ST_U) {
#pragma unroll
      for (int u = 0; u < ST_U; ++u) {
        if (!(r + u < r1 + R && load_row(p, x, lo, hi, r + u + R, coff, w[2 * R + u]))) {
#pragma unroll
          for (int e = 0; e < V; ++e) w[2 * R + u].v[e] = (T)0;
        }
      }
#pragma unroll
      for (int u = 0; u < ST_U; ++u) {
        if (r + u < r1) {
          Vec16<T> o;
#pragma unroll
          for (int e = 0; e < V; ++e) {
            T acc = (T)0;
#pragma unroll
            for (int k = 0; k < NT; ++k)
              if (MASK & (1 << k)) acc = fma(c[k], w[u + k].v[e], acc);
            o.v[e] = do_scale ? acc * scale : acc;
          }
          store_vec(y + (size_t)(r + u) * p.ncols + coff, o);
        }
      }
#pragma unroll
      for (int j = 0; j < 2 * R; ++j) w[j] = w[j + ST_U];
    }
  } else {
    // chunk touches a global edge: per-row taps, zero taps are skipped (never read)
    for (long long r = r0; r < r1; ++r) {
      const long long gi = p.row0 + r;
      const double* sp = special_taps(p, gi);
      Vec16<T> o;
#pragma unroll
      for (int e = 0; e < V; ++e) o.v[e] = (T)0;
#pragma unroll
      for (int k = 0; k < NT; ++k) {
        const double tk = sp ? sp[k] : p.interior[k];
        if (tk != 0.0) {
          Vec16<T> v;
          if (load_row(p, x, lo, hi, r + k - R, coff, v)) {
#pragma unroll
            for (int e = 0; e < V; ++e) o.v[e] = fma((T)tk, v.v[e], o.v[e]);
          }
        }
      }
      if (do_scale) {
#pragma unroll
        for (int e = 0; e < V; ++e) o.v[e] *= scale;
      }
      store_vec(y + (size_t)r * p.ncols + coff, o);
    }
  }
  if constexpr (PEER) {
    if (edge_cta) {
      // the last of the edge CTAs closes the exchange: it has seen BOTH neighbours' flags (so no rank can run
      // more than one exchange ahead of a neighbour: the two parities never collide) and this rank's push
      __syncthreads();
      if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&hp.tickets[1], 1u);
        if (t == (unsigned int)(n_edge * n_ct) - 1u) {
          const int par = (int)(seq & 1ull);
          const HaloBox* me = reinterpret_cast<const HaloBox*>(hp.mine);
          if (hp.prev) while (ld_acquire_sys_u64(&me->flag[par][0]) < seq) { }
          if (hp.next) while (ld_acquire_sys_u64(&me->flag[par][1]) < seq) { }
          while (*reinterpret_cast<volatile unsigned int*>(&hp.tickets[2]) == 0u) { }
          hp.tickets[0] = 0u;
          hp.tickets[1] = 0u;
          hp.tickets[2] = 0u;
          __threadfence();
          *reinterpret_cast<volatile unsigned long long*>(hp.seq) = seq;
        }
      }
    }
  }
}

// -------------------------------------------------------------------------
// generic path: one thread per output element (any ncols / alignment)
// -------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
stencil_generic_kernel(const T* x, T* y, const T* __restrict__ lo,
                       const T* __restrict__ hi, const __grid_constant__ StencilParams p) {
  const size_t per = (size_t)p.nloc * (size_t)p.ncols;
  const size_t total = per * (size_t)p.nbatch;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const T scale = (T)p.scale;
  const T* x0 = x;
  T* y0 = y;
  for (size_t gidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; gidx < total; gidx += stride) {
    const size_t b = gidx / per, idx = gidx - b * per;
    x = x0 + b * (size_t)p.batch_stride;
    y = y0 + b * (size_t)p.batch_stride;
    const long long r = (long long)(idx / (size_t)p.ncols);
    const long long j = (long long)(idx - (size_t)r * (size_t)p.ncols);
    const long long gi = p.row0 + r;
    const double* sp = special_taps(p, gi);
    T acc = (T)0;
#pragma unroll
    for (int k = 0; k < NT; ++k) {
      const double tk = sp ? sp[k] : p.interior[k];
      if (tk != 0.0) {
        const T* rp = row_ptr(p, x, lo, hi, r + k - R);
        if (rp) acc = fma((T)tk, rp[j], acc);
      }
    }
    y[idx] = (p.scale != 1.0) ? acc * scale : acc;
  }
}

int interior_mask(const double t[NT]) {
  int m = 0;
  for (int k = 0; k < NT; ++k)
    if (t[k] != 0.0) m |= (1 << k);
  return m;
}

template <typename T, int ROWS, int U, int COLS>
int launch_vec(const void* x, void* y, const void* lo, const void* hi, const StencilParams& p, int mask,
               cudaStream_t st) {
  constexpr int V = Vec16<T>::N;
  const long long nblk = ((p.ncols / V + COLS - 1) / COLS) * ((p.nloc + ROWS - 1) / ROWS);
  if (nblk > 0x7fffffffLL) return B2_ERR_ARG;
  const dim3 grid((unsigned)nblk, (unsigned)p.nbatch);
  if (p.nbatch > 65535) return B2_ERR_ARG;
#define B2_ST_CASE(M)                                                                              \
  case M:                                                                                          \
    stencil_vec_kernel<T, M, ROWS, U, COLS><<<grid, COLS, 0, st>>>((const T*)x, (T*)y, (const T*)lo, \
                                                                   (const T*)hi, p);               \
    break;
  switch (mask) {
    B2_ST_CASE(0x0c)  // taps {0,+1}
    B2_ST_CASE(0x06)  // taps {-1,0}
    B2_ST_CASE(0x0a)  // taps {-1,+1}
    B2_ST_CASE(0x1b)  // taps {-2,-1,+1,+2}
    default:
      stencil_vec_kernel<T, 0x1f, ROWS, U, COLS><<<grid, COLS, 0, st>>>((const T*)x, (T*)y, (const T*)lo,
                                                                        (const T*)hi, p);
  }
#undef B2_ST_CASE
  B2_LAUNCH_CHECK();
  return B2_OK;
}

// peer-memory halo mode: the tuned (4, 4, 128) variant with the exchange fused in
template <typename T>
int launch_vec_peer(const void* x, void* y, const StencilParams& p, const HaloPeer& hp, int mask, cudaStream_t st) {
  constexpr int V = Vec16<T>::N;
  constexpr int ROWS = 4, U = 4, COLS = 128;
  const long long nblk = ((p.ncols / V + COLS - 1) / COLS) * ((p.nloc + ROWS - 1) / ROWS);
  if (nblk > 0x7fffffffLL) return B2_ERR_ARG;
  const dim3 grid((unsigned)nblk, 1u);
#define B2_ST_CASE(M)                                                                                       \
  case M:                                                                                                   \
    stencil_vec_kernel<T, M, ROWS, U, COLS, true><<<grid, COLS, 0, st>>>((const T*)x, (T*)y, nullptr, nullptr, p, hp); \
    break;
  switch (mask) {
    B2_ST_CASE(0x0c)
    B2_ST_CASE(0x06)
    B2_ST_CASE(0x0a)
    B2_ST_CASE(0x1b)
    default:
      stencil_vec_kernel<T, 0x1f, ROWS, U, COLS, true><<<grid, COLS, 0, st>>>((const T*)x, (T*)y, nullptr, nullptr, p, hp);
  }
#undef B2_ST_CASE
  B2_LAUNCH_CHECK();
  return B2_OK;
}

template <typename T>
int launch_stencil(b2_ctx* ctx, const void* x, void* y, const void* lo, const void* hi,
                   const StencilParams& p, cudaStream_t st) {
  constexpr int V = Vec16<T>::N;
  const bool vec_ok = (p.ncols % V == 0) && b2_aligned16(x) && b2_aligned16(y) &&
                      (!lo || b2_aligned16(lo)) && (!hi || b2_aligned16(hi)) &&
                      (p.ncols / V >= 8);
  const int mask = interior_mask(p.interior);
  if (vec_ok) {
    static int variant = -1;
    if (variant < 0) {
      const char* e = getenv("B2_STENCIL_VARIANT");
      variant = e ? atoi(e) : 0;
    }
    switch (variant) {
      case 1: return launch_vec<T, 8, 4, 128>(x, y, lo, hi, p, mask, st);
      case 2: return launch_vec<T, 8, 2, 128>(x, y, lo, hi, p, mask, st);
      case 3: return launch_vec<T, 4, 4, 128>(x, y, lo, hi, p, mask, st);
      case 4: return launch_vec<T, 4, 2, 128>(x, y, lo, hi, p, mask, st);
      case 5: return launch_vec<T, 8, 8, 128>(x, y, lo, hi, p, mask, st);
      case 6: return launch_vec<T, 8, 4, 256>(x, y, lo, hi, p, mask, st);
      case 7: return launch_vec<T, 8, 4, 64>(x, y, lo, hi, p, mask, st);
      case 8: return launch_vec<T, 4, 4, 256>(x, y, lo, hi, p, mask, st);
      case 9: return launch_vec<T, 8, 1, 128>(x, y, lo, hi, p, mask, st);
      case 10: return launch_vec<T, 2, 2, 128>(x, y, lo, hi, p, mask, st);
      case 11: return launch_vec<T, 8, 2, 256>(x, y, lo, hi, p, mask, st);
      case 12: return launch_vec<T, 64, 4, 128>(x, y, lo, hi, p, mask, st);   // first design (r01 baseline)
      default: return launch_vec<T, 4, 4, 128>(x, y, lo, hi, p, mask, st);    // tuned: see profiles/r01_stencil_tuning.md
    }
  } else {
    size_t total = (size_t)p.nloc * (size_t)p.ncols * (size_t)p.nbatch;
    size_t need = (total + 255) / 256;
    size_t cap = (size_t)ctx->sm_count * 8;
    int grid = (int)(need < cap ? need : cap);
    stencil_generic_kernel<T><<<grid, 256, 0, st>>>((const T*)x, (T*)y, (const T*)lo, (const T*)hi, p);
  }
  B2_LAUNCH_CHECK();
  return B2_OK;
}

}  // namespace

extern "C" int b2_first_derivative_halo(int kind, int order, int adjoint, int* need_lo,
                                        int* need_hi) {
  int lo, hi;
  if (kind == B2_FD_FORWARD) { lo = adjoint ? 1 : 0; hi = adjoint ? 0 : 1; }
  else if (kind == B2_FD_BACKWARD) { lo = adjoint ? 0 : 1; hi = adjoint ? 1 : 0; }
  else if (kind == B2_FD_CENTERED && order == 3) { lo = hi = 1; }
  else if (kind == B2_FD_CENTERED && order == 5) { lo = hi = 2; }
  else return B2_ERR_UNSUPPORTED;
  if (need_lo) *need_lo = lo;
  if (need_hi) *need_hi = hi;
  return B2_OK;
}

int b2_fd_build_params(StencilParams* p, int n_lo, int n_hi, size_t nrows_local, size_t ncols,
                       size_t row0, size_t nrows_global, int kind, int order, int edge,
                       double sampling, int adjoint, int deriv = 1) {
  if (kind != B2_FD_FORWARD && kind != B2_FD_BACKWARD && kind != B2_FD_CENTERED)
    return B2_ERR_UNSUPPORTED;
  if (deriv == 1 && kind == B2_FD_CENTERED && order != 3 && order != 5) return B2_ERR_UNSUPPORTED;
  if (row0 + nrows_local > nrows_global) return B2_ERR_ARG;
  const long long N = (long long)nrows_global;
  // interior taps = taps of a row far from both edges of a very long axis
  row_taps(1000, 2000, deriv, kind, order, edge, adjoint, p->interior);
  for (int s = 0; s < SPECIAL; ++s) {
    row_taps(s, N, deriv, kind, order, edge, adjoint, p->top[s]);
    row_taps(N - 1 - s, N, deriv, kind, order, edge, adjoint, p->bot[s]);
  }
  p->scale = (deriv == 2) ? 1.0 / (sampling * sampling) : 1.0 / sampling;
  p->batch_stride = 0;
  p->nbatch = 1;
  p->nloc = (long long)nrows_local;
  p->ncols = (long long)ncols;
  p->row0 = (long long)row0;
  p->nglob = N;
  p->n_lo = n_lo;
  p->n_hi = n_hi;
  return B2_OK;
}

extern "C" int b2_first_derivative(b2_ctx* ctx, const void* x, void* y, const void* halo_lo,
                                   int n_lo, const void* halo_hi, int n_hi, size_t nrows_local,
                                   size_t ncols, size_t row0, size_t nrows_global, int kind,
                                   int order, int edge, double sampling, int adjoint, int dtype,
                                   void* stream) {
  if (!ctx) return B2_ERR_ARG;
  if (nrows_local == 0 || ncols == 0) return B2_OK;
  if (!x || !y) return B2_ERR_ARG;
  if (n_lo < 0 || n_hi < 0 || n_lo > 8 || n_hi > 8) return B2_ERR_ARG;
  if (!halo_lo) n_lo = 0;
  if (!halo_hi) n_hi = 0;
  int need_lo, need_hi;
  int rc = b2_first_derivative_halo(kind, order, adjoint, &need_lo, &need_hi);
  if (rc) return rc;
  // a missing halo is only legal where the stencil would read outside the global array
  const long long avail_lo = (long long)row0, avail_hi = (long long)(nrows_global - row0 - nrows_local);
  if (n_lo < (need_lo < avail_lo ? need_lo : avail_lo)) return B2_ERR_HALO;
  if (n_hi < (need_hi < avail_hi ? need_hi : avail_hi)) return B2_ERR_HALO;
  StencilParams p;
  rc = b2_fd_build_params(&p, n_lo, n_hi, nrows_local, ncols, row0, nrows_global, kind, order,
                          edge, sampling, adjoint);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case B2_F32: return launch_stencil<float>(ctx, x, y, halo_lo, halo_hi, p, st);
    case B2_F64: return launch_stencil<double>(ctx, x, y, halo_lo, halo_hi, p, st);
    default: return B2_ERR_DTYPE;
  }
}


extern "C" int b2_second_derivative_halo(int kind, int edge, int adjoint, int* need_lo, int* need_hi);

// ---- peer-memory halo handle ------------------------------------------------------------------------------------
struct b2_halo {
  int rank, size;
  char* box[3];               // [0] rank-1's box (peer mapping or NULL), [1] mine, [2] rank+1's
  size_t cap;
  unsigned long long* seq;    // device
  unsigned int* tickets;      // device, 4 uints
};

extern "C" size_t b2_halo_bytes(size_t cap_bytes) { return HALO_HDR + 4 * cap_bytes; }

// boxes_host[r]: rank r's box as mapped in THIS process (own pointer for r == rank; only rank +/- 1 are used).
// Zeroes this rank's flags: callers barrier on the host between creation and the first apply.
extern "C" int b2_halo_create(int rank, int size, void* const* boxes_host, size_t cap_bytes, b2_halo** out) {
  if (!out || !boxes_host || size < 1 || rank < 0 || rank >= size || cap_bytes == 0 || (cap_bytes % 16)) return B2_ERR_ARG;
  b2_halo* h = new b2_halo();
  h->rank = rank;
  h->size = size;
  h->cap = cap_bytes;
  h->box[0] = rank > 0 ? (char*)boxes_host[rank - 1] : nullptr;
  h->box[1] = (char*)boxes_host[rank];
  h->box[2] = rank < size - 1 ? (char*)boxes_host[rank + 1] : nullptr;
  h->seq = nullptr;
  h->tickets = nullptr;
  cudaError_t e = cudaMalloc((void**)&h->seq, sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->tickets, 4 * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(h->seq, 0, sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(h->tickets, 0, 4 * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(h->box[1], 0, HALO_HDR);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    if (h->seq) cudaFree(h->seq);
    if (h->tickets) cudaFree(h->tickets);
    delete h;
    return (int)e;
  }
  *out = h;
  return B2_OK;
}

extern "C" int b2_halo_destroy(b2_halo* h) {
  if (!h) return B2_OK;
  if (h->seq) cudaFree(h->seq);
  if (h->tickets) cudaFree(h->tickets);
  delete h;
  return B2_OK;
}

// One-launch distributed stencil: deriv = 1 (MPIFirstDerivative) or 2 (MPISecondDerivative); the halo rows
// travel through the peer boxes inside the kernel.  Collective over the ranks of the handle (same call
// sequence on every rank, one stream); every rank must own at least max(need_lo, need_hi) rows.
extern "C" int b2_derivative_peer(b2_ctx* ctx, b2_halo* h, const void* x, void* y, size_t nrows_local, size_t ncols,
                                  size_t row0, size_t nrows_global, int deriv, int kind, int order, int edge,
                                  double sampling, int adjoint, int dtype, void* stream) {
  if (!ctx || !h || !x || !y || (deriv != 1 && deriv != 2)) return B2_ERR_ARG;
  if (dtype != B2_F32 && dtype != B2_F64) return B2_ERR_DTYPE;
  int need_lo, need_hi;
  int rc = deriv == 1 ? b2_first_derivative_halo(kind, order, adjoint, &need_lo, &need_hi)
                      : b2_second_derivative_halo(kind, edge, adjoint, &need_lo, &need_hi);
  if (rc) return rc;
  const size_t esz = b2_dtype_size(dtype), V = 16 / esz;
  if ((long long)nrows_local < (need_lo > need_hi ? need_lo : need_hi)) return B2_ERR_HALO;
  if (ncols % V || ncols / V < 8 || !b2_aligned16(x) || !b2_aligned16(y)) return B2_ERR_ALIGN;
  if ((size_t)(need_lo > need_hi ? need_lo : need_hi) * ncols * esz > h->cap) return B2_ERR_WORKSPACE;
  const int n_lo = h->box[0] ? need_lo : 0, n_hi = h->box[2] ? need_hi : 0;
  StencilParams p;
  rc = b2_fd_build_params(&p, n_lo, n_hi, nrows_local, ncols, row0, nrows_global, kind, deriv == 1 ? order : 3, edge,
                          sampling, adjoint, deriv);
  if (rc) return rc;
  HaloPeer hp;
  hp.mine = h->box[1];
  hp.prev = h->box[0];
  hp.next = h->box[2];
  hp.cap = h->cap;
  hp.seq = h->seq;
  hp.tickets = h->tickets;
  hp.send_lo = h->box[0] ? need_hi : 0;    // rank-1 needs my first need_hi rows as ITS hi halo
  hp.send_hi = h->box[2] ? need_lo : 0;    // rank+1 needs my last need_lo rows as ITS lo halo
  const int mask = interior_mask(p.interior);
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == B2_F32 ? launch_vec_peer<float>(x, y, p, hp, mask, st) : launch_vec_peer<double>(x, y, p, hp, mask, st);
}

// ---- MPISecondDerivative per-rank apply (basicoperators/SecondDerivative.py:125-257) -----------------
extern "C" int b2_second_derivative_halo(int kind, int edge, int adjoint, int* need_lo, int* need_hi) {
  int lo, hi;
  if (kind == B2_FD_FORWARD) { lo = adjoint ? 2 : 0; hi = adjoint ? 0 : 2; }
  else if (kind == B2_FD_BACKWARD) { lo = adjoint ? 0 : 2; hi = adjoint ? 2 : 0; }
  else if (kind == B2_FD_CENTERED) { lo = hi = edge ? 2 : 1; }   // the edge rows reach two rows away
  else return B2_ERR_UNSUPPORTED;
  if (need_lo) *need_lo = lo;
  if (need_hi) *need_hi = hi;
  return B2_OK;
}

extern "C" int b2_second_derivative(b2_ctx* ctx, const void* x, void* y, const void* halo_lo, int n_lo,
                                    const void* halo_hi, int n_hi, size_t nrows_local, size_t ncols, size_t row0,
                                    size_t nrows_global, int kind, int edge, double sampling, int adjoint,
                                    int dtype, void* stream) {
  if (!ctx) return B2_ERR_ARG;
  if (nrows_local == 0 || ncols == 0) return B2_OK;
  if (!x || !y) return B2_ERR_ARG;
  if (n_lo < 0 || n_hi < 0 || n_lo > 8 || n_hi > 8) return B2_ERR_ARG;
  if (!halo_lo) n_lo = 0;
  if (!halo_hi) n_hi = 0;
  int need_lo, need_hi;
  int rc = b2_second_derivative_halo(kind, edge, adjoint, &need_lo, &need_hi);
  if (rc) return rc;
  const long long avail_lo = (long long)row0, avail_hi = (long long)(nrows_global - row0 - nrows_local);
  // interior rows only need the taps' reach; be strict with the tap reach of this kind
  int reach_lo = need_lo, reach_hi = need_hi;
  if (kind == B2_FD_CENTERED) reach_lo = reach_hi = 1;   // 2 only matters next to a global edge (checked by taps)
  if (n_lo < (reach_lo < avail_lo ? reach_lo : avail_lo)) return B2_ERR_HALO;
  if (n_hi < (reach_hi < avail_hi ? reach_hi : avail_hi)) return B2_ERR_HALO;
  StencilParams p;
  rc = b2_fd_build_params(&p, n_lo, n_hi, nrows_local, ncols, row0, nrows_global, kind, 3, edge, sampling, adjoint, 2);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case B2_F32: return launch_stencil<float>(ctx, x, y, halo_lo, halo_hi, p, st);
    case B2_F64: return launch_stencil<double>(ctx, x, y, halo_lo, halo_hi, p, st);
    default: return B2_ERR_DTYPE;
  }
}

// ---- rank-local derivative along the MIDDLE axis of a C-ordered [n_outer][n_axis][n_inner] block ------
// (the non-partitioned directions of MPILaplacian / MPIGradient: Laplacian.py:97-126, Gradient.py:101-119
//  wrap a serial pylops First/SecondDerivative per rank; here the same stencil kernel runs batched)
extern "C" int b2_derivative_axis(b2_ctx* ctx, const void* x, void* y, size_t n_outer, size_t n_axis, size_t n_inner,
                                  int deriv, int kind, int order, int edge, double sampling, int adjoint, int dtype,
                                  void* stream) {
  if (!ctx || (deriv != 1 && deriv != 2)) return B2_ERR_ARG;
  if (n_outer == 0 || n_axis == 0 || n_inner == 0) return B2_OK;
  if (!x || !y) return B2_ERR_ARG;
  StencilParams p;
  int rc = b2_fd_build_params(&p, 0, 0, n_axis, n_inner, 0, n_axis, kind, order, edge, sampling, adjoint, deriv);
  if (rc) return rc;
  p.nbatch = 1;
  p.batch_stride = (long long)(n_axis * n_inner);
  cudaStream_t st = (cudaStream_t)stream;
  for (size_t done = 0; done < n_outer; done += 65535) {
    p.nbatch = (int)(n_outer - done < 65535 ? n_outer - done : 65535);
    const size_t off = done * n_axis * n_inner * b2_dtype_size(dtype);
    switch (dtype) {
      case B2_F32: rc = launch_stencil<float>(ctx, (const char*)x + off, (char*)y + off, nullptr, nullptr, p, st); break;
      case B2_F64: rc = launch_stencil<double>(ctx, (const char*)x + off, (char*)y + off, nullptr, nullptr, p, st); break;
      default: return B2_ERR_DTYPE;
    }
    if (rc) return rc;
  }
  return B2_OK;
}
