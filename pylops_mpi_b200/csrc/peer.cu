// One-shot all-reduce of a few float64 scalars over NVLink PEER MEMORY (no NCCL call):
// every rank stores its partial values into a slot of every peer's (IPC-mapped) mailbox, publishes a
// sequence flag with system-scope release, spins on its own mailbox until all P flags of this
// sequence number arrived, and folds the P contributions in rank order (bit-identical on every rank).
// This is the collective half of DistributedArray.dot / norm (DistributedArray.py:684-686, 714-757)
// and of the CGLS step scalars (cls_basic.py:389-401): latency ~ one NVLink round trip instead of an
// NCCL launch + ring/tree protocol.  Double-buffered by sequence parity (a rank can be at most one
// call ahead of its slowest peer, because call n+1 cannot complete before every peer has entered it).
#include <string.h>
#include "common.cuh"

namespace {
constexpr int PEER_MAX = 8, VAL_MAX = 8;
struct Slots {
  double data[2][PEER_MAX][VAL_MAX];
  unsigned long long flag[2][PEER_MAX];
};
struct PeerPtrs {
  Slots* p[PEER_MAX];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_volatile(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__global__ void peer_allreduce_kernel(PeerPtrs pp, int rank, int P, double* __restrict__ vals, int k,
                                      unsigned long long seq, int op) {
  const int par = (int)(seq & 1ull);
  const int t = threadIdx.x;
  if (t < P) {
    Slots* dst = pp.p[t];
    for (int j = 0; j < k; ++j) dst->data[par][rank][j] = vals[j];
    __threadfence_system();
    st_release_sys(&dst->flag[par][rank], seq);
  }
  __syncthreads();
  Slots* me = pp.p[rank];
  if (t < P) {
    while (ld_acquire_sys(&me->flag[par][t]) < seq) { }
  }
  __syncthreads();
  if (t < k) {
    double acc = ld_volatile(&me->data[par][0][t]);
    for (int r = 1; r < P; ++r) {
      const double v = ld_volatile(&me->data[par][r][t]);
      acc = (op == B2_SUM) ? acc + v : (op == B2_MAX ? fmax(acc, v) : fmin(acc, v));
    }
    vals[t] = acc;
  }
}
}  // namespace

struct b2_peer {
  int rank, size;
  PeerPtrs pp;
  unsigned long long seq;
};

extern "C" size_t b2_peer_slots_bytes(void) { return sizeof(Slots); }

// local_slots: this rank's mailbox (b2_symm_alloc'ed, b2_peer_slots_bytes() bytes, ZEROED by this call);
// slots_host[r]: pointer to rank r's mailbox as mapped in THIS process (own pointer for r == rank).
// Collective: all ranks must have zeroed their mailbox before anybody's first b2_peer_allreduce
// (callers barrier on the host after b2_peer_create).
extern "C" int b2_peer_create(int rank, int size, void* const* slots_host, b2_peer** out) {
  if (!out || !slots_host || size < 1 || size > PEER_MAX || rank < 0 || rank >= size) return B2_ERR_ARG;
  b2_peer* h = new b2_peer();
  h->rank = rank;
  h->size = size;
  h->seq = 0;
  for (int r = 0; r < PEER_MAX; ++r) h->pp.p[r] = r < size ? (Slots*)slots_host[r] : nullptr;
  cudaError_t e = cudaMemset(slots_host[rank], 0, sizeof(Slots));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { delete h; return (int)e; }
  *out = h;
  return B2_OK;
}

extern "C" int b2_peer_destroy(b2_peer* h) {
  delete h;
  return B2_OK;
}

// in-place all-reduce of k <= 8 float64 values resident on the device
extern "C" int b2_peer_allreduce(b2_peer* h, double* vals_dev, int k, int op, void* stream) {
  if (!h || !vals_dev || k < 1 || k > VAL_MAX) return B2_ERR_ARG;
  if (op != B2_SUM && op != B2_MAX && op != B2_MIN) return B2_ERR_ARG;
  h->seq += 1;
  peer_allreduce_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(h->pp, h->rank, h->size, vals_dev, k, h->seq, op);
  B2_LAUNCH_CHECK();
  return B2_OK;
}
