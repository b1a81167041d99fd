// NCCL-over-NVLink collectives replacing the reference's mpi4py / cupy.cuda.nccl
// layer (pylops_mpi/utils/_nccl.py:98-403, utils/_mpi.py:21-344, dispatched by
// Distributed.py:35-349).  One ncclComm_t per (sub)group, created once; every
// call is enqueued on the caller's stream.  Complex payloads travel as 2x real
// (as _nccl.py:32-63 does).
#include <nccl.h>
#include <string.h>
#include <vector>
#include "common.cuh"

struct b2_comm {
  ncclComm_t comm;
  int rank, size, device;
};

#define B2_NCCL(call)                                    \
  do {                                                   \
    ncclResult_t r__ = (call);                           \
    if (r__ != ncclSuccess) return 1000 + (int)r__;      \
  } while (0)

namespace {
bool map_dtype(int dtype, ncclDataType_t* dt, size_t* mult) {
  *mult = 1;
  switch (dtype) {
    case B2_F32: *dt = ncclFloat32; return true;
    case B2_F64: *dt = ncclFloat64; return true;
    case B2_C64: *dt = ncclFloat32; *mult = 2; return true;
    case B2_C128: *dt = ncclFloat64; *mult = 2; return true;
    case B2_BF16: *dt = ncclBfloat16; return true;
    case B2_I64: *dt = ncclInt64; return true;
    default: return false;
  }
}
bool map_op(int op, ncclRedOp_t* o) {
  switch (op) {
    case B2_SUM: *o = ncclSum; return true;
    case B2_MAX: *o = ncclMax; return true;
    case B2_MIN: *o = ncclMin; return true;
    default: return false;
  }
}
}  // namespace

extern "C" int b2_get_unique_id(void* id128_host) {
  if (!id128_host) return B2_ERR_ARG;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  B2_NCCL(ncclGetUniqueId(&id));
  memcpy(id128_host, &id, sizeof id);
  return B2_OK;
}

extern "C" int b2_comm_create(int rank, int size, const void* id128_host, int device,
                              b2_comm** out) {
  if (!out || !id128_host || rank < 0 || rank >= size) return B2_ERR_ARG;
  B2_CUDA(cudaSetDevice(device));
  ncclUniqueId id;
  memcpy(&id, id128_host, sizeof id);
  ncclComm_t c;
  B2_NCCL(ncclCommInitRank(&c, size, id, rank));
  b2_comm* h = new b2_comm{c, rank, size, device};
  *out = h;
  return B2_OK;
}

extern "C" int b2_comm_split(b2_comm* comm, int color, int key, b2_comm** out) {
  if (!comm || !out) return B2_ERR_ARG;
  ncclComm_t c = nullptr;
  B2_NCCL(ncclCommSplit(comm->comm, color < 0 ? NCCL_SPLIT_NOCOLOR : color, key, &c, nullptr));
  if (!c) { *out = nullptr; return B2_OK; }
  int r = 0, s = 0;
  B2_NCCL(ncclCommUserRank(c, &r));
  B2_NCCL(ncclCommCount(c, &s));
  *out = new b2_comm{c, r, s, comm->device};
  return B2_OK;
}

extern "C" int b2_comm_destroy(b2_comm* comm) {
  if (!comm) return B2_OK;
  ncclResult_t r = ncclCommDestroy(comm->comm);
  delete comm;
  return r == ncclSuccess ? B2_OK : 1000 + (int)r;
}

extern "C" int b2_comm_rank(const b2_comm* comm, int* rank, int* size) {
  if (!comm) return B2_ERR_ARG;
  if (rank) *rank = comm->rank;
  if (size) *size = comm->size;
  return B2_OK;
}

extern "C" int b2_allreduce(b2_comm* comm, const void* send, void* recv, size_t n, int dtype,
                            int op, void* stream) {
  if (!comm) return B2_ERR_ARG;
  if (n == 0) return B2_OK;
  ncclDataType_t dt; size_t mult; ncclRedOp_t o;
  if (!map_dtype(dtype, &dt, &mult)) return B2_ERR_DTYPE;
  if (!map_op(op, &o)) return B2_ERR_ARG;
  if (mult == 2 && op != B2_SUM) return B2_ERR_DTYPE;
  B2_NCCL(ncclAllReduce(send, recv, n * mult, dt, o, comm->comm, (cudaStream_t)stream));
  return B2_OK;
}

extern "C" int b2_allgather(b2_comm* comm, const void* send, void* recv, size_t n_per_rank,
                            int dtype, void* stream) {
  if (!comm) return B2_ERR_ARG;
  if (n_per_rank == 0) return B2_OK;
  ncclDataType_t dt; size_t mult;
  if (!map_dtype(dtype, &dt, &mult)) return B2_ERR_DTYPE;
  B2_NCCL(ncclAllGather(send, recv, n_per_rank * mult, dt, comm->comm, (cudaStream_t)stream));
  return B2_OK;
}

extern "C" int b2_allgatherv(b2_comm* comm, const void* send, void* recv,
                             const size_t* counts_host, int dtype, void* stream) {
  if (!comm || !counts_host) return B2_ERR_ARG;
  ncclDataType_t dt; size_t mult;
  if (!map_dtype(dtype, &dt, &mult)) return B2_ERR_DTYPE;
  const size_t esz = b2_dtype_size(dtype);
  bool equal = true;
  for (int r = 1; r < comm->size; ++r) equal = equal && (counts_host[r] == counts_host[0]);
  if (equal) return b2_allgather(comm, send, recv, counts_host[0], dtype, stream);
  // one grouped broadcast per contributing rank: no padding, no extra copies
  B2_NCCL(ncclGroupStart());
  size_t off = 0;
  for (int r = 0; r < comm->size; ++r) {
    if (counts_host[r]) {
      void* dst = (char*)recv + off * esz;
      const void* src = (r == comm->rank) ? send : dst;
      ncclResult_t res = ncclBroadcast(src, dst, counts_host[r] * mult, dt, r, comm->comm,
                                       (cudaStream_t)stream);
      if (res != ncclSuccess) { ncclGroupEnd(); return 1000 + (int)res; }
    }
    off += counts_host[r];
  }
  B2_NCCL(ncclGroupEnd());
  return B2_OK;
}

extern "C" int b2_bcast(b2_comm* comm, void* buf, size_t n, int dtype, int root, void* stream) {
  if (!comm) return B2_ERR_ARG;
  if (n == 0) return B2_OK;
  ncclDataType_t dt; size_t mult;
  if (!map_dtype(dtype, &dt, &mult)) return B2_ERR_DTYPE;
  B2_NCCL(ncclBroadcast(buf, buf, n * mult, dt, root, comm->comm, (cudaStream_t)stream));
  return B2_OK;
}

extern "C" int b2_send(b2_comm* comm, const void* buf, size_t n, int dtype, int peer,
                       void* stream) {
  if (!comm) return B2_ERR_ARG;
  if (n == 0) return B2_OK;   // NCCL hangs on 0-byte p2p (DistributedArray.py:910-912)
  ncclDataType_t dt; size_t mult;
  if (!map_dtype(dtype, &dt, &mult)) return B2_ERR_DTYPE;
  B2_NCCL(ncclSend(buf, n * mult, dt, peer, comm->comm, (cudaStream_t)stream));
  return B2_OK;
}

extern "C" int b2_recv(b2_comm* comm, void* buf, size_t n, int dtype, int peer, void* stream) {
  if (!comm) return B2_ERR_ARG;
  if (n == 0) return B2_OK;
  ncclDataType_t dt; size_t mult;
  if (!map_dtype(dtype, &dt, &mult)) return B2_ERR_DTYPE;
  B2_NCCL(ncclRecv(buf, n * mult, dt, peer, comm->comm, (cudaStream_t)stream));
  return B2_OK;
}

extern "C" int b2_group_start(void) { B2_NCCL(ncclGroupStart()); return B2_OK; }
extern "C" int b2_group_end(void) { B2_NCCL(ncclGroupEnd()); return B2_OK; }

// Allgather with explicit placement: rank r's counts[r] elements land at recv + offsets[r]
// (element units).  One grouped set of broadcasts; the local contribution may already sit in
// place (send == recv + offsets[rank]).  Lets a producer kernel write its slice straight into the
// gathered buffer and lets callers gather in chunks while the next chunk is being computed
// (MPIFredholm1: signalprocessing/Fredholm1.py:131-132 gathers only after ALL slices are done).
extern "C" int b2_allgatherv_at(b2_comm* comm, const void* send, void* recv, const size_t* counts_host,
                                const size_t* offsets_host, int dtype, void* stream) {
  if (!comm || !counts_host || !offsets_host) return B2_ERR_ARG;
  ncclDataType_t dt; size_t mult;
  if (!map_dtype(dtype, &dt, &mult)) return B2_ERR_DTYPE;
  const size_t esz = b2_dtype_size(dtype);
  B2_NCCL(ncclGroupStart());
  for (int r = 0; r < comm->size; ++r) {
    if (!counts_host[r]) continue;
    void* dst = (char*)recv + offsets_host[r] * esz;
    const void* src = (r == comm->rank) ? send : dst;
    ncclResult_t res = ncclBroadcast(src, dst, counts_host[r] * mult, dt, r, comm->comm, (cudaStream_t)stream);
    if (res != ncclSuccess) { ncclGroupEnd(); return 1000 + (int)res; }
  }
  B2_NCCL(ncclGroupEnd());
  return B2_OK;
}
