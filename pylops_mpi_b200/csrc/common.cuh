// Shared helpers for libb200lops (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/b200lops.h"

#define B2_CUDA(call)                                  \
  do {                                                 \
    cudaError_t e__ = (call);                          \
    if (e__ != cudaSuccess) return (int)e__;           \
  } while (0)

#define B2_LAUNCH_CHECK()                              \
  do {                                                 \
    cudaError_t e__ = cudaGetLastError();              \
    if (e__ != cudaSuccess) return (int)e__;           \
  } while (0)

struct b2_ctx {
  int device;
  int sm_count;
  // reduction workspace: partial sums + ticket counters (device memory)
  double* red_partials;     // B2_RED_MAX_BLOCKS * B2_RED_MAX_OUT doubles
  unsigned int* tickets;    // B2_TICKETS uints, zero between launches
  float* gemv_partials;     // scratch for transposed gemv (bytes = gemv_partials_bytes)
  size_t gemv_partials_bytes;
  // host-buffer pipeline (b2_first_derivative_host)
  void* pipe_buf[3][2];     // [slot][in/out]
  size_t pipe_bytes;
  cudaStream_t pipe_stream[3];
  cudaEvent_t pipe_ev[3][3];
};

constexpr int B2_RED_MAX_BLOCKS = 2048;
constexpr int B2_RED_MAX_OUT = 16;     // doubles per block (k<=8 complex dots)
constexpr int B2_TICKETS = 4096;

static inline size_t b2_dtype_size(int dt) {
  switch (dt) {
    case B2_F32: return 4;
    case B2_F64: return 8;
    case B2_C64: return 8;
    case B2_C128: return 16;
    case B2_BF16: return 2;
    case B2_I64: return 8;
    default: return 0;
  }
}

static inline bool b2_aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

// ---- streaming 16-byte global loads / stores (read-once data: bypass L1) ----
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// coherent variant: safe when the kernel writes the same buffer (in-place ops)
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void stg_stream16(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

template <typename T>
struct Vec16;  // 16-byte vector of T
template <>
struct Vec16<float> {
  static constexpr int N = 4;
  float v[4];
};
template <>
struct Vec16<double> {
  static constexpr int N = 2;
  double v[2];
};

template <typename T>
__device__ __forceinline__ Vec16<T> load_vec(const T* p) {
  uint4 r = ldg_stream16(p);
  Vec16<T> o;
  *reinterpret_cast<uint4*>(&o) = r;
  return o;
}
template <typename T>
__device__ __forceinline__ Vec16<T> load_vec_coherent(const T* p) {
  uint4 r = ld_stream16(p);
  Vec16<T> o;
  *reinterpret_cast<uint4*>(&o) = r;
  return o;
}
template <typename T>
__device__ __forceinline__ void store_vec(T* p, const Vec16<T>& v) {
  stg_stream16(p, *reinterpret_cast<const uint4*>(&v));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
