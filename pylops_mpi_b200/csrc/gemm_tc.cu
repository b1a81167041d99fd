// bf16 x bf16 -> fp32 dense tile product on the 5th-generation tensor cores:
//   C[m x n] (+)= op(A) B,   A: m x k (op=N) or k x m (op=T/H), B: k x n, all row-major.
// The multi-column tile product of MPIMatrixMult (reference:
// pylops_mpi/basicoperators/MatrixMult.py:366-370, 409-413, 663-670, 742-763 --
// `ncp.matmul` / `ncp.dot` on the per-rank tiles).
//
// Hand-written sm_100a kernel (inline PTX, no CUTLASS):
//   * persistent CTAs (one per SM), static round-robin tile scheduler with
//     grouped rasterisation for L2 reuse,
//   * warp 0 = TMA producer (cp.async.bulk.tensor, 128B swizzle, zero-filled
//     out-of-bounds => arbitrary m, n, k),
//   * warp 1 = single-thread tcgen05.mma issuer (M=128, N=256, K=16 atoms,
//     accumulators in TMEM, 2 accumulator stages = 512 TMEM columns),
//   * warps 2..5 = epilogue (tcgen05.ld TMEM -> registers -> global, optional
//     C += accumulate),
//   * 4-stage smem ring (48 KB / stage) with mbarrier full/empty pairs,
//     tcgen05.commit releases stages and publishes finished accumulators.
// Row-major operands map onto UMMA "major" modes without any transposition copy:
//   A op=N  -> K-major,   A op=T -> MN-major,   B (k x n row-major) -> MN-major.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr uint32_t BM = 128, BN = 256, BK = 64, STAGES = 4, UMMA_K = 16;
constexpr uint32_t A_STAGE_BYTES = BM * BK * 2;          // 16 KB
constexpr uint32_t B_STAGE_BYTES = BK * BN * 2;          // 32 KB
constexpr uint32_t STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr uint32_t ATOM_BYTES = 64 * BK * 2;             // one 64-wide MN block x BK k-rows: 8 KB
constexpr uint32_t TMEM_COLS = 512;                      // 2 accumulator stages x 256 fp32 columns
constexpr uint32_t NUM_THREADS = 192;
constexpr uint32_t GROUP_M = 8;
constexpr size_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

// ---- PTX wrappers -------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t tx) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(tx)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar,
                                            int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- descriptors -----------------------------------------------------------------------
// shared-memory matrix descriptor (sm_100 "version 1"), SWIZZLE_128B:
//   bits [0,14) start address >> 4, [16,30) leading byte offset >> 4, [32,46) stride byte
//   offset >> 4, [46,48) version = 1, [61,64) layout type (2 = 128B swizzle)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor for kind::f16: c=F32 (bits 4-5 = 1), a=b=BF16 (bits 7-9, 10-12 = 1),
// a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(uint32_t M, uint32_t N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((N >> 3) << 17) | ((M >> 4) << 24);
}

struct TileCoord {
  uint32_t m_blk, n_blk;
};
__device__ __forceinline__ TileCoord tile_coord(uint32_t tile, uint32_t num_m, uint32_t num_n) {
  const uint32_t per_group = GROUP_M * num_n;
  const uint32_t group = tile / per_group;
  const uint32_t first_m = group * GROUP_M;
  const uint32_t gsize = (num_m - first_m < GROUP_M) ? (num_m - first_m) : GROUP_M;
  const uint32_t in_group = tile % per_group;
  return {first_m + in_group % gsize, in_group / gsize};
}

template <bool A_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    float* __restrict__ C, size_t ldc, uint32_t m, uint32_t n, uint32_t k,
                    int accumulate, int vec_ok) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t num_m = (m + BM - 1) / BM, num_n = (n + BN - 1) / BN;
  const uint32_t num_tiles = num_m * num_n;
  const uint32_t num_kb = (k + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (uint32_t s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileCoord tc = tile_coord(tile, num_m, num_n);
        for (uint32_t kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          if (!A_MN) {
            // A row-major m x k: box = 64 k (128 B) x 128 m-rows
            tma_load_2d(sa, &tmA, &full_bar[stage], (int32_t)(kb * BK), (int32_t)(tc.m_blk * BM));
          } else {
            // A stored k x m: boxes of 64 m (128 B) x 64 k-rows
#pragma unroll
            for (uint32_t j = 0; j < BM / 64; ++j)
              tma_load_2d(sa + j * ATOM_BYTES, &tmA, &full_bar[stage], (int32_t)(tc.m_blk * BM + j * 64),
                          (int32_t)(kb * BK));
          }
#pragma unroll
          for (uint32_t j = 0; j < BN / 64; ++j)
            tma_load_2d(sb + j * ATOM_BYTES, &tmB, &full_bar[stage], (int32_t)(tc.n_blk * BN + j * 64),
                        (int32_t)(kb * BK));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN, A_MN, true);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (uint32_t kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (uint32_t kk = 0; kk < BK / UMMA_K; ++kk) {
            // K-major: 8-row groups 1024 B apart (SBO), k advances 32 B inside the 128 B swizzle row
            // MN-major: 8 k-rows = 1024 B (SBO), next 64-wide MN block = ATOM_BYTES (LBO),
            //           k advances UMMA_K rows x 128 B
            const uint64_t adesc = A_MN ? make_smem_desc(sa + kk * UMMA_K * 128, ATOM_BYTES, 1024)
                                        : make_smem_desc(sa + kk * UMMA_K * 2, 0, 1024);
            const uint64_t bdesc = make_smem_desc(sb + kk * UMMA_K * 128, ATOM_BYTES, 1024);
            umma_bf16(tmem_d, adesc, bdesc, idesc, (kb | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);           // frees the smem stage when the MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);           // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const uint32_t g = warp & 3;                    // TMEM lane group this warp may access
    uint32_t acc = 0, acc_phase = 0;
    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const TileCoord tc = tile_coord(tile, num_m, num_n);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const size_t row = (size_t)tc.m_blk * BM + g * 32 + lane;
      float* crow = C + row * ldc;
#pragma unroll 1
      for (uint32_t c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((g * 32u) << 16) + acc * BN + c0, v);
        tmem_ld_wait();
        const size_t col0 = (size_t)tc.n_blk * BN + c0;
        if (row < m && col0 < n) {
          if (vec_ok && col0 + 32 <= n) {
#pragma unroll
            for (uint32_t j = 0; j < 32; j += 4) {
              float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                     __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
              float4* p = reinterpret_cast<float4*>(crow + col0 + j);
              if (accumulate) {
                const float4 c = *p;
                o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
              }
              *p = o;
            }
          } else {
#pragma unroll
            for (uint32_t j = 0; j < 32; ++j) {
              if (col0 + j < n) {
                float o = __uint_as_float(v[j]);
                if (accumulate) o += crow[col0 + j];
                crow[col0 + j] = o;
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host side ------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D bf16 tensor map over a row-major [rows x cols] matrix with leading dimension ld (elements);
// box = box_cols (inner, 64 = 128 B) x box_rows
int make_tmap(CUtensorMap* tm, const void* base, size_t rows, size_t cols, size_t ld, uint32_t box_cols,
              uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return B2_ERR_UNSUPPORTED;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)(ld * 2)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? B2_OK : B2_ERR_ARG;
}

__global__ void zero_c_kernel(float* C, size_t ldc, size_t m, size_t n) {
  const size_t total = m * n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    C[(i / n) * ldc + (i % n)] = 0.f;
}

}  // namespace

int b2_gemm_bf16_2cta(b2_ctx* ctx, const void* A, size_t lda, const void* B, size_t ldb, float* C, size_t ldc,
                      size_t m, size_t n, size_t k, int op_a, int accumulate, cudaStream_t st);

extern "C" int b2_gemm_bf16(b2_ctx* ctx, const void* A, size_t lda, const void* B, size_t ldb, float* C,
                            size_t ldc, size_t m, size_t n, size_t k, int op_a, int accumulate,
                            void* stream) {
  if (!ctx) return B2_ERR_ARG;
  if (op_a != B2_OP_N && op_a != B2_OP_T && op_a != B2_OP_H) return B2_ERR_ARG;
  if (m == 0 || n == 0) return B2_OK;
  if (!C) return B2_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (k == 0) {
    if (!accumulate) {
      zero_c_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(C, ldc, m, n);
      B2_LAUNCH_CHECK();
    }
    return B2_OK;
  }
  if (!A || !B) return B2_ERR_ARG;
  if (m > 0x7fffffffu || n > 0x7fffffffu || k > 0x7fffffffu) return B2_ERR_ARG;
  // TMA: 16-byte aligned bases, row pitches multiple of 16 bytes
  if (!b2_aligned16(A) || !b2_aligned16(B) || (lda % 8) || (ldb % 8)) return B2_ERR_ALIGN;
  {
    // kernel selection: default = cta_group::2 pair kernel (gemm_tc2.cu, 1.49 PF/s at 8192^3);
    // B2_GEMM_2CTA=0 -> the 1-CTA kernel below (1.33 PF/s), also used when m <= 128
    static int use2 = -1;
    if (use2 < 0) {
      const char* e = getenv("B2_GEMM_2CTA");
      use2 = e ? atoi(e) : 1;
    }
    if (use2 && m > BM) return b2_gemm_bf16_2cta(ctx, A, lda, B, ldb, C, ldc, m, n, k, op_a, accumulate, st);
  }
  const bool a_mn = (op_a != B2_OP_N);
  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn) rc = make_tmap(&tmA, A, m, k, lda, 64, BM);   // rows = m, cols = k
  else rc = make_tmap(&tmA, A, k, m, lda, 64, BK);          // rows = k, cols = m
  if (rc) return rc;
  rc = make_tmap(&tmB, B, k, n, ldb, 64, BK);               // rows = k, cols = n
  if (rc) return rc;
  const uint32_t num_tiles = (uint32_t)(((m + BM - 1) / BM) * ((n + BN - 1) / BN));
  const uint32_t grid = num_tiles < (uint32_t)ctx->sm_count ? num_tiles : (uint32_t)ctx->sm_count;
  const int vec_ok = (((uintptr_t)C & 15u) == 0 && (ldc % 4) == 0) ? 1 : 0;
  static bool attr_set[2] = {false, false};
  if (a_mn) {
    if (!attr_set[1]) {
      B2_CUDA(cudaFuncSetAttribute(gemm_bf16_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
      attr_set[1] = true;
    }
    gemm_bf16_tc_kernel<true><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmA, tmB, C, ldc, (uint32_t)m, (uint32_t)n,
                                                                      (uint32_t)k, accumulate, vec_ok);
  } else {
    if (!attr_set[0]) {
      B2_CUDA(cudaFuncSetAttribute(gemm_bf16_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
      attr_set[0] = true;
    }
    gemm_bf16_tc_kernel<false><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmA, tmB, C, ldc, (uint32_t)m, (uint32_t)n,
                                                                       (uint32_t)k, accumulate, vec_ok);
  }
  B2_LAUNCH_CHECK();
  return B2_OK;
}
