// 2-CTA (cta_group::2) variant of the bf16 -> fp32 tensor-core tile product.
// A CTA pair (cluster of 2 along M) computes a 256 x 256 output tile with UMMA M=256:
// each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 columns), the
// leader CTA's single MMA thread issues tcgen05.mma.cta_group::2 reading both CTAs' shared
// memory, and each CTA's TMEM holds its 128 rows of the accumulator.  Versus the 1-CTA kernel
// (gemm_tc.cu) this halves the B bytes staged per flop (32 KB instead of 48 KB of L2->SMEM
// traffic per 128x256x64 MMA block) and frees smem for a 6-stage ring.
// Same operand-major mapping as gemm_tc.cu: A op=N -> K-major, A op=T and B -> MN-major.
#include <cuda.h>
#include "common.cuh"

namespace {

constexpr uint32_t BM = 128, BN = 256, BNH = 128, BK = 64, STAGES = 6, UMMA_K = 16;
constexpr uint32_t A_STAGE_BYTES = BM * BK * 2;           // 16 KB (this CTA's rows)
constexpr uint32_t B_STAGE_BYTES = BK * BNH * 2;          // 16 KB (this CTA's half of the columns)
constexpr uint32_t STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr uint32_t ATOM_BYTES = 64 * BK * 2;              // 8 KB
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t NUM_THREADS = 192;
constexpr uint32_t GROUP_M = 4;                           // in units of 256-row cluster tiles
constexpr size_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr size_t SMEM_BYTES_SEG = SMEM_BYTES + 4 * 32 * 33 * sizeof(float);   // + epilogue transpose staging

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t tx) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(tx) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP2:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE2;\n\t"
      "bra WAIT_LOOP2;\n\t"
      "DONE2:\n\t"
      "}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// TMA load whose completion bytes are signalled on an mbarrier that may live in the PEER CTA
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* tmap, uint32_t bar_cluster_addr,
                                                int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this smem offset in BOTH CTAs of the pair once the MMAs retire
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar) {
  const uint16_t mask = 0x3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask)
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
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

// column-segmented output (stationary-A MPIMatrixMult, round 2): output columns [c*seg_cols, (c+1)*seg_cols) go to
// base[c] (leading dimension ldc) -- in general a buffer in ANOTHER GPU's memory (IPC-mapped, written over NVLink
// by the epilogue's 16-byte stores), so the reduce-scatter of the partial products needs no separate collective.
struct SegOut {
  float* base[8];
  uint32_t seg_cols;     // multiple of 32
  uint32_t nseg;         // 0 = plain C
};

template <bool A_MN, bool SEG = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     float* __restrict__ C, size_t ldc, uint32_t m, uint32_t n, uint32_t k, int accumulate,
                     int vec_ok, const SegOut seg = SegOut{}) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = (cta_rank == 0);
  const uint32_t cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const uint32_t num_m = (m + 2 * BM - 1) / (2 * BM), num_n = (n + BN - 1) / BN;
  const uint32_t num_tiles = num_m * num_n;
  const uint32_t num_kb = (k + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);      // leader: ONE arrive.expect_tx for the bytes of BOTH CTAs; the peer's TMA
                                       // only completes tx bytes (a blocking remote release-arrive per stage
                                       // from the peer's producer paced the whole pipeline: 721 TF/s)
      mbar_init(&empty_bar[s], 1);     // multicast commit from the leader's MMA thread
    }
    for (uint32_t s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);      // multicast commit
      mbar_init(&tmem_empty_bar[s], 8);     // leader only: 4 epilogue warps of each CTA (one arrive per warp)
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_ptr_smem, TMEM_COLS);
  tcgen05_fence_before();
  cluster_sync();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer (one thread per CTA) =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (uint32_t tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const TileCoord tc = tile_coord(tile, num_m, num_n);
        const int32_t m0 = (int32_t)(tc.m_blk * 2 * BM + cta_rank * BM);
        const int32_t n0 = (int32_t)(tc.n_blk * BN + cta_rank * BNH);
        for (uint32_t kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          const uint32_t leader_full = mapa(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          if (!A_MN) {
            tma_load_2d_2sm(sa, &tmA, leader_full, (int32_t)(kb * BK), m0);
          } else {
#pragma unroll
            for (uint32_t j = 0; j < BM / 64; ++j)
              tma_load_2d_2sm(sa + j * ATOM_BYTES, &tmA, leader_full, m0 + (int32_t)(j * 64), (int32_t)(kb * BK));
          }
#pragma unroll
          for (uint32_t j = 0; j < BNH / 64; ++j)
            tma_load_2d_2sm(sb + j * ATOM_BYTES, &tmB, leader_full, n0 + (int32_t)(j * 64), (int32_t)(kb * BK));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: ONE thread of the LEADER CTA drives both SMs =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc(2 * BM, BN, A_MN, true);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (uint32_t tile = cluster_id; tile < num_tiles; tile += num_clusters) {
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
            const uint64_t adesc = A_MN ? make_smem_desc(sa + kk * UMMA_K * 128, ATOM_BYTES, 1024)
                                        : make_smem_desc(sa + kk * UMMA_K * 2, 0, 1024);
            const uint64_t bdesc = make_smem_desc(sb + kk * UMMA_K * 128, ATOM_BYTES, 1024);
            umma_bf16_2cta(tmem_d, adesc, bdesc, idesc, (kb | kk) != 0 ? 1u : 0u);
          }
          umma_commit_multicast(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_multicast(&tmem_full_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (both CTAs): own 128 rows x 256 columns =====================
    const uint32_t g = warp & 3;
    uint32_t acc = 0, acc_phase = 0;
    for (uint32_t tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const TileCoord tc = tile_coord(tile, num_m, num_n);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const size_t row = (size_t)tc.m_blk * 2 * BM + cta_rank * BM + g * 32 + lane;
      float* crow = C + row * ldc;
#pragma unroll 1
      for (uint32_t c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((g * 32u) << 16) + acc * BN + c0, v);
        tmem_ld_wait();
        size_t col0 = (size_t)tc.n_blk * BN + c0;
        const bool in_range = row < m && col0 < n;
        if constexpr (SEG) {
          // Segmented output = (mostly) PEER memory: transpose the 32 x 32 chunk through shared memory so that each
          // store instruction writes one full 128-byte line of one row (lane = column) -- 16-byte pieces scattered
          // over 32 rows waste most of every NVLink packet (measured: stationary-A at 0.62 of peak with direct stores).
          float* stg = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256) + (warp - 2) * (32 * 33);
#pragma unroll
          for (uint32_t j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(v[j]);
          __syncwarp();
          if (col0 < n) {      // this 32-column chunk lies inside ONE segment (seg_cols % 32 == 0)
            const uint32_t sidx = (uint32_t)(col0 / seg.seg_cols);
            const size_t row_w0 = (size_t)tc.m_blk * 2 * BM + cta_rank * BM + g * 32;
            const uint32_t nrow = row_w0 < m ? (uint32_t)((m - row_w0 < 32) ? m - row_w0 : 32) : 0u;
            float* dst = seg.base[sidx] + row_w0 * ldc + (col0 - (size_t)sidx * seg.seg_cols) + lane;
            for (uint32_t r = 0; r < nrow; ++r, dst += ldc) *dst = stg[r * 33 + lane];
          }
          __syncwarp();
          continue;
        }
        if (in_range) {
          const size_t nlim = (size_t)n;
          if (vec_ok && col0 + 32 <= nlim) {
#pragma unroll
            for (uint32_t j = 0; j < 32; j += 4) {
              float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                     __uint_as_float(v[j + 3]));
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
              if (col0 + j < nlim) {
                float o = __uint_as_float(v[j]);
                if (accumulate) o += crow[col0 + j];
                crow[col0 + j] = o;
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa(smem_u32(&tmem_empty_bar[acc]), 0));   // leader's MMA thread waits
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  cluster_sync();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc2(tmem_base, TMEM_COLS);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap2(CUtensorMap* tm, const void* base, size_t rows, size_t cols, size_t ld, uint32_t box_cols,
               uint32_t box_rows) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
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

}  // namespace

// internal entry (dispatched from b2_gemm_bf16); preconditions already checked there
int b2_gemm_bf16_2cta(b2_ctx* ctx, const void* A, size_t lda, const void* B, size_t ldb, float* C, size_t ldc,
                      size_t m, size_t n, size_t k, int op_a, int accumulate, cudaStream_t st) {
  const bool a_mn = (op_a != B2_OP_N);
  CUtensorMap tmA, tmB;
  int rc = a_mn ? make_tmap2(&tmA, A, k, m, lda, 64, BK) : make_tmap2(&tmA, A, m, k, lda, 64, BM);
  if (rc) return rc;
  rc = make_tmap2(&tmB, B, k, n, ldb, 64, BK);
  if (rc) return rc;
  const uint32_t num_tiles = (uint32_t)(((m + 2 * BM - 1) / (2 * BM)) * ((n + BN - 1) / BN));
  uint32_t clusters = (uint32_t)ctx->sm_count / 2;
  if (num_tiles < clusters) clusters = num_tiles;
  const int vec_ok = (((uintptr_t)C & 15u) == 0 && (ldc % 4) == 0) ? 1 : 0;
  static bool attr_set[2] = {false, false};
  if (a_mn) {
    if (!attr_set[1]) {
      B2_CUDA(cudaFuncSetAttribute(gemm_bf16_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
      attr_set[1] = true;
    }
    gemm_bf16_tc2_kernel<true><<<2 * clusters, NUM_THREADS, SMEM_BYTES, st>>>(tmA, tmB, C, ldc, (uint32_t)m, (uint32_t)n,
                                                                              (uint32_t)k, accumulate, vec_ok);
  } else {
    if (!attr_set[0]) {
      B2_CUDA(cudaFuncSetAttribute(gemm_bf16_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
      attr_set[0] = true;
    }
    gemm_bf16_tc2_kernel<false><<<2 * clusters, NUM_THREADS, SMEM_BYTES, st>>>(tmA, tmB, C, ldc, (uint32_t)m, (uint32_t)n,
                                                                               (uint32_t)k, accumulate, vec_ok);
  }
  B2_LAUNCH_CHECK();
  return B2_OK;
}


// C segments: out[:, c*seg_cols:(c+1)*seg_cols) = (op(A) B)[:, same columns] written to segs_host[c] (ld = ldc);
// n must equal nseg * seg_cols, seg_cols % 32 == 0, every segment base 16-byte aligned, ldc % 4 == 0.
extern "C" int b2_gemm_bf16_seg(b2_ctx* ctx, const void* A, size_t lda, const void* B, size_t ldb, float* const* segs_host,
                                int nseg, size_t seg_cols, size_t ldc, size_t m, size_t n, size_t k, int op_a, void* stream) {
  if (!ctx || !A || !B || !segs_host || nseg < 1 || nseg > 8) return B2_ERR_ARG;
  if (op_a != B2_OP_N && op_a != B2_OP_T && op_a != B2_OP_H) return B2_ERR_ARG;
  if (m == 0 || n == 0 || k == 0) return B2_ERR_ARG;
  if (seg_cols % 32 || n != (size_t)nseg * seg_cols || (ldc % 4)) return B2_ERR_ARG;
  if (m > 0x7fffffffu || n > 0x7fffffffu || k > 0x7fffffffu) return B2_ERR_ARG;
  if (!b2_aligned16(A) || !b2_aligned16(B) || (lda % 8) || (ldb % 8)) return B2_ERR_ALIGN;
  SegOut so;
  so.seg_cols = (uint32_t)seg_cols;
  so.nseg = (uint32_t)nseg;
  for (int i = 0; i < 8; ++i) {
    so.base[i] = i < nseg ? segs_host[i] : nullptr;
    if (i < nseg && (!segs_host[i] || !b2_aligned16(segs_host[i]))) return B2_ERR_ALIGN;
  }
  const bool a_mn = (op_a != B2_OP_N);
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tmA, tmB;
  int rc = a_mn ? make_tmap2(&tmA, A, k, m, lda, 64, BK) : make_tmap2(&tmA, A, m, k, lda, 64, BM);
  if (rc) return rc;
  rc = make_tmap2(&tmB, B, k, n, ldb, 64, BK);
  if (rc) return rc;
  const uint32_t num_tiles = (uint32_t)(((m + 2 * BM - 1) / (2 * BM)) * ((n + BN - 1) / BN));
  uint32_t clusters = (uint32_t)ctx->sm_count / 2;
  if (num_tiles < clusters) clusters = num_tiles;
  static bool attr_set[2] = {false, false};
  if (a_mn) {
    if (!attr_set[1]) {
      B2_CUDA(cudaFuncSetAttribute(gemm_bf16_tc2_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES_SEG));
      attr_set[1] = true;
    }
    gemm_bf16_tc2_kernel<true, true><<<2 * clusters, NUM_THREADS, SMEM_BYTES_SEG, st>>>(tmA, tmB, nullptr, ldc, (uint32_t)m, (uint32_t)n,
                                                                                    (uint32_t)k, 0, 1, so);
  } else {
    if (!attr_set[0]) {
      B2_CUDA(cudaFuncSetAttribute(gemm_bf16_tc2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES_SEG));
      attr_set[0] = true;
    }
    gemm_bf16_tc2_kernel<false, true><<<2 * clusters, NUM_THREADS, SMEM_BYTES_SEG, st>>>(tmA, tmB, nullptr, ldc, (uint32_t)m, (uint32_t)n,
                                                                                     (uint32_t)k, 0, 1, so);
  }
  B2_LAUNCH_CHECK();
  return B2_OK;
}
