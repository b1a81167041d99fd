// MPIFredholm1 per-rank batched product on the tcgen05 tensor cores (reference:
// pylops_mpi/signalprocessing/Fredholm1.py:119-132 forward `ncp.matmul(self.G, x)`, :147-170 adjoint
// `ncp.matmul(GT, x)` / `G.conj().transpose @ x`).
//
//   y[s] = op(G[s]) x[s],   s < nsl,   G[s]: nx x ny,  x[s]: (ny | nx) x nz,  float32 or complex64.
//
// float32-class accuracy on bf16 tensor cores ("bf16x3" split): every fp32 number v is written as
// v = v0 + v1 + v2 with bf16 v0 = rn(v), v1 = rn(v - v0), v2 = rn(v - v0 - v1) (24 significant bits), and
//   a*b ~= a0 b0 + a0 b1 + a1 b0 + a0 b2 + a1 b1 + a2 b0            (dropped terms <= 2^-24 |a||b|)
// is six tensor-core products.  The tensor core's fp32 accumulate is lossier than an FMA chain (measured:
// 8x the error of the SIMT kernel with ONE accumulator, growing with the number of accumulator updates), so
// the leading term a0 b0 and the five correction terms (2^-8 smaller) go to TWO TMEM accumulators that the
// epilogue adds: only K/16 updates touch the large accumulator instead of 6K/16.  Each operand tile is staged
// once per k-block and used by up to three of the six MMAs, so the L2->SMEM traffic per MMA is HALF that of
// a plain 128x128 bf16 GEMM tile.
//
// complex64 as one REAL product: G[s] viewed as floats is the real matrix A (nx x 2ny, columns = re,im
// interleaved); with X' (2ny x 2nz) built from x as
//      X'[2k  ,2z] =  re x[k,z]   X'[2k  ,2z+1] = im x[k,z]
//      X'[2k+1,2z] = -im x[k,z]   X'[2k+1,2z+1] = re x[k,z]
// A X' (nx x 2nz) IS the interleaved complex64 result.  Same flops as the complex product (8 nx ny nz).
//
// Operator state vs per-apply data: G is operator state -> its planes (and those of G^H, the reference's
// `saveGt`) are split ONCE at plan creation; x changes every apply -> `pack_x_kernel` builds the three
// bf16 planes of X'^T (K-major, so both MMA operands are the canonical "TN" form) right before the product.
//
// Product kernel: persistent CTAs, warp 0 = TMA producer (3-D tensor maps [k', row, slice*3+plane], OOB
// zero fill => arbitrary nx, ny, nz), warp 1 = single-thread tcgen05.mma issuer (M=128, N<=128, K=16; two
// TMEM accumulator stages), warps 2-5 = epilogue (tcgen05.ld -> registers -> y, and -- fused all-gather of
// Fredholm1.py:131-132 -- the same 16-byte stores into every peer GPU's IPC-mapped output over NVLink).
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "tc_ptx.cuh"

using namespace tcptx;

namespace {

constexpr uint32_t BM = 128, BN = 128, UMMA_K = 16, NPL = 3;
constexpr uint32_t NUM_THREADS = 192;
constexpr uint32_t TMEM_COLS = 512;   // 2 accumulator stages x (main | small-terms) x 128 fp32 columns
constexpr uint32_t ACC_COLS = 2 * BN; // columns of one accumulator stage

template <uint32_t BK>
struct Cfg {
  static constexpr uint32_t TILE_BYTES = 128 * BK * 2;             // one plane tile (128 rows x BK bf16)
  static constexpr uint32_t STAGE_BYTES = 2 * NPL * TILE_BYTES;     // A0..A2, B0..B2
  static constexpr uint32_t STAGES = (BK == 64) ? 2 : 4;            // 192 KB of operand ring either way
  static constexpr uint32_t SBO = 8 * BK * 2;                       // 8 rows of one swizzle span
  static constexpr uint64_t LAYOUT = (BK == 64) ? 2 : 4;            // SWIZZLE_128B : SWIZZLE_64B
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 + 256;
};

struct PeerOut {
  float* p[8];
  int n;
};

template <uint32_t BK>
__global__ void __launch_bounds__(NUM_THREADS, 1)
fredholm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   float* __restrict__ Y, const PeerOut peers, uint32_t nsl, uint32_t m, uint32_t n, uint32_t kpad,
                   uint32_t n_umma, int vec_ok) {
  using C = Cfg<BK>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full_bar = empty_bar + C::STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t num_m = (m + BM - 1) / BM, num_n = (n + BN - 1) / BN;
  const uint32_t num_tiles = nsl * num_m * num_n;
  const uint32_t num_kb = (kpad + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t s = 0; s < C::STAGES; ++s) {
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
      const uint32_t tx_bytes = NPL * (BM * BK * 2) + NPL * (n_umma * BK * 2);
      for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const uint32_t n_blk = tile % num_n, m_blk = (tile / num_n) % num_m, s = tile / (num_n * num_m);
        for (uint32_t kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + NPL * C::TILE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
#pragma unroll
          for (uint32_t p = 0; p < NPL; ++p)
            tma_load_3d(sa + p * C::TILE_BYTES, &tmA, &full_bar[stage], (int32_t)(kb * BK), (int32_t)(m_blk * BM),
                        (int32_t)(s * NPL + p));
#pragma unroll
          for (uint32_t p = 0; p < NPL; ++p)
            tma_load_3d(sb + p * C::TILE_BYTES, &tmB, &full_bar[stage], (int32_t)(kb * BK), (int32_t)(n_blk * BN),
                        (int32_t)(s * NPL + p));
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(BM, n_umma, false, false);   // both operands K-major
      // small terms first: (A2,B0) (A1,B1) (A0,B2) (A1,B0) (A0,B1) (A0,B0)
      const uint32_t pa[6] = {2, 1, 0, 1, 0, 0}, pb[6] = {0, 1, 2, 0, 1, 0};
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_main = tmem_base + acc * ACC_COLS, tmem_small = tmem_main + BN;
        for (uint32_t kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + NPL * C::TILE_BYTES;
#pragma unroll
          for (uint32_t kk = 0; kk < BK / UMMA_K; ++kk) {
#pragma unroll
            for (uint32_t q = 0; q < 6; ++q) {
              // K-major, swizzled: 8-row groups SBO apart, k advances 32 B inside the swizzle span
              const uint64_t adesc = make_smem_desc(sa + pa[q] * C::TILE_BYTES + kk * UMMA_K * 2, 0, C::SBO, C::LAYOUT);
              const uint64_t bdesc = make_smem_desc(sb + pb[q] * C::TILE_BYTES + kk * UMMA_K * 2, 0, C::SBO, C::LAYOUT);
              if (q == 5) umma_bf16(tmem_main, adesc, bdesc, idesc, (kb | kk) != 0 ? 1u : 0u);
              else umma_bf16(tmem_small, adesc, bdesc, idesc, (kb | kk | q) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> y (+ peers over NVLink) =====================
    const uint32_t g = warp & 3;
    uint32_t acc = 0, acc_phase = 0;
    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const uint32_t n_blk = tile % num_n, m_blk = (tile / num_n) % num_m, s = tile / (num_n * num_m);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t row = m_blk * BM + g * 32 + lane;
      const size_t roff = ((size_t)s * m + row) * n;
#pragma unroll 1
      for (uint32_t c0 = 0; c0 < n_umma; c0 += 32) {
        uint32_t v[32], w[32];
        tmem_ld_32x32b_x32(tmem_base + ((g * 32u) << 16) + acc * ACC_COLS + c0, v);
        tmem_ld_32x32b_x32(tmem_base + ((g * 32u) << 16) + acc * ACC_COLS + BN + c0, w);
        tmem_ld_wait();
#pragma unroll
        for (uint32_t j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
        const uint32_t col0 = n_blk * BN + c0;
        if (row < m && col0 < n) {
          const size_t off = roff + col0;
          if (vec_ok && col0 + 32 <= n) {
#pragma unroll
            for (uint32_t j = 0; j < 32; j += 4) {
              const float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                           __uint_as_float(v[j + 3]));
              *reinterpret_cast<float4*>(Y + off + j) = o;
              for (int d = 0; d < peers.n; ++d) *reinterpret_cast<float4*>(peers.p[d] + off + j) = o;
            }
          } else {
#pragma unroll
            for (uint32_t j = 0; j < 32; ++j) {
              if (col0 + j < n) {
                const float o = __uint_as_float(v[j]);
                Y[off + j] = o;
                for (int d = 0; d < peers.n; ++d) peers.p[d][off + j] = o;
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

// ---- bf16x3 split --------------------------------------------------------------------------------------
struct Split3 {
  __nv_bfloat16 p[3];
};
__device__ __forceinline__ Split3 split3(float v) {
  Split3 r;
  r.p[0] = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(r.p[0]);
  r.p[1] = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(r.p[1]);
  r.p[2] = __float2bfloat16_rn(r2);
  return r;
}

// operator state, once per plan: planes[s][p][r][c'] (c' < kpad, zero padded) of op(G[s]) as a real matrix.
//   dir 0:  r = i (nx rows),  c' = cx ? 2j+cc : j   <- G[s][i][j]            (cc: 0 = re, 1 = im)
//   dir 1:  r = j (ny rows),  c' = cx ? 2i+cc : i   <- conj(G[s][i][j])      (G^H)
__global__ void pack_g_kernel(const float* __restrict__ G, __nv_bfloat16* __restrict__ out, size_t nsl, size_t nx,
                              size_t ny, int cx, int dir, size_t rows, size_t kpad) {
  const size_t total = nsl * rows * kpad;
  const size_t kp = (dir == 0 ? ny : nx) * (cx ? 2 : 1);
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t c = e % kpad, r = (e / kpad) % rows, s = e / (kpad * rows);
    float v = 0.f;
    if (c < kp) {
      const size_t inner = cx ? c / 2 : c;
      const int cc = cx ? (int)(c & 1) : 0;
      const size_t i = dir == 0 ? r : inner, j = dir == 0 ? inner : r;
      const size_t idx = (s * nx + i) * ny + j;
      v = cx ? G[2 * idx + cc] : G[idx];
      if (dir == 1 && cc == 1) v = -v;
    }
    const Split3 sp = split3(v);
    const size_t base = ((s * NPL) * rows + r) * kpad + c;
#pragma unroll
    for (int p = 0; p < 3; ++p) out[base + (size_t)p * rows * kpad] = sp.p[p];
  }
}

// per apply: planes of X'^T,  BT[s][p][n'][k'] (k' < kp; padding columns stay zero from plan creation)
//   complex: n' = 2z+d, k' = 2k+c:  (c,d) = (0,0) re, (1,0) -im, (0,1) im, (1,1) re
//   real   : n' = z,    k' = k
template <bool CX>
__global__ void __launch_bounds__(256)
pack_x_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ BT, uint32_t K, uint32_t nz, uint32_t nrows,
              uint32_t kpad) {
  __shared__ float2 tile[32][33];
  const uint32_t s = blockIdx.z, k0 = blockIdx.y * 32, z0 = blockIdx.x * 32;
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (uint32_t kk = ty; kk < 32; kk += 8) {
    const uint32_t k = k0 + kk, z = z0 + tx;
    float2 v = make_float2(0.f, 0.f);
    if (k < K && z < nz) {
      const size_t idx = ((size_t)s * K + k) * nz + z;
      if (CX) v = reinterpret_cast<const float2*>(x)[idx];
      else v.x = x[idx];
    }
    tile[kk][tx] = v;
  }
  __syncthreads();
  const size_t plane = (size_t)nrows * kpad;
  __nv_bfloat16* base = BT + (size_t)s * NPL * plane;
  const uint32_t k = k0 + tx;
  for (uint32_t zz = ty; zz < 32; zz += 8) {
    const uint32_t z = z0 + zz;
    if (k >= K || z >= nz) continue;
    const float2 v = tile[tx][zz];
    const Split3 re = split3(v.x);
    if (CX) {
      const Split3 im = split3(v.y);
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        __nv_bfloat162 r0, r1;
        r0.x = re.p[p];  r0.y = __hneg(im.p[p]);      // row 2z  : k' = 2k -> re, 2k+1 -> -im
        r1.x = im.p[p];  r1.y = re.p[p];              // row 2z+1: k' = 2k -> im, 2k+1 -> re
        __nv_bfloat16* q = base + p * plane;
        *reinterpret_cast<__nv_bfloat162*>(q + (size_t)(2 * z) * kpad + 2 * k) = r0;
        *reinterpret_cast<__nv_bfloat162*>(q + (size_t)(2 * z + 1) * kpad + 2 * k) = r1;
      }
    } else {
#pragma unroll
      for (int p = 0; p < 3; ++p) base[p * plane + (size_t)z * kpad + k] = re.p[p];
    }
  }
}

int make_tmap3(CUtensorMap* tm, const void* base, uint64_t kpad, uint64_t rows, uint64_t nmat, uint32_t box_k,
               uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return B2_ERR_UNSUPPORTED;
  cuuint64_t gdim[3] = {kpad, rows, nmat};
  cuuint64_t gstr[2] = {kpad * 2, rows * kpad * 2};
  cuuint32_t box[3] = {box_k, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = box_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? B2_OK : B2_ERR_ARG;
}

size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct b2_fredholm_plan {
  b2_ctx* ctx;
  size_t nsl, nx, ny, nz;
  int cx;                       // complex64 (1) or float32 (0)
  uint32_t bk;                  // 64 (128B swizzle, 2 stages) or 32 (64B swizzle, 4 stages)
  // per direction d (0 forward, 1 adjoint): output rows m[d], contraction length kp[d] (real), padded kpad[d]
  size_t m[2], kp[2], kpad[2];
  __nv_bfloat16* A[2];          // planes of op(G): [nsl][3][m][kpad]
  __nv_bfloat16* BT[2];         // planes of X'^T : [nsl][3][n][kpad]   (per-apply workspace, padding kept zero)
  uint32_t n, n_umma;           // output columns (real), UMMA N
  CUtensorMap tmA[2], tmB[2];
};

extern "C" int b2_fredholm_plan_destroy(b2_fredholm_plan* pl) {
  if (!pl) return B2_OK;
  for (int d = 0; d < 2; ++d) {
    if (pl->A[d]) cudaFree(pl->A[d]);
    if (pl->BT[d]) cudaFree(pl->BT[d]);
  }
  delete pl;
  return B2_OK;
}

extern "C" int b2_fredholm_plan_create(b2_ctx* ctx, const void* G, size_t nsl, size_t nx, size_t ny, size_t nz, int dtype,
                                       b2_fredholm_plan** out) {
  if (!ctx || !out || !G) return B2_ERR_ARG;
  if (dtype != B2_F32 && dtype != B2_C64) return B2_ERR_DTYPE;
  if (nsl == 0 || nx == 0 || ny == 0 || nz == 0) return B2_ERR_ARG;
  if (nsl * NPL > 0x7fffffffull || nx > 0x3fffffffull || ny > 0x3fffffffull || nz > 0x3fffffffull) return B2_ERR_ARG;
  if (!b2_aligned16(G)) return B2_ERR_ALIGN;
  b2_fredholm_plan* pl = new b2_fredholm_plan();
  memset(pl, 0, sizeof(*pl));
  pl->ctx = ctx;
  pl->nsl = nsl; pl->nx = nx; pl->ny = ny; pl->nz = nz;
  pl->cx = dtype == B2_C64;
  {
    static int bk = -1;
    if (bk < 0) {
      const char* e = getenv("B2_FREDHOLM_BK");
      bk = e ? atoi(e) : 32;
      if (bk != 32 && bk != 64) bk = 32;
    }
    pl->bk = (uint32_t)bk;
  }
  const size_t mul = pl->cx ? 2 : 1;
  pl->n = (uint32_t)(nz * mul);
  pl->n_umma = pl->n >= BN ? BN : (uint32_t)round_up(pl->n, 16);
  pl->m[0] = nx; pl->kp[0] = ny * mul;
  pl->m[1] = ny; pl->kp[1] = nx * mul;
  int rc = B2_OK;
  for (int d = 0; d < 2 && rc == B2_OK; ++d) {
    pl->kpad[d] = round_up(pl->kp[d], 8);
    const size_t a_elems = nsl * NPL * pl->m[d] * pl->kpad[d], b_elems = nsl * NPL * (size_t)pl->n * pl->kpad[d];
    cudaError_t e = cudaMalloc((void**)&pl->A[d], a_elems * 2);
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->BT[d], b_elems * 2);
    if (e == cudaSuccess) e = cudaMemset(pl->BT[d], 0, b_elems * 2);
    if (e != cudaSuccess) { rc = (int)e; break; }
    const size_t total = nsl * pl->m[d] * pl->kpad[d];
    size_t blocks = (total + 255) / 256;
    if (blocks > (size_t)ctx->sm_count * 32) blocks = (size_t)ctx->sm_count * 32;
    pack_g_kernel<<<(unsigned)blocks, 256>>>((const float*)G, pl->A[d], nsl, nx, ny, pl->cx, d, pl->m[d], pl->kpad[d]);
    e = cudaGetLastError();
    if (e != cudaSuccess) { rc = (int)e; break; }
    rc = make_tmap3(&pl->tmA[d], pl->A[d], pl->kpad[d], pl->m[d], nsl * NPL, pl->bk, BM);
    if (rc == B2_OK) rc = make_tmap3(&pl->tmB[d], pl->BT[d], pl->kpad[d], pl->n, nsl * NPL, pl->bk, pl->n_umma);
  }
  if (rc == B2_OK) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = (int)e;
  }
  if (rc != B2_OK) {
    b2_fredholm_plan_destroy(pl);
    return rc;
  }
  *out = pl;
  return B2_OK;
}

template <uint32_t BK>
static int launch_product(b2_fredholm_plan* pl, int d, float* y, const PeerOut& po, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    B2_CUDA(cudaFuncSetAttribute(fredholm_tc_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<BK>::SMEM_BYTES));
    attr_set = true;
  }
  const uint32_t m = (uint32_t)pl->m[d];
  const uint32_t num_tiles = (uint32_t)(pl->nsl * ((m + BM - 1) / BM) * ((pl->n + BN - 1) / BN));
  const uint32_t grid = num_tiles < (uint32_t)pl->ctx->sm_count ? num_tiles : (uint32_t)pl->ctx->sm_count;
  int vec_ok = (b2_aligned16(y) && (pl->n % 4) == 0) ? 1 : 0;
  for (int i = 0; i < po.n; ++i)
    if (!b2_aligned16(po.p[i])) vec_ok = 0;
  fredholm_tc_kernel<BK><<<grid, NUM_THREADS, Cfg<BK>::SMEM_BYTES, st>>>(pl->tmA[d], pl->tmB[d], y, po, (uint32_t)pl->nsl, m,
                                                                        pl->n, (uint32_t)pl->kpad[d], pl->n_umma, vec_ok);
  B2_LAUNCH_CHECK();
  return B2_OK;
}

// y[s] = op(G[s]) x[s] for all slices of the plan; peers_host (npeers <= 8, may be NULL/0): the same logical
// output position in peer GPUs' IPC-mapped buffers -- the epilogue stores every element there too (fused all-gather).
// Applies of one plan must be stream-ordered (they share the X' workspace).
extern "C" int b2_fredholm_apply(b2_fredholm_plan* pl, const void* x, void* y, void* const* peers_host, int npeers,
                                 int adjoint, void* stream) {
  if (!pl || !x || !y || npeers < 0 || npeers > 8 || (npeers && !peers_host)) return B2_ERR_ARG;
  const int d = adjoint ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  const uint32_t K = (uint32_t)(d == 0 ? pl->ny : pl->nx);
  dim3 grid((unsigned)((pl->nz + 31) / 32), (unsigned)((K + 31) / 32), (unsigned)pl->nsl);
  if (grid.y > 65535u || grid.z > 65535u) return B2_ERR_ARG;
  if (pl->cx)
    pack_x_kernel<true><<<grid, 256, 0, st>>>((const float*)x, pl->BT[d], K, (uint32_t)pl->nz, pl->n, (uint32_t)pl->kpad[d]);
  else
    pack_x_kernel<false><<<grid, 256, 0, st>>>((const float*)x, pl->BT[d], K, (uint32_t)pl->nz, pl->n, (uint32_t)pl->kpad[d]);
  B2_LAUNCH_CHECK();
  PeerOut po;
  po.n = npeers;
  for (int i = 0; i < 8; ++i) po.p[i] = i < npeers ? (float*)peers_host[i] : nullptr;
  return pl->bk == 64 ? launch_product<64>(pl, d, (float*)y, po, st) : launch_product<32>(pl, d, (float*)y, po, st);
}
