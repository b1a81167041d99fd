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
// Second operand format, "fp16x2" (B2_FREDHOLM_MODE=h2, default): v*2^e = hi + lo*2^-11 with fp16 hi = rn(v 2^e),
// lo = rn((v 2^e - hi) 2^11) (22 significant bits; e = power-of-two scale per G slice / per 32-column strip of
// x so that the largest element sits just below 2^15 -- undone exactly in the epilogue), and
//   a*b ~= hi_a hi_b + (hi_a lo_b + lo_a hi_b) 2^-11                  (dropped term <= 2^-22 |a||b|)
// is THREE products on two planes per operand: half the tensor-pipe time, 2/3 of the bytes of bf16x3
// (A planes = the 4 bytes/element of the float32 original).  Error-compensated split GEMM after Ootomo & Yokota.
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
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_ptx.cuh"

using namespace tcptx;

namespace {

constexpr uint32_t BM = 128, BN = 128, UMMA_K = 16;
constexpr uint32_t NUM_THREADS = 192;
constexpr uint32_t TMEM_COLS = 512;   // 2 accumulator stages x (main | small-terms) x 128 fp32 columns
constexpr uint32_t ACC_COLS = 2 * BN; // columns of one accumulator stage
constexpr int MODE_B3 = 0, MODE_H2 = 1;
constexpr uint32_t ZSTRIP = 32;       // columns of x per pack block / per fp16 scale

__host__ __device__ constexpr uint32_t npl_of(int mode) { return mode == MODE_B3 ? 3u : 2u; }

template <int MODE, uint32_t BK>
struct Cfg {
  static constexpr uint32_t NPL = npl_of(MODE);
  static constexpr uint32_t NQ = MODE == MODE_B3 ? 6 : 3;           // tensor-core products per k-step
  static constexpr uint32_t TILE_BYTES = 128 * BK * 2;             // one plane tile (128 rows x BK 16-bit)
  static constexpr uint32_t STAGE_BYTES = 2 * NPL * TILE_BYTES;     // A planes, B planes
  static constexpr uint32_t STAGES = (192u * 1024u) / STAGE_BYTES;  // 192 KB operand ring: 2/4 (b3), 3/6 (h2)
  static constexpr uint32_t SBO = 8 * BK * 2;                       // 8 rows of one swizzle span
  static constexpr uint64_t LAYOUT = (BK == 64) ? 2 : 4;            // SWIZZLE_128B : SWIZZLE_64B
  static constexpr size_t STAGING_BYTES = 4 * 32 * 33 * sizeof(float);   // epilogue transpose: [warp][32 rows][33]
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 + 256 + STAGING_BYTES;
};

struct PeerOut {
  float* p[8];
  int n;
};

__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// kind::f16 instruction descriptor, both operands K-major; fmt: 1 = bf16, 0 = fp16
__device__ __forceinline__ uint32_t make_idesc(uint32_t M, uint32_t N, uint32_t fmt) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

template <int MODE, uint32_t BK>
__global__ void __launch_bounds__(NUM_THREADS, 1)
fredholm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   float* __restrict__ Y, const PeerOut peers, const float* __restrict__ invA,
                   const float* __restrict__ invB, uint32_t nz, uint32_t zdiv, uint32_t nsl, uint32_t m,
                   uint32_t n, uint32_t kpad, uint32_t n_umma, int vec_ok, int concat, int staged) {
  using C = Cfg<MODE, BK>;
  constexpr uint32_t NPL = C::NPL;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full_bar = empty_bar + C::STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* stage_all = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + 256);

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
      bool dep_ready = false;     // the B planes are written by the pack kernel launched just before (PDL)
      const uint32_t tx_bytes = NPL * (BM * BK * 2) + NPL * (n_umma * BK * 2);
      for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const uint32_t n_blk = tile % num_n, m_blk = (tile / num_n) % num_m, s = tile / (num_n * num_m);
        for (uint32_t kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + NPL * C::TILE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
#pragma unroll
          for (uint32_t p = 0; p < NPL; ++p)      // operator state: independent of the pack kernel
            tma_load_3d(sa + p * C::TILE_BYTES, &tmA, &full_bar[stage], (int32_t)(kb * BK), (int32_t)(m_blk * BM),
                        (int32_t)(s * NPL + p));
          if (!dep_ready) { grid_dep_wait(); dep_ready = true; }
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
      const uint32_t idesc = make_idesc(BM, n_umma, MODE == MODE_B3 ? 1u : 0u);
      // small terms first; the LAST product of the list is the leading term (main accumulator)
      //   b3: (A2,B0) (A1,B1) (A0,B2) (A1,B0) (A0,B1) | (A0,B0)      h2: (hi,lo) (lo,hi) | (hi,hi)
      constexpr uint32_t NQ = C::NQ;
      const uint32_t pa[6] = {MODE == MODE_B3 ? 2u : 0u, 1u, 0u, 1u, 0u, 0u};
      const uint32_t pb[6] = {MODE == MODE_B3 ? 0u : 1u, MODE == MODE_B3 ? 1u : 0u, MODE == MODE_B3 ? 2u : 0u, 0u, 1u, 0u};
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
          if (MODE == MODE_H2 && concat) {
            // fp16x2, full-width tiles: the hi and lo planes of B sit back to back in the stage, so ONE N = 256 MMA
            // computes hi_a x [hi_b | lo_b] straight into [main | small]; lo_a x hi_b follows into small.  Two MMAs
            // per k-step instead of three: the same tensor time, 17 % fewer operand bytes read from shared memory.
            const uint32_t idesc2 = make_idesc(BM, 2 * BN, 0u);
#pragma unroll
            for (uint32_t kk = 0; kk < BK / UMMA_K; ++kk) {
              const uint64_t a_hi = make_smem_desc(sa + kk * UMMA_K * 2, 0, C::SBO, C::LAYOUT);
              const uint64_t a_lo = make_smem_desc(sa + C::TILE_BYTES + kk * UMMA_K * 2, 0, C::SBO, C::LAYOUT);
              const uint64_t b_hi = make_smem_desc(sb + kk * UMMA_K * 2, 0, C::SBO, C::LAYOUT);
              umma_bf16(tmem_main, a_hi, b_hi, idesc2, (kb | kk) != 0 ? 1u : 0u);
              umma_bf16(tmem_small, a_lo, b_hi, idesc, 1u);
            }
          } else
#pragma unroll
          for (uint32_t kk = 0; kk < BK / UMMA_K; ++kk) {
#pragma unroll
            for (uint32_t q = 0; q < NQ; ++q) {
              // K-major, swizzled: 8-row groups SBO apart, k advances 32 B inside the swizzle span
              const uint64_t adesc = make_smem_desc(sa + pa[q] * C::TILE_BYTES + kk * UMMA_K * 2, 0, C::SBO, C::LAYOUT);
              const uint64_t bdesc = make_smem_desc(sb + pb[q] * C::TILE_BYTES + kk * UMMA_K * 2, 0, C::SBO, C::LAYOUT);
              if (q == NQ - 1) umma_bf16(tmem_main, adesc, bdesc, idesc, (kb | kk) != 0 ? 1u : 0u);
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
    grid_dep_wait();     // y (and, in h2 mode, the scales of x) may only be touched once the previous kernel is done
    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const uint32_t n_blk = tile % num_n, m_blk = (tile / num_n) % num_m, s = tile / (num_n * num_m);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t row = m_blk * BM + g * 32 + lane;
      const size_t roff = ((size_t)s * m + row) * n;
      const float sa_inv = (MODE == MODE_H2 && row < m) ? invA[(size_t)s * m + row] : 1.f;   // per output row
#pragma unroll 1
      for (uint32_t c0 = 0; c0 < n_umma; c0 += 32) {
        uint32_t v[32], w[32];
        tmem_ld_32x32b_x32(tmem_base + ((g * 32u) << 16) + acc * ACC_COLS + c0, v);
        tmem_ld_32x32b_x32(tmem_base + ((g * 32u) << 16) + acc * ACC_COLS + BN + c0, w);
        tmem_ld_wait();
        const uint32_t col0 = n_blk * BN + c0;
        if (MODE == MODE_H2) {
          // undo the power-of-two operand scales (exact: per output row of op(G), per column of x) and the 2^11
          // of the correction terms
          const uint32_t zc = col0 / zdiv, nzc = 32 / zdiv;          // nzc distinct column scales in this chunk
          const float* sb_inv = invB + (size_t)s * nz + zc;
          float cs[32];
          if (zc + nzc <= nz && ((reinterpret_cast<uintptr_t>(sb_inv) & 15u) == 0)) {
#pragma unroll
            for (uint32_t t = 0; t < 32; t += 4) {                   // 16-byte broadcast loads (same address in every lane)
              if (t < nzc) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(sb_inv + t));
                cs[t] = q.x; cs[t + 1] = q.y; cs[t + 2] = q.z; cs[t + 3] = q.w;
              }
            }
          } else {
#pragma unroll
            for (uint32_t t = 0; t < 32; ++t)
              if (t < nzc) cs[t] = (zc + t < nz) ? __ldg(sb_inv + t) : 0.f;
          }
          if (zdiv == 2) {                                             // complex: columns (re, im) share a scale
#pragma unroll
            for (uint32_t j = 0; j < 32; ++j)
              v[j] = __float_as_uint(fmaf(__uint_as_float(w[j]), 1.f / 2048.f, __uint_as_float(v[j])) * (sa_inv * cs[j >> 1]));
          } else {
#pragma unroll
            for (uint32_t j = 0; j < 32; ++j)
              v[j] = __float_as_uint(fmaf(__uint_as_float(w[j]), 1.f / 2048.f, __uint_as_float(v[j])) * (sa_inv * cs[j]));
          }
        } else {
#pragma unroll
          for (uint32_t j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
        }
        // Transpose the 32 x 32 chunk through shared memory so that every store instruction writes ONE full 128-byte
        // line of ONE row (lane = column): the TMEM layout (lane = row) would give 32 scattered 16-byte pieces per
        // instruction -- tolerable in local HBM, but over NVLink (fused all-gather into the peers' outputs) 16-byte
        // writes waste most of every packet.
        if (!staged) {
          // single-GPU default: direct 16-byte stores of this thread's row piece
          if (row < m && col0 < n) {
            const size_t off = roff + col0;
            if (vec_ok && col0 + 32 <= n) {
#pragma unroll
              for (uint32_t j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(Y + off + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                     __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            } else {
#pragma unroll
              for (uint32_t j = 0; j < 32; ++j)
                if (col0 + j < n) Y[off + j] = __uint_as_float(v[j]);
            }
          }
          continue;
        }
        float* stg = stage_all + (warp - 2) * (32 * 33);
#pragma unroll
        for (uint32_t j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(v[j]);
        __syncwarp();
        const uint32_t ccol = col0 + lane;
        if (ccol < n) {
          const uint32_t row_w0 = m_blk * BM + g * 32;
          const uint32_t nrow = row_w0 < m ? (m - row_w0 < 32 ? m - row_w0 : 32) : 0;
          size_t off = ((size_t)s * m + row_w0) * n + ccol;
          for (uint32_t r = 0; r < nrow; ++r, off += n) {
            const float o = stg[r * 33 + lane];
            Y[off] = o;
            for (int d = 0; d < peers.n; ++d) peers.p[d][off] = o;
          }
        }
        __syncwarp();
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

// ---- operand splits (16-bit planes stored as raw ushort) ---------------------------------------------------
template <int MODE> struct Split { unsigned short p[npl_of(MODE)]; };

template <int MODE>
__device__ __forceinline__ Split<MODE> split(float v, float scale) {
  Split<MODE> r;
  if (MODE == MODE_B3) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(h0);
    const __nv_bfloat16 h1 = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(h1);
    r.p[0] = __bfloat16_as_ushort(h0);
    r.p[1] = __bfloat16_as_ushort(h1);
    r.p[npl_of(MODE) - 1] = __bfloat16_as_ushort(__float2bfloat16_rn(r2));
  } else {
    const float vs = v * scale;                      // |vs| < 2^15: no fp16 overflow
    const __half h0 = __float2half_rn(vs);
    r.p[0] = __half_as_ushort(h0);
    r.p[1] = __half_as_ushort(__float2half_rn((vs - __half2float(h0)) * 2048.f));
  }
  return r;
}
__device__ __forceinline__ unsigned short neg16(unsigned short h) { return h ^ 0x8000u; }   // bf16 and fp16: sign bit

// power-of-two scale that puts amax just below 2^15 (exponent clamped so scale and 1/scale stay normal floats)
__device__ __forceinline__ void pow2_scale(float amax, float* scale, float* inv) {
  int ex = 0;
  if (amax > 0.f && amax < INFINITY) frexpf(amax, &ex);      // amax = f * 2^ex, f in [0.5, 1)
  else ex = 15;
  int e = 15 - ex;
  e = e > 100 ? 100 : (e < -100 ? -100 : e);
  *scale = ldexpf(1.f, e);
  *inv = ldexpf(1.f, -e);
}

__device__ __forceinline__ float block_max(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < nw; ++i) r = fmaxf(r, red[i]);
  __syncthreads();
  return r;
}

// h2 only: power-of-two scale of every output ROW of op(G[s]) (operator state, once per plan and direction):
//   dir 0: row i of G[s] (contiguous);  dir 1: row j of G[s]^H = column j of G[s]
__global__ void __launch_bounds__(128) row_scale_kernel(const float* __restrict__ G, size_t nx, size_t ny, int cx, int dir,
                                                         float* scA, float* invA) {
  __shared__ float red[4];
  const size_t rows = dir == 0 ? nx : ny, inner = dir == 0 ? ny : nx;
  const size_t s = blockIdx.x / rows, r = blockIdx.x % rows;
  const size_t mul = cx ? 2 : 1;
  const float* g = G + s * nx * ny * mul;
  float am = 0.f;
  for (size_t q = threadIdx.x; q < inner * mul; q += blockDim.x) {
    const size_t e = q / mul, c = q % mul;
    const size_t idx = dir == 0 ? (r * ny + e) : (e * ny + r);
    am = fmaxf(am, fabsf(g[idx * mul + c]));
  }
  am = block_max(am, red);
  if (threadIdx.x == 0) pow2_scale(am, &scA[blockIdx.x], &invA[blockIdx.x]);
}

// operator state, once per plan: planes[s][p][r][c'] (c' < kpad, zero padded) of op(G[s]) as a real matrix.
//   dir 0:  r = i (nx rows),  c' = cx ? 2j+cc : j   <- G[s][i][j]            (cc: 0 = re, 1 = im)
//   dir 1:  r = j (ny rows),  c' = cx ? 2i+cc : i   <- conj(G[s][i][j])      (G^H)
template <int MODE>
__global__ void pack_g_kernel(const float* __restrict__ G, unsigned short* __restrict__ out, const float* __restrict__ scA,
                              size_t nsl, size_t nx, size_t ny, int cx, int dir, size_t rows, size_t kpad) {
  constexpr uint32_t NPL = npl_of(MODE);
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
    const Split<MODE> sp = split<MODE>(v, MODE == MODE_H2 ? scA[s * rows + r] : 1.f);
    const size_t base = ((s * NPL) * rows + r) * kpad + c;
#pragma unroll
    for (uint32_t p = 0; p < NPL; ++p) out[base + (size_t)p * rows * kpad] = sp.p[p];
  }
}

// per apply: planes of X'^T,  BT[s][p][n'][k'] (k' < kpad; columns in [kp, kpad) are written as zeros)
//   complex: n' = 2z+d, k' = 2k+c:  (c,d) = (0,0) re, (1,0) -im, (0,1) im, (1,1) re
//   real   : n' = z,    k' = k
// One 1024-thread block per (32-column strip of x, slice, 128-row tile of k): h2 first takes every COLUMN's amax over
// all k (power-of-two scale per column, its inverse goes to invB for the epilogue), then the 128 x 32 tile is
// transposed through shared memory and written as 16-byte (complex) / 8-byte (real) vectors along k'.
constexpr uint32_t PK_THREADS = 1024, PK_ROWS = 128;
template <bool CX, int MODE>
__global__ void __launch_bounds__(PK_THREADS)
pack_x_kernel(const float* __restrict__ x, unsigned short* __restrict__ BT, float* __restrict__ invB, uint32_t K,
              uint32_t nz, uint32_t nrows, uint32_t kpad) {
  constexpr uint32_t NPL = npl_of(MODE);
  __shared__ float2 tile[PK_ROWS][33];
  __shared__ float colmax[32][33];
  __shared__ float scale_s[32];
  grid_dep_launch();               // PDL: the product kernel may start its prologue / A loads now
  const uint32_t s = blockIdx.y, z0 = blockIdx.x * ZSTRIP;
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* xs = x + (size_t)s * K * nz * (CX ? 2 : 1);
  if (MODE == MODE_H2) {
    float am = 0.f;
    const uint32_t z = z0 + tx;
    if (z < nz)
      for (uint32_t k = ty; k < K; k += 32) {
        if (CX) {
          const float2 v = reinterpret_cast<const float2*>(xs)[(size_t)k * nz + z];
          am = fmaxf(am, fmaxf(fabsf(v.x), fabsf(v.y)));
        } else {
          am = fmaxf(am, fabsf(xs[(size_t)k * nz + z]));
        }
      }
    colmax[ty][tx] = am;
    __syncthreads();
    if (ty == 0) {
      float m = colmax[0][tx];
      for (int i = 1; i < 32; ++i) m = fmaxf(m, colmax[i][tx]);
      float sc, inv;
      pow2_scale(m, &sc, &inv);
      scale_s[tx] = sc;
      if (z < nz && blockIdx.z == 0) invB[(size_t)s * nz + z] = inv;
    }
    __syncthreads();
  }
  const size_t plane = (size_t)nrows * kpad;
  unsigned short* base = BT + (size_t)s * NPL * plane;
  const uint32_t zz = threadIdx.x >> 5, kq = threadIdx.x & 31;     // write phase: one z, four consecutive k
  const float scale = MODE == MODE_H2 ? scale_s[zz] : 1.f;
  // blockIdx.z selects ONE 128-row tile of k (the column scales above are recomputed by every k-block: a few L2
  // reads per thread, in exchange for twice the CTAs in flight on the config-5 shape)
  {
    const uint32_t k0 = blockIdx.z * PK_ROWS;
    for (uint32_t kk = ty; kk < PK_ROWS; kk += 32) {
      const uint32_t k = k0 + kk, z = z0 + tx;
      float2 v = make_float2(0.f, 0.f);
      if (k < K && z < nz) {
        if (CX) v = reinterpret_cast<const float2*>(xs)[(size_t)k * nz + z];
        else v.x = xs[(size_t)k * nz + z];
      }
      tile[kk][tx] = v;
    }
    __syncthreads();
    const uint32_t z = z0 + zz, kb = k0 + 4 * kq;
    if (z < nz && kb * (CX ? 2 : 1) < kpad) {
      Split<MODE> re[4], im[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 v = tile[4 * kq + j][zz];
        re[j] = split<MODE>(v.x, scale);
        if (CX) im[j] = split<MODE>(v.y, scale);
      }
#pragma unroll
      for (uint32_t p = 0; p < NPL; ++p) {
        unsigned short* q = base + p * plane;
        if (CX) {
          uint4 r0, r1;      // row 2z: (re, -im) pairs; row 2z+1: (im, re) pairs
          r0.x = re[0].p[p] | ((uint32_t)neg16(im[0].p[p]) << 16);  r1.x = im[0].p[p] | ((uint32_t)re[0].p[p] << 16);
          r0.y = re[1].p[p] | ((uint32_t)neg16(im[1].p[p]) << 16);  r1.y = im[1].p[p] | ((uint32_t)re[1].p[p] << 16);
          r0.z = re[2].p[p] | ((uint32_t)neg16(im[2].p[p]) << 16);  r1.z = im[2].p[p] | ((uint32_t)re[2].p[p] << 16);
          r0.w = re[3].p[p] | ((uint32_t)neg16(im[3].p[p]) << 16);  r1.w = im[3].p[p] | ((uint32_t)re[3].p[p] << 16);
          *reinterpret_cast<uint4*>(q + (size_t)(2 * z) * kpad + 2 * kb) = r0;
          *reinterpret_cast<uint4*>(q + (size_t)(2 * z + 1) * kpad + 2 * kb) = r1;
        } else {
          uint2 r0;
          r0.x = re[0].p[p] | ((uint32_t)re[1].p[p] << 16);
          r0.y = re[2].p[p] | ((uint32_t)re[3].p[p] << 16);
          *reinterpret_cast<uint2*>(q + (size_t)z * kpad + kb) = r0;
        }
      }
    }
    __syncthreads();
  }
}

// Single-pass variant for short contractions (k' columns covered by <= PS_K values of k: config 5 has K = 256):
// one 256-thread block per (16-column strip of x, slice) reads its whole K x 16 strip ONCE -- 16 independent loads
// per thread in flight -- into shared memory, takes the column scales from the values it already holds and writes
// the planes.  Versus the generic kernel above: no second read of x, a quarter of the threads per block (faster
// block launch), 256 blocks for config 5.  Shared tile columns are rotated by k/4 so that both the row-wise fill
// and the 4-consecutive-k reads of the write phase are bank-conflict free.
constexpr uint32_t PS_THREADS = 256, PS_K = 256, PS_Z = 16;
template <bool CX, int MODE>
__global__ void __launch_bounds__(PS_THREADS)
pack_x_small_kernel(const float* __restrict__ x, unsigned short* __restrict__ BT, float* __restrict__ invB, uint32_t K,
                    uint32_t nz, uint32_t nrows, uint32_t kpad) {
  constexpr uint32_t NPL = npl_of(MODE);
  __shared__ float2 tile[PS_K][PS_Z];
  __shared__ float colmax[PS_THREADS / PS_Z][PS_Z + 1];
  __shared__ float scale_s[PS_Z];
  grid_dep_launch();               // PDL: the product kernel may start its prologue / A loads now
  const uint32_t s = blockIdx.y, z0 = blockIdx.x * PS_Z;
  const uint32_t tx = threadIdx.x & (PS_Z - 1), ty = threadIdx.x / PS_Z;
  const float* xs = x + (size_t)s * K * nz * (CX ? 2 : 1);
  {
    const uint32_t z = z0 + tx;
    float2 v[PS_K / (PS_THREADS / PS_Z)];
#pragma unroll
    for (uint32_t i = 0; i < PS_K / (PS_THREADS / PS_Z); ++i) {
      const uint32_t k = ty + (PS_THREADS / PS_Z) * i;
      v[i] = make_float2(0.f, 0.f);
      if (k < K && z < nz) {
        if (CX) v[i] = reinterpret_cast<const float2*>(xs)[(size_t)k * nz + z];
        else v[i].x = xs[(size_t)k * nz + z];
      }
    }
    float am = 0.f;
#pragma unroll
    for (uint32_t i = 0; i < PS_K / (PS_THREADS / PS_Z); ++i) {
      const uint32_t k = ty + (PS_THREADS / PS_Z) * i;
      tile[k][(tx + (k >> 2)) & (PS_Z - 1)] = v[i];
      am = fmaxf(am, fmaxf(fabsf(v[i].x), fabsf(v[i].y)));
    }
    if (MODE == MODE_H2) colmax[ty][tx] = am;
  }
  __syncthreads();
  if (MODE == MODE_H2) {
    if (threadIdx.x < PS_Z) {
      float m = colmax[0][threadIdx.x];
#pragma unroll
      for (uint32_t i = 1; i < PS_THREADS / PS_Z; ++i) m = fmaxf(m, colmax[i][threadIdx.x]);
      float sc, inv;
      pow2_scale(m, &sc, &inv);
      scale_s[threadIdx.x] = sc;
      if (z0 + threadIdx.x < nz) invB[(size_t)s * nz + z0 + threadIdx.x] = inv;
    }
    __syncthreads();
  }
  const size_t plane = (size_t)nrows * kpad;
  unsigned short* base = BT + (size_t)s * NPL * plane;
#pragma unroll
  for (uint32_t it = 0; it < PS_Z * (PS_K / 4) / PS_THREADS; ++it) {
    const uint32_t item = threadIdx.x + PS_THREADS * it;
    const uint32_t kq = item & (PS_K / 4 - 1), zz = item / (PS_K / 4);
    const uint32_t z = z0 + zz, kb = 4 * kq;
    if (z >= nz || kb * (CX ? 2 : 1) >= kpad) continue;
    const float scale = MODE == MODE_H2 ? scale_s[zz] : 1.f;
    Split<MODE> re[4], im[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 v = tile[kb + j][(zz + kq) & (PS_Z - 1)];
      re[j] = split<MODE>(v.x, scale);
      if (CX) im[j] = split<MODE>(v.y, scale);
    }
#pragma unroll
    for (uint32_t p = 0; p < NPL; ++p) {
      unsigned short* q = base + p * plane;
      if (CX) {
        uint4 r0, r1;      // row 2z: (re, -im) pairs; row 2z+1: (im, re) pairs
        r0.x = re[0].p[p] | ((uint32_t)neg16(im[0].p[p]) << 16);  r1.x = im[0].p[p] | ((uint32_t)re[0].p[p] << 16);
        r0.y = re[1].p[p] | ((uint32_t)neg16(im[1].p[p]) << 16);  r1.y = im[1].p[p] | ((uint32_t)re[1].p[p] << 16);
        r0.z = re[2].p[p] | ((uint32_t)neg16(im[2].p[p]) << 16);  r1.z = im[2].p[p] | ((uint32_t)re[2].p[p] << 16);
        r0.w = re[3].p[p] | ((uint32_t)neg16(im[3].p[p]) << 16);  r1.w = im[3].p[p] | ((uint32_t)re[3].p[p] << 16);
        *reinterpret_cast<uint4*>(q + (size_t)(2 * z) * kpad + 2 * kb) = r0;
        *reinterpret_cast<uint4*>(q + (size_t)(2 * z + 1) * kpad + 2 * kb) = r1;
      } else {
        uint2 r0;
        r0.x = re[0].p[p] | ((uint32_t)re[1].p[p] << 16);
        r0.y = re[2].p[p] | ((uint32_t)re[3].p[p] << 16);
        *reinterpret_cast<uint2*>(q + (size_t)z * kpad + kb) = r0;
      }
    }
  }
}

int make_tmap3(CUtensorMap* tm, const void* base, uint64_t kpad, uint64_t rows, uint64_t nmat, uint32_t box_k,
               uint32_t box_rows, bool fp16) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return B2_ERR_UNSUPPORTED;
  cuuint64_t gdim[3] = {kpad, rows, nmat};
  cuuint64_t gstr[2] = {kpad * 2, rows * kpad * 2};
  cuuint32_t box[3] = {box_k, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = box_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = fn(tm, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base),
                  gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? B2_OK : B2_ERR_ARG;
}

size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct b2_fredholm_plan {
  b2_ctx* ctx;
  size_t nsl, nx, ny, nz;
  int cx;                       // complex64 (1) or float32 (0)
  int mode;                     // MODE_B3 (bf16x3) or MODE_H2 (fp16x2)
  uint32_t bk;                  // 64 (128B swizzle) or 32 (64B swizzle, twice the stages)
  // per direction d (0 forward, 1 adjoint): output rows m[d], contraction length kp[d] (real), padded kpad[d]
  size_t m[2], kp[2], kpad[2];
  unsigned short* A[2];         // planes of op(G): [nsl][npl][m][kpad]
  unsigned short* BT[2];        // planes of X'^T : [nsl][npl][n][kpad]   (per-apply workspace)
  float *scA[2], *invA[2], *invB;   // h2: power-of-two scale per (slice, output row) of op(G) and its inverse (per
                                    // direction), inverse scale per (slice, column of x)
  uint32_t n, n_umma, nstrips;  // output columns (real), UMMA N, 32-column strips of x
  int concat;                   // fp16x2 with full 128-column tiles: hi_a x [hi_b | lo_b] as one N = 256 MMA
  int pack_small;               // B2_FREDHOLM_PACK_SMALL=0: always use the generic two-pass pack kernel
  int stage_always;             // B2_FREDHOLM_STAGE=1: coalescing epilogue also without peers (default: only with peers)
  CUtensorMap tmA[2], tmB[2];
};

extern "C" int b2_fredholm_plan_destroy(b2_fredholm_plan* pl) {
  if (!pl) return B2_OK;
  for (int d = 0; d < 2; ++d) {
    if (pl->A[d]) cudaFree(pl->A[d]);
    if (pl->BT[d]) cudaFree(pl->BT[d]);
  }
  for (int d = 0; d < 2; ++d) {
    if (pl->scA[d]) cudaFree(pl->scA[d]);
    if (pl->invA[d]) cudaFree(pl->invA[d]);
  }
  if (pl->invB) cudaFree(pl->invB);
  delete pl;
  return B2_OK;
}

extern "C" int b2_fredholm_plan_create(b2_ctx* ctx, const void* G, size_t nsl, size_t nx, size_t ny, size_t nz, int dtype,
                                       b2_fredholm_plan** out) {
  if (!ctx || !out || !G) return B2_ERR_ARG;
  if (dtype != B2_F32 && dtype != B2_C64) return B2_ERR_DTYPE;
  if (nsl == 0 || nx == 0 || ny == 0 || nz == 0) return B2_ERR_ARG;
  if (nsl * 3 > 0x7fffffffull || nsl > 65535 || nx > 0x3fffffffull || ny > 0x3fffffffull || nz > 0x3fffffffull) return B2_ERR_ARG;
  if (!b2_aligned16(G)) return B2_ERR_ALIGN;
  b2_fredholm_plan* pl = new b2_fredholm_plan();
  memset(pl, 0, sizeof(*pl));
  pl->ctx = ctx;
  pl->nsl = nsl; pl->nx = nx; pl->ny = ny; pl->nz = nz;
  pl->cx = dtype == B2_C64;
  {
    const char* e = getenv("B2_FREDHOLM_BK");
    const int bk = e ? atoi(e) : 32;
    pl->bk = bk == 64 ? 64u : 32u;
    const char* mo = getenv("B2_FREDHOLM_MODE");
    pl->mode = (mo && (mo[0] == 'b' || mo[0] == 'B')) ? MODE_B3 : MODE_H2;
    const char* cc = getenv("B2_FREDHOLM_CONCAT");
    pl->concat = cc ? atoi(cc) : 1;
    const char* sg = getenv("B2_FREDHOLM_STAGE");
    pl->stage_always = sg ? atoi(sg) : 0;
    const char* ps = getenv("B2_FREDHOLM_PACK_SMALL");
    pl->pack_small = ps ? atoi(ps) : 1;
  }
  const uint32_t npl = npl_of(pl->mode);
  const size_t mul = pl->cx ? 2 : 1;
  pl->n = (uint32_t)(nz * mul);
  pl->n_umma = pl->n >= BN ? BN : (uint32_t)round_up(pl->n, 16);
  if (pl->n_umma != BN || pl->mode != MODE_H2) pl->concat = 0;
  pl->nstrips = (uint32_t)((nz + ZSTRIP - 1) / ZSTRIP);
  pl->m[0] = nx; pl->kp[0] = ny * mul;
  pl->m[1] = ny; pl->kp[1] = nx * mul;
  int rc = B2_OK;
  cudaError_t e = cudaMalloc((void**)&pl->invB, nsl * nz * sizeof(float));
  if (e != cudaSuccess) rc = (int)e;
  for (int d = 0; d < 2 && rc == B2_OK; ++d) {
    e = cudaMalloc((void**)&pl->scA[d], nsl * pl->m[d] * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->invA[d], nsl * pl->m[d] * sizeof(float));
    if (e != cudaSuccess) { rc = (int)e; break; }
    if (pl->mode == MODE_H2) {
      if (nsl * pl->m[d] > 0x7fffffffull) { rc = B2_ERR_ARG; break; }
      row_scale_kernel<<<(unsigned)(nsl * pl->m[d]), 128>>>((const float*)G, nx, ny, pl->cx, d, pl->scA[d], pl->invA[d]);
      e = cudaGetLastError();
      if (e != cudaSuccess) { rc = (int)e; break; }
    }
    pl->kpad[d] = round_up(pl->kp[d], 8);
    const size_t a_elems = nsl * npl * pl->m[d] * pl->kpad[d], b_elems = nsl * npl * (size_t)pl->n * pl->kpad[d];
    e = cudaMalloc((void**)&pl->A[d], a_elems * 2);
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->BT[d], b_elems * 2);
    if (e == cudaSuccess) e = cudaMemset(pl->BT[d], 0, b_elems * 2);
    if (e != cudaSuccess) { rc = (int)e; break; }
    const size_t total = nsl * pl->m[d] * pl->kpad[d];
    size_t blocks = (total + 255) / 256;
    if (blocks > (size_t)ctx->sm_count * 32) blocks = (size_t)ctx->sm_count * 32;
    if (pl->mode == MODE_B3)
      pack_g_kernel<MODE_B3><<<(unsigned)blocks, 256>>>((const float*)G, pl->A[d], pl->scA[d], nsl, nx, ny, pl->cx, d, pl->m[d], pl->kpad[d]);
    else
      pack_g_kernel<MODE_H2><<<(unsigned)blocks, 256>>>((const float*)G, pl->A[d], pl->scA[d], nsl, nx, ny, pl->cx, d, pl->m[d], pl->kpad[d]);
    e = cudaGetLastError();
    if (e != cudaSuccess) { rc = (int)e; break; }
    const bool fp16 = pl->mode == MODE_H2;
    rc = make_tmap3(&pl->tmA[d], pl->A[d], pl->kpad[d], pl->m[d], nsl * npl, pl->bk, BM, fp16);
    if (rc == B2_OK) rc = make_tmap3(&pl->tmB[d], pl->BT[d], pl->kpad[d], pl->n, nsl * npl, pl->bk, pl->n_umma, fp16);
  }
  if (rc == B2_OK) {
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = (int)e;
  }
  if (rc != B2_OK) {
    b2_fredholm_plan_destroy(pl);
    return rc;
  }
  *out = pl;
  return B2_OK;
}

template <int MODE, uint32_t BK>
static int launch_product(b2_fredholm_plan* pl, int d, float* y, const PeerOut& po, cudaStream_t st) {
  using C = Cfg<MODE, BK>;
  static bool attr_set = false;
  if (!attr_set) {
    B2_CUDA(cudaFuncSetAttribute(fredholm_tc_kernel<MODE, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
    attr_set = true;
  }
  const uint32_t m = (uint32_t)pl->m[d];
  const uint32_t num_tiles = (uint32_t)(pl->nsl * ((m + BM - 1) / BM) * ((pl->n + BN - 1) / BN));
  const uint32_t grid = num_tiles < (uint32_t)pl->ctx->sm_count ? num_tiles : (uint32_t)pl->ctx->sm_count;
  int vec_ok = (b2_aligned16(y) && (pl->n % 4) == 0) ? 1 : 0;
  for (int i = 0; i < po.n; ++i)
    if (!b2_aligned16(po.p[i])) vec_ok = 0;
  // programmatic dependent launch: this kernel's prologue and its loads of the (static) A planes overlap the
  // tail of the pack kernel; griddepcontrol.wait guards everything that depends on the pack kernel's output
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  B2_CUDA(cudaLaunchKernelEx(&cfg, fredholm_tc_kernel<MODE, BK>, pl->tmA[d], pl->tmB[d], y, po, (const float*)pl->invA[d],
                             (const float*)pl->invB, (uint32_t)pl->nz, pl->cx ? 2u : 1u, (uint32_t)pl->nsl, m, pl->n,
                             (uint32_t)pl->kpad[d], pl->n_umma, vec_ok, pl->concat,
                             (po.n > 0 || pl->stage_always) ? 1 : 0));
  return B2_OK;
}

// y[s] = op(G[s]) x[s] for all slices of the plan; peers_host (npeers <= 8, may be NULL/0): the same logical
// output position in peer GPUs' IPC-mapped buffers -- the epilogue stores every element there too (fused all-gather).
// Applies of one plan must be stream-ordered (they share the X' workspace).
static int fredholm_apply_impl(b2_fredholm_plan* pl, const void* x, void* y, void* const* peers_host, int npeers,
                               int adjoint, int parts, void* stream) {
  if (!pl || !x || !y || npeers < 0 || npeers > 8 || (npeers && !peers_host)) return B2_ERR_ARG;
  const int d = adjoint ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  const uint32_t K = (uint32_t)(d == 0 ? pl->ny : pl->nx);
  const uint32_t nz = (uint32_t)pl->nz, kpad = (uint32_t)pl->kpad[d];
  const uint32_t kcover = kpad / (pl->cx ? 2u : 1u);          // k values whose k' columns exist (incl. padding)
  dim3 grid(pl->nstrips, (unsigned)pl->nsl, (kcover + PK_ROWS - 1) / PK_ROWS);
  if (grid.z > 65535u) return B2_ERR_ARG;
  const float* xf = (const float*)x;
  if ((parts & 1) && pl->pack_small && kcover <= PS_K) {
    dim3 gs((nz + PS_Z - 1) / PS_Z, (unsigned)pl->nsl);
    if (pl->mode == MODE_B3) {
      if (pl->cx) pack_x_small_kernel<true, MODE_B3><<<gs, PS_THREADS, 0, st>>>(xf, pl->BT[d], pl->invB, K, nz, pl->n, kpad);
      else pack_x_small_kernel<false, MODE_B3><<<gs, PS_THREADS, 0, st>>>(xf, pl->BT[d], pl->invB, K, nz, pl->n, kpad);
    } else {
      if (pl->cx) pack_x_small_kernel<true, MODE_H2><<<gs, PS_THREADS, 0, st>>>(xf, pl->BT[d], pl->invB, K, nz, pl->n, kpad);
      else pack_x_small_kernel<false, MODE_H2><<<gs, PS_THREADS, 0, st>>>(xf, pl->BT[d], pl->invB, K, nz, pl->n, kpad);
    }
    B2_LAUNCH_CHECK();
  } else if (parts & 1) {
    if (pl->mode == MODE_B3) {
      if (pl->cx) pack_x_kernel<true, MODE_B3><<<grid, PK_THREADS, 0, st>>>(xf, pl->BT[d], pl->invB, K, nz, pl->n, kpad);
      else pack_x_kernel<false, MODE_B3><<<grid, PK_THREADS, 0, st>>>(xf, pl->BT[d], pl->invB, K, nz, pl->n, kpad);
    } else {
      if (pl->cx) pack_x_kernel<true, MODE_H2><<<grid, PK_THREADS, 0, st>>>(xf, pl->BT[d], pl->invB, K, nz, pl->n, kpad);
      else pack_x_kernel<false, MODE_H2><<<grid, PK_THREADS, 0, st>>>(xf, pl->BT[d], pl->invB, K, nz, pl->n, kpad);
    }
    B2_LAUNCH_CHECK();
  }
  if (!(parts & 2)) return B2_OK;
  PeerOut po;
  po.n = npeers;
  for (int i = 0; i < 8; ++i) po.p[i] = i < npeers ? (float*)peers_host[i] : nullptr;
  if (pl->mode == MODE_B3)
    return pl->bk == 64 ? launch_product<MODE_B3, 64>(pl, d, (float*)y, po, st) : launch_product<MODE_B3, 32>(pl, d, (float*)y, po, st);
  return pl->bk == 64 ? launch_product<MODE_H2, 64>(pl, d, (float*)y, po, st) : launch_product<MODE_H2, 32>(pl, d, (float*)y, po, st);
}

extern "C" int b2_fredholm_apply(b2_fredholm_plan* pl, const void* x, void* y, void* const* peers_host, int npeers,
                                 int adjoint, void* stream) {
  return fredholm_apply_impl(pl, x, y, peers_host, npeers, adjoint, 3, stream);
}

// profiling aid: parts = 1 packs x only, 2 runs the product on the planes of the previous pack, 3 = both
extern "C" int b2_fredholm_apply_parts(b2_fredholm_plan* pl, const void* x, void* y, int adjoint, int parts, void* stream) {
  if (parts < 1 || parts > 3) return B2_ERR_ARG;
  return fredholm_apply_impl(pl, x, y, nullptr, 0, adjoint, parts, stream);
}
