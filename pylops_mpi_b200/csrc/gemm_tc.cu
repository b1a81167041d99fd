// bf16 x bf16 -> fp32 tile product on the 5th-gen tensor cores (tcgen05 + TMEM,
// operands staged by TMA).  Placeholder translation unit until the kernel lands:
// the entry point exists so the ABI is stable, and reports UNSUPPORTED loudly.
#include "common.cuh"

extern "C" int b2_gemm_bf16(b2_ctx* ctx, const void* A, size_t lda, const void* B, size_t ldb,
                            float* C, size_t ldc, size_t m, size_t n, size_t k, int op_a,
                            int accumulate, void* stream) {
  (void)ctx; (void)A; (void)lda; (void)B; (void)ldb; (void)C; (void)ldc; (void)m; (void)n;
  (void)k; (void)op_a; (void)accumulate; (void)stream;
  return B2_ERR_UNSUPPORTED;
}
