// Host-buffer entry point for MPIFirstDerivative: the "plugin call with host
// arrays" (what a NumPy caller of FirstDerivative.py:129-133 hands over after
// DistributedArray.to_dist, which keeps the GLOBAL array replicated in host
// memory on every rank, DistributedArray.py:440-459).  This rank processes rows
// [row_begin, row_end) of the global [nrows_global x ncols] host array; its halo
// rows are read straight from the replicated host array, so no collective is
// needed.  Rows are cut into chunks; each chunk is uploaded (with its stencil
// overlap rows), differentiated by the same kernel the device path uses, and
// downloaded, on three rotating streams so H2D, kernel and D2H of neighbouring
// chunks overlap (PCIe is full duplex).  Pinned host memory gives true async
// copies; pageable memory still works (the runtime stages it).
#include "common.cuh"

extern "C" int b2_first_derivative_host(b2_ctx* ctx, const void* x_host, void* y_host,
                                        size_t nrows_global, size_t ncols, size_t row_begin,
                                        size_t row_end, int kind, int order, int edge,
                                        double sampling, int adjoint, int dtype) {
  if (!ctx) return B2_ERR_ARG;
  if (row_end > nrows_global || row_begin > row_end) return B2_ERR_ARG;
  if (row_begin == row_end || ncols == 0) return B2_OK;
  if (!x_host || !y_host) return B2_ERR_ARG;
  const size_t es = b2_dtype_size(dtype);
  if (dtype != B2_F32 && dtype != B2_F64) return B2_ERR_DTYPE;
  int need_lo, need_hi;
  int rc = b2_first_derivative_halo(kind, order, adjoint, &need_lo, &need_hi);
  if (rc) return rc;
  B2_CUDA(cudaSetDevice(ctx->device));
  const size_t row_bytes = ncols * es;
  const size_t nrows = row_end - row_begin;
  // ~32 MiB chunks, at least 64 rows, at most the whole block
  size_t rows_per = (32u << 20) / row_bytes;
  if (rows_per < 64) rows_per = 64;
  if (rows_per > nrows) rows_per = nrows;
  const size_t in_bytes = (rows_per + 4) * row_bytes, out_bytes = rows_per * row_bytes;
  const size_t need = in_bytes > out_bytes ? in_bytes : out_bytes;
  if (ctx->pipe_bytes < need) {
    for (int s = 0; s < 3; ++s)
      for (int k = 0; k < 2; ++k) {
        if (ctx->pipe_buf[s][k]) B2_CUDA(cudaFree(ctx->pipe_buf[s][k]));
        ctx->pipe_buf[s][k] = nullptr;
      }
    ctx->pipe_bytes = 0;
    for (int s = 0; s < 3; ++s)
      for (int k = 0; k < 2; ++k) B2_CUDA(cudaMalloc(&ctx->pipe_buf[s][k], need));
    ctx->pipe_bytes = need;
  }
  for (int s = 0; s < 3; ++s)
    if (!ctx->pipe_stream[s]) B2_CUDA(cudaStreamCreateWithFlags(&ctx->pipe_stream[s], cudaStreamNonBlocking));

  size_t chunk = 0;
  for (size_t r0 = row_begin; r0 < row_end; r0 += rows_per, ++chunk) {
    const int s = (int)(chunk % 3);
    cudaStream_t st = ctx->pipe_stream[s];
    const size_t r1 = r0 + rows_per < row_end ? r0 + rows_per : row_end;
    const size_t lo = r0 < (size_t)need_lo ? r0 : (size_t)need_lo;
    const size_t hi = nrows_global - r1 < (size_t)need_hi ? nrows_global - r1 : (size_t)need_hi;
    char* din = (char*)ctx->pipe_buf[s][0];
    char* dout = (char*)ctx->pipe_buf[s][1];
    // stream order on slot s already serialises reuse of its buffers
    B2_CUDA(cudaMemcpyAsync(din, (const char*)x_host + (r0 - lo) * row_bytes,
                            (r1 - r0 + lo + hi) * row_bytes, cudaMemcpyHostToDevice, st));
    rc = b2_first_derivative(ctx, din + lo * row_bytes, dout, lo ? din : nullptr, (int)lo,
                             hi ? din + (lo + (r1 - r0)) * row_bytes : nullptr, (int)hi, r1 - r0,
                             ncols, r0, nrows_global, kind, order, edge, sampling, adjoint, dtype, st);
    if (rc) return rc;
    B2_CUDA(cudaMemcpyAsync((char*)y_host + r0 * row_bytes, dout, (r1 - r0) * row_bytes,
                            cudaMemcpyDeviceToHost, st));
  }
  for (int s = 0; s < 3; ++s) B2_CUDA(cudaStreamSynchronize(ctx->pipe_stream[s]));
  return B2_OK;
}
