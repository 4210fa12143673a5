// k_sampler.cu — K6: model-error grids generated directly in HBM (sampler.cuh), in the reference file's own layout:
// out[(a*c + b) * (N_sys*n_err) + j*n_err + i] == error_X[a, b, j, i] of utils.py:826-847 (C order), which is also
// the engine's SoA operand layout — the sweep consumes the buffer without a transpose or a host round trip.
#include "engine.h"
#include "sampler.cuh"

namespace {

struct SamplerArgs {
  uint64_t seed;
  int which;
  int64_t N_sys, j_first, n_boundary;
  int n_err, norm_type;
  const double* levels;     // device [n_err]
  double* out;              // device [r*c][N_sys*n_err]
  unsigned long long* stats;  // device [2]: rejected attempts, projected samples
};

template <int r, int c>
__global__ void __launch_bounds__(256) sampler_kernel(const SamplerArgs a) {
  const int64_t S = a.N_sys * a.n_err;
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long rej = 0, proj = 0;
  if (s < S) {
    const int64_t jl = s / a.n_err;
    const int i = (int)(s % a.n_err);
    const int64_t j = a.j_first + jl;
    double T[r * c];
    const int rc = lq::sample_error_matrix<r, c>(a.seed, a.which, j, i, a.levels[i], j < a.n_boundary, a.norm_type, T);
    if (rc < 0) { proj = 1; rej = (unsigned long long)(-rc); } else rej = (unsigned long long)rc;
#pragma unroll
    for (int q = 0; q < r * c; ++q) a.out[(int64_t)q * S + s] = T[q];
  }
  // block-level tally (diagnostics only)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    rej += __shfl_xor_sync(0xffffffffu, rej, o);
    proj += __shfl_xor_sync(0xffffffffu, proj, o);
  }
  if ((threadIdx.x & 31) == 0 && (rej | proj)) {
    atomicAdd(a.stats, rej);
    atomicAdd(a.stats + 1, proj);
  }
}

template <int r, int c>
int launch_sampler_t(lqmpc_ctx* ctx, const SamplerArgs& a) {
  const int64_t S = a.N_sys * a.n_err;
  const int64_t blocks = (S + 255) / 256;
  if (blocks > 0x7fffffffLL) return lq_set_error(ctx, LQMPC_EINVAL, "grid too large for one launch");
  sampler_kernel<r, c><<<(unsigned)blocks, 256, 0, ctx->stream>>>(a);
  ctx->launches++;
  return lq_check_cuda(ctx, cudaGetLastError(), "sampler_kernel launch");
}

}  // namespace

int lq_launch_sampler(lqmpc_ctx* ctx, uint64_t seed, int which, int rows, int cols, int64_t N_sys, int64_t j_first,
                      int n_err, const double* levels_host, int64_t n_boundary, int norm_type, double* out,
                      int64_t* stats_host) {
  int rc = lq_reserve_ws(ctx, (size_t)n_err * sizeof(double) + 64);
  if (rc) return rc;
  double* lev = reinterpret_cast<double*>(ctx->ws);
  unsigned long long* st = reinterpret_cast<unsigned long long*>(lev + n_err + (n_err & 1));
  rc = lq_check_cuda(ctx, cudaMemcpyAsync(lev, levels_host, (size_t)n_err * sizeof(double), cudaMemcpyHostToDevice,
                                          ctx->stream), "H2D levels");
  if (rc) return rc;
  cudaMemsetAsync(st, 0, 16, ctx->stream);
  SamplerArgs a{seed, which, N_sys, j_first, n_boundary, n_err, norm_type, lev, out, st};
  rc = -100;
#define X(N_, M_)                                                                    \
  if (rc == -100 && rows == N_ && cols == N_) rc = launch_sampler_t<N_, N_>(ctx, a); \
  if (rc == -100 && rows == N_ && cols == M_) rc = launch_sampler_t<N_, M_>(ctx, a);
  LQ_FOR_EACH_DIM(X)
#undef X
  if (rc == -100) return lq_set_error(ctx, LQMPC_EINVAL, "unsupported sampler shape (rows x cols must be n x n or n x m of a compiled pair)");
  if (rc) return rc;
  if (stats_host) {
    unsigned long long h[2] = {0, 0};
    rc = lq_check_cuda(ctx, cudaMemcpyAsync(h, st, 16, cudaMemcpyDeviceToHost, ctx->stream), "D2H sampler stats");
    if (rc) return rc;
    rc = lq_check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "sampler sync");
    if (rc) return rc;
    stats_host[0] = (int64_t)h[0];
    stats_host[1] = (int64_t)h[1];
  }
  return 0;
}
