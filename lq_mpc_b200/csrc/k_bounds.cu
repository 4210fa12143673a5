// k_bounds.cu — K3: batched bound coefficients (alpha, beta, xi, eta, J_bound) and batched dlqr.
// One sample per thread. The extreme eigenvalues of the (N m) x (N m) Gram operators come from the matrix-free
// bisection of gramspec.cuh (registers only, no scratch); only the literal-kron ordering with non-scalar weights (and
// LQMPC_K3_DENSE=1) takes the dense Householder route with its [element][thread] workspace.
#include "bounds.cuh"
#include <stdlib.h>

#include "engine.h"

namespace {

__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// CTAs per SM the register allocation of the n m <= 2 instantiation is held to (1 = unconstrained; A/B switch).
// Measured on cfg-sweep-f (bounds phase of 5e7 evals): unconstrained (162 registers, 3 CTAs/SM) 0.314 s,
// 4 (128 registers, 96 B of stack) 0.293 s, 5 (96 registers, 200 B) 0.303 s.
#ifndef LQ_K3_MINB_SMALL
#define LQ_K3_MINB_SMALL 4
#endif
template <int n, int m>
__global__ void __launch_bounds__(128, (n * m <= 2) ? LQ_K3_MINB_SMALL : 1) bounds_kernel(const __grid_constant__ lq::Problem<n, m> pb,
                                                     const __grid_constant__ BoundsArgs a) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const lq::WsView ws{a.ws + tid, nthreads};
  for (int64_t s = tid; s < a.S; s += nthreads) {
    double Ah[n * n], Bh[n * m], K[m * n], x[n];
#pragma unroll
    for (int e = 0; e < n * n; ++e) Ah[e] = pb.A[e] + (a.dA ? ld_stream(a.dA + (int64_t)e * a.S + s) : 0.0);
#pragma unroll
    for (int e = 0; e < n * m; ++e) Bh[e] = pb.B[e] + (a.dB ? ld_stream(a.dB + (int64_t)e * a.S + s) : 0.0);
    int flags = 0;
    if (a.K_in) {
#pragma unroll
      for (int e = 0; e < m * n; ++e) K[e] = a.K_in[(int64_t)e * a.S + s];
    } else if (a.K_shared) {
#pragma unroll
      for (int e = 0; e < m * n; ++e) K[e] = a.K_shared[e];
    } else {
      double X[n * n];
      if (!lq::dare_sda<n, m>(Ah, Bh, pb.Q, pb.R, X)) flags |= lq::FLAG_DARE_NOCONV;
      lq::dlqr_gain<n, m>(Ah, Bh, pb.R, X, K);
#pragma unroll
      for (int e = 0; e < m * n; ++e) K[e] = -K[e];   // callers pass -K_dlqr (utils_class.py:843-844)
      if (a.P_out) {
#pragma unroll
        for (int e = 0; e < n * n; ++e) a.P_out[(int64_t)e * a.S + s] = X[e];
      }
    }
#pragma unroll
    for (int i = 0; i < n; ++i) x[i] = a.x_shared ? a.x_shared[i] : a.x[(int64_t)i * a.S + s];
    lq::BoundsScalars sc;
    sc.N = a.N;
    sc.e_A = a.eA ? a.eA[s] : a.eA_s;
    sc.e_B = a.eB ? a.eB[s] : a.eB_s;
    sc.M_V = a.MV ? a.MV[s] : a.MV_s;
    sc.p[0] = a.p[0]; sc.p[1] = a.p[1]; sc.p[2] = a.p[2];
    sc.V_expert = a.V_expert;
    sc.bar_u = a.bar_u; sc.bar_d_u = a.bar_d_u;
    sc.strict_reference = a.strict;
    sc.polyF = a.polyF; sc.polyP = a.polyP;
    sc.force_dense = a.dense;
    double out[lq::BF_COUNT];
    flags |= lq::bounds_sample<n, m>(pb, Ah, Bh, K, x, sc, ws, out);
    if (a.alpha) a.alpha[s] = out[lq::BF_ALPHA];
    if (a.beta) a.beta[s] = out[lq::BF_BETA];
    if (a.xi) a.xi[s] = out[lq::BF_XI];
    if (a.eta) a.eta[s] = out[lq::BF_ETA];
    if (a.bound) a.bound[s] = out[lq::BF_BOUND];
    if (a.detail) {
#pragma unroll
      for (int f = 0; f < lq::BF_COUNT; ++f) a.detail[(int64_t)f * a.S + s] = out[f];
    }
    if (a.K_out) {
#pragma unroll
      for (int e = 0; e < m * n; ++e) a.K_out[(int64_t)e * a.S + s] = K[e];
    }
    if (a.flags) a.flags[s] = flags;
  }
}

template <int n, int m>
__global__ void __launch_bounds__(128) dlqr_kernel(const __grid_constant__ lq::Problem<n, m> pb, int64_t S,
                                                   const double* dA, const double* dB, double* K_out, double* P_out,
                                                   int32_t* flags) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double Ah[n * n], Bh[n * m], K[m * n], X[n * n];
#pragma unroll
  for (int e = 0; e < n * n; ++e) Ah[e] = pb.A[e] + (dA ? ld_stream(dA + (int64_t)e * S + s) : 0.0);
#pragma unroll
  for (int e = 0; e < n * m; ++e) Bh[e] = pb.B[e] + (dB ? ld_stream(dB + (int64_t)e * S + s) : 0.0);
  const bool ok = lq::dare_sda<n, m>(Ah, Bh, pb.Q, pb.R, X);
  lq::dlqr_gain<n, m>(Ah, Bh, pb.R, X, K);
  if (K_out) {
#pragma unroll
    for (int e = 0; e < m * n; ++e) K_out[(int64_t)e * S + s] = K[e];
  }
  if (P_out) {
#pragma unroll
    for (int e = 0; e < n * n; ++e) P_out[(int64_t)e * S + s] = X[e];
  }
  if (flags) flags[s] = ok ? 0 : lq::FLAG_DARE_NOCONV;
}

template <int n, int m>
int launch_bounds_t(lqmpc_ctx* ctx, BoundsArgs a) {
  const lq::Problem<n, m>& pb = *reinterpret_cast<const lq::Problem<n, m>*>(ctx->pb);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  const int threads = 128;
  int64_t blocks = (a.S + threads - 1) / threads;
  a.dense = getenv("LQMPC_K3_DENSE") != nullptr ? 1 : 0;
  a.ws = nullptr;
  if (!lq::bounds_matrix_free(pb.qr_scalar, a.strict, a.dense)) {
    // dense route: (N m)^2 doubles of scratch per thread — cap the grid and stride over the batch
    const int64_t cap = (int64_t)sms * 8;
    if (blocks > cap) blocks = cap;
    const int64_t per = lq::bounds_ws_doubles<n, m>(a.N);
    int rc = lq_reserve_ws(ctx, (size_t)(per * blocks * threads) * sizeof(double));
    if (rc) return rc;
    a.ws = reinterpret_cast<double*>(ctx->ws);
  } else if (blocks > 0x7fffffffLL) {
    return lq_set_error(ctx, LQMPC_EINVAL, "batch too large for one launch");
  }
  bounds_kernel<n, m><<<(unsigned)blocks, threads, 0, ctx->stream>>>(pb, a);
  ctx->launches++;
  return lq_check_cuda(ctx, cudaGetLastError(), "bounds_kernel launch");
}

template <int n, int m>
int launch_dlqr_t(lqmpc_ctx* ctx, int64_t S, const double* dA, const double* dB, double* K, double* P,
                  int32_t* flags) {
  const lq::Problem<n, m>& pb = *reinterpret_cast<const lq::Problem<n, m>*>(ctx->pb);
  const int threads = 128;
  const int64_t blocks = (S + threads - 1) / threads;
  dlqr_kernel<n, m><<<(unsigned)blocks, threads, 0, ctx->stream>>>(pb, S, dA, dB, K, P, flags);
  ctx->launches++;
  return lq_check_cuda(ctx, cudaGetLastError(), "dlqr_kernel launch");
}

}  // namespace

int lq_launch_bounds(lqmpc_ctx* ctx, const BoundsArgs& a) {
#define X(N_, M_) \
  if (ctx->n == N_ && ctx->m == M_) return launch_bounds_t<N_, M_>(ctx, a);
  LQ_FOR_EACH_DIM(X)
#undef X
  return lq_set_error(ctx, LQMPC_EINVAL, "unsupported (n, m); see lqmpc_supported_dims()");
}

int lq_launch_dlqr(lqmpc_ctx* ctx, int64_t S, const double* dA, const double* dB, double* K, double* P,
                   int32_t* flags) {
#define X(N_, M_) \
  if (ctx->n == N_ && ctx->m == M_) return launch_dlqr_t<N_, M_>(ctx, S, dA, dB, K, P, flags);
  LQ_FOR_EACH_DIM(X)
#undef X
  return lq_set_error(ctx, LQMPC_EINVAL, "unsupported (n, m); see lqmpc_supported_dims()");
}
