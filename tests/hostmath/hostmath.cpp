// TEST FIXTURE ONLY — compiles the engine's __host__ __device__ math templates with g++ so that the per-sample
// numerics can be compared with the oracle inside the GPU-less build container.  It is built and loaded by
// tests/ only; nothing in lq_mpc_b200/ links or loads it (the product has no CPU path).
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../lq_mpc_b200/csrc/bounds.cuh"
#include "../../lq_mpc_b200/csrc/pclqr.cuh"
#include "../../lq_mpc_b200/csrc/sampler.cuh"

#define HM_FOR_EACH_DIM(X) X(1, 1) X(2, 1) X(2, 2) X(3, 1) X(3, 2) X(3, 3) X(4, 1) X(4, 2) X(4, 4) X(6, 2) X(8, 2)

namespace {

lq::Refs g_refs;   // set by hm_set_refs (test fixture state)
lq::Poly g_poly;   // set by hm_set_poly: general input polytope F_u u <= 1 (p = 0: the box of lo/hi)

template <int n, int m>
void fill_problem(lq::Problem<n, m>& pb, const double* A, const double* B, const double* Q, const double* R,
                  const double* P, const double* lo, const double* hi, int N_opc) {
  memset(&pb, 0, sizeof(pb));
  memcpy(pb.A, A, sizeof(pb.A));
  memcpy(pb.B, B, sizeof(pb.B));
  memcpy(pb.Q, Q, sizeof(pb.Q));
  memcpy(pb.R, R, sizeof(pb.R));
  memcpy(pb.Pt, P, sizeof(pb.Pt));
  pb.has_bounds = (lo || hi) ? 1 : 0;
  for (int j = 0; j < m; ++j) {
    pb.ulo[j] = lo ? lo[j] : -HUGE_VAL;
    pb.uhi[j] = hi ? hi[j] : HUGE_VAL;
  }
  lq::prepare_problem<n, m>(pb, N_opc);
}

struct HostSink {
  int64_t S, s;
  int mn;
  double *J, *rho, *ratio, *Vn, *JT, *K0;
  int32_t* flags;
  void operator()(int h, double j, double r, double ra, double vn, double jt, int fl, const double* K) const {
    const int64_t o = (int64_t)h * S + s;
    if (J) J[o] = j;
    if (rho) rho[o] = r;
    if (ratio) ratio[o] = ra;
    if (Vn) Vn[o] = vn;
    if (JT) JT[o] = jt;
    if (flags) flags[o] = fl;
    if (K0)
      for (int e = 0; e < mn; ++e) K0[((int64_t)h * mn + e) * S + s] = K[e];
  }
};

template <int n, int m>
int eval_t(const double* A, const double* B, const double* Q, const double* R, const double* P, int N_opc,
           int64_t S, const double* dA, const double* dB, const double* x0, int Nmin, int Nmax, int T, double* J,
           double* rho, double* ratio, double* Vn, double* JT, int32_t* flags, double* K0, double* prepared) {
  lq::Problem<n, m> pb;
  fill_problem<n, m>(pb, A, B, Q, R, P, nullptr, nullptr, N_opc);
  if (prepared) {
    memcpy(prepared, pb.Pexp, sizeof(pb.Pexp));
    memcpy(prepared + n * n, pb.Qinv, sizeof(pb.Qinv));
    prepared[2 * n * n + 0] = pb.maxQ; prepared[2 * n * n + 1] = pb.minQ;
    prepared[2 * n * n + 2] = pb.maxR; prepared[2 * n * n + 3] = pb.minR;
  }
  for (int64_t s = 0; s < S; ++s) {
    double a[n * n], b[n * m], x[n];
    for (int e = 0; e < n * n; ++e) a[e] = dA[e * S + s];
    for (int e = 0; e < n * m; ++e) b[e] = dB[e * S + s];
    for (int e = 0; e < n; ++e) x[e] = x0[e * S + s];
    HostSink sink{S, s, m * n, J, rho, ratio, Vn, T > 0 ? JT : nullptr, K0, flags};
    lq::eval_sample<n, m>(pb, a, b, x, Nmin, Nmax, T, Vn != nullptr, sink);
  }
  return 0;
}

struct HostTraj {
  int n, m;
  int64_t S, s;
  double *X, *U;
  void state(int t, const double* x) const {
    if (X) for (int i = 0; i < n; ++i) X[((int64_t)t * n + i) * S + s] = x[i];
  }
  void input(int t, const double* u) const {
    if (U) for (int j = 0; j < m; ++j) U[((int64_t)t * m + j) * S + s] = u[j];
  }
};

// mode 0: open-loop solves over `npts` shared points (pts != NULL) or the per-sample x0; mode 1: closed-loop simulate
template <int n, int m>
int mpc_t(int mode, const double* A, const double* B, const double* Q, const double* R, const double* P,
          const double* lo, const double* hi, int64_t S, const double* dA, const double* dB, int N, int T, int npts,
          const double* pts, const double* x0, double* V, double* u0, double* M_V, double* J_T, double* X, double* U,
          int32_t* flags, int32_t* n_active) {
  lq::Problem<n, m> pb;
  fill_problem<n, m>(pb, A, B, Q, R, P, lo, hi, 1);
  std::vector<double> wsbuf((size_t)lq::clqr_ws_doubles<n, m>(N));
  const lq::WsView ws{wsbuf.data(), 1};
  for (int64_t s = 0; s < S; ++s) {
    lq::Plan<n, m> pl;
    for (int e = 0; e < n * n; ++e) pl.Ah[e] = pb.A[e] + (dA ? dA[e * S + s] : 0.0);
    for (int e = 0; e < n * m; ++e) pl.Bh[e] = pb.B[e] + (dB ? dB[e * S + s] : 0.0);
    const int pf = lq::plan_prepare<n, m>(pb, pl, N, ws);
    if (mode == 0) {
      const int Pn = pts ? npts : 1;
      double mvv = -HUGE_VAL;
      for (int p = 0; p < Pn; ++p) {
        double xx[n], uu[m], v;
        for (int i = 0; i < n; ++i) xx[i] = pts ? pts[p * n + i] : x0[i * S + s];
        const int f = pf | (g_poly.p > 0 ? lq::pclqr_solve<n, m>(pb, pl, N, xx, g_poly, ws, uu, &v, g_refs)
                                         : lq::clqr_solve<n, m>(pb, pl, N, xx, ws, uu, &v, g_refs));
        if (v > mvv) mvv = v;
        if (V) V[(int64_t)p * S + s] = v;
        if (u0) for (int j = 0; j < m; ++j) u0[((int64_t)p * m + j) * S + s] = uu[j];
        if (flags) flags[(int64_t)p * S + s] = f;
      }
      if (M_V) M_V[s] = mvv;
    } else {
      double xx[n], jt;
      int act;
      for (int i = 0; i < n; ++i) xx[i] = pts ? pts[i] : x0[i * S + s];
      HostTraj traj{n, m, S, s, X, U};
      const int f = pf | (g_poly.p > 0
                              ? lq::psimulate_sample<n, m>(pb, pl, N, T, xx, g_poly, ws, &jt, &act, traj, g_refs)
                              : lq::simulate_sample<n, m>(pb, pl, N, T, xx, ws, &jt, &act, traj, g_refs));
      if (J_T) J_T[s] = jt;
      if (flags) flags[s] = f;
      if (n_active) n_active[s] = act;
    }
  }
  return 0;
}

// bounds for S samples (SoA operands); K_in may be NULL (-> -K_dlqr by SDA). scal: N,eA,eB,MV (per sample arrays)
template <int n, int m>
int bounds_t(const double* A, const double* B, const double* Q, const double* R, const double* P, const double* lo,
             const double* hi, int64_t S, const double* dA, const double* dB, int N, const double* eA,
             const double* eB, const double* MV, const double* x, const double* K_in, const double* p3,
             double V_expert, int strict, double* detail, double* K_out, double* P_out, int32_t* flags) {
  lq::Problem<n, m> pb;
  fill_problem<n, m>(pb, A, B, Q, R, P, lo, hi, 1);
  std::vector<double> wsbuf((size_t)lq::bounds_ws_doubles<n, m>(N));
  const lq::WsView ws{wsbuf.data(), 1};
  for (int64_t s = 0; s < S; ++s) {
    double Ah[n * n], Bh[n * m], K[m * n], xx[n], X[n * n];
    for (int e = 0; e < n * n; ++e) Ah[e] = pb.A[e] + (dA ? dA[e * S + s] : 0.0);
    for (int e = 0; e < n * m; ++e) Bh[e] = pb.B[e] + (dB ? dB[e * S + s] : 0.0);
    int fl = 0;
    if (K_in) {
      for (int e = 0; e < m * n; ++e) K[e] = K_in[e * S + s];
    } else {
      if (!lq::dare_sda<n, m>(Ah, Bh, pb.Q, pb.R, X)) fl |= lq::FLAG_DARE_NOCONV;
      lq::dlqr_gain<n, m>(Ah, Bh, pb.R, X, K);
      for (int e = 0; e < m * n; ++e) K[e] = -K[e];
      if (P_out) for (int e = 0; e < n * n; ++e) P_out[e * S + s] = X[e];
    }
    for (int i = 0; i < n; ++i) xx[i] = x[i * S + s];
    lq::BoundsScalars sc;
    sc.N = N; sc.e_A = eA[s]; sc.e_B = eB[s]; sc.M_V = MV[s];
    sc.p[0] = p3[0]; sc.p[1] = p3[1]; sc.p[2] = p3[2];
    sc.V_expert = V_expert; sc.bar_u = -1.0; sc.bar_d_u = -1.0; sc.strict_reference = strict & 1;
    sc.force_dense = (strict >> 1) & 1;            // bit 1 of `strict`: dense Householder route (A/B tests)
    double out[lq::BF_COUNT];
    fl |= lq::bounds_sample<n, m>(pb, Ah, Bh, K, xx, sc, ws, out);
    for (int f = 0; f < lq::BF_COUNT; ++f) detail[(int64_t)f * S + s] = out[f];
    if (K_out) for (int e = 0; e < m * n; ++e) K_out[e * S + s] = K[e];
    flags[s] = fl;
  }
  return 0;
}

template <int n>
int rho_t(int64_t S, const double* M, double* rho, int32_t* okf) {
  for (int64_t s = 0; s < S; ++s) {
    bool ok;
    rho[s] = lq::spectral_radius<n>(M + s * n * n, &ok);
    okf[s] = ok ? 1 : 0;
  }
  return 0;
}

// fast (closed-form) path alone: rho and whether it was trusted (1) or would fall back to QR (0)
template <int n>
int rho_poly_t(int64_t S, const double* M, double* rho, int32_t* trusted) {
  for (int64_t s = 0; s < S; ++s) {
    double r = 0.0;
    trusted[s] = lq::RhoFast<n>::run(M + s * n * n, &r) ? 1 : 0;
    rho[s] = r;
  }
  return 0;
}

template <int n>
int symeig_t(int64_t S, const double* M, double* lo, double* hi) {
  for (int64_t s = 0; s < S; ++s) lq::sym_eig_minmax<n>(M + s * n * n, lo + s, hi + s);
  return 0;
}

}  // namespace

extern "C" {

int hm_eval(int n, int m, const double* A, const double* B, const double* Q, const double* R, const double* P,
            int N_opc, int64_t S, const double* dA, const double* dB, const double* x0, int Nmin, int Nmax, int T,
            double* J, double* rho, double* ratio, double* Vn, double* JT, int32_t* flags, double* K0,
            double* prepared) {
#define X(N_, M_) \
  if (n == N_ && m == M_) \
    return eval_t<N_, M_>(A, B, Q, R, P, N_opc, S, dA, dB, x0, Nmin, Nmax, T, J, rho, ratio, Vn, JT, flags, K0, prepared);
  HM_FOR_EACH_DIM(X)
#undef X
  return -1;
}

int hm_mpc(int mode, int n, int m, const double* A, const double* B, const double* Q, const double* R, const double* P,
           const double* lo, const double* hi, int64_t S, const double* dA, const double* dB, int N, int T, int npts,
           const double* pts, const double* x0, double* V, double* u0, double* M_V, double* J_T, double* X, double* U,
           int32_t* flags, int32_t* n_active) {
#define X_(N_, M_) \
  if (n == N_ && m == M_) \
    return mpc_t<N_, M_>(mode, A, B, Q, R, P, lo, hi, S, dA, dB, N, T, npts, pts, x0, V, u0, M_V, J_T, X, U, flags, n_active);
  HM_FOR_EACH_DIM(X_)
#undef X_
  return -1;
}

// references shared by every following hm_mpc call: x_ref (n x ld), u_ref (m x ld) row-major; NULLs clear them
void hm_set_refs(const double* x_ref, const double* u_ref, int ld) {
  g_refs.xr = x_ref;
  g_refs.ur = u_ref;
  g_refs.ld = ld;
}

// general input polytope for every following hm_mpc call: F (p x m) row-major; p = 0 clears it
void hm_set_poly(const double* F, int p) {
  g_poly.F = F;
  g_poly.p = p;
}

// extremes of a symmetric tridiagonal (d[k], e[k] with e[0] unused) by K3's Sturm search
void hm_tridiag_extremes(const double* d, const double* e, int k, double* out2) {
  std::vector<double> buf((size_t)2 * k);
  for (int i = 0; i < k; ++i) { buf[i] = d[i]; buf[k + i] = e[i]; }
  const lq::WsView ws{buf.data(), 1};
  lq::ws_tridiag_extremes(ws, 0, k, k, out2, out2 + 1);
}

int hm_bounds_fields(void) { return (int)lq::BF_COUNT; }

int hm_bounds(int n, int m, const double* A, const double* B, const double* Q, const double* R, const double* P,
              const double* lo, const double* hi, int64_t S, const double* dA, const double* dB, int N,
              const double* eA, const double* eB, const double* MV, const double* x, const double* K_in,
              const double* p3, double V_expert, int strict, double* detail, double* K_out, double* P_out,
              int32_t* flags) {
#define X_(N_, M_) \
  if (n == N_ && m == M_) \
    return bounds_t<N_, M_>(A, B, Q, R, P, lo, hi, S, dA, dB, N, eA, eB, MV, x, K_in, p3, V_expert, strict, detail, K_out, P_out, flags);
  HM_FOR_EACH_DIM(X_)
#undef X_
  return -1;
}

// matrices packed AoS here: M[s][n*n]
int hm_spectral_radius(int n, int64_t S, const double* M, double* rho, int32_t* ok) {
  switch (n) {
    case 1: return rho_t<1>(S, M, rho, ok);
    case 2: return rho_t<2>(S, M, rho, ok);
    case 3: return rho_t<3>(S, M, rho, ok);
    case 4: return rho_t<4>(S, M, rho, ok);
    case 5: return rho_t<5>(S, M, rho, ok);
    case 6: return rho_t<6>(S, M, rho, ok);
    case 8: return rho_t<8>(S, M, rho, ok);
  }
  return -1;
}

int hm_spectral_radius_poly(int n, int64_t S, const double* M, double* rho, int32_t* trusted) {
  switch (n) {
    case 3: return rho_poly_t<3>(S, M, rho, trusted);
    case 4: return rho_poly_t<4>(S, M, rho, trusted);
  }
  return -1;
}

void hm_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  const lq::Philox4 x = lq::philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
  for (int i = 0; i < 4; ++i) out[i] = x.v[i];
}

// seeded synthetic evaluation samples (lqmpc_eval_seeded's generator), AoS output: dA [S][n*n], dB [S][n*m], x0 [S][n]
int hm_seeded_samples(int n, int m, uint64_t seed, int64_t first, int64_t S, double e_A, double e_B, double* dA,
                      double* dB, double* x0) {
  for (int64_t s = 0; s < S; ++s) {
    double* a = dA + s * n * n; double* b = dB + s * n * m; double* x = x0 + s * n;
    if (n == 4 && m == 2) lq::seeded_sample<4, 2>(seed, first + s, e_A, e_B, a, b, x);
    else if (n == 2 && m == 1) lq::seeded_sample<2, 1>(seed, first + s, e_A, e_B, a, b, x);
    else if (n == 3 && m == 3) lq::seeded_sample<3, 3>(seed, first + s, e_A, e_B, a, b, x);
    else if (n == 1 && m == 1) lq::seeded_sample<1, 1>(seed, first + s, e_A, e_B, a, b, x);
    else return -1;
  }
  return 0;
}

// the sampler core for the shapes of the reference example (2x2, 2x1) and the synthetic one (4x4, 4x2); SoA output
int hm_sample_error_grid(uint64_t seed, int which, int rows, int cols, int64_t N_sys, int64_t j_first, int n_err,
                         const double* levels, int64_t n_boundary, int norm_type, double* out, int64_t* stats) {
  const int64_t S = N_sys * n_err;
  stats[0] = stats[1] = 0;
  for (int64_t s = 0; s < S; ++s) {
    const int64_t j = j_first + s / n_err;
    const int i = (int)(s % n_err);
    double T[16];
    int rc;
    if (rows == 2 && cols == 2) rc = lq::sample_error_matrix<2, 2>(seed, which, j, i, levels[i], j < n_boundary, norm_type, T);
    else if (rows == 2 && cols == 1) rc = lq::sample_error_matrix<2, 1>(seed, which, j, i, levels[i], j < n_boundary, norm_type, T);
    else if (rows == 4 && cols == 4) rc = lq::sample_error_matrix<4, 4>(seed, which, j, i, levels[i], j < n_boundary, norm_type, T);
    else if (rows == 4 && cols == 2) rc = lq::sample_error_matrix<4, 2>(seed, which, j, i, levels[i], j < n_boundary, norm_type, T);
    else return -1;
    if (rc < 0) { stats[1] += 1; stats[0] += -rc; } else stats[0] += rc;
    for (int q = 0; q < rows * cols; ++q) out[(int64_t)q * S + s] = T[q];
  }
  return 0;
}

int hm_sym_eig_minmax(int n, int64_t S, const double* M, double* lo, double* hi) {
  switch (n) {
    case 1: return symeig_t<1>(S, M, lo, hi);
    case 2: return symeig_t<2>(S, M, lo, hi);
    case 3: return symeig_t<3>(S, M, lo, hi);
    case 4: return symeig_t<4>(S, M, lo, hi);
    case 6: return symeig_t<6>(S, M, lo, hi);
    case 8: return symeig_t<8>(S, M, lo, hi);
  }
  return -1;
}

}  // extern "C"
