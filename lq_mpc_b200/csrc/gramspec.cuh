// gramspec.cuh — extreme eigenvalues of the horizon-N Gram operators WITHOUT forming them (matrix-free, O(N n^3) per probe).
//
// energy_bound needs ||Gamma||_2 = sqrt(lambda_max(Gamma'Gamma)) and lambda_min(H^), H^ = barR + Gamma' barQ Gamma
// (reference utils.py:248-258: np.linalg.norm(Gamma, 2); utils.py:316-322: np.min(np.linalg.eigvals(hatH))) where
// Gamma (utils.py:145-174) is the block lower-triangular Toeplitz map from the stacked inputs u_0..u_{N-1} to the
// stacked states x_0..x_N of x+ = A x + B u, x_0 = 0. Both matrices are (N m) x (N m); round 1 assembled the Gram
// matrix and reduced it by Householder reflections ((N m)^3 flops, (N m)^2 doubles of scratch per sample).
//
// Gamma is a SIMULATION OPERATOR, so the quadratic form it induces is an N-stage LQ cost,
//     u' (barR - x I + Gamma' barQ Gamma) u  =  sum_t  x_t' Q x_t + u_t' (R - x I) u_t ,   x_{t+1} = A x_t + B u_t, x_0 = 0,
// and dynamic programming from the last stage eliminates u_{N-1}, ..., u_0 one block at a time:
//     G_s = (R - x I) + B' P B,     P <- Q + A' (P - P B G_s^-1 B' P) A,     P_0 = Q        (s = 1..N)
// This IS the block LDL' factorisation of the (N m) x (N m) matrix in that elimination order (Sylvester: its inertia is
// the sum of the inertias of the G_s). Hence
//     H^ - x I  positive definite   <=>   every G_s is positive definite              (lambda_min(H^) > x)
//     x I - Gamma'Gamma  pos. def.  <=>   every G'_s = x I - B' P B is pos. def. with P <- I + A'(P + P B G'_s^-1 B' P) A
// Each probe is a Cholesky factorisation in disguise: when every pivot is positive the computed factors are the exact
// factors of a matrix within a few ulps ||.|| of the probed one, so the yes/no answer is reliable down to the same
// backward error LAPACK's eigvals has on the assembled matrix. The two extremes are located by bisection on these
// one-sided predicates: ~52 probes x N stages x O(n^3) flops, all in registers, no scratch memory, one sample per
// thread with two independent recursions (the two searches) interleaved for instruction-level parallelism.
// General (non-scalar) weights in the time-major convention cost nothing extra: Q and R simply enter the recursion.
// (The literal kron(Q, I) ordering of utils.py:317-318 with non-scalar Q is NOT a stage-wise cost; it stays on the
// dense path of bounds.cuh.)
#pragma once
#include "riccati.cuh"

namespace lq {

// Square-root-free factorisation test of a small symmetric matrix: G = L D L' (unit lower L left below the diagonal),
// reciprocals of the pivots in dinv (MUFU.RCP64H + 2 Newton steps instead of the ~30-instruction FP64 rsqrt of a
// Cholesky: ncu showed 8 % of the kernel's instructions there). Returns "every pivot is positive" (a positive normal
// number: a denormal pivot counts as zero); a non-positive or NaN pivot is replaced by 1 so that nothing downstream overflows (the probe has failed anyway).
template <int m>
LQ_HD bool ldl_pos(double* G, double* dinv) {
  bool ok = true;
  double dv[m];
  LQ_UNROLL for (int j = 0; j < m; ++j) {
    double w[m];
    double d = G[j * m + j];
    LQ_UNROLL for (int k = 0; k < j; ++k) {
      w[k] = G[j * m + k] * dv[k];
      d = fma(-G[j * m + k], w[k], d);
    }
    const bool pos = is_pos_normal(d);           // integer pipe: the probes are FP64-pipe bound
    ok = ok && pos;
    d = pos ? d : 1.0;
    dv[j] = d;
    dinv[j] = rcp(d);
    LQ_UNROLL for (int i = j + 1; i < m; ++i) {
      double s = G[i * m + j];
      LQ_UNROLL for (int k = 0; k < j; ++k) s = fma(-G[i * m + k], w[k], s);
      G[i * m + j] = s * dinv[j];
    }
  }
  return ok;
}

// One elimination stage. G = Rd + sigma B'PB (Rd already carries the shift), positivity test, and unless `last`
// P <- Qw + A' (P - sigma (P B) G^-1 (P B)') A. Returns "G is positive definite".
template <int n, int m>
LQ_HD bool gs_stage(const double* Ah, const double* Bh, const double* Qw, const double* Rd, double sigma, bool last,
                    double* P) {
  double Y[n * m], L[m * m], Di[m];
  mm<n, n, m>(P, Bh, Y);
  LQ_UNROLL for (int i = 0; i < m; ++i)
    LQ_UNROLL for (int j = 0; j <= i; ++j) {
      double acc = 0.0;
      LQ_UNROLL for (int k = 0; k < n; ++k) acc = fma(Bh[k * m + i], Y[k * m + j], acc);
      L[i * m + j] = fma(sigma, acc, Rd[i * m + j]);
    }
  const bool ok = ldl_pos<m>(L, Di);
  if (last) return ok;
  // Z = P B L^-T (unit triangular: no divisions), then P - sigma Z D^-1 Z'
  LQ_UNROLL for (int i = 0; i < n; ++i)
    LQ_UNROLL for (int j = 1; j < m; ++j) {
      double s = Y[i * m + j];
      LQ_UNROLL for (int k = 0; k < j; ++k) s = fma(-Y[i * m + k], L[j * m + k], s);
      Y[i * m + j] = s;
    }
  double Mx[n * n], MA[n * n];
  LQ_UNROLL for (int i = 0; i < n; ++i) {
    double yd[m];
    LQ_UNROLL for (int k = 0; k < m; ++k) yd[k] = Y[i * m + k] * Di[k];
    LQ_UNROLL for (int j = i; j < n; ++j) {
      double acc = 0.0;
      LQ_UNROLL for (int k = 0; k < m; ++k) acc = fma(yd[k], Y[j * m + k], acc);
      const double v = fma(-sigma, acc, P[i * n + j]);
      Mx[i * n + j] = v;
      Mx[j * n + i] = v;
    }
  }
  mm<n, n, n>(Mx, Ah, MA);
  sym_add_mtm<n, n>(Qw, Ah, MA, P);
  return ok;
}

// n = 2 (the shipped example and every sweep of the reference): the congruence P -> A' P A acts on the three unique
// entries (p00, p01, p11) of a symmetric 2 x 2 matrix as ONE 3 x 3 map that depends on A^ only,
//     M3 = [ a^2, 2ac, c^2 ; ab, ad + bc, cd ; b^2, 2bd, d^2 ],   A^ = [a b; c d],
// formed once per sample: 9 FMAs per stage instead of the 14 of the two 2 x 2 products (the stage drops from ~35 to
// ~28 FP64 instructions at m = 1; the probes are FP64-pipe bound). Same elimination as gs_stage; P is read and written
// through its full 2 x 2 storage so that callers do not change.
LQ_HD void gs_congruence2(const double* Ah, double* M3) {
  const double a = Ah[0], b = Ah[1], c = Ah[2], d = Ah[3];
  M3[0] = a * a; M3[1] = 2.0 * a * c; M3[2] = c * c;
  M3[3] = a * b; M3[4] = fma(a, d, b * c); M3[5] = c * d;
  M3[6] = b * b; M3[7] = 2.0 * b * d; M3[8] = d * d;
}

template <int m, bool LAST>
LQ_HD bool gs_stage2(const double* M3, const double* Bh, const double* Qw, const double* Rd, double sigma, double* P) {
  constexpr int n = 2;
  double Y[n * m], L[m * m], Di[m];
  LQ_UNROLL for (int j = 0; j < m; ++j) {
    Y[j] = fma(P[1], Bh[m + j], P[0] * Bh[j]);
    Y[m + j] = fma(P[3], Bh[m + j], P[1] * Bh[j]);
  }
  LQ_UNROLL for (int i = 0; i < m; ++i)
    LQ_UNROLL for (int j = 0; j <= i; ++j) {
      const double acc = fma(Bh[m + i], Y[m + j], Bh[i] * Y[j]);
      L[i * m + j] = fma(sigma, acc, Rd[i * m + j]);
    }
  const bool ok = ldl_pos<m>(L, Di);
  if (LAST) return ok;                     // the last stage only decides its pivot (compile-time: the loop is peeled)
  LQ_UNROLL for (int i = 0; i < n; ++i)
    LQ_UNROLL for (int j = 1; j < m; ++j) {
      double sacc = Y[i * m + j];
      LQ_UNROLL for (int k = 0; k < j; ++k) sacc = fma(-Y[i * m + k], L[j * m + k], sacc);
      Y[i * m + j] = sacc;
    }
  double p0 = P[0], p1 = P[1], p2 = P[3];
  LQ_UNROLL for (int k = 0; k < m; ++k) {
    const double w0 = sigma * (Y[k] * Di[k]), w1 = sigma * (Y[m + k] * Di[k]);
    p0 = fma(-w0, Y[k], p0);
    p1 = fma(-w0, Y[m + k], p1);
    p2 = fma(-w1, Y[m + k], p2);
  }
  const double t0 = fma(M3[2], p2, fma(M3[1], p1, fma(M3[0], p0, Qw[0])));
  const double t1 = fma(M3[5], p2, fma(M3[4], p1, fma(M3[3], p0, Qw[1])));
  const double t2 = fma(M3[8], p2, fma(M3[7], p1, fma(M3[6], p0, Qw[3])));
  P[0] = t0; P[1] = t1; P[2] = t1; P[3] = t2;
  return ok;
}

// Relative width at which a bisection stops: 2^-42. The spectrum enters alpha / beta / J_bound, which are compared at
// 1e-9 (north_star); 2.3e-13 leaves three orders of margin and saves a fifth of the probes of a full-precision search.
constexpr double kGramSpectrumTol = 2.2737367544323206e-13;

struct GramSpectrum {
  double min_H;    // lambda_min(barR + Gamma' barQ Gamma), time-major weights (== R[0] + Q[0] lambda_min(Gamma'Gamma) for scalar weights)
  double cmax;     // lambda_max(Gamma'Gamma)
  int probes;      // bisection passes used (diagnostics)
};

// Both extremes for the estimated model (Ah, Bh) and horizon N. Q, R: stage weights (R positive definite, Q >= 0);
// minR = lambda_min(R).
template <int n, int m>
LQ_HD GramSpectrum gram_spectrum(const double* Ah, const double* Bh, const double* Q, const double* R, double minR,
                                 int N) {
  double eye[n * n];
  LQ_UNROLL for (int i = 0; i < n; ++i)
    LQ_UNROLL for (int j = 0; j < n; ++j) eye[i * n + j] = (i == j) ? 1.0 : 0.0;
  // ---- brackets. lambda_max(Gamma'Gamma): Rayleigh quotients of the unit inputs at stage 0 from below, the l1 bound
  //      on a Toeplitz operator norm (sum of the impulse-response norms) from above.
  double colsq[m], sumF = 0.0;
  LQ_UNROLL for (int j = 0; j < m; ++j) colsq[j] = 0.0;
  {
    double Gd[n * m];
    LQ_UNROLL for (int e = 0; e < n * m; ++e) Gd[e] = Bh[e];
    for (int d = 0; d < N; ++d) {
      double fro = 0.0;
      LQ_UNROLL for (int r = 0; r < n; ++r)
        LQ_UNROLL for (int j = 0; j < m; ++j) {
          const double g = Gd[r * m + j];
          colsq[j] = fma(g, g, colsq[j]);
          fro = fma(g, g, fro);
        }
      sumF += sqrt(fro);
      double Gn[n * m];
      mm<n, n, m>(Ah, Gd, Gn);
      LQ_UNROLL for (int e = 0; e < n * m; ++e) Gd[e] = Gn[e];
    }
  }
  double loC = 0.0;
  LQ_UNROLL for (int j = 0; j < m; ++j) loC = dmax(loC, colsq[j]);
  loC *= (1.0 - 1e-14);
  double hiC = sumF * sumF * (1.0 + 1e-14);
  // lambda_min(H^): H^ >= barR from below; the Rayleigh quotient of a unit input at the LAST stage (it only moves x_N)
  // from above: min_j (R + B'QB)_jj.
  double loH = minR * (1.0 - 1e-15), hiH = HUGE_VAL;
  {
    double QB[n * m];
    mm<n, n, m>(Q, Bh, QB);
    LQ_UNROLL for (int j = 0; j < m; ++j) {
      double acc = R[j * m + j];
      LQ_UNROLL for (int k = 0; k < n; ++k) acc = fma(Bh[k * m + j], QB[k * m + j], acc);
      hiH = dmin(hiH, acc);
    }
    hiH *= (1.0 + 1e-14);
  }
  GramSpectrum out;
  bool liveH = (hiH > loH), liveC = (hiC > loC) && (hiC == hiC) && (hiC < 1.7e308);
  if (!(hiC == hiC) || !(hiC < 1.7e308)) { loC = hiC = sumF * sumF; }      // overflow / NaN operands: propagate
  // Plain bisection on the two predicates. (Safeguarded regula falsi on the last pivot — a smooth monotone function
  // of the shift whose zero is the eigenvalue — was measured on the host harness: 14-32 probes per sample on average
  // instead of 43, but its slow tail, samples whose next pole sits within ~1e-6 of the root, puts the MAXIMUM over the
  // 32 samples of a warp at 33-48 probes: no gain in SIMT, so the simpler search stays.)
  double M3[9];
  if (n == 2) gs_congruence2(Ah, M3);
  int pass = 0;
  for (; pass < 128 && (liveH || liveC); ++pass) {
    const double xH = 0.5 * (loH + hiH), xC = 0.5 * (loC + hiC);
    double RdH[m * m], RdC[m * m], PH[n * n], PC[n * n];
    LQ_UNROLL for (int i = 0; i < m; ++i)
      LQ_UNROLL for (int j = 0; j < m; ++j) {
        RdH[i * m + j] = R[i * m + j] - ((i == j) ? xH : 0.0);
        RdC[i * m + j] = (i == j) ? xC : 0.0;
      }
    LQ_UNROLL for (int e = 0; e < n * n; ++e) { PH[e] = Q[e]; PC[e] = eye[e]; }
    bool okH = true, okC = true;
    if (n == 2) {                         // as below, with the 3 x 3 congruence map
      for (int s = 1; s < N; ++s) {
        okH = gs_stage2<m, false>(M3, Bh, Q, RdH, 1.0, PH) && okH;
        okC = gs_stage2<m, false>(M3, Bh, eye, RdC, -1.0, PC) && okC;
        if (!okH && !okC) break;
      }
      if (okH || okC) {                   // stage N (peeled): pivots only
        okH = gs_stage2<m, true>(M3, Bh, Q, RdH, 1.0, PH) && okH;
        okC = gs_stage2<m, true>(M3, Bh, eye, RdC, -1.0, PC) && okC;
      }
    } else if (n <= 4) {                  // two independent recursions interleaved: instruction-level parallelism
      for (int s = 1; s <= N; ++s) {
        const bool last = (s == N);
        okH = gs_stage<n, m>(Ah, Bh, Q, RdH, 1.0, last, PH) && okH;
        okC = gs_stage<n, m>(Ah, Bh, eye, RdC, -1.0, last, PC) && okC;
        if (!okH && !okC) break;          // both probes already decided
      }
    } else {                              // n >= 5: one recursion at a time (an n x n block is 2 n^2 registers)
      for (int s = 1; s <= N && okH; ++s) okH = gs_stage<n, m>(Ah, Bh, Q, RdH, 1.0, s == N, PH);
      for (int s = 1; s <= N && okC; ++s) okC = gs_stage<n, m>(Ah, Bh, eye, RdC, -1.0, s == N, PC);
    }
    if (liveH) {
      if (okH) loH = xH; else hiH = xH;
      const double mid = 0.5 * (loH + hiH);
      liveH = (hiH - loH > kGramSpectrumTol * hiH) && (mid > loH) && (mid < hiH);
    }
    if (liveC) {
      if (okC) hiC = xC; else loC = xC;
      const double mid = 0.5 * (loC + hiC);
      liveC = (hiC - loC > kGramSpectrumTol * hiC) && (mid > loC) && (mid < hiC);
    }
  }
  out.min_H = 0.5 * (loH + hiH);
  out.cmax = 0.5 * (loC + hiC);
  out.probes = pass;
  return out;
}

}  // namespace lq
