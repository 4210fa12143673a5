// eig.cuh — spectral radius of a small general (non-symmetric) real matrix, one matrix per thread.
//
// Replaces `np.max(np.abs(np.linalg.eigvals(A_cl)))` (reference utils.py:358 and the stability check of the
// closed loop).  Algorithm: Householder reduction to upper Hessenberg form followed by the Francis implicit
// double-shift QR iteration with deflation (the classical EISPACK "hqr" scheme), written here for tiny fixed n
// with every loop bound a compile-time constant and run-time *predicates* selecting the active window, so that
// after unrolling all indices are static and the matrix stays in registers on the device.
#pragma once
#include "small_la.cuh"

namespace lq {

// Householder reduction to upper Hessenberg form (similarity transform; eigenvalues preserved).
template <int n>
LQ_HD void hessenberg(double* a) {
  LQ_UNROLL for (int k = 0; k < n - 2; ++k) {
    double alpha = 0.0;
    LQ_UNROLL for (int i = k + 1; i < n; ++i) alpha = fma(a[i * n + k], a[i * n + k], alpha);
    double tail = alpha - a[(k + 1) * n + k] * a[(k + 1) * n + k];
    if (tail > 0.0) {  // something below the sub-diagonal to annihilate
      const double x0 = a[(k + 1) * n + k];
      const double nrm = sqrt(alpha);
      const double beta = (x0 >= 0.0) ? -nrm : nrm;
      double v[n];
      LQ_UNROLL for (int i = 0; i < n; ++i) v[i] = (i > k) ? a[i * n + k] : 0.0;
      v[k + 1] = x0 - beta;
      double vv = 0.0;
      LQ_UNROLL for (int i = k + 1; i < n; ++i) vv = fma(v[i], v[i], vv);
      const double tau = 2.0 / vv;
      // A <- (I - tau v v^T) A
      LQ_UNROLL for (int j = 0; j < n; ++j) {
        double s = 0.0;
        LQ_UNROLL for (int i = k + 1; i < n; ++i) s = fma(v[i], a[i * n + j], s);
        s *= tau;
        LQ_UNROLL for (int i = k + 1; i < n; ++i) a[i * n + j] = fma(-s, v[i], a[i * n + j]);
      }
      // A <- A (I - tau v v^T)
      LQ_UNROLL for (int i = 0; i < n; ++i) {
        double s = 0.0;
        LQ_UNROLL for (int j = k + 1; j < n; ++j) s = fma(a[i * n + j], v[j], s);
        s *= tau;
        LQ_UNROLL for (int j = k + 1; j < n; ++j) a[i * n + j] = fma(-s, v[j], a[i * n + j]);
      }
      LQ_UNROLL for (int i = k + 2; i < n; ++i) a[i * n + k] = 0.0;
    }
  }
}

// moduli of the two eigenvalues of the 2x2 block [[y, b],[c, x]] (w = b*c), max of them.
LQ_HD double block2_rho(double y, double x, double w) {
  const double p = 0.5 * (y - x);
  const double q = fma(p, p, w);
  const double z = sqrt(fabs(q));
  if (q >= 0.0) {
    const double zz = p + dsign(z, p);
    const double r1 = x + zz;
    const double r2 = (zz != 0.0) ? x - w / zz : r1;
    return dmax(fabs(r1), fabs(r2));
  }
  const double re = x + p;
  return sqrt(fma(re, re, z * z));
}

// Get/set with a run-time index over a register array, resolved by predicated selects (n is tiny).
template <int len>
LQ_HD double rget(const double* a, int idx) {
  double r = 0.0;
  LQ_UNROLL for (int i = 0; i < len; ++i) r = (i == idx) ? a[i] : r;
  return r;
}

// Spectral radius; *ok=false if the QR iteration did not converge within the iteration budget.
template <int n>
LQ_HD double spectral_radius(const double* Ain, bool* ok) {
  *ok = true;
  if (n == 1) return fabs(Ain[0]);
  if (n == 2) return block2_rho(Ain[0], Ain[3], Ain[1] * Ain[2]);
  double a[n * n];
  LQ_UNROLL for (int i = 0; i < n * n; ++i) a[i] = Ain[i];
  hessenberg<n>(a);
  double anorm = 0.0;
  LQ_UNROLL for (int i = 0; i < n; ++i)
    LQ_UNROLL for (int j = 0; j < n; ++j)
      if (j >= i - 1) anorm += fabs(a[i * n + j]);
  double rho = 0.0;
  double t = 0.0;       // accumulated exceptional shifts
  int nn = n - 1;       // active block is rows/cols l..nn
  int its = 0;
  int guard = 0;
  while (nn >= 0 && guard < 150 * n) {
    ++guard;
    // ---- find the start l of the active unreduced block (small sub-diagonal test)
    int l = 0;
    LQ_UNROLL for (int i = n - 1; i >= 1; --i) {
      if (i <= nn && l == 0) {
        double s = fabs(a[(i - 1) * n + (i - 1)]) + fabs(a[i * n + i]);
        if (s == 0.0) s = anorm;
        if (fabs(a[i * n + (i - 1)]) + s == s) { a[i * n + (i - 1)] = 0.0; l = i; }
      }
    }
    // x = a[nn][nn], y = a[nn-1][nn-1], w = a[nn][nn-1]*a[nn-1][nn]
    double x = 0.0, y = 0.0, w = 0.0;
    LQ_UNROLL for (int i = 0; i < n; ++i)
      if (i == nn) {
        x = a[i * n + i];
        if (i >= 1) { y = a[(i - 1) * n + (i - 1)]; w = a[i * n + (i - 1)] * a[(i - 1) * n + i]; }
      }
    if (l == nn) {  // one real root
      rho = dmax(rho, fabs(x + t));
      nn -= 1; its = 0;
      continue;
    }
    if (l == nn - 1) {  // a 2x2 block: two roots
      rho = dmax(rho, block2_rho(y + t, x + t, w));
      nn -= 2; its = 0;
      continue;
    }
    // ---- no deflation: one Francis double-shift sweep on rows/cols l..nn  (nn - l >= 2)
    if (its >= 120) { *ok = false; break; }
    if (its == 10 || its == 20) {  // exceptional shift (EISPACK schedule)
      t += x;
      LQ_UNROLL for (int i = 0; i < n; ++i) if (i <= nn) a[i * n + i] -= x;
      double s = 0.0;
      LQ_UNROLL for (int i = 2; i < n; ++i)
        if (i == nn) s = fabs(a[i * n + (i - 1)]) + fabs(a[(i - 1) * n + (i - 2)]);
      x = y = 0.75 * s;
      w = -0.4375 * s * s;
    }
    ++its;
    // choose the start row mst of the bulge (two consecutive small sub-diagonals), scanning up from nn-2
    int mst = l;
    double p = 0.0, q = 0.0, r = 0.0;
    {
      bool done = false;
      LQ_UNROLL for (int mm = n - 3; mm >= 0; --mm) {
        if (mm <= nn - 2 && mm >= l && !done) {
          const double z = a[mm * n + mm];
          const double rr = x - z, ss = y - z;
          double pp = (rr * ss - w) / a[(mm + 1) * n + mm] + a[mm * n + (mm + 1)];
          double qq = a[(mm + 1) * n + (mm + 1)] - z - rr - ss;
          double r3 = a[(mm + 2) * n + (mm + 1)];
          const double sc = fabs(pp) + fabs(qq) + fabs(r3);
          pp /= sc; qq /= sc; r3 /= sc;
          p = pp; q = qq; r = r3; mst = mm;
          if (mm == l) {
            done = true;
          } else {
            const double u = fabs(a[mm * n + (mm - 1)]) * (fabs(qq) + fabs(r3));
            const double v = fabs(pp) * (fabs(a[(mm - 1) * n + (mm - 1)]) + fabs(z) +
                                         fabs(a[(mm + 1) * n + (mm + 1)]));
            if (u + v == v) done = true;
          }
        }
      }
    }
    LQ_UNROLL for (int i = 2; i < n; ++i)
      if (i >= mst + 2 && i <= nn) {
        a[i * n + (i - 2)] = 0.0;
        if (i >= 3 && i != mst + 2) a[i * n + (i - 3)] = 0.0;
      }
    // bulge chase
    LQ_UNROLL for (int k = 0; k < n - 1; ++k) {
      if (k >= mst && k <= nn - 1) {
        const bool last = (k == nn - 1);
        double xs = 0.0;
        if (k != mst) {
          p = a[k * n + (k - 1 < 0 ? 0 : k - 1)];
          q = a[(k + 1) * n + (k - 1 < 0 ? 0 : k - 1)];
          r = 0.0;
          if (!last && k + 2 < n) r = a[(k + 2 < n ? k + 2 : n - 1) * n + (k - 1 < 0 ? 0 : k - 1)];
          xs = fabs(p) + fabs(q) + fabs(r);
          if (xs != 0.0) { p /= xs; q /= xs; r /= xs; }
        }
        const double s = dsign(sqrt(p * p + q * q + r * r), p);
        if (s != 0.0) {
          if (k == mst) {
            if (l != mst && k >= 1) a[k * n + (k - 1)] = -a[k * n + (k - 1)];
          } else if (k >= 1) {
            a[k * n + (k - 1)] = -s * xs;
          }
          p += s;
          const double hx = p / s, hy = q / s, hz = r / s;
          q /= p; r /= p;
          // row modification (columns k..nn)
          LQ_UNROLL for (int j = 0; j < n; ++j)
            if (j >= k && j <= nn) {
              double pp = a[k * n + j] + q * a[(k + 1) * n + j];
              if (!last && k + 2 < n) {
                pp += r * a[(k + 2 < n ? k + 2 : n - 1) * n + j];
                a[(k + 2 < n ? k + 2 : n - 1) * n + j] -= pp * hz;
              }
              a[(k + 1) * n + j] -= pp * hy;
              a[k * n + j] -= pp * hx;
            }
          // column modification (rows l..min(nn, k+3))
          LQ_UNROLL for (int i = 0; i < n; ++i)
            if (i >= l && i <= nn && i <= k + 3) {
              double pp = hx * a[i * n + k] + hy * a[i * n + (k + 1)];
              if (!last && k + 2 < n) {
                pp += hz * a[i * n + (k + 2 < n ? k + 2 : n - 1)];
                a[i * n + (k + 2 < n ? k + 2 : n - 1)] -= pp * r;
              }
              a[i * n + (k + 1)] -= pp * q;
              a[i * n + k] -= pp;
            }
        }
      }
    }
  }
  if (nn >= 0) *ok = false;
  return rho;
}

}  // namespace lq
