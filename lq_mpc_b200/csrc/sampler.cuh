// sampler.cuh — counter-based model-error sampler (K6), one (perturbation j, error level i) pair per thread.
//
// Restates the reference's `random_matrix` (utils.py:779-823): r x c matrices with entries uniform in [-e, e],
// accepted when ||T|| <= e in the Frobenius ('f') or spectral ('2') norm; the first `n_boundary` perturbations of
// every level sit ON the norm boundary. The reference draws from Python's unseeded `random.uniform` (utils.py:774)
// and obtains the boundary ones by rejection with np.isclose — here (as in lq_mpc_b200/sampling.py) they are
// rescaled onto ||T|| = e, and every draw comes from Philox4x32-10 keyed by the seed with the counter
// (j, i | which << 16, attempt, pair index): any sharding / launch shape sees identical samples, and the integer
// stream is bit-exact against the numpy restatement (oracle/np_sampler.py, pinned on Random123's known answers).
// A uniform is 53 bits of two 32-bit words (numpy's construction); the entry is e * (2u - 1): 2u - 1 is exact in
// FP64, so the value has ONE rounding and cannot depend on FMA contraction.
#pragma once
#include "small_la.cuh"

namespace lq {

struct Philox4 {
  uint32_t v[4];
};

LQ_HD void philox_mulhilo(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
  const uint64_t p = (uint64_t)a * (uint64_t)b;
  *hi = (uint32_t)(p >> 32);
  *lo = (uint32_t)p;
}

LQ_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  LQ_UNROLL for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    philox_mulhilo(0xD2511F53u, c0, &hi0, &lo0);
    philox_mulhilo(0xCD9E8D57u, c2, &hi1, &lo1);
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{{c0, c1, c2, c3}};
}

// symmetric uniform in [-1, 1): exact (k - 2^52) 2^-52 from 53 random bits
LQ_HD double philox_symm(uint32_t w0, uint32_t w1) {
  const double u53 = (double)(w0 >> 5) * 67108864.0 + (double)(w1 >> 6);   // integer in [0, 2^53)
  return (u53 - 4503599627370496.0) * (1.0 / 4503599627370496.0);
}

constexpr uint32_t kSamplerStream = 0x4C514D50u;   // "LQMP"
constexpr int kSamplerMaxAttempts = 256;

// One r x c perturbation. norm_type 0: Frobenius, 1: spectral. Returns the number of rejected attempts, or
// -attempts when none was accepted and the last draw was projected onto the ball (never seen for 2 x 2 / 2 x 1).
template <int r, int c>
LQ_HD int sample_error_matrix(uint64_t seed, int which, int64_t j, int level, double e, bool boundary,
                              int norm_type, double* T) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ kSamplerStream;
  const uint32_t c0 = (uint32_t)j;
  const uint32_t c1 = (uint32_t)level | ((uint32_t)which << 16) | ((uint32_t)((uint64_t)j >> 32) << 20);
  double nv = 0.0;
  for (int a = 0; a < kSamplerMaxAttempts; ++a) {
    LQ_UNROLL for (int p = 0; p < (r * c + 1) / 2; ++p) {
      const Philox4 x = philox4x32_10(c0, c1, (uint32_t)a, (uint32_t)p, k0, k1);
      T[2 * p] = e * philox_symm(x.v[0], x.v[1]);
      if (2 * p + 1 < r * c) T[2 * p + 1] = e * philox_symm(x.v[2], x.v[3]);
    }
    if (norm_type == 0) {
      double ss = 0.0;
      LQ_UNROLL for (int q = 0; q < r * c; ++q) ss = fma(T[q], T[q], ss);
      nv = sqrt(ss);
    } else {
      nv = norm2<r, c>(T);
    }
    if (boundary) {
      if (nv > 0.0) {
        const double sc = e / nv;
        LQ_UNROLL for (int q = 0; q < r * c; ++q) T[q] *= sc;
        return a;
      }
    } else if (nv <= e) {
      return a;
    }
  }
  const double sc = e / nv;
  LQ_UNROLL for (int q = 0; q < r * c; ++q) T[q] *= sc;
  return -kSamplerMaxAttempts;
}

// ---- seeded synthetic evaluation samples (lqmpc_eval_seeded): sample s draws its (dA, dB, x0) from Philox4x32-10 keyed
// by the seed with the counter (s lo, s hi, pair index, 0): dA, dB entries e (2u - 1) uniform in [-e, e) (one rounding,
// as above), x0 ~ N(0, I) by Box-Muller on the next pairs. The GLOBAL sample index is the counter, so any sharding of
// [first, first + S) over ranks sees identical samples. Restated in oracle/np_sampler.seeded_samples:
// the uniforms are bit-exact, the normals agree to the last digits of log / cos / sin.
constexpr uint32_t kSeededStream = 0x53454544u;   // "SEED"

template <int n, int m>
LQ_HD void seeded_sample(uint64_t seed, int64_t s, double e_A, double e_B, double* dA, double* dB, double* x0) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ kSeededStream;
  const uint32_t c0 = (uint32_t)s, c1 = (uint32_t)((uint64_t)s >> 32);
  constexpr int nu = n * n + n * m;
  LQ_UNROLL for (int p = 0; p < (nu + 1) / 2; ++p) {
    const Philox4 x = philox4x32_10(c0, c1, (uint32_t)p, 0u, k0, k1);
    const double u0 = philox_symm(x.v[0], x.v[1]), u1 = philox_symm(x.v[2], x.v[3]);
    const int q0 = 2 * p, q1 = 2 * p + 1;
    if (q0 < n * n) dA[q0] = e_A * u0; else dB[q0 - n * n] = e_B * u0;
    if (q1 < nu) { if (q1 < n * n) dA[q1] = e_A * u1; else dB[q1 - n * n] = e_B * u1; }
  }
  LQ_UNROLL for (int p = 0; p < (n + 1) / 2; ++p) {
    const Philox4 x = philox4x32_10(c0, c1, (uint32_t)((nu + 1) / 2 + p), 0u, k0, k1);
    // u in (0, 1] for the logarithm, v in [0, 1) for the angle: 53-bit uniforms
    const double u = ((double)(x.v[0] >> 5) * 67108864.0 + (double)(x.v[1] >> 6) + 1.0) * (1.0 / 9007199254740992.0);
    const double v = ((double)(x.v[2] >> 5) * 67108864.0 + (double)(x.v[3] >> 6)) * (1.0 / 9007199254740992.0);
    const double r = sqrt(-2.0 * log(u)), th = 6.283185307179586476925286766559 * v;
    x0[2 * p] = r * cos(th);
    if (2 * p + 1 < n) x0[2 * p + 1] = r * sin(th);
  }
}

}  // namespace lq
