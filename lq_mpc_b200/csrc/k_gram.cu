// k_gram.cu — K3g: extreme eigenvalues of the Gram matrix Gamma'Gamma, ONE WARP PER SAMPLE, matrix in shared memory.
//
// energy_bound needs ||Gamma||_2 and lambda_min(H^) (reference utils.py:248-258, 316-322: np.linalg.norm(Gamma, 2)
// and np.min(np.linalg.eigvals(hatH)) on (N m) x (N m) matrices). For Q = qI, R = rI (every shipped scenario)
// H^ = rI + q Gamma'Gamma, so both come from the extreme eigenvalues of C = Gamma'Gamma (k = N m).
// The thread-per-sample kernel (bounds.cuh) keeps C in a per-thread slice of a global workspace; that is fine for
// k ~ 7 but at k = 50 (cfg-sweep, N = 50) the Householder reduction re-reads 20 kB per sample k times: ncu showed
// 108 GB of DRAM traffic per launch of 2e5 samples (61.6 ms, fp64 pipe 4 %) against ~20 MB of algorithmic bytes.
// Here a warp owns the sample: G_d = A^d B and C ((k+1)-strided, both triangles: conflict-free row access) live in
// shared memory, the Householder tridiagonalisation (EISPACK tred1 arithmetic) runs lane-parallel over rows, and
// the tridiagonal form (2k doubles per sample, SoA) goes back to the thread-per-sample kernel for the Sturm bisection.
// HBM traffic drops to the operands plus that tridiagonal (n^2 + n m doubles in, 2k doubles out per sample).
#include "engine.h"

namespace {

constexpr int kWarps = 4;      // default CTA size in warps (eligibility test); the launcher picks 1..8 per launch
constexpr int kMaxWarps = 8;

__host__ __device__ inline int lq_gram_ldc(int k) {
  int ld = (k + 1) & ~1;              // even, >= k
  if (((ld >> 1) & 1) == 0) ld += 2;  // ld / 2 odd
  return ld;
}

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct GramArgs {
  int64_t S;
  const double* dA;   // SoA [n*n][S] or NULL
  const double* dB;   // SoA [n*m][S] or NULL
  int N;
  double* tri;        // [2k][S]: diagonal d (k rows) then off-diagonal |e| (k rows; e[0] = 0) of the tridiagonal form
};

template <int n, int m>
__global__ void __launch_bounds__(kMaxWarps * 32) gram_extremes_kernel(const __grid_constant__ lq::Problem<n, m> pb,
                                                                    const GramArgs a, const int per_warp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // leading dimension: even with LDC/2 odd — rows are 16-byte aligned and the 128-bit loads of 8 consecutive rows fall
  // into 8 distinct bank groups (no conflicts), so the Householder loops move TWO entries per shared-memory
  // instruction (the kernel is bound by instruction issue: ~1 LDS/STS per FMA with 64-bit accesses)
  const int N = a.N, k = N * m, LDC = lq_gram_ldc(k);
  double* base = reinterpret_cast<double*>(smem_raw) + (int64_t)wib * per_warp;
  const int ke = (k + 1) & ~1;         // vectors padded to an even length: 16-byte aligned starts
  double* C = base;                    // k x LDC
  double* v = C + k * LDC;             // k
  double* q = v + ke;                  // k
  double* dd = q + ke;                 // k  diagonal of the tridiagonal form
  double* e2 = dd + ke;                // k  squared off-diagonal
  double* G = e2 + ke;                 // N x (n*m): G_d = A^d B
  const int wpc = blockDim.x >> 5;     // warps per CTA, chosen by the launcher
  const int64_t wid = (int64_t)blockIdx.x * wpc + wib, nw = (int64_t)gridDim.x * wpc;
#define C_(i, j) C[(i) * LDC + (j)]
  for (int64_t s = wid; s < a.S; s += nw) {
    // ---- G_d = A^^d B^ (every lane runs the tiny recurrence; lane d % 32 stores G_d)
    {
      double Ah[n * n], Gd[n * m];
#pragma unroll
      for (int e = 0; e < n * n; ++e) Ah[e] = pb.A[e] + (a.dA ? a.dA[(int64_t)e * a.S + s] : 0.0);
#pragma unroll
      for (int e = 0; e < n * m; ++e) Gd[e] = pb.B[e] + (a.dB ? a.dB[(int64_t)e * a.S + s] : 0.0);
      for (int d = 0; d < N; ++d) {
        if ((d & 31) == lane) {
#pragma unroll
          for (int e = 0; e < n * m; ++e) G[d * (n * m) + e] = Gd[e];
        }
        double Gn[n * m];
        lq::mm<n, n, m>(Ah, Gd, Gn);
#pragma unroll
        for (int e = 0; e < n * m; ++e) Gd[e] = Gn[e];
      }
    }
    __syncwarp();
    // ---- C = Gamma'Gamma along its block diagonals: C[j][j-dl] = C[j+1][j+1-dl] + G_{N-1-j}' G_{N-1-j+dl}
    for (int dl = lane; dl < N; dl += 32) {
      double acc[m * m];
#pragma unroll
      for (int e = 0; e < m * m; ++e) acc[e] = 0.0;
      for (int j = N - 1; j >= dl; --j) {
        const int jp = j - dl;
        const double* ga = G + (N - 1 - j) * (n * m);
        const double* gb = G + (N - 1 - jp) * (n * m);
#pragma unroll
        for (int aa = 0; aa < m; ++aa)
#pragma unroll
          for (int bb = 0; bb < m; ++bb) {
            double t = acc[aa * m + bb];
#pragma unroll
            for (int r = 0; r < n; ++r) t = fma(ga[r * m + aa], gb[r * m + bb], t);
            acc[aa * m + bb] = t;
            C_(j * m + aa, jp * m + bb) = t;
            C_(jp * m + bb, j * m + aa) = t;
          }
      }
    }
    __syncwarp();
    // ---- Householder tridiagonalisation (tred1 arithmetic), rows distributed over lanes
    for (int i = k - 1; i >= 1; --i) {
      const int l = i - 1;
      if (l == 0) {
        if (lane == 0) e2[i] = C_(i, 0) * C_(i, 0);
        continue;
      }
      double sc = 0.0;
      for (int j = lane; j <= l; j += 32) sc += fabs(C_(i, j));
      sc = wsum(sc);
      if (sc == 0.0) {
        if (lane == 0) e2[i] = C_(i, l) * C_(i, l);
        continue;
      }
      const double isc = 1.0 / sc;
      double h = 0.0;
      for (int j = lane; j <= l; j += 32) {
        const double t = C_(i, j) * isc;
        v[j] = t;
        h = fma(t, t, h);
      }
      h = wsum(h);
      __syncwarp();
      const double f0 = v[l];
      const double g0 = (f0 >= 0.0) ? -sqrt(h) : sqrt(h);
      if (lane == 0) e2[i] = (sc * g0) * (sc * g0);
      h -= f0 * g0;
      __syncwarp();
      if (lane == 0) v[l] = f0 - g0;
      __syncwarp();
      // p = C v / h on the leading (l+1) block; f = p . v   (two columns per shared-memory instruction)
      const double ih = 1.0 / h;
      const int lp = (l + 1) & ~1;                          // paired part of 0..l
      double fl = 0.0;
      for (int j = lane; j <= l; j += 32) {
        double g0 = 0.0, g1 = 0.0;
        const double* row = C + j * LDC;
        for (int kk = 0; kk < lp; kk += 2) {
          const double2 c2 = *reinterpret_cast<const double2*>(row + kk);
          const double2 v2 = *reinterpret_cast<const double2*>(v + kk);
          g0 = fma(c2.x, v2.x, g0);
          g1 = fma(c2.y, v2.y, g1);
        }
        if (lp <= l) g0 = fma(row[l], v[l], g0);
        const double g = (g0 + g1) * ih;
        q[j] = g;
        fl = fma(g, v[j], fl);
      }
      fl = wsum(fl);
      const double hh = fl / (h + h);
      __syncwarp();
      for (int j = lane; j <= l; j += 32) q[j] -= hh * v[j];
      __syncwarp();
      // rank-2 update of the whole leading block (both triangles are kept)
      for (int j = lane; j <= l; j += 32) {
        const double nvj = -v[j], nqj = -q[j];
        double* row = C + j * LDC;
        for (int kk = 0; kk < lp; kk += 2) {
          double2 c2 = *reinterpret_cast<const double2*>(row + kk);
          const double2 v2 = *reinterpret_cast<const double2*>(v + kk);
          const double2 q2 = *reinterpret_cast<const double2*>(q + kk);
          c2.x = fma(nvj, q2.x, fma(nqj, v2.x, c2.x));
          c2.y = fma(nvj, q2.y, fma(nqj, v2.y, c2.y));
          *reinterpret_cast<double2*>(row + kk) = c2;
        }
        if (lp <= l) row[l] = fma(nvj, q[l], fma(nqj, v[l], row[l]));
      }
      __syncwarp();
    }
    for (int i = lane; i < k; i += 32) dd[i] = C_(i, i);
    if (lane == 0) e2[0] = 0.0;
    __syncwarp();
    // ---- hand the tridiagonal form (d, e) to the thread-per-sample bounds kernel, SoA [2k][S]: the Sturm bisection
    //      is a sequential recurrence per shift — one sample per LANE (k_bounds.cu) uses every lane, one sample per
    //      warp would idle 31 (a 32-way multisection variant here spent 70 % of this kernel's instructions on it)
    for (int i = lane; i < k; i += 32) {
      a.tri[(int64_t)i * a.S + s] = dd[i];
      a.tri[(int64_t)(k + i) * a.S + s] = sqrt(e2[i]);
    }
    __syncwarp();
  }
#undef C_
}

template <int n, int m>
int launch_gram_t(lqmpc_ctx* ctx, int64_t S, const double* dA, const double* dB, int N, double* tri) {
  const lq::Problem<n, m>& pb = *reinterpret_cast<const lq::Problem<n, m>*>(ctx->pb);
  const int k = N * m;
  const int per_warp = ((k * lq_gram_ldc(k) + N * n * m + 4 * ((k + 1) & ~1)) + 1) & ~1;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  // The kernel is latency-bound (profile r01k3g2: 8 warps/SM at k = 50, issue slots 28 %): pick the CTA size that
  // puts the most warps on an SM for this k (shared memory per warp grows with k^2; 1 KB is reserved per CTA).
  int best_w = 1, best_per_sm = 1, best_total = 0;
  for (int w = kMaxWarps; w >= 1; --w) {
    const size_t sm_w = (size_t)w * per_warp * sizeof(double);
    if (sm_w > 200 * 1024) continue;
    cudaFuncSetAttribute(gram_extremes_kernel<n, m>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_w);
    int ps = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ps, gram_extremes_kernel<n, m>, w * 32, sm_w);
    if (ps * w > best_total) { best_total = ps * w; best_w = w; best_per_sm = ps; }
  }
  const size_t smem = (size_t)best_w * per_warp * sizeof(double);
  cudaFuncSetAttribute(gram_extremes_kernel<n, m>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int64_t blocks = (int64_t)sms * best_per_sm;
  const int64_t want = (S + best_w - 1) / best_w;
  if (blocks > want) blocks = want;
  GramArgs a{S, dA, dB, N, tri};
  gram_extremes_kernel<n, m><<<(unsigned)blocks, best_w * 32, smem, ctx->stream>>>(pb, a, per_warp);
  ctx->launches++;
  return lq_check_cuda(ctx, cudaGetLastError(), "gram_extremes_kernel launch");
}

}  // namespace

// shared-memory footprint decides eligibility: 4 warps x (k (k+1) + ...) doubles must fit one CTA
bool lq_gram_warp_eligible(int n, int m, int N) {
  const int k = N * m;
  const size_t per_warp = (size_t)(k * lq_gram_ldc(k) + N * n * m + 4 * ((k + 1) & ~1) + 2);
  return k >= 12 && (size_t)kWarps * per_warp * sizeof(double) <= 200 * 1024;
}

int lq_launch_gram(lqmpc_ctx* ctx, int64_t S, const double* dA, const double* dB, int N, double* tri) {
#define X(N_, M_) \
  if (ctx->n == N_ && ctx->m == M_) return launch_gram_t<N_, M_>(ctx, S, dA, dB, N, tri);
  LQ_FOR_EACH_DIM(X)
#undef X
  return lq_set_error(ctx, LQMPC_EINVAL, "unsupported (n, m); see lqmpc_supported_dims()");
}
