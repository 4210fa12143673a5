// Development probe (GPU box): how the FP64 pipe of one SM sub-partition is shared between DFMA chains and FP64
// tensor-core MMAs (mma.sync.m8n8k4.f64). One CTA of 1024 threads on one SM; role per warp from a table.
//   role 0 idle, 1 timed dependent DFMA chain, 2 untimed stream of independent DMMAs, 3 timed dependent DMMA chain,
//   4 timed chain of dependent rsqrt-like MUFU+DFMA sequences
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probes/fp64_pipe_probe scripts/probes/fp64_pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct Roles { int r[32]; };

__global__ void __launch_bounds__(1024) probe(Roles roles, int iters, double a, double b, long long* cyc, double* sink,
                                               unsigned* smsp) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int role = roles.r[w];
  __shared__ volatile int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  double x = a + lane * 1e-9, acc = 0.0;
  long long t0 = 0, t1 = 0;
  if (role == 1) {
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 16; ++u) x = fma(x, b, a);
    }
    t1 = clock64();
    acc = x;
    __threadfence_block();
    if (lane == 0) stop = 1;
  } else if (role == 3) {
    double c0 = lane, c1 = -lane;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 16; ++u) dmma(c0, c1, a, b);
    }
    t1 = clock64();
    acc = c0 + c1;
    __threadfence_block();
    if (lane == 0) stop = 1;
  } else if (role == 4) {
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        double y;
        asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        const double h = 0.5 * x;
        double e = fma(-h * y, y, 0.5);
        y = fma(y, e, y);
        e = fma(-h * y, y, 0.5);
        x = fma(y, e, y) + a;
      }
    }
    t1 = clock64();
    acc = x;
    __threadfence_block();
    if (lane == 0) stop = 1;
  } else if (role == 2) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = lane + i; c[i][1] = lane - i; }
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += c[i][0] + c[i][1];
  }
  if (lane == 0) {
    cyc[w] = t1 - t0;
    unsigned id;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(id));
    smsp[w] = id;
  }
  if (acc == 123.456) sink[0] = acc;
}

int main() {
  long long* cyc; double* sink; unsigned* smsp;
  cudaMallocManaged(&cyc, 32 * sizeof(long long));
  cudaMallocManaged(&sink, 8);
  cudaMallocManaged(&smsp, 32 * sizeof(unsigned));
  const int iters = 2000;
  struct Case { const char* name; int timed_role; int n_comp; int comp_stride; int comp_first; } cases[] = {
      {"DFMA chain alone", 1, 0, 0, 0},
      {"DFMA chain + 1 DMMA warp, warp 4 (same w%4)", 1, 1, 4, 4},
      {"DFMA chain + 4 DMMA warps 4,8,12,16 (same w%4)", 1, 4, 4, 4},
      {"DFMA chain + 1 DMMA warp, warp 1", 1, 1, 4, 1},
      {"DFMA chain + 4 DMMA warps 1,5,9,13 (w%4 = 1)", 1, 4, 4, 1},
      {"DFMA chain + 12 DMMA warps w%4 != 0", 1, 12, 0, 0},
      {"DMMA chain alone", 3, 0, 0, 0},
      {"DMMA chain + 4 DMMA warps same w%4", 3, 4, 4, 4},
      {"rsqrt(seed + 2 Newton) chain alone", 4, 0, 0, 0},
      {"rsqrt chain + 4 DMMA warps same w%4", 4, 4, 4, 4},
  };
  for (auto& cs : cases) {
    Roles r{};
    r.r[0] = cs.timed_role;
    if (cs.n_comp == 12) { for (int w = 1; w < 16; ++w) if (w % 4) r.r[w] = 2; }
    else for (int i = 0; i < cs.n_comp; ++i) r.r[cs.comp_first + i * cs.comp_stride] = 2;
    probe<<<1, 1024>>>(r, iters, 1.0000001, 0.9999999, cyc, sink, smsp);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    const double ops = (cs.timed_role == 4) ? iters * 4.0 : iters * 16.0;
    printf("%-52s %8.2f cycles per dependent op\n", cs.name, cyc[0] / ops);
  }
  printf("%%warpid of warps 0..7:");
  for (int w = 0; w < 8; ++w) printf(" %u", smsp[w]);
  printf("\n");
  return 0;
}
