// Micro-benchmarks that set the ceilings for the PCG kernels on B200 (run under gpurun):
//   1. grid barrier latency: cooperative_groups grid.sync() vs a hand-rolled atomic+spin barrier
//   2. random 32 B gather rate from an L2-resident 3.2 MB table (the SpMV inner operation)
//   3. L2-resident streaming bandwidth at the PCG vector sizes
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/microbench tools/microbench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ double4 ld256(const double4* p) {
  double4 v;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st256(double4* p, double4 v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}

// ---- 1. barriers -----------------------------------------------------------------------------
__global__ void k_gridsync(int iters, long long* out) {
  cg::grid_group g = cg::this_grid();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) g.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = clock64() - t0;
}
__device__ __forceinline__ void my_barrier(unsigned int* ctr, unsigned int& phase) {
  __syncthreads();
  if (threadIdx.x == 0) {
    phase += gridDim.x;
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned int v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < phase);
  }
  __syncthreads();
}
__global__ void k_mybarrier(int iters, unsigned int* ctr, long long* out) {
  unsigned int phase = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) my_barrier(ctr, phase);
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = clock64() - t0;
}

// ---- 2. gathers ---------------------------------------------------------------------------------
// each thread sums `per` random 32 B records, UNR independent loads in flight
template <int UNR, int MODE>
__global__ void k_gather(const double4* __restrict__ tab, const int* __restrict__ idx, int per, double* out) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthreads = gridDim.x * blockDim.x;
  double ax = 0, ay = 0, az = 0;
  for (int j = 0; j < per; j += UNR) {
    int c[UNR]; double4 v[UNR];
#pragma unroll
    for (int q = 0; q < UNR; ++q) c[q] = idx[(size_t)(j + q) * nthreads + tid];
#pragma unroll
    for (int q = 0; q < UNR; ++q) {
      if (MODE == 0) v[q] = ld256(tab + c[q]);
      else if (MODE == 1) asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[q].x), "=d"(v[q].y), "=d"(v[q].z), "=d"(v[q].w) : "l"(tab + c[q]));
      else if (MODE == 2) asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[q].x), "=d"(v[q].y), "=d"(v[q].z), "=d"(v[q].w) : "l"(tab + c[q]));
      else if (MODE == 3) { const double2* p2 = (const double2*)(tab + c[q]); double2 a = __ldg(p2), b = __ldg(p2 + 1); v[q] = make_double4(a.x, a.y, b.x, b.y); }
      else { const double* p1 = (const double*)(tab + c[q]); v[q] = make_double4(__ldg(p1), __ldg(p1 + 1), __ldg(p1 + 2), 0.0); }
    }
#pragma unroll
    for (int q = 0; q < UNR; ++q) { ax += v[q].x; ay += v[q].y; az += v[q].z; }
  }
  if (ax + ay + az == 1.2345) out[tid] = ax;
}

// ---- 3. streams -----------------------------------------------------------------------------------
__global__ void k_axpy4(double4* __restrict__ x, const double4* __restrict__ a, const double4* __restrict__ b, double4* __restrict__ y, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double4 va = ld256(a + i), vb = ld256(b + i), vx = ld256(x + i);
    vx.x += va.x * vb.x; vx.y += va.y * vb.y; vx.z += va.z * vb.z;
    st256(x + i, vx);
    st256(y + i, vb);
  }
}

template <class F>
float time_us(F&& f, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); f();
  CK(cudaEventRecord(a));
  for (int r = 0; r < reps; ++r) f();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  return ms * 1000.f / reps;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("device %s sms %d clock %d kHz\n", prop.name, sms, khz);
  long long* d_out; CK(cudaMalloc(&d_out, 64));
  unsigned int* d_ctr; CK(cudaMalloc(&d_ctr, 4));

  // 1. barriers
  for (int threads : {256, 512, 768, 1024}) {
    for (int bps : {1, 2}) {
      if (threads * bps > 2048) continue;
      int iters = 200; int grid = sms * bps;
      void* args[] = {&iters, &d_out};
      float us = time_us([&] { CK(cudaLaunchCooperativeKernel((void*)k_gridsync, dim3(grid), dim3(threads), args, 0, 0)); }, 5);
      long long cyc; CK(cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost));
      printf("grid.sync   grid %4d x %4d : %.3f us/barrier (event) %.0f cycles/barrier\n", grid, threads, us / iters, (double)cyc / iters);
      void* args2[] = {&iters, &d_ctr, &d_out};
      float us2 = time_us([&] { CK(cudaMemsetAsync(d_ctr, 0, 4)); CK(cudaLaunchCooperativeKernel((void*)k_mybarrier, dim3(grid), dim3(threads), args2, 0, 0)); }, 5);
      CK(cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost));
      printf("my barrier  grid %4d x %4d : %.3f us/barrier (event) %.0f cycles/barrier\n", grid, threads, us2 / iters, (double)cyc / iters);
    }
  }
  for (int grid : {1, 5, 16, 64}) {
    int iters = 200, threads = 768;
    void* args[] = {&iters, &d_out};
    float us = time_us([&] { CK(cudaLaunchCooperativeKernel((void*)k_gridsync, dim3(grid), dim3(threads), args, 0, 0)); }, 5);
    printf("grid.sync   grid %4d x %4d : %.3f us/barrier\n", grid, threads, us / iters);
  }
  {
    float us = time_us([&] { k_axpy4<<<1, 32>>>(nullptr, nullptr, nullptr, nullptr, 0); }, 200);
    printf("empty kernel launch, back to back: %.3f us\n", us);
  }

  // 2. gathers: table of n records, total 2M gathers
  const int n = 100000; const size_t total = 2u << 20;
  double4* tab; CK(cudaMalloc(&tab, sizeof(double4) * n)); CK(cudaMemset(tab, 0, sizeof(double4) * n));
  std::vector<int> h(total);
  srand(1);
  for (size_t i = 0; i < total; ++i) h[i] = (int)(((unsigned)rand() * 32768u + (unsigned)rand()) % n);
  int* idx; CK(cudaMalloc(&idx, 4 * total)); CK(cudaMemcpy(idx, h.data(), 4 * total, cudaMemcpyHostToDevice));
  double* gout; CK(cudaMalloc(&gout, 8 * total));
  printf("random 32 B gathers: %zu gathers from a %.1f MB table\n", total, n * 32 / 1e6);
  for (int threads_total : {65536, 131072, 262144}) {
    const int per = (int)(total / threads_total);
    const int block = 256, grid = threads_total / block;
#define RUNG(U, M, name) { float us = time_us([&] { k_gather<U, M><<<grid, block>>>(tab, idx, per, gout); }, 20); \
      printf("  threads %6d per %3d unroll %d %-14s : %7.2f us  (%.1f Ggather/s)\n", threads_total, per, U, name, us, total / us / 1e3); }
    RUNG(4, 0, "ld.v4.f64"); RUNG(8, 0, "ld.v4.f64"); RUNG(16, 0, "ld.v4.f64");
    RUNG(8, 1, "ld.nc.v4.f64"); RUNG(8, 2, "no_allocate"); RUNG(8, 3, "2 x ldg.128"); RUNG(8, 4, "3 x ldg.64");
  }
  // sequential (coalesced) indices for comparison: same kernel, idx = tid-contiguous records
  for (size_t i = 0; i < total; ++i) h[i] = (int)(i % n);
  CK(cudaMemcpy(idx, h.data(), 4 * total, cudaMemcpyHostToDevice));
  { const int threads_total = 131072, per = (int)(total / threads_total), block = 256, grid = threads_total / block;
    RUNG(8, 0, "coalesced idx"); }

  // 3. streams
  for (int nn : {100000, 1000000}) {
    double4 *x, *a, *b, *y;
    CK(cudaMalloc(&x, 32 * (size_t)nn)); CK(cudaMalloc(&a, 32 * (size_t)nn)); CK(cudaMalloc(&b, 32 * (size_t)nn)); CK(cudaMalloc(&y, 32 * (size_t)nn));
    CK(cudaMemset(x, 0, 32 * (size_t)nn)); CK(cudaMemset(a, 0, 32 * (size_t)nn)); CK(cudaMemset(b, 0, 32 * (size_t)nn));
    for (int grid : {sms, sms * 2, sms * 4, sms * 8}) {
      for (int block : {256, 1024}) {
        float us = time_us([&] { k_axpy4<<<grid, block>>>(x, a, b, y, nn); }, 50);
        printf("stream n=%7d (5 x 32 B per node = %.1f MB) grid %4d x %4d: %7.2f us  %.2f TB/s\n", nn, nn * 160 / 1e6, grid, block, us, nn * 160.0 / us / 1e6);
      }
    }
    cudaFree(x); cudaFree(a); cudaFree(b); cudaFree(y);
  }
  return 0;
}
