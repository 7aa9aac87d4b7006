// FP64 peak probe for B200 (sm_100a): DMMA (mma.sync f64) issue-rate per shape, DFMA rate,
// and dependent-chain latency.  Output: one JSON object per line.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/dmma_probe tools/dmma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int SHAPE> struct Frag;

__device__ __forceinline__ void mma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double (&c)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// NACC independent accumulators per warp; iters outer iterations.
template <int NACC>
__global__ void k_884(double* out, int iters, double a, double b) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) mma884(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_1684(double* out, int iters, double a, double b) {
  double c[NACC][4]; double av[2] = {a, a + 1};
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) mma1684(c[i], av, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_1688(double* out, int iters, double a, double b) {
  double c[NACC][4]; double av[4] = {a, a + 1, a + 2, a + 3}; double bv[2] = {b, b + 1};
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) mma1688(c[i], av, bv);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_16816(double* out, int iters, double a, double b) {
  double c[NACC][4]; double av[8]; double bv[4];
#pragma unroll
  for (int i = 0; i < 8; i++) av[i] = a + i;
#pragma unroll
  for (int i = 0; i < 4; i++) bv[i] = b + i;
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) mma16816(c[i], av, bv);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// Shared-memory fed m16n8k4 loop, register-blocked WM=16*MB x WN=8*NB per warp: measures whether LDS feeding keeps the pipe full.
template <int MB, int NB>
__global__ void k_smemfed(double* out, int iters) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1e-3 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  double c[MB][NB][4];
#pragma unroll
  for (int i = 0; i < MB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) c[i][j][0] = c[i][j][1] = c[i][j][2] = c[i][j][3] = 0;
  for (int it = 0; it < iters; it++) {
    const double* base = sm + ((it & 15) * 256);
    double a[MB][2], b[NB];
#pragma unroll
    for (int i = 0; i < MB; i++) {
      double2 v = *reinterpret_cast<const double2*>(base + ((t * 2) ^ 0) * 16 + i * 128 + g * 2);
      a[i][0] = v.x; a[i][1] = v.y;
    }
#pragma unroll
    for (int j = 0; j < NB; j++) b[j] = base[2048 + j * 32 + t * 8 + g];
#pragma unroll
    for (int i = 0; i < MB; i++)
#pragma unroll
      for (int j = 0; j < NB; j++) mma1684(c[i][j], a[i], b[j]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < MB; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) s += c[i][j][0] + c[i][j][1] + c[i][j][2] + c[i][j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch();
  { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { fprintf(stderr, "launch failed: %s\n", cudaGetErrorString(e)); return -1.f; } }
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("{\"probe\":\"device\",\"name\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  const int iters = 20000;
  int warps_list[] = {4, 8, 16, 32};
  for (int wi = 0; wi < 4; wi++) {
    int warps = warps_list[wi]; int threads = warps * 32;
    int grid = sms * (warps <= 16 ? 2 : 1);
    double nw = (double)grid * warps;
#define REPORT(NAME, NACC, FLOPS_PER_MMA, KERNEL)                                                     \
    { float ms = time_ms([&] { KERNEL<<<grid, threads>>>(out, iters, 1.0000001, 0.5); });              \
      double fl = nw * (double)iters * NACC * FLOPS_PER_MMA;                                           \
      printf("{\"probe\":\"%s\",\"nacc\":%d,\"warps_per_cta\":%d,\"grid\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n", NAME, NACC, warps, grid, ms, fl / ms * 1e-9); }
    REPORT("dmma_m8n8k4", 8, 512.0, k_884<8>)
    REPORT("dmma_m8n8k4", 16, 512.0, k_884<16>)
    REPORT("dmma_m16n8k4", 4, 1024.0, k_1684<4>)
    REPORT("dmma_m16n8k4", 8, 1024.0, k_1684<8>)
    REPORT("dmma_m16n8k4", 16, 1024.0, k_1684<16>)
    REPORT("dmma_m16n8k8", 8, 2048.0, k_1688<8>)
    REPORT("dmma_m16n8k16", 4, 4096.0, k_16816<4>)
    REPORT("dmma_m16n8k16", 8, 4096.0, k_16816<8>)
    REPORT("dfma", 16, 64.0, k_dfma<16>)
    fflush(stdout);
  }
  // latency: 1 warp, 1 accumulator chain
  {
    float ms = time_ms([&] { k_1684<1><<<1, 32>>>(out, 200000, 1.0000001, 0.5); });
    printf("{\"probe\":\"dmma_m16n8k4_chain_latency\",\"ns_per_mma\":%.2f}\n", ms * 1e6 / 200000);
    ms = time_ms([&] { k_884<1><<<1, 32>>>(out, 200000, 1.0000001, 0.5); });
    printf("{\"probe\":\"dmma_m8n8k4_chain_latency\",\"ns_per_mma\":%.2f}\n", ms * 1e6 / 200000);
  }
  // smem-fed variants
  {
    int it2 = 20000;
#define SREPORT(MB, NB, WARPS)                                                                         \
    { int grid = sms; int threads = WARPS * 32;                                                        \
      CK(cudaFuncSetAttribute(k_smemfed<MB, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536)); \
      float ms = time_ms([&] { k_smemfed<MB, NB><<<grid, threads, 65536>>>(out, it2); });              \
      double fl = (double)grid * WARPS * it2 * MB * NB * 1024.0;                                       \
      printf("{\"probe\":\"smemfed_m16n8k4\",\"mb\":%d,\"nb\":%d,\"warps_per_cta\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n", MB, NB, WARPS, ms, fl / ms * 1e-9); }
    SREPORT(1, 13, 8)
    SREPORT(2, 13, 4)
    SREPORT(2, 13, 8)
    SREPORT(2, 7, 8)
    SREPORT(2, 7, 16)
    SREPORT(4, 4, 8)
    SREPORT(4, 7, 8)
    SREPORT(1, 13, 16)
  }
  // sustained: 3 s of the best config to see power-capped clocks
  {
    int grid = sms * 2, threads = 256;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int n = 0;
    for (; n < 60; n++) k_1684<8><<<grid, threads>>>(out, 200000, 1.0000001, 0.5);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = (double)grid * 8 * 200000.0 * 8 * 1024.0 * n;
    printf("{\"probe\":\"dmma_m16n8k4_sustained\",\"seconds\":%.2f,\"tflops\":%.3f}\n", ms * 1e-3, fl / ms * 1e-9);
  }
  return 0;
}
