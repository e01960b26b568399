// Microbenchmark (measurement tool, not on the product path): per-SM throughput of the shared-memory operations the
// partition / dedup kernels are built from, with spread (random) addresses.  nvcc -arch=sm_100a -O3 -o ubench_smem ubench_smem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32;
#define N_IT 4096
__device__ __forceinline__ u32 rnd(u32& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE> __global__ void __launch_bounds__(512) k(u32* out, u32 nslots) {
  extern __shared__ u32 sm[];
  for (u32 i = threadIdx.x; i < nslots; i += blockDim.x) sm[i] = MODE == 1 ? 0xFFFFFFFFu : 0u;
  __syncthreads();
  u32 s = threadIdx.x * 2654435761u + blockIdx.x, acc = 0;
  const u32 mask = nslots - 1;
  for (int it = 0; it < N_IT; ++it) {
    const u32 a = rnd(s) & mask;
    if (MODE == 0) acc += atomicAdd(&sm[a], 1u);                       // ATOMS.ADD with return
    if (MODE == 1) acc += atomicCAS(&sm[a], 0xFFFFFFFFu, a);            // ATOMS.CAS
    if (MODE == 2) { sm[a] = threadIdx.x; acc += sm[(a + 7) & mask]; }  // STS + LDS
    if (MODE == 3) acc += __match_any_sync(0xffffffffu, a & 511);       // MATCH.ANY
    if (MODE == 4) { u32 p = __match_any_sync(0xffffffffu, a & 511); u32 lane = threadIdx.x & 31; u32 lead = __ffs(p) - 1; u32 old = 0;
                     if (lane == lead) { old = sm[a & 511]; sm[a & 511] = old + __popc(p); } old = __shfl_sync(0xffffffffu, old, lead); acc += old + __popc(p & ((1u << lane) - 1)); }
    if (MODE == 5) atomicAdd(&sm[a], 1u);                               // no return (RED-like)
    if (MODE == 6) acc += a;                                            // baseline loop
    if (MODE == 7) { u32 peers = 0xffffffffu; u32 b = a & 511;          // match by 9 ballots
#pragma unroll
                     for (int i = 0; i < 9; ++i) { u32 m = __ballot_sync(0xffffffffu, (b >> i) & 1); peers &= ((b >> i) & 1) ? m : ~m; } acc += peers; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + sm[threadIdx.x & mask];
}
template <int MODE> void run(const char* name, int ctas_per_sm, u32 nslots) {
  u32* out; cudaMalloc(&out, 148 * 8 * 512 * 4);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * ctas_per_sm, 512, nslots * 4>>>(out, nslots);
  cudaEventRecord(e0);
  k<MODE><<<148 * ctas_per_sm, 512, nslots * 4>>>(out, nslots);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops_per_sm = (double)ctas_per_sm * 512 * N_IT;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-28s ctas/SM=%d slots=%5u  %.3f ms  %.2f cycles/lane-op/SM (at %d MHz)  err=%s\n", name, ctas_per_sm, nslots, ms, ms * 1e-3 * clk * 1e3 / ops_per_sm, clk / 1000,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}
int main() {
  for (int c = 1; c <= 3; c += 2) {
    run<6>("baseline loop", c, 8192);
    run<0>("ATOMS.ADD ret spread", c, 512);
    run<0>("ATOMS.ADD ret spread", c, 8192);
    run<5>("ATOMS.ADD noret spread", c, 512);
    run<1>("ATOMS.CAS spread", c, 8192);
    run<2>("STS+LDS spread", c, 8192);
    run<3>("MATCH.ANY", c, 512);
    run<4>("MATCH + leader LDS/STS + SHFL", c, 512);
    run<7>("match by 9 ballots", c, 512);
  }
  return 0;
}
