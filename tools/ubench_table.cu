// Micro-benchmark (development aid): rate of CAS-inserts into an open-addressing table as a function of the
// table size (L2-resident vs HBM-resident), and of bitmap atomicOr scatter.  nvcc -O3 -arch=sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64; typedef unsigned int u32;
__device__ __forceinline__ u64 mix(u64 h) { h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33; return h; }
__global__ void k_insert(const u64* __restrict__ keys, u64 n, u64* table, u64 mask, u32* __restrict__ out) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 key = keys[i];
  u64 idx = (key >> 24) & mask, entry = (key << 40) | i;
  for (;;) { u64 cur = atomicCAS(table + idx, ~0ull, entry); if (cur == ~0ull) break; idx = (idx + 1) & mask; }
  out[i] = (u32)idx;
}
__global__ void k_gen(u64* keys, u64 n, u64 seed) { u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) keys[i] = mix(i * 0x9E3779B97F4A7C15ull + seed); }
__global__ void k_bits(const u64* __restrict__ keys, u64 n, u32* bits, u64 nbits_mask) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 p = keys[i] & nbits_mask;
  atomicOr(bits + (p >> 5), 1u << (p & 31));
}
int main() {
  const u64 n = 1ull << 21;   // records per "bucket"
  u64* keys; u32* out; cudaMalloc(&keys, n * 8); cudaMalloc(&out, n * 4);
  k_gen<<<(n + 255) / 256, 256>>>(keys, n, 1);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int lg = 19; lg <= 30; lg += (lg < 26 ? 1 : 2)) {
    u64 cap = 1ull << lg; if (cap < 2 * n && lg < 22) continue;
    u64* table; if (cudaMalloc(&table, cap * 8) != cudaSuccess) break;
    float best = 1e9, tot = 0; int reps = 20;
    for (int r = 0; r < reps; ++r) {
      cudaMemsetAsync(table, 0xFF, cap * 8);
      cudaEventRecord(a);
      k_insert<<<(n + 255) / 256, 256>>>(keys, n, table, cap - 1, out);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); if (r > 2) { tot += ms; if (ms < best) best = ms; }
    }
    printf("insert n=%llu cap=2^%d (%.0f MB): best %.1f us avg %.1f us -> %.2f G inserts/s\n", n, lg, cap * 8 / 1e6, best * 1e3, tot / (reps - 3) * 1e3, n / (best * 1e-3) / 1e9);
    cudaFree(table);
  }
  // bitmap scatter
  for (int lg = 24; lg <= 32; lg += 2) {
    u64 nbits = 1ull << lg; u32* bits; cudaMalloc(&bits, nbits / 8);
    cudaMemset(bits, 0, nbits / 8);
    float best = 1e9;
    for (int r = 0; r < 10; ++r) { cudaEventRecord(a); k_bits<<<(n + 255) / 256, 256>>>(keys, n, bits, nbits - 1); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    printf("atomicOr scatter n=%llu bitmap=%.0f MB: best %.1f us -> %.2f G/s\n", n, nbits / 8 / 1e6, best * 1e3, n / (best * 1e-3) / 1e9);
    cudaFree(bits);
  }
  return 0;
}
