// superpang_b200 -- platform layer.
//
// The library is CUDA for sm_100a.  The same kernel sources can ALSO be compiled by g++ with
// -DSPB_EMU, where every "kernel" is executed thread-by-thread on the host.  That build
// (tests/emu/libspb_emu.so) exists only so that the CPU-only CI tier (`pytest -m "not gpu"`) can
// check the kernels' *logic* against the oracle; it is never built into, or loaded by, the product
// package, which fails loudly when libsuperpang_b200.so (the nvcc build) is missing.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <cstdlib>

typedef unsigned long long u64;
static unsigned long long g_spb_launches = 0;   // number of this library's own kernel launches (host side)
typedef uint32_t u32;
typedef uint8_t u8;

#ifdef SPB_EMU
// ---------------------------------------------------------------- host emulation of the kernels
#include <algorithm>
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define SPB_LAUNCH_BOUNDS(n)
#define SPB_LAUNCH_BOUNDS2(n, m)
struct spb_emu_dim3 { unsigned x = 1, y = 1, z = 1; };
static spb_emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
typedef void* spb_stream_t;
// Threads run one after the other.  SPB_EMU_ORDER=reverse runs them in descending order and =shuffle in a
// pseudo-random block order, so that the tests also exercise the paths that only trigger when a larger position
// reaches the table before a smaller one (on the GPU that depends on scheduling).
template <class F> static inline void spb_emu_launch(size_t nb, unsigned nt, F f) {
  gridDim.x = (unsigned)nb; blockDim.x = nt;
  const char* ord = getenv("SPB_EMU_ORDER");
  int mode = !ord ? 0 : (ord[0] == 'r' ? 1 : (ord[0] == 's' ? 2 : 0));
  for (size_t i = 0; i < nb; ++i) {
    size_t b = mode == 0 ? i : mode == 1 ? nb - 1 - i : (size_t)(((unsigned long long)i * 2654435761ull + 12345) % nb);
    if (mode == 2) { b = i; }   // shuffle handled below per thread to stay a permutation
    blockIdx.x = (unsigned)b;
    for (unsigned j = 0; j < nt; ++j) {
      unsigned t = mode == 0 ? j : mode == 1 ? nt - 1 - j : (unsigned)((j * 167u + 13u) % nt);
      threadIdx.x = t; f();
    }
  }
}
#define SPB_LAUNCH(K, NB, NT, STREAM, ...) \
  do { if ((NB) > 0) { ++g_spb_launches; spb_emu_launch((size_t)(NB), (unsigned)(NT), [&]() { K(__VA_ARGS__); }); } } while (0)
// Block-cooperative kernels are written as phases separated by block barriers; anything that crosses a barrier lives
// in SPB_SHARED memory.  The emulator runs every phase for all threads of a block before the next phase starts.
static int spb_emu_phase = 0;
#define SPB_SHARED static
#define SPB_PHASE_BEGIN(n) if (spb_emu_phase == (n)) {
#define SPB_PHASE_END }
template <class F> static inline void spb_emu_launch_phased(int nph, size_t nb, unsigned nt, F f) {
  gridDim.x = (unsigned)nb; blockDim.x = nt;
  const char* ord = getenv("SPB_EMU_ORDER");
  int mode = !ord ? 0 : (ord[0] == 'r' ? 1 : (ord[0] == 's' ? 2 : 0));
  for (size_t i = 0; i < nb; ++i) {
    blockIdx.x = (unsigned)(mode == 1 ? nb - 1 - i : i);
    for (int ph = 0; ph < nph; ++ph) {
      spb_emu_phase = ph;
      for (unsigned j = 0; j < nt; ++j) { threadIdx.x = mode == 0 ? j : mode == 1 ? nt - 1 - j : (unsigned)((j * 167u + 13u) % nt); f(); }
    }
  }
  spb_emu_phase = 0;
}
#define SPB_LAUNCH_PHASED(K, NPH, NB, NT, STREAM, ...) \
  do { if ((NB) > 0) { ++g_spb_launches; spb_emu_launch_phased((NPH), (size_t)(NB), (unsigned)(NT), [&]() { K(__VA_ARGS__); }); } } while (0)
// block-cooperative kernel whose shared array exceeds the static limit (dynamic shared memory on the GPU)
#define SPB_DYN_SHARED(type, name, n) static type name[n]
#define SPB_LAUNCH_PHASED_DSMEM(K, NPH, NB, NT, SMEM, STREAM, ...) SPB_LAUNCH_PHASED(K, NPH, NB, NT, STREAM, __VA_ARGS__)
static inline u64 atomicCAS(u64* a, u64 cmp, u64 val) { u64 o = *a; if (o == cmp) *a = val; return o; }
static inline u32 atomicCAS(u32* a, u32 cmp, u32 val) { u32 o = *a; if (o == cmp) *a = val; return o; }
static inline u64 atomicMin(u64* a, u64 v) { u64 o = *a; if (v < o) *a = v; return o; }
static inline u32 atomicMin(u32* a, u32 v) { u32 o = *a; if (v < o) *a = v; return o; }
static inline u64 atomicAdd(u64* a, u64 v) { u64 o = *a; *a = o + v; return o; }
static inline u32 atomicAdd(u32* a, u32 v) { u32 o = *a; *a = o + v; return o; }
static inline u32 atomicOr(u32* a, u32 v) { u32 o = *a; *a = o | v; return o; }
static inline int __all_sync(unsigned, int p) { return p; }     // one emulated thread at a time
static inline int __popcll(u64 x) { return __builtin_popcountll(x); }
static inline u32 __funnelshift_l(u32 lo, u32 hi, u32 sh) { sh &= 31; return sh ? (hi << sh) | (lo >> (32 - sh)) : hi; }
static inline u32 __funnelshift_r(u32 lo, u32 hi, u32 sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
static inline int __popc(u32 x) { return __builtin_popcount(x); }
static inline u32 __brev(u32 x) { u32 y = 0; for (int i = 0; i < 32; ++i) y |= ((x >> i) & 1u) << (31 - i); return y; }
struct uint2 { u32 x, y; };
static inline uint2 make_uint2(u32 a, u32 b) { uint2 r; r.x = a; r.y = b; return r; }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
#else
// ---------------------------------------------------------------- CUDA (sm_100a)
#include <cuda_runtime.h>
#define SPB_LAUNCH_BOUNDS(n) __launch_bounds__(n)
#define SPB_LAUNCH_BOUNDS2(n, m) __launch_bounds__(n, m)
typedef cudaStream_t spb_stream_t;
#define SPB_LAUNCH(K, NB, NT, STREAM, ...) \
  do { if ((NB) > 0) { ++g_spb_launches; K<<<(unsigned)(NB), (unsigned)(NT), 0, (STREAM)>>>(__VA_ARGS__); } } while (0)
#define SPB_SHARED __shared__
#define SPB_PHASE_BEGIN(n) { if ((n) > 0) __syncthreads();
#define SPB_PHASE_END }
#define SPB_LAUNCH_PHASED(K, NPH, NB, NT, STREAM, ...) SPB_LAUNCH(K, NB, NT, STREAM, __VA_ARGS__)
#define SPB_DYN_SHARED(type, name, n) extern __shared__ type name[]
#define SPB_LAUNCH_PHASED_DSMEM(K, NPH, NB, NT, SMEM, STREAM, ...) \
  do { if ((NB) > 0) { ++g_spb_launches; cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM)); \
       cudaFuncSetAttribute(K, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
       K<<<(unsigned)(NB), (unsigned)(NT), (SMEM), (STREAM)>>>(__VA_ARGS__); } } while (0)
#endif

// global linear thread id (64-bit)
#define SPB_GTID() ((u64)blockIdx.x * blockDim.x + threadIdx.x)
