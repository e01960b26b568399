// superpang_b200 -- host orchestration + C-ABI (include/superpang_b200.h).
//
// Built by nvcc for sm_100a into libsuperpang_b200.so (the product), or by g++ -DSPB_EMU into
// tests/emu/libspb_emu.so (kernel-logic emulator for the CPU-only test tier; never shipped).
#include "spb_kernels.cuh"
#include "spb_partition.cuh"
#include "../../include/superpang_b200.h"

#include <cstdio>
#include <cstdarg>
#include <vector>
#include <string>
#include <algorithm>
#include <map>

#ifndef SPB_EMU
#include <cub/cub.cuh>
#endif

// =========================================================================================== runtime
struct spb_ctx {
  int device = 0;
  spb_stream_t stream = 0;
  char err[512] = {0};
  int state = 0;                       // 0 created, 1 loaded, 2 dbg built, 3 compacted
  u32 k = 0, n_seq = 0;
  u64 T = 0, n_inst = 0, npos_pad = 0, nw = 0;
  std::vector<u64> seq_off_h;
  spb_counts cnt;
  std::map<void*, u64> live;           // live device blocks -> size
  std::multimap<u64, void*> cache;     // freed blocks kept for reuse (steady state: no cudaMalloc at all)
  u64* pinned = 0;                     // pinned host scratch for scalar read-backs
#ifndef SPB_EMU
  cudaStream_t stream2[3] = {0, 0, 0}; cudaEvent_t ev_fork = 0, ev_join[3] = {0, 0, 0};   // extra streams for the bucket loops
#endif
  // ---- persistent device state
  u64* seq_off = 0;
  u32 *F = 0, *R = 0, *vbits = 0, *obits = 0, *fbits = 0, *sp = 0, *v2pos = 0, *seqlim = 0, *errflag = 0;
  bool mg_owned = false;                            // spb_mg_owned_links ran (the orientation bitmap is complete on this rank)
  u8* vflags = 0;
  u32 nV = 0, nJ = 0;
  u64* Jk = 0; u8* Js = 0;
  u32* edges = 0;
  // compaction
  u32 nI = 0, nP = 0, nD = 0, nN = 0, nO = 0, nM = 0, nIv = 0;
  u32 *init_list = 0, *ibase = 0, *plo = 0, *phi = 0, *dart_nbp = 0, *vnbp = 0, *parent = 0, *label_of_root = 0;
  NbpTab nt; int32_t* nbp_comp = 0;
  u64 *voff = 0, *soff = 0, *ooff = 0;
  u32 *verts = 0, *origins = 0; u8* seqs = 0;
  u32* iv = 0; u64* iv_off = 0;
  u64* g_off = 0; u32 *g_succ = 0, *g_rev = 0; u64 g_edges = 0;      // NBP graph (lazy)
  u32 *occ_counts = 0, *occ_ids = 0; u64* occ_off = 0; u64 occ_pairs = 0;   // NBP occurrences in the sequences (lazy)
  u64* ivp = 0; u32 n_ivp = 0, ivp_sb = 0;                          // (init vertex, sequence) pairs (lazy)
  // staged dedup state (records sent by this rank)
  u64 *dd_keys = 0, *dd_bstart = 0; u32* dd_pos = 0; u64 dd_n = 0, dd_lo = 0, dd_hi = 0; u32 dd_bb = 0; bool dd_fbits_ready = false;
  std::vector<u64> dd_bstart_h;
  u64* mg_links = 0; u32 *mg_seen = 0, *mg_dup = 0;                 // multi-GPU links protocol
  // rank tables of the first-appearance bitmap: exclusive block offsets (256 positions) and per-word prefixes
  u32* blkoff = 0; u8* wpre = 0; u64 nb = 0;
  // Sliced multi-GPU tail: what THIS context computes of the per-position, per-vertex and per-NBP
  // passes.  on = false (one GPU, or the replicated tail): everything.
  struct Slice {
    bool on = false; u32 rank = 0, G = 1;
    u64 pos_lo = 0, pos_hi = 0, ids_lo = 0, ids_hi = 0;    // own positions; positions whose ids are filled (one more on either side)
    u32 v_lo = 0, v_hi = 0;                                // vertices whose first appearance lies in [pos_lo, pos_hi): v2pos is filled for them
    u32 t_lo = 0, t_hi = 0;                                // own tiles of 8 vertices: classification, piece lists, edge rows, vertex -> NBP map
    u32 n_lo = 0, n_hi = 0;                                // own raw NBPs: vertex lists and sequences
    u64 vert_lo = 0, vert_n = 0, seq_lo = 0, seq_n = 0, edge_n = 0;
  } sl;
  // NBP sequences straight to the caller's page-locked host buffer (spb_set_host_output): copied on a second stream as
  // soon as they are emitted, while origins / components / edge rows are still being computed
  u8* out_seq_host = 0; u64 out_seq_cap = 0; bool d2h_pending = false;
#ifndef SPB_EMU
  cudaStream_t copy_stream = 0; cudaEvent_t ev_emit = 0, ev_d2h = 0;
#endif
  // explicit edge rows: counted at DBG build (n_edges), written after the NBP emission (or at the first getter that needs them)
  u32* edge_off = 0; bool edges_pending = false;
  u32* shard_tmp = 0;                                     // last spb_mg_shard(1) result
  u8* seqs_base = 0;                                      // allocation behind `seqs` (the slice keeps the 4-byte phase of the global offsets)
  // compaction state that lives across the phases (classification -> contracted graph -> NBP table -> emission -> origins)
  u32 *pl_nbr = 0, *pr_nbr = 0, *nxt = 0, *rk_ptr[2] = {0, 0}, *rk_acc[2] = {0, 0}, *rk_root[2] = {0, 0}, *rk_mnv[2] = {0, 0};
  u8 *pl_side = 0, *pr_side = 0;
  int rk_cur = 0; u32 ncyc = 0;
  u64 *op = 0, *mp = 0; u64 n_op = 0, n_mp = 0, op_cap = 0, mp_cap = 0; u8* has01 = 0;
  // per-stage statistics: scalar values and (CUDA) event pairs resolved lazily by spb_get_stage_stats, so that timing a
  // stage never synchronises the host with the stream
  struct StatRec { std::string name; double val; int ev; };
  std::vector<StatRec> stat;
  size_t ev_used = 0;
  u64 n_syncs = 0;                     // host <-> stream synchronisations since the last stats reset
#ifndef SPB_EMU
  std::vector<cudaEvent_t> evpool;
#endif
};

static int fail(spb_ctx* c, int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(c->err, sizeof c->err, fmt, ap); va_end(ap);
  return code;
}

#ifdef SPB_EMU
#define CK(x) do { (void)(x); } while (0)
static void* backend_alloc(u64 bytes) { return malloc(bytes); }
static void backend_free(void* p) { free(p); }
static void dev_memset(spb_ctx*, void* p, int v, u64 bytes) { memset(p, v, bytes); }
static void dev_copy(spb_ctx*, void* dst, const void* src, u64 bytes) { memcpy(dst, src, bytes); }
static void dev_sync(spb_ctx* c) { ++c->n_syncs; }
static int t_begin(spb_ctx*) { return 0; }
static void t_end(spb_ctx* c, int, const char* name) { c->stat.push_back({name, 0.0, -1}); }
#else
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(c, SPB_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); } while (0)
static void* backend_alloc(u64 bytes) {
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
static void backend_free(void* p) { cudaFree(p); }
static void dev_memset(spb_ctx* c, void* p, int v, u64 bytes) { cudaMemsetAsync(p, v, bytes, c->stream); }
static void dev_copy(spb_ctx* c, void* dst, const void* src, u64 bytes) { cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, c->stream); }
static void dev_sync(spb_ctx* c) { ++c->n_syncs; cudaStreamSynchronize(c->stream); }
static int t_begin(spb_ctx* c) {                      // start of a timed stage: returns the index of its event pair
  if (c->ev_used + 2 > c->evpool.size()) { cudaEvent_t e; for (int i = 0; i < 2; ++i) { cudaEventCreate(&e); c->evpool.push_back(e); } }
  const int t = (int)c->ev_used; c->ev_used += 2;
  cudaEventRecord(c->evpool[t], c->stream);
  return t;
}
static void t_end(spb_ctx* c, int t, const char* name) { cudaEventRecord(c->evpool[t + 1], c->stream); c->stat.push_back({name, 0.0, t}); }
#endif
static void stat_set(spb_ctx* c, const char* name, double v) { c->stat.push_back({name, v, -1}); }
static void stats_reset(spb_ctx* c) { c->stat.clear(); c->ev_used = 0; }

// Caching block allocator.  Every buffer of the pipeline is used on the context's single stream, so a
// freed block can be handed out again immediately (stream order == program order).  After the first
// pass over a workload the allocation sequence repeats exactly and no cudaMalloc / cudaFree happens.
static void cache_release(spb_ctx* c) {
  for (auto& kv : c->cache) backend_free(kv.second);
  c->cache.clear();
}
static void* dev_alloc_raw(spb_ctx* c, u64 bytes) {
  bytes = (std::max<u64>(bytes, 1) + 511) & ~511ull;
  auto it = c->cache.lower_bound(bytes);
  if (it != c->cache.end() && it->first <= bytes + bytes / 4 + (1u << 20)) {
    void* p = it->second; u64 sz = it->first;
    c->cache.erase(it);
    c->live[p] = sz;
    return p;
  }
  void* p = backend_alloc(bytes);
  if (!p) { dev_sync(c); cache_release(c); p = backend_alloc(bytes); }
  if (p) c->live[p] = bytes;
  return p;
}
static void dev_free_raw(spb_ctx* c, void* p) {
  auto it = c->live.find(p);
  if (it == c->live.end()) return;
  c->cache.insert({it->second, p});
  c->live.erase(it);
}

// Small read-backs (counters, scan totals) into the context's pinned scratch.  On the GPU they are written by a tiny kernel
// through the mapped host pointer instead of a copy-engine transfer: a cudaMemcpy would queue behind the bulk device-to-host
// copy of the NBP sequences that spb_compact keeps in flight on its second stream, and every host sync behind it would
// wait for the whole gigabyte.
#ifdef SPB_EMU
static void pin_read(spb_ctx*, void* pinned_dst, const void* src, u64 bytes) { memcpy(pinned_dst, src, bytes); }
#else
__global__ void k_pin_copy(const u32* __restrict__ src, u32* dst, u32 nwords) {
  for (u32 i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
}
static void pin_read(spb_ctx* c, void* pinned_dst, const void* src, u64 bytes) {      // bytes: a multiple of 4, src 4-byte aligned
  ++g_spb_launches;
  k_pin_copy<<<1, 128, 0, c->stream>>>((const u32*)src, (u32*)pinned_dst, (u32)(bytes / 4));
}
#endif
template <class T> static T* dalloc(spb_ctx* c, u64 n) { return (T*)dev_alloc_raw(c, n * sizeof(T) + 64); }
template <class T> static void dfree(spb_ctx* c, T*& p) {
  if (!p) return;
  dev_free_raw(c, (void*)p);
  p = nullptr;
}
#define ALLOC(ptr, T, n) do { (ptr) = dalloc<T>(c, (n)); if (!(ptr)) return fail(c, SPB_ENOMEM, "device allocation of %llu bytes failed (%s)", (unsigned long long)((n) * sizeof(T)), #ptr); } while (0)
// scalar read-back through pinned host memory (one stream sync)
template <class T> static T dread(spb_ctx* c, const T* p) {
  static_assert(sizeof(T) % 4 == 0, "");
  pin_read(c, c->pinned, p, sizeof(T)); dev_sync(c);
  T v; memcpy(&v, c->pinned, sizeof(T)); return v;
}

// The bucket loops keep two L2-resident tables in flight: even buckets on the context's stream, odd ones on a second
// stream (each bucket: clear its table, then insert).  fork/join order the second stream against the first.
struct TwoStreams {
  spb_stream_t st[4];
  int n = 2;                                         // tables / streams in flight (SPB_STREAMS=1..4, default 2: 2 x 32 MB fit in L2)
#ifdef SPB_EMU
  explicit TwoStreams(spb_ctx* c) { for (auto& x : st) x = c->stream; if (const char* e = getenv("SPB_STREAMS")) n = std::max(1, std::min(4, atoi(e))); }
  void clear(spb_ctx*, int, void* p, u64 bytes) { memset(p, 0xFF, bytes); }
  void join(spb_ctx*) {}
#else
  explicit TwoStreams(spb_ctx* c) {
    if (const char* e = getenv("SPB_STREAMS")) n = std::max(1, std::min(4, atoi(e)));
    if (!c->stream2[0]) {
      for (int i = 0; i < 3; ++i) { cudaStreamCreateWithFlags(&c->stream2[i], cudaStreamNonBlocking); cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming); }
      cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    }
    st[0] = c->stream; for (int i = 1; i < 4; ++i) st[i] = c->stream2[i - 1];
    cudaEventRecord(c->ev_fork, c->stream);
    for (int i = 1; i < n; ++i) cudaStreamWaitEvent(st[i], c->ev_fork, 0);
  }
  void clear(spb_ctx*, int i, void* p, u64 bytes) { cudaMemsetAsync(p, 0xFF, bytes, st[i]); }
  void join(spb_ctx* c) { for (int i = 1; i < n; ++i) { cudaEventRecord(c->ev_join[i - 1], st[i]); cudaStreamWaitEvent(c->stream, c->ev_join[i - 1], 0); } }
#endif
};

static inline u64 nblk(u64 n, u32 t) { return (n + t - 1) / t; }
// grid of a grid-stride kernel: enough resident blocks for every SM, never millions of tiny ones
static inline u64 gs_grid(u64 n) { u64 b = (n + 255) / 256; return b < 148 * 32 ? b : 148 * 32; }
#define TPB 256

// =========================================================================================== primitives
// Device-wide scan / sort / select.  CUDA: CUB (CCCL shipped with CUDA 12.9).  Emulator: <algorithm>.
#ifdef SPB_EMU
static int prim_scan_u32(spb_ctx*, const u32* in, u32* out, u64 n, u64* total) {
  u64 s = 0; for (u64 i = 0; i < n; ++i) { u32 v = in[i]; out[i] = (u32)s; s += v; } *total = s; return 0;
}
static int prim_scan_u64(spb_ctx*, const u64* in, u64* out, u64 n, u64* total) {
  u64 s = 0; for (u64 i = 0; i < n; ++i) { u64 v = in[i]; out[i] = s; s += v; } *total = s; return 0;
}
static int prim_sort_u64_u8(spb_ctx*, u64*& k, u8*& v, u64*& k2, u8*& v2, u64 n) {
  std::vector<u64> idx(n); for (u64 i = 0; i < n; ++i) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](u64 a, u64 b) { return k[a] < k[b]; });
  for (u64 i = 0; i < n; ++i) { k2[i] = k[idx[i]]; v2[i] = v[idx[i]]; }
  std::swap(k, k2); std::swap(v, v2); return 0;
}
static int prim_sort_u64_u64(spb_ctx*, u64*& k, u64*& v, u64*& k2, u64*& v2, u64 n) {
  std::vector<u64> idx(n); for (u64 i = 0; i < n; ++i) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](u64 a, u64 b) { return k[a] < k[b]; });
  for (u64 i = 0; i < n; ++i) { k2[i] = k[idx[i]]; v2[i] = v[idx[i]]; }
  std::swap(k, k2); std::swap(v, v2); return 0;
}
static int prim_sort_u64(spb_ctx*, u64*& k, u64* k2, u64 n, int end_bit = 64) { std::sort(k, k + n); (void)k2; (void)end_bit; return 0; }
static int prim_partition_u64_u32(spb_ctx*, u64*& k, u32*& v, u64*& k2, u32*& v2, u64 n, int b0, int b1) {   // stable, by key bits [b0, b1)
  std::vector<u64> idx(n); for (u64 i = 0; i < n; ++i) idx[i] = i;
  u64 mask = (b1 - b0 >= 64) ? ~0ull : ((1ull << (b1 - b0)) - 1);
  std::stable_sort(idx.begin(), idx.end(), [&](u64 a, u64 b) { return ((k[a] >> b0) & mask) < ((k[b] >> b0) & mask); });
  for (u64 i = 0; i < n; ++i) { k2[i] = k[idx[i]]; v2[i] = v[idx[i]]; }
  std::swap(k, k2); std::swap(v, v2); return 0;
}
static int prim_select_flag(spb_ctx*, const u8* flags, u8 mask, u64 n, u32* out, u32* count) {
  u32 m = 0; for (u64 i = 0; i < n; ++i) if (flags[i] & mask) out[m++] = (u32)i; *count = m; return 0;
}
static int prim_count_flag(spb_ctx*, const u8* flags, u8 mask, u64 n, u32* count) {
  u32 m = 0; for (u64 i = 0; i < n; ++i) if (flags[i] & mask) ++m; *count = m; return 0;
}
static int prim_segmax_u64(spb_ctx*, const u64* in, u64* out, u64 n) {   // running max within equal high-32 segments
  for (u64 i = 0; i < n; ++i) out[i] = (i && (in[i] >> 32) == (out[i - 1] >> 32) && out[i - 1] > in[i]) ? out[i - 1] : in[i];
  return 0;
}
#else
struct FlagPred { const u8* f; u8 m; __device__ bool operator()(u32 i) const { return (f[i] & m) != 0; } };
struct SegMaxOp { __device__ u64 operator()(u64 a, u64 b) const { return ((a >> 32) == (b >> 32) && a > b) ? a : b; } };
#define CUB2(call_with_tmp) do { void* tmp = nullptr; size_t tb = 0; CK(call_with_tmp); tmp = dev_alloc_raw(c, tb); \
    if (!tmp) return fail(c, SPB_ENOMEM, "cub temp storage (%zu bytes)", tb); cudaError_t e2_ = (call_with_tmp); dev_free_raw(c, tmp); CK(e2_); } while (0)
static int prim_scan_u32(spb_ctx* c, const u32* in, u32* out, u64 n, u64* total) {
  if (n == 0) { *total = 0; return 0; }
  CUB2(cub::DeviceScan::ExclusiveSum(tmp, tb, in, out, (size_t)n, c->stream));
  u32* pin = (u32*)c->pinned;                        // read back through pinned memory: one stream sync, no staging copy
  pin_read(c, pin, in + n - 1, 4); pin_read(c, pin + 1, out + n - 1, 4); dev_sync(c);
  *total = (u64)pin[0] + pin[1]; return 0;
}
static int prim_scan_u64(spb_ctx* c, const u64* in, u64* out, u64 n, u64* total) {
  if (n == 0) { *total = 0; return 0; }
  CUB2(cub::DeviceScan::ExclusiveSum(tmp, tb, in, out, (size_t)n, c->stream));
  pin_read(c, c->pinned, in + n - 1, 8); pin_read(c, c->pinned + 1, out + n - 1, 8); dev_sync(c);
  *total = c->pinned[0] + c->pinned[1]; return 0;
}
static int prim_sort_u64_u8(spb_ctx* c, u64*& k, u8*& v, u64*& k2, u8*& v2, u64 n) {
  if (n == 0) return 0;
  cub::DoubleBuffer<u64> dk(k, k2); cub::DoubleBuffer<u8> dv(v, v2);
  CUB2(cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, (size_t)n, 0, 64, c->stream));
  if (dk.Current() != k) { std::swap(k, k2); }
  if (dv.Current() != v) { std::swap(v, v2); }
  return 0;
}
static int prim_sort_u64_u64(spb_ctx* c, u64*& k, u64*& v, u64*& k2, u64*& v2, u64 n) {
  if (n == 0) return 0;
  cub::DoubleBuffer<u64> dk(k, k2); cub::DoubleBuffer<u64> dv(v, v2);
  CUB2(cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, (size_t)n, 0, 64, c->stream));
  if (dk.Current() != k) { std::swap(k, k2); }
  if (dv.Current() != v) { std::swap(v, v2); }
  return 0;
}
static int prim_partition_u64_u32(spb_ctx* c, u64*& k, u32*& v, u64*& k2, u32*& v2, u64 n, int b0, int b1) {
  if (n == 0) return 0;
  cub::DoubleBuffer<u64> dk(k, k2); cub::DoubleBuffer<u32> dv(v, v2);
  CUB2(cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, (size_t)n, b0, b1, c->stream));
  if (dk.Current() != k) { std::swap(k, k2); }
  if (dv.Current() != v) { std::swap(v, v2); }
  return 0;
}
static int prim_sort_u64(spb_ctx* c, u64*& k, u64* k2, u64 n, int end_bit = 64) {   // keys are known to fit in end_bit bits
  if (n == 0) return 0;
  cub::DoubleBuffer<u64> dk(k, k2);
  CUB2(cub::DeviceRadixSort::SortKeys(tmp, tb, dk, (size_t)n, 0, end_bit, c->stream));
  if (dk.Current() != k) {            // keep the result in the caller's primary buffer
    CK(cudaMemcpyAsync(k, k2, n * 8, cudaMemcpyDeviceToDevice, c->stream));
  }
  return 0;
}
static int prim_select_flag(spb_ctx* c, const u8* flags, u8 mask, u64 n, u32* out, u32* count) {
  *count = 0;
  if (n == 0) return 0;
  u32* dcount = (u32*)dev_alloc_raw(c, 8);
  if (!dcount) return fail(c, SPB_ENOMEM, "select count");
  cub::CountingInputIterator<u32> it(0);
  FlagPred pred{flags, mask};
  CUB2(cub::DeviceSelect::If(tmp, tb, it, out, dcount, (int)n, pred, c->stream));
  *count = dread(c, dcount);
  dev_free_raw(c, dcount);
  return 0;
}
__global__ void k_count_flag(const u8* __restrict__ f, u8 m, u64 n, u32* cnt) {
  u64 i = SPB_GTID();
  bool b = i < n && (f[i] & m);
  u32 bal = __ballot_sync(0xffffffffu, b);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(cnt, (u32)__popc(bal));
}
static int prim_count_flag(spb_ctx* c, const u8* flags, u8 mask, u64 n, u32* count) {
  u32* d = (u32*)dev_alloc_raw(c, 8);
  if (!d) return fail(c, SPB_ENOMEM, "count");
  dev_memset(c, d, 0, 4);
  SPB_LAUNCH(k_count_flag, nblk(n, TPB), TPB, c->stream, flags, mask, n, d);
  *count = dread(c, d);
  dev_free_raw(c, d);
  return 0;
}
static int prim_segmax_u64(spb_ctx* c, const u64* in, u64* out, u64 n) {
  if (n == 0) return 0;
  CUB2(cub::DeviceScan::InclusiveScan(tmp, tb, in, out, SegMaxOp(), (size_t)n, c->stream));
  return 0;
}
#endif
#define PRIM(x) do { int r_ = (x); if (r_) return r_; } while (0)

// sorted u64 keys (+ optional u8 payload) -> unique, in place; checks payload agreement
static int unique_sorted(spb_ctx* c, u64*& keys, u8*& pay, u64 n, u32* n_out) {
  *n_out = 0;
  if (n == 0) return 0;
  u32 *head, *pos; u64* ok; u8* op = nullptr;
  ALLOC(head, u32, n); ALLOC(pos, u32, n); ALLOC(ok, u64, n);
  if (pay) ALLOC(op, u8, n);
  SPB_LAUNCH(k_unique_heads, nblk(n, TPB), TPB, c->stream, keys, pay, n, head, c->errflag);
  u64 total = 0;
  PRIM(prim_scan_u32(c, head, pos, n, &total));
  SPB_LAUNCH(k_unique_scatter, nblk(n, TPB), TPB, c->stream, keys, pay, n, head, pos, ok, op);
  dfree(c, head); dfree(c, pos); dfree(c, keys); keys = ok;
  if (pay) { dfree(c, pay); pay = op; }
  *n_out = (u32)total;
  return 0;
}

static int check_err(spb_ctx* c, const char* where) {
  u32 e = dread(c, c->errflag);
  if (e & ERR_BADCHAR) return fail(c, SPB_EINVAL, "%s: input contains characters other than A/C/G/T", where);
  if (e & ERR_DEGENERATE)
    return fail(c, SPB_EDEGENERATE, "%s: degenerate input (k-mer overlaps are not orientation-unique; the reference asserts on it, lib/Assembler.py:1088/:1104)", where);
  if (e & ERR_OVERFLOW) return fail(c, SPB_ENOMEM, "%s: internal list overflow", where);
#ifndef SPB_EMU
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return fail(c, SPB_ECUDA, "%s: %s", where, cudaGetErrorString(ce));
#endif
  return 0;
}

// one range of positions through the dedup table; anchored = pass 1 (every 32nd position) + pass 2 (extension test)
template <int W> static void launch_extract(spb_ctx* c, u64 lo, u64 hi, u64* slots, u64 cap, u32* seenpos, u32* wbits, u32* dbits, u64* arec,
                                            u64* ext_count) {
  if (hi <= lo) return;
  if (arec) SPB_LAUNCH(k_anchor_insert<W>, nblk(nblk(hi - lo, 32), TPB), TPB, c->stream, c->F, c->R, c->vbits, c->T, c->k, slots, cap - 1, seenpos,
                       dbits, arec, lo, hi);
  SPB_LAUNCH(k_extract_insert<W>, nblk(hi - lo, TPB), TPB, c->stream, c->F, c->R, c->vbits, c->T, c->k, slots, cap - 1, seenpos, c->obits, wbits,
             dbits, (const u64*)arec, lo, hi, ext_count);
}
static void launch_extract_k(spb_ctx* c, u64 lo, u64 hi, u64* slots, u64 cap, u32* seenpos, u32* wbits, u32* dbits, u64* arec, u64* ext_count) {
  switch ((2 * c->k + 31) / 32) {
    case 7: launch_extract<7>(c, lo, hi, slots, cap, seenpos, wbits, dbits, arec, ext_count); break;       // k = 101
    case 19: launch_extract<19>(c, lo, hi, slots, cap, seenpos, wbits, dbits, arec, ext_count); break;     // k = 301
    case 32: launch_extract<32>(c, lo, hi, slots, cap, seenpos, wbits, dbits, arec, ext_count); break;     // k = 501
    default: launch_extract<0>(c, lo, hi, slots, cap, seenpos, wbits, dbits, arec, ext_count); break;
  }
}

// Entry points make the context's device current and give the caller's device back when they return (a host process
// that drives another GPU with its own runtime calls must not find its current device switched underneath it).
#ifndef SPB_EMU
struct DevGuard {
  int prev = -1;
  explicit DevGuard(int dev) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define SPB_ON_DEVICE(c) DevGuard spb_dev_guard_((c)->device)
#else
#define SPB_ON_DEVICE(c) do {} while (0)
#endif
// =========================================================================================== C-ABI
extern "C" {

const char* spb_build_kind(void) {
#ifdef SPB_EMU
  return "emu";
#else
  return "cuda sm_100a";
#endif
}

int spb_create(spb_ctx** out, int device, void* cuda_stream) {
  if (!out) return SPB_EINVAL;
  *out = nullptr;
#ifndef SPB_EMU
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) { cudaGetLastError(); return SPB_ECUDA; }
  if (cudaSetDevice(device) != cudaSuccess) return SPB_ECUDA;
  // random 8-byte table probes: do not let L2 over-fetch neighbouring sectors from HBM
  cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
  cudaGetLastError();
#endif
  spb_ctx* c = new spb_ctx();
  c->device = device;
  c->stream = (spb_stream_t)cuda_stream;
  memset(&c->cnt, 0, sizeof c->cnt);
  memset(&c->nt, 0, sizeof c->nt);
#ifdef SPB_EMU
  c->pinned = (u64*)malloc(4096);
#else
  if (cudaHostAlloc((void**)&c->pinned, 4096, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { delete c; return SPB_ENOMEM; }
#endif
  c->errflag = dalloc<u32>(c, 4);
  if (!c->errflag) { delete c; return SPB_ENOMEM; }
  dev_memset(c, c->errflag, 0, 16);
  *out = c;
  return SPB_OK;
}

// return every live block except the error flag to the cache and reset the per-input state
static void reset_state(spb_ctx* c) {
  std::vector<void*> blocks;
  for (auto& kv : c->live) if (kv.first != (void*)c->errflag) blocks.push_back(kv.first);
  for (void* p : blocks) dev_free_raw(c, p);
  spb_ctx fresh;
  fresh.device = c->device; fresh.stream = c->stream; fresh.errflag = c->errflag; fresh.pinned = c->pinned;
  fresh.out_seq_host = c->out_seq_host; fresh.out_seq_cap = c->out_seq_cap;
#ifndef SPB_EMU
  for (int i = 0; i < 3; ++i) { fresh.stream2[i] = c->stream2[i]; fresh.ev_join[i] = c->ev_join[i]; }
  fresh.ev_fork = c->ev_fork;
  fresh.copy_stream = c->copy_stream; fresh.ev_emit = c->ev_emit; fresh.ev_d2h = c->ev_d2h;
  if (c->d2h_pending) cudaEventSynchronize(c->ev_d2h);          // a copy into the caller's buffer is still in flight
  fresh.evpool.swap(c->evpool);
#endif
  fresh.live.swap(c->live); fresh.cache.swap(c->cache);
  memset(&fresh.cnt, 0, sizeof fresh.cnt); memset(&fresh.nt, 0, sizeof fresh.nt);
  *c = fresh;
}

void spb_destroy(spb_ctx* c) {
  if (!c) return;
  SPB_ON_DEVICE(c);
  dev_sync(c);
  for (auto& kv : c->live) backend_free(kv.first);
  c->live.clear();
  cache_release(c);
#ifdef SPB_EMU
  free(c->pinned);
#else
  cudaFreeHost(c->pinned);
  for (auto e : c->evpool) cudaEventDestroy(e);
  if (c->stream2[0]) { for (int i = 0; i < 3; ++i) { cudaStreamDestroy(c->stream2[i]); cudaEventDestroy(c->ev_join[i]); } cudaEventDestroy(c->ev_fork); }
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); cudaEventDestroy(c->ev_emit); cudaEventDestroy(c->ev_d2h); }
#endif
  delete c;
}

const char* spb_last_error(const spb_ctx* c) { return c ? c->err : "null context"; }

int spb_load(spb_ctx* c, const uint8_t* ascii, const uint64_t* seq_off, uint32_t n_seq, uint32_t k) {
  if (!c) return SPB_EINVAL;
  SPB_ON_DEVICE(c);
  if (!ascii || !seq_off || n_seq == 0) return fail(c, SPB_EINVAL, "empty input");
  if (k < 3 || (k & 1) == 0) return fail(c, SPB_EINVAL, "k must be odd and >= 3 (got %u); the reference assumes odd k (SURVEY Q8)", k);
  if (k > 4000) return fail(c, SPB_EINVAL, "k = %u exceeds the reference's limit (Compressor.pyx:35)", k);
  u64 T = seq_off[n_seq];
  if (seq_off[0] != 0) return fail(c, SPB_EINVAL, "seq_off[0] must be 0");
  for (u32 s = 0; s < n_seq; ++s) if (seq_off[s + 1] < seq_off[s]) return fail(c, SPB_EINVAL, "seq_off not monotone");
  if (T == 0) return fail(c, SPB_EINVAL, "empty input");
  if (T >= (1ull << 32) - 4096) return fail(c, SPB_EINVAL, "input of %llu bases exceeds the 32-bit position range of one device", (unsigned long long)T);
  reset_state(c);
  c->n_syncs = 0;
  dev_memset(c, c->errflag, 0, 16);
  c->k = k; c->n_seq = n_seq; c->T = T;
  c->seq_off_h.assign(seq_off, seq_off + n_seq + 1);
  u64 ninst = 0;
  for (u32 s = 0; s < n_seq; ++s) { u64 len = seq_off[s + 1] - seq_off[s]; if (len > k) ninst += len - k + 1; }
  c->n_inst = ninst;
  const u32 W = (2 * k + 31) / 32;
  c->nw = (T + 15) / 16;
  c->npos_pad = nblk(T, TPB) * TPB + TPB;
  u8* d_ascii;
  ALLOC(d_ascii, u8, T);
  dev_copy(c, d_ascii, ascii, T);
  ALLOC(c->seq_off, u64, n_seq + 1);
  dev_copy(c, c->seq_off, seq_off, (n_seq + 1) * 8ull);
  u64 nwp = c->nw + W + 8;
  ALLOC(c->F, u32, nwp); ALLOC(c->R, u32, nwp);
  dev_memset(c, c->F + c->nw, 0, (nwp - c->nw) * 4); dev_memset(c, c->R + c->nw, 0, (nwp - c->nw) * 4);
  SPB_LAUNCH(k_pack, nblk(c->nw, TPB), TPB, c->stream, d_ascii, T, c->F, c->nw, c->errflag);
  SPB_LAUNCH(k_pack_rc, nblk(c->nw, TPB), TPB, c->stream, c->F, T, c->R, c->nw);
  u64 nw32 = c->npos_pad / 32;
  ALLOC(c->vbits, u32, nw32 + 8);
  SPB_LAUNCH(k_valid_bits, nblk(nw32 + 8, TPB), TPB, c->stream, c->seq_off, n_seq, T, k, c->vbits, nw32 + 8);
  dfree(c, d_ascii);
  int r = check_err(c, "spb_load");
  if (r) return r;
  c->cnt.n_seq = n_seq; c->cnt.n_bases = T; c->cnt.n_inst = ninst;
  c->state = 1;
  return SPB_OK;
}

static u64 smem_limit() {                             // tests force the fallbacks with a tiny limit
  if (const char* e = getenv("SPB_SMEM_LIMIT")) return (u64)std::max(1, atoi(e));
  return SPB_ST_LIMIT;
}
static u32 default_bucket_bits(u64 n_inst, u32 n_owners) {
  u32 bb = 0;
  while (bb < 14 && (n_inst >> bb) > 1200000ull) ++bb;      // ~1 M records per bucket: 16 MB scratch table, L2 resident
  u32 need = 0; while ((1u << need) < n_owners) ++need;
  return bb > need ? bb : need;
}

static int alloc_position_state(spb_ctx* c, u64 min_positions = 0) {
  if (c->sp) return 0;
  if (min_positions + TPB > c->npos_pad) c->npos_pad = nblk(min_positions, TPB) * TPB + TPB;
  u64 nw32 = c->npos_pad / 32;
  ALLOC(c->sp, u32, c->npos_pad + 8);
  ALLOC(c->obits, u32, nw32 + 8); ALLOC(c->fbits, u32, nw32 + 8);
  dev_memset(c, c->obits, 0, (nw32 + 8) * 4); dev_memset(c, c->fbits, 0, (nw32 + 8) * 4);
  return 0;
}

// rolling polynomial hashes: two odd multipliers = 5 (mod 8) (multiplicative order 2^30), inverses by Newton iteration
static RollK roll_constants(u32 k) {
  RollK K;
  K.B[0] = 0x9E3779B5u; K.B[1] = 0x85EBCA6Du;
  for (int h = 0; h < 2; ++h) {
    const u32 B = K.B[h];
    u32 inv = B; for (int it = 0; it < 5; ++it) inv *= 2u - B * inv;
    u32 pw = 1, sq = B; for (u32 e = k - 1; e; e >>= 1) { if (e & 1) pw *= sq; sq *= sq; }
    K.Binv[h] = inv; K.Bk1[h] = pw; K.negBk[h] = 0u - pw * B; K.negBinv[h] = 0u - inv; K.B4[h] = B * B * B * B;
  }
  return K;
}

// S1: records of the positions [pos_lo, pos_hi) (pos_lo a multiple of 256), partitioned by the top `bb` hash bits
static int dd_records(spb_ctx* c, u64 pos_lo, u64 pos_hi, u32 bb, u64 min_positions = 0, u32 spread = 0) {
  int r0 = alloc_position_state(c, min_positions); if (r0) return r0;
  if (pos_hi < pos_lo) pos_hi = pos_lo;
  u64 n = nblk(pos_hi - pos_lo, TPB) * TPB;
  u64 *keys, *keys2; u32 *pos, *pos2;
  ALLOC(keys, u64, n); ALLOC(pos, u32, n);
  u64 grid = n / TPB;
  if (!getenv("SPB_NO_ROLL") && c->k >= 17) {        // (the rolling kernels keep 16-base windows of both strand ends: k >= 16)
    const RollK K = roll_constants(c->k);
    { const int tq_ = t_begin(c); SPB_LAUNCH(k_make_records_roll, nblk(n / SPB_RUN, 128), 128, c->stream, c->F, c->R, c->vbits, c->T, c->k, pos_lo, pos_hi, n, keys, pos, c->obits,
               spread, K, 0u, 0u, (u64*)nullptr, 0ull); t_end(c, tq_, "k_make_records_roll_ms"); }
  } else
  switch ((2 * c->k + 31) / 32) {
    case 7:  SPB_LAUNCH(k_make_records<7>, grid, TPB, c->stream, c->F, c->R, c->vbits, c->T, c->k, pos_lo, pos_hi, keys, pos, c->obits, spread); break;
    case 19: SPB_LAUNCH(k_make_records<19>, grid, TPB, c->stream, c->F, c->R, c->vbits, c->T, c->k, pos_lo, pos_hi, keys, pos, c->obits, spread); break;
    case 32: SPB_LAUNCH(k_make_records<32>, grid, TPB, c->stream, c->F, c->R, c->vbits, c->T, c->k, pos_lo, pos_hi, keys, pos, c->obits, spread); break;
    default: SPB_LAUNCH(k_make_records<0>, grid, TPB, c->stream, c->F, c->R, c->vbits, c->T, c->k, pos_lo, pos_hi, keys, pos, c->obits, spread); break;
  }
  if (bb > 0) {
    ALLOC(keys2, u64, n); ALLOC(pos2, u32, n);
    PRIM(prim_partition_u64_u32(c, keys, pos, keys2, pos2, n, 64 - bb, 64));
    dfree(c, keys2); dfree(c, pos2);
  }
  const u32 B = 1u << bb;
  u64* coff; ALLOC(coff, u64, 2);
  u64 co[2] = {0, n};
  memcpy(c->pinned, co, 16); dev_copy(c, coff, c->pinned, 16);
  ALLOC(c->dd_bstart, u64, B + 2);
  SPB_LAUNCH(k_bucket_bounds, nblk(B + 1, TPB), TPB, c->stream, keys, coff, 1u, bb, 0u, B, c->dd_bstart);
  c->dd_bstart_h.resize(B + 1);
  dev_copy(c, c->dd_bstart_h.data(), c->dd_bstart, (B + 1) * 8ull); dev_sync(c);
  dfree(c, coff);
  c->dd_keys = keys; c->dd_pos = pos; c->dd_n = n; c->dd_lo = pos_lo; c->dd_hi = pos_hi; c->dd_bb = bb;
  return 0;
}

// S2: deduplicate the records of buckets [b_lo, b_hi) received in `nchunks` bucket-sorted chunks; rep[i] = first position
static int dd_owner(spb_ctx* c, const u64* keys, const u32* pos, const u64* chunk_counts, u32 nchunks, u32 bb, u32 b_lo, u32 b_hi, u32* rep) {
  const u32 B = b_hi - b_lo;
  std::vector<u64> coff_h(nchunks + 1, 0);
  for (u32 r = 0; r < nchunks; ++r) coff_h[r + 1] = coff_h[r] + chunk_counts[r];
  if (coff_h[nchunks] == 0) return 0;
  u64 *coff, *sub;
  ALLOC(coff, u64, nchunks + 1); ALLOC(sub, u64, (u64)nchunks * (B + 1));
  dev_copy(c, coff, coff_h.data(), (nchunks + 1) * 8ull); dev_sync(c);
  SPB_LAUNCH(k_bucket_bounds, nblk((u64)nchunks * (B + 1), TPB), TPB, c->stream, keys, coff, nchunks, bb, b_lo, B, sub);
  std::vector<u64> sub_h((u64)nchunks * (B + 1));
  dev_copy(c, sub_h.data(), sub, sub_h.size() * 8); dev_sync(c);
  std::vector<u64> nb(B, 0); u64 nmax = 0;
  for (u32 b = 0; b < B; ++b) { for (u32 r = 0; r < nchunks; ++r) nb[b] += sub_h[(u64)r * (B + 1) + b + 1] - sub_h[(u64)r * (B + 1) + b]; nmax = std::max(nmax, nb[b]); }
  u64 cap = 1024; while (cap < 2 * nmax) cap <<= 1;
  c->cnt.table_slots = cap;
  u64* scratch[2] = {nullptr, nullptr};
  ALLOC(scratch[0], u64, cap);
  if (B > 1) ALLOC(scratch[1], u64, cap); else scratch[1] = scratch[0];
  dev_memset(c, scratch[0], 0xFF, cap * 8);
  u32* dbits; ALLOC(dbits, u32, c->npos_pad / 32 + 8);
  dev_memset(c, dbits, 0, (c->npos_pad / 32 + 8) * 4);
  BktView v{keys, pos, rep, coff, sub, nchunks, B, 0};
  u32 step = 0;
  for (u32 b = 0; b < B; ++b) {
    if (nb[b] == 0) continue;
    v.b = b;
    u64* cur = scratch[step & 1]; u64* other = scratch[(step & 1) ^ 1];
    SPB_LAUNCH(k_bkt_insert, nblk(nb[b], TPB), TPB, c->stream, v, nb[b], c->F, c->R, c->T, c->k, cur, cap - 1, dbits);
    u64 g = std::max<u64>(nblk(nb[b], TPB), std::min<u64>(nblk(cap, TPB), 148 * 8));
    SPB_LAUNCH(k_bkt_final, g, TPB, c->stream, v, nb[b], c->F, c->R, c->T, c->k, cur, cap - 1, dbits, other, B > 1 ? cap : 0);
    ++step;
  }
  if (B > 1) dfree(c, scratch[1]);
  dfree(c, scratch[0]); dfree(c, coff); dfree(c, sub); dfree(c, dbits);
  return 0;
}

// S3: scatter the representatives of the records this rank produced back to position order
static int dd_scatter(spb_ctx* c, const u32* rep) {
  const u32 B = 1u << c->dd_bb;
  u64 maxlen = 0;
  for (u32 b = 0; b < B; ++b) maxlen = std::max(maxlen, c->dd_bstart_h[b + 1] - c->dd_bstart_h[b]);
  SPB_LAUNCH(k_unpartition, nblk(maxlen * B, TPB), TPB, c->stream, c->dd_pos, rep, c->dd_bstart, B, c->sp, c->fbits);
  dfree(c, c->dd_keys); dfree(c, c->dd_pos); dfree(c, c->dd_bstart);
  return 0;
}

// vertex ids for the positions [lo, hi) + the vertex table (needs the complete fbits bitmap)
// ---- views for the kernels that may run on a slice (single GPU: the slice is everything, the views are plain arrays)
static VPos vpos_view(spb_ctx* c) {
  return VPos{c->v2pos, c->sl.on ? c->sl.v_lo : 0u, c->sl.on ? c->sl.v_hi : c->nV, c->fbits, c->blkoff, c->wpre, (u32)c->nb};
}
static VidAt vid_view(spb_ctx* c) {
  return VidAt{c->sp, c->sl.on ? c->sl.ids_lo : 0ull, c->sl.on ? c->sl.ids_hi : c->T, c->fbits, c->mg_seen, c->blkoff, c->wpre};
}
static VNbp vnbp_view(spb_ctx* c) {
  const u32 lo = c->sl.on ? 8u * c->sl.t_lo : 0u;
  const u32 hi = c->sl.on ? (u32)std::min<u64>(8ull * (c->sl.t_hi + 4), c->nV) : c->nV;     // first appearances of the slice reach up to 31 ids beyond the cut
  if (!c->vnbp) return VNbp{nullptr, 0u, 0u, c->plo, c->phi, c->nP, c->dart_nbp};     // not materialised: every lookup goes through the piece list
  return VNbp{c->vnbp, lo, hi, c->plo, c->phi, c->nP, c->dart_nbp};
}
// vertex -> NBP map of ALL vertices (one GPU / completed state), built from the sorted piece list when a getter needs the
// array and the compaction did not write it (mostly-unique inputs, see cp_emit)
static int ensure_vnbp(spb_ctx* c) {
  if (c->vnbp || c->state < 3) return 0;
  const u32 nV = c->nV;
  ALLOC(c->vnbp, u32, (u64)nV + 1);
  dev_memset(c, c->vnbp + nV, 0xFF, 4);
  SPB_LAUNCH(k_fill_vnbp, nblk(nblk(nV, 1024) * 32, TPB), TPB, c->stream, c->plo, c->phi, c->nP, c->dart_nbp, 0u, nV, c->vnbp);
  return 0;
}

// vertex ids of the positions [lo, hi); vertex -> first position for the first appearances in [tlo, thi); the flag byte
// (implicit-edge bit) of every vertex.  The rank tables of the first-appearance bitmap stay in the context.
static int dd_ids(spb_ctx* c, u64 lo, u64 hi, const u64* slots, u64 cap = 1, const u32* sp_in = nullptr, int follow = 0, u64 tlo = 0, u64 thi = ~0ull) {
  if (!sp_in) sp_in = c->sp;
  const u64 T = c->T;
  u64 nb = c->npos_pad / 256;
  u32* blkcnt; ALLOC(blkcnt, u32, nb);
  dfree(c, c->blkoff); dfree(c, c->wpre);
  ALLOC(c->blkoff, u32, nb + 1); ALLOC(c->wpre, u8, nb * 8 + 8);
  c->nb = nb;
  SPB_LAUNCH(k_block_counts, nblk(nb, TPB), TPB, c->stream, c->fbits, blkcnt, c->wpre, nb);
  u64 nV = 0;
  PRIM(prim_scan_u32(c, blkcnt, c->blkoff, nb, &nV));
  dfree(c, blkcnt);
  c->nV = (u32)nV;
  dfree(c, c->v2pos); dfree(c, c->vflags);
  ALLOC(c->v2pos, u32, nV + 1); ALLOC(c->vflags, u8, nV + 64);
  dev_memset(c, c->vflags + (nV & ~7ull), 0, 64);
  const u32* blkoff = c->blkoff; const u8* wpre = c->wpre;
  const int tq_ids = t_begin(c);
  switch ((2 * c->k + 31) / 32) {
    case 7:  SPB_LAUNCH(k_assign_ids<7>, gs_grid(T), TPB, c->stream, sp_in, c->sp, c->vbits, T, lo, hi, c->fbits, blkoff, wpre, slots, cap - 1, c->F, c->R, c->k, c->v2pos, c->vflags, follow, c->errflag, tlo, thi); break;
    case 19: SPB_LAUNCH(k_assign_ids<19>, gs_grid(T), TPB, c->stream, sp_in, c->sp, c->vbits, T, lo, hi, c->fbits, blkoff, wpre, slots, cap - 1, c->F, c->R, c->k, c->v2pos, c->vflags, follow, c->errflag, tlo, thi); break;
    case 32: SPB_LAUNCH(k_assign_ids<32>, gs_grid(T), TPB, c->stream, sp_in, c->sp, c->vbits, T, lo, hi, c->fbits, blkoff, wpre, slots, cap - 1, c->F, c->R, c->k, c->v2pos, c->vflags, follow, c->errflag, tlo, thi); break;
    default: SPB_LAUNCH(k_assign_ids<0>, gs_grid(T), TPB, c->stream, sp_in, c->sp, c->vbits, T, lo, hi, c->fbits, blkoff, wpre, slots, cap - 1, c->F, c->R, c->k, c->v2pos, c->vflags, follow, c->errflag, tlo, thi); break;
  }
  t_end(c, tq_ids, "k_assign_ids_ms");
  return 0;
}

// raw junction entries (both directions, unsorted) of the pairs (p, p + 1) in the bitmap words [w_lo, w_hi)
static int graph_junctions(spb_ctx* c, u64 w_lo, u64 w_hi, u64** Jk_out, u8** Js_out, u64* nj_out) {
  const u64 T = c->T; const u64 nV = c->nV;
  u64* jcount; ALLOC(jcount, u64, 2);
  const u64 npos = 32 * (w_hi - w_lo);
  u64 jcap = std::max<u64>(4096, npos / 8);
  u64* Jk = nullptr; u8* Js = nullptr; u64 nj = 0;
  u32* vobits = nullptr;
  if (!c->sl.on && c->n_inst - nV > nV / 4) {          // many repeated instances: per-vertex orientation bits pay off
    ALLOC(vobits, u32, nV / 32 + 8);
    dev_memset(c, vobits, 0, (nV / 32 + 8) * 4);
    SPB_LAUNCH(k_vertex_obits, std::min<u64>(nblk(nV, TPB), 148 * 16), TPB, c->stream, c->v2pos, c->obits, (u32)nV, vobits);
  }
  for (int attempt = 0; attempt < 2; ++attempt) {
    ALLOC(Jk, u64, jcap); ALLOC(Js, u8, jcap);
    dev_memset(c, jcount, 0, 16);
    { const int tq_ = t_begin(c); SPB_LAUNCH(k_junction_emit, std::min<u64>(nblk(npos, TPB), 148 * 16), TPB, c->stream, c->sp, T, w_lo, w_hi, c->fbits, c->vflags, vpos_view(c), c->obits, vobits, Jk, Js, jcount, jcap, c->errflag); t_end(c, tq_, "k_junction_emit_ms"); }
    nj = dread(c, jcount);
    if (nj <= jcap) break;
    dfree(c, Jk); dfree(c, Js); jcap = nj;
  }
  dfree(c, jcount); dfree(c, vobits);
  *Jk_out = Jk; *Js_out = Js; *nj_out = nj;
  return 0;
}

// all junction entries (takes ownership of Jk / Js) -> sorted unique J, junction / sequence-limit flags, and the explicit
// edge rows of the vertex tiles [t_lo, t_hi) (one GPU: all of them)
static int graph_finish(spb_ctx* c, u64* Jk, u8* Js, u64 nj) {
  const u32 k = c->k; const u64 nV = c->nV;
  {
    u64* k2; u8* v2; ALLOC(k2, u64, nj); ALLOC(v2, u8, nj);
    PRIM(prim_sort_u64_u8(c, Jk, Js, k2, v2, nj));
    dfree(c, k2); dfree(c, v2);
  }
  PRIM(unique_sorted(c, Jk, Js, nj, &c->nJ));
  c->Jk = Jk; c->Js = Js;
  SPB_LAUNCH(k_mark_hasj, nblk(c->nJ, TPB), TPB, c->stream, c->Jk, c->nJ, c->vflags);
  ALLOC(c->seqlim, u32, 2ull * c->n_seq);
  SPB_LAUNCH(k_mark_seqlim, nblk(c->n_seq, TPB), TPB, c->stream, c->seq_off, c->n_seq, k, vid_view(c), c->vflags, c->seqlim);
  GView g{c->vflags, c->Jk, c->Js, c->nJ, c->nV};
  const u32 nTall = (u32)((nV + 7) / 8);
  const u32 t_lo = c->sl.on ? c->sl.t_lo : 0u, t_hi = c->sl.on ? c->sl.t_hi : nTall;
  const u64 nT = t_hi - t_lo;
  u32 *ecnt, *eoff; ALLOC(ecnt, u32, nT + 1); ALLOC(eoff, u32, nT + 1);
  SPB_LAUNCH(k_edge_counts8, nblk(nT, TPB), TPB, c->stream, g, ecnt, t_lo, t_hi);
  u64 nE = 0;
  PRIM(prim_scan_u32(c, ecnt, eoff, nT, &nE));
  dfree(c, ecnt);
  c->edge_off = eoff; c->edges_pending = true;          // rows are written by ensure_edges (after the NBP emission, so that the
  c->sl.edge_n = nE;                                    // copy of the sequences to the host overlaps with it)
  c->cnt.n_vertices = nV; c->cnt.n_edges = nE; c->cnt.n_junction = c->nJ;     // (sliced: n_edges is set to the total by the caller)
  return 0;
}

// explicit edge rows of the own vertex tiles (reference lib/Assembler.py:247-249): deferred part of graph_finish
static int ensure_edges(spb_ctx* c) {
  if (!c->edges_pending) return 0;
  GView g{c->vflags, c->Jk, c->Js, c->nJ, c->nV};
  const u32 nTall = (u32)(((u64)c->nV + 7) / 8);
  const u32 t_lo = c->sl.on ? c->sl.t_lo : 0u, t_hi = c->sl.on ? c->sl.t_hi : nTall;
  const u64 nT = t_hi - t_lo;
  ALLOC(c->edges, u32, 2 * c->sl.edge_n + 2);
  { const int tq_ = t_begin(c); SPB_LAUNCH(k_edge_write, std::min<u64>(nblk(8 * nT, TPB), 148 * 32), TPB, c->stream, g, c->edge_off, c->edges, t_lo, t_hi); t_end(c, tq_, "k_edge_write_ms"); }
  dfree(c, c->edge_off);
  c->edges_pending = false;
  return 0;
}

// junction edges, sequence limits, explicit edge list (needs the complete sp / obits arrays)
static int dd_graph(spb_ctx* c) {
  u64* Jk; u8* Js; u64 nj;
  PRIM(graph_junctions(c, 0, (c->T + 31) / 32, &Jk, &Js, &nj));
  return graph_finish(c, Jk, Js, nj);
}

// Partitioned dedup, "links" form (single GPU, mostly-unique inputs): records -> 2^bb buckets -> one L2-resident scratch
// table per bucket; only duplicates / displaced positions write anything (a link to a smaller equal position).
static int dedup_staged_links(spb_ctx* c) {
  const u64 T = c->T, P = c->npos_pad;
  u32 bb = 0; while (bb < 8 && (c->n_inst >> bb) > 2200000ull) ++bb;       // <= 256 buckets: one radix pass
  if (const char* e = getenv("SPB_BUCKET_BITS")) bb = (u32)atoi(e);
  int tk = t_begin(c);
  const u32 B = 1u << bb;
  std::vector<u64> blo(B), bn(B);
  // SPB_PARTITION_FUSED=1 selects the hand-written single-pass partition (k_partition_records).  Measured slower than
  // record generation + one library radix pass in round 1 (20.0 vs 8.7 ms on config 3: 8-record runs per bucket and
  // tile are too short for the scattered stores), so it stays an experiment; the emulator tests cover both.
  bool fused = getenv("SPB_PARTITION_FUSED") && bb <= 8;       // 256-entry histogram in shared memory
  if (fused) {
    // ---- fused hash + single-pass partition into fixed-capacity bucket regions
    int r0 = alloc_position_state(c); if (r0) return r0;
    u64 bcap = (c->n_inst / B) + (c->n_inst / B) / 16 + 8192;
    bcap = (bcap + 255) & ~255ull;
    u64* keys; u32 *pos, *cursor; ALLOC(keys, u64, B * bcap); ALLOC(pos, u32, B * bcap); ALLOC(cursor, u32, 256 + 8);
    dev_memset(c, cursor, 0, (256 + 8) * 4);
    u64 tiles = nblk(P, 256 * SPB_PT);
    switch ((2 * c->k + 31) / 32) {
      case 7:  SPB_LAUNCH_PHASED(k_partition_records<7>, 4, tiles, TPB, c->stream, c->F, c->R, c->vbits, T, c->k, bb, keys, pos, bcap, cursor, c->obits, c->errflag); break;
      case 19: SPB_LAUNCH_PHASED(k_partition_records<19>, 4, tiles, TPB, c->stream, c->F, c->R, c->vbits, T, c->k, bb, keys, pos, bcap, cursor, c->obits, c->errflag); break;
      case 32: SPB_LAUNCH_PHASED(k_partition_records<32>, 4, tiles, TPB, c->stream, c->F, c->R, c->vbits, T, c->k, bb, keys, pos, bcap, cursor, c->obits, c->errflag); break;
      default: SPB_LAUNCH_PHASED(k_partition_records<0>, 4, tiles, TPB, c->stream, c->F, c->R, c->vbits, T, c->k, bb, keys, pos, bcap, cursor, c->obits, c->errflag); break;
    }
    std::vector<u32> cnt_h(B);
    pin_read(c, c->pinned, cursor, B * 4ull); dev_sync(c);
    memcpy(cnt_h.data(), c->pinned, B * 4ull);
    u32 e = dread(c, c->errflag);
    dfree(c, cursor);
    if (e & ERR_OVERFLOW) {                          // skewed hashes (masses of identical k-mers): exact two-pass partition instead
      dev_memset(c, c->errflag, 0, 4);
      dfree(c, keys); dfree(c, pos);
      fused = false;
    } else {
      for (u32 b = 0; b < B; ++b) { blo[b] = (u64)b * bcap; bn[b] = cnt_h[b]; }
      c->dd_keys = keys; c->dd_pos = pos;
    }
  }
  if (!fused) {
    PRIM(dd_records(c, 0, T, bb));
    for (u32 b = 0; b < B; ++b) { blo[b] = c->dd_bstart_h[b]; bn[b] = c->dd_bstart_h[b + 1] - blo[b]; }
    dfree(c, c->dd_bstart);
  }
  t_end(c, tk, "records_partition_ms");
  tk = t_begin(c);
  u64 nmax = 0;
  for (u32 b = 0; b < B; ++b) nmax = std::max(nmax, bn[b]);
  u64 cap = 1024; while (cap < 2 * nmax) cap <<= 1;
  c->cnt.table_slots = cap;
  u64* scratch[4] = {nullptr, nullptr, nullptr, nullptr};
  u32 *seenpos, *dupbits; ALLOC(seenpos, u32, P + 8); ALLOC(dupbits, u32, P / 32 + 8);
  dev_memset(c, dupbits, 0, (P / 32 + 8) * 4);
  {
    TwoStreams ts(c);
    for (int i = 0; i < ts.n; ++i) ALLOC(scratch[i], u64, cap);
    u32 step = 0;
    for (u32 b = 0; b < B; ++b) {
      u64 lo = blo[b], n = bn[b];
      if (n == 0) continue;
      const int i = step % ts.n;
      ts.clear(c, i, scratch[i], cap * 8);
      SPB_LAUNCH(k_bkt_insert_links, nblk(n, TPB), TPB, ts.st[i], c->dd_keys, c->dd_pos, lo, n, c->F, c->R, T, c->k, scratch[i], cap - 1, seenpos,
                 dupbits, (u64*)nullptr, 0);
      ++step;
    }
    ts.join(c);
  }
  for (auto& x : scratch) dfree(c, x);
  dfree(c, c->dd_keys); dfree(c, c->dd_pos);
  t_end(c, tk, "bucket_dedup_ms");
  u64 nb = P / 256, nw32 = P / 32;
  SPB_LAUNCH(k_first_bits, nblk(nb, TPB), TPB, c->stream, c->vbits, dupbits, c->fbits, nb);     // first = valid and not linked
  dev_memset(c, c->fbits + nw32, 0, 32);
  dfree(c, dupbits);
  int r = dd_ids(c, 0, T, nullptr, 1, seenpos, 1);
  dfree(c, seenpos);
  return r;
}

// Two-level partitioned dedup with shared-memory tables (see k_smem_dedup): records radix-sorted on the top hash bits
// into sub-buckets of ~7.6 K records, one CTA per sub-bucket.  Returns 1 when the input does not suit it (a sub-bucket
// beyond the table: masses of identical k-mers) -- the caller then takes the L2-table path.
static int dedup_smem_links(spb_ctx* c) {
  const u64 T = c->T, P = c->npos_pad;
  u64 avg = 7680;
  if (const char* e = getenv("SPB_SMEM_AVG")) avg = std::max(1, atoi(e));       // tests: many tiny sub-buckets
  u32 nbits = 0; while ((P >> nbits) > avg) ++nbits;
  if (nbits > 24) return 1;
  int tk = t_begin(c);
  PRIM(dd_records(c, 0, T, nbits, 0, 1));
  t_end(c, tk, "records_partition_ms");
  tk = t_begin(c);
  const u64 NB = 1ull << nbits;
  u32 *seenpos, *dupbits; ALLOC(seenpos, u32, P + 8); ALLOC(dupbits, u32, P / 32 + 8);
  dev_memset(c, dupbits, 0, (P / 32 + 8) * 4);
  u64 nmax = 0;
  for (u64 b = 0; b < NB; ++b) nmax = std::max(nmax, c->dd_bstart_h[b + 1] - c->dd_bstart_h[b]);
  if (nmax <= smem_limit())
    { const int tq_ = t_begin(c); SPB_LAUNCH_PHASED_DSMEM(k_smem_dedup<false>, 2, NB, 512, SPB_ST_SLOTS * 4, c->stream, c->dd_keys, c->dd_pos, c->dd_bstart, c->F, c->R, T, c->k, nbits,
                            seenpos, dupbits, c->errflag, (u64*)nullptr, (u64*)nullptr, 0ull); t_end(c, tq_, "k_smem_dedup_ms"); }
  dfree(c, c->dd_keys); dfree(c, c->dd_pos); dfree(c, c->dd_bstart);
  if (nmax > smem_limit()) { dfree(c, seenpos); dfree(c, dupbits); return 1; }
  c->cnt.table_slots = SPB_ST_SLOTS;
  t_end(c, tk, "bucket_dedup_ms"); stat_set(c, "smem_tables", (double)NB);
  u64 nb = P / 256, nw32 = P / 32;
  SPB_LAUNCH(k_first_bits, nblk(nb, TPB), TPB, c->stream, c->vbits, dupbits, c->fbits, nb);     // first = valid and not linked
  dev_memset(c, c->fbits + nw32, 0, 32);
  dfree(c, dupbits);
  int r = dd_ids(c, 0, T, nullptr, 1, seenpos, 1);
  dfree(c, seenpos);
  return r;
}

// Launch of k_roll_partition in one of its compiled shapes (threads per CTA x positions per thread and staging group).
static void launch_roll_partition(spb_ctx* c, u64 pos_lo, u64 pos_hi, u32 run, u64 n_runs, u32 own_bits, u32 own_id, u32 l1, u32 copies, u64* rec1, u64 cap1, u32* cursor1) {
  const RollK K = roll_constants(c->k);
  int shape = 0;
  if (const char* e = getenv("SPB_P1_SHAPE")) shape = atoi(e);
#define SPB_P1_LAUNCH(T1, MINB) do { if (own_bits) SPB_P1_LAUNCH2(T1, MINB, true); else SPB_P1_LAUNCH2(T1, MINB, false); } while (0)
#define SPB_P1_LAUNCH2(T1, MINB, MG) SPB_LAUNCH_PHASED_DSMEM((k_roll_partition<T1, MINB, MG>), 1, nblk(n_runs, T1), T1, (T1) * SPB_P1_GS * 8, c->stream, c->F, c->R, c->vbits, c->T, \
                                         c->k, pos_lo, pos_hi, c->npos_pad, run, n_runs, own_bits, own_id, l1, copies, rec1, cap1, ((u64)copies << l1) * cap1, cursor1, c->obits, K, c->errflag)
  switch (shape) {
    case 1: SPB_P1_LAUNCH(256, 4); break;
    case 2: SPB_P1_LAUNCH(1024, 1); break;
    default: SPB_P1_LAUNCH(512, 2); break;
  }
#undef SPB_P1_LAUNCH2
#undef SPB_P1_LAUNCH
}

// Two-level partitioned dedup, hand-written end to end (spb_partition.cuh): rolling hashes fused with the level-1 partition,
// level-2 partition, one CTA per final bucket with a bulk-copied record block and a shared-memory table.  Returns 1 when
// the input does not suit it (a bucket beyond its region: masses of identical k-mers) -- the caller then takes the
// sort-based path.
static int dedup_part_links(spb_ctx* c) {
  const u64 T = c->T;
  if (c->k < 17) return 1;                           // the rolling kernels keep 16-base windows of both strand ends: other paths for tiny k
  int r0 = alloc_position_state(c); if (r0) return r0;
  const u64 P = c->npos_pad;
  u64 avg = SPB_D_AVG;
  if (const char* e = getenv("SPB_PART_AVG")) avg = std::max(1, atoi(e));        // tests: many tiny buckets
  u32 nbits = 0; while ((c->n_inst >> nbits) > avg) ++nbits;
  if (nbits > 18) return 1;
  u32 l1 = (nbits + 1) / 2;
  if (const char* e = getenv("SPB_PART_L1")) l1 = std::min<u32>(nbits, std::min(9, std::max(0, atoi(e))));
  const u32 l2 = nbits - l1;
  if (l2 > 9) return 1;
  u32 run = 256;                                     // positions per thread of k_roll_partition (longer runs: the threads of a warp stream too far apart for L1)
  if (const char* e = getenv("SPB_P1_RUN")) run = std::max(32, atoi(e) & ~31);
  u32 copies = 1;                                    // regions per level-1 bucket (CTA x of k_roll_partition appends to copy x % copies); measured: no gain on one GPU
  if (const char* e = getenv("SPB_P1_COPIES")) copies = std::max(1, std::min(64, atoi(e)));
  if (c->n_inst < (1ull << 24) && !getenv("SPB_P1_COPIES")) copies = 1;
  const u64 NB1 = 1ull << l1, NBF = 1ull << nbits, NR1 = NB1 * copies;
  const u64 n1 = nblk(c->n_inst, NR1);
  const u64 cap1 = (n1 + n1 / 32 + 2048 + 1) & ~1ull;                              // mean + slack (hashes are uniform: sd = sqrt(mean))
  const u64 cap2 = SPB_D_LIMIT;
  int tk = t_begin(c);
  u64 *rec1, *rec2; u32* cursors;
  ALLOC(rec1, u64, NR1 * cap1 + 8192 + 2); ALLOC(rec2, u64, NBF * cap2 + SPB_P2_CHUNK + 2); ALLOC(cursors, u32, NR1 + NBF + 8);
  dev_memset(c, cursors, 0, (NR1 + NBF + 8) * 4);
  u32 *cursor1 = cursors, *cursor2 = cursors + NR1;
  { const int tq_ = t_begin(c); launch_roll_partition(c, 0, T, run, nblk(P, run), 0, 0, l1, copies, rec1, cap1, cursor1); t_end(c, tq_, "k_roll_partition_ms"); }
  const u32 cpr = (u32)nblk(cap1, SPB_P2_CHUNK);
#ifdef SPB_EMU
  const u64 grid2 = NR1 * cpr, gridd = NBF;          // one chunk / bucket per block
#else
  const u64 grid2 = std::min<u64>(NR1 * cpr, 2 * 148), gridd = std::min<u64>(NBF, 2 * 148);   // persistent: two CTAs per SM
#endif
  { const int tq_ = t_begin(c); SPB_LAUNCH_PHASED_DSMEM(k_partition2, 1, grid2, SPB_P2_T, 3 * SPB_P2_CHUNK * 8, c->stream, rec1, cap1, cursor1, cpr, (u32)NR1, (u32)NB1, l2, rec2, cap2, NBF * cap2, cursor2, c->errflag, 0u);
    t_end(c, tq_, "k_partition2_ms"); }
  dfree(c, rec1);
  t_end(c, tk, "records_partition_ms");
  tk = t_begin(c);
  u32 *seenpos, *dupbits; ALLOC(seenpos, u32, P + 8); ALLOC(dupbits, u32, P / 32 + 8);
  dev_memset(c, dupbits, 0, (P / 32 + 8) * 4);
  // Optimistic links (default): records whose known hash bits (bucket + 32 stored) agree are linked without touching the genome and every link
  // is verified afterwards in one streaming pass (SPB_D_EXACT=1: the inline, warp-cooperative verification instead).
  const bool opt = !getenv("SPB_D_EXACT");
  u32 hmask = 0xFFFFFFFFu;                             // tests clear stored hash bits to provoke collisions (both forms stay exact)
  if (const char* e = getenv("SPB_D_HMASK")) hmask = (u32)strtoul(e, nullptr, 16);
  {
    const int tq_ = t_begin(c);
    if (opt)                                                                // optimistic links, verified below
      SPB_LAUNCH_PHASED_DSMEM((k_bucket_dedup<false, 1, true>), 3, NBF, SPB_D_T, SPB_D_SMEM_U64(1) * 8, c->stream, rec2, cap2, cursor2, (u32)NBF, l2, c->F, c->R, c->obits, T, c->k,
                              seenpos, dupbits, c->errflag, (u64*)nullptr, (u64*)nullptr, 0ull, hmask);
    else if (!(getenv("SPB_D_NBUF") && atoi(getenv("SPB_D_NBUF")) == 2))   // measured: three resident CTAs beat two persistent ones (2.6 vs 3.1 ms, config 3)
      SPB_LAUNCH_PHASED_DSMEM((k_bucket_dedup<false, 1, false>), 3, NBF, SPB_D_T, SPB_D_SMEM_U64(1) * 8, c->stream, rec2, cap2, cursor2, (u32)NBF, l2, c->F, c->R, c->obits, T, c->k,
                              seenpos, dupbits, c->errflag, (u64*)nullptr, (u64*)nullptr, 0ull, hmask);
    else
      SPB_LAUNCH_PHASED_DSMEM((k_bucket_dedup<false, 2, false>), 3, gridd, SPB_D_T, SPB_D_SMEM_U64(2) * 8, c->stream, rec2, cap2, cursor2, (u32)NBF, l2, c->F, c->R, c->obits, T, c->k,
                              seenpos, dupbits, c->errflag, (u64*)nullptr, (u64*)nullptr, 0ull, hmask);
    t_end(c, tq_, "k_bucket_dedup_ms");
  }
  dfree(c, rec2); dfree(c, cursors);
  const u32 e = dread(c, c->errflag);
  if (e & ERR_OVERFLOW) {                            // skewed buckets: the caller falls back to the exact sort-based partition
    dev_memset(c, c->errflag, 0, 4);
    dev_memset(c, c->obits, 0, (P / 32 + 8) * 4);
    dfree(c, seenpos); dfree(c, dupbits);
    return 1;
  }
  c->cnt.table_slots = SPB_D_SLOTS;
  if (opt) {
    const u64 FCAP = 1u << 20;
    const int tq_ = t_begin(c);
    u32 *failbits, *flist; u64* nfail;
    ALLOC(failbits, u32, P / 32 + 8); ALLOC(flist, u32, FCAP + 1); ALLOC(nfail, u64, 2);
    dev_memset(c, failbits, 0, (P / 32 + 8) * 4); dev_memset(c, nfail, 0, 16);
    const u64 nwd = P / 32;
    SPB_LAUNCH(k_verify_links, std::min<u64>(nblk(nwd, TPB), 148 * 16), TPB, c->stream, seenpos, dupbits, nwd, c->F, c->R, c->obits, T, c->k, failbits, flist, nfail, FCAP);
    u64 nf = dread(c, nfail);
    if (nf && nf <= FCAP) {                            // links that leaned on a failed predecessor: compared in full
      SPB_LAUNCH(k_verify_followups, nblk(nf, TPB), TPB, c->stream, (u32)nf, seenpos, dupbits, c->F, c->R, c->obits, T, c->k, failbits, flist, nfail, FCAP);
      nf = dread(c, nfail);
    }
    t_end(c, tq_, "k_verify_links_ms");
    stat_set(c, "failed_links", (double)nf);
    if (nf > FCAP) {                                 // far more hash collisions than any real input produces: exact path instead
      dfree(c, failbits); dfree(c, flist); dfree(c, nfail);
      dev_memset(c, c->obits, 0, (P / 32 + 8) * 4);
      dfree(c, seenpos); dfree(c, dupbits);
      return 1;
    }
    if (nf) {
      u64 *keys, *keys2; u32* newt; ALLOC(keys, u64, nf + 1); ALLOC(keys2, u64, nf + 1); ALLOC(newt, u32, nf + 1);
      SPB_LAUNCH(k_repair_roots, nblk(nf, TPB), TPB, c->stream, flist, (u32)nf, seenpos, dupbits, keys);
      PRIM(prim_sort_u64(c, keys, keys2, nf));
      SPB_LAUNCH(k_repair_find, nblk(nf, TPB), TPB, c->stream, keys, (u32)nf, seenpos, dupbits, c->F, c->R, c->obits, T, c->k, newt);
      SPB_LAUNCH(k_repair_apply, nblk(nf, TPB), TPB, c->stream, keys, (u32)nf, newt, seenpos, dupbits);
      dfree(c, keys); dfree(c, keys2); dfree(c, newt);
    }
    dfree(c, failbits); dfree(c, flist); dfree(c, nfail);
  }
  t_end(c, tk, "bucket_dedup_ms"); stat_set(c, "partition2", (double)NBF);
  u64 nb = P / 256, nw32 = P / 32;
  SPB_LAUNCH(k_first_bits, nblk(nb, TPB), TPB, c->stream, c->vbits, dupbits, c->fbits, nb);     // first = valid and not linked
  dev_memset(c, c->fbits + nw32, 0, 32);
  dfree(c, dupbits);
  int r = dd_ids(c, 0, T, nullptr, 1, seenpos, 1);
  dfree(c, seenpos);
  return r;
}

// the older single-table dedup (two random HBM accesses per instance); kept selectable for A/B measurements
// The choice between the dedup schemes is made on two sample chunks of positions -- the first CH positions and CH
// positions from the middle of the input (another genome, in a genome set) -- through a table sized for them alone
// (the full-size table of the global scheme is several GB: clearing it just to find out that the partitioned scheme
// applies would cost more than the probe).  Returns the number of extended / duplicate positions of the second chunk.
static int probe_sample_chunks(spb_ctx* c, u64 CH, u64* n_ext, u64* n_dup, u64* chunk_n) {
  const u64 P = c->npos_pad;
  u64 mid = (P / 2) / CH * CH;                       // chunk-aligned, so the bitmaps shift by whole words
  if (mid < CH) mid = CH;
  const u64 hi = std::min(P, mid + CH);
  if (hi <= mid) { *n_ext = *n_dup = 0; *chunk_n = 0; return 0; }
  const u64 shift = mid - CH;                        // position p of the second chunk is kept at index p - shift
  u64 cap = 1024; while (cap < 4 * CH) cap <<= 1;
  u64 *slots, *arec, *extc; u32 *wbits, *dbits, *seenpos;
  const u64 nbw = 2 * CH / 32 + 8;
  ALLOC(slots, u64, cap); ALLOC(arec, u64, nbw); ALLOC(extc, u64, 2);
  ALLOC(wbits, u32, nbw); ALLOC(dbits, u32, nbw); ALLOC(seenpos, u32, 2 * CH + 8);
  dev_memset(c, slots, 0xFF, cap * 8); dev_memset(c, extc, 0, 16);
  dev_memset(c, dbits, 0, nbw * 4); dev_memset(c, wbits, 0, nbw * 4);
  u32 *obits_save = c->obits, *otmp; ALLOC(otmp, u32, nbw);           // the probe must not touch the real orientation bitmap
  c->obits = otmp;
  launch_extract_k(c, 0, CH, slots, cap, seenpos, wbits, dbits, nullptr, nullptr);
  c->obits = otmp - shift / 32;
  launch_extract_k(c, mid, hi, slots, cap, seenpos - shift, wbits - shift / 32, dbits - shift / 32, arec - shift / 32, extc);
  c->obits = obits_save;
  *n_ext = dread(c, extc); *n_dup = dread(c, extc + 1); *chunk_n = hi - mid;
  dfree(c, slots); dfree(c, arec); dfree(c, extc); dfree(c, wbits); dfree(c, dbits); dfree(c, seenpos); dfree(c, otmp);
  return 0;
}

static int dedup_global_table(spb_ctx* c, bool allow_switch) {
  const u64 T = c->T;
  int r0 = alloc_position_state(c); if (r0) return r0;
  int decided = 0;
  u64 CH = 1ull << 22, switch_min = 8ull << 20;
  if (const char* e = getenv("SPB_CHUNK_BITS")) { int b = atoi(e); if (b >= 8 && b <= 32) CH = 1ull << b; }   // tests: tiny chunks
  if (const char* e = getenv("SPB_SWITCH_MIN")) switch_min = strtoull(e, nullptr, 10);                        // tests: small inputs switch too
  if (allow_switch && c->n_inst >= switch_min && !getenv("SPB_ANCHOR") && c->npos_pad > CH) {
    u64 n_ext = 0, n_dup = 0, chunk_n = 1;
    u64 CHp = 256;                                     // sample chunks: a power of two, at most CH / 2 positions and 1/64 of the input
    while (2 * CHp <= CH / 2 && 2 * CHp <= c->n_inst / 64) CHp *= 2;
    PRIM(probe_sample_chunks(c, CHp, &n_ext, &n_dup, &chunk_n));
    decided = (double)n_ext >= 0.25 * (double)chunk_n ? 2 : -1;
    // partitioned dedup only pays when almost nothing is duplicated (every duplicate costs it a scattered write)
    if (decided < 0 && (double)n_dup < 0.02 * (double)chunk_n) return 1;
  }
  u64 cap = 1024; while (cap < 2 * c->n_inst) cap <<= 1;
  c->cnt.table_slots = cap;
  u64* slots; ALLOC(slots, u64, cap);
  dev_memset(c, slots, 0xFF, cap * 8);
  u64 nw32 = c->npos_pad / 32;
  u32 *wbits, *dbits; ALLOC(wbits, u32, nw32 + 8); ALLOC(dbits, u32, nw32 + 8);
  dev_memset(c, dbits, 0, (nw32 + 8) * 4); dev_memset(c, wbits, 0, (nw32 + 8) * 4);
  // Positions go through the table in chunks, in order.  Chunk 0 is plain.  From chunk 1 on, the "anchor + extend"
  // scheme is tried: if the second chunk extends well (genome sets that share long stretches: the realistic case), all
  // remaining chunks use it, each seeing the complete table of the chunks before it; otherwise (nearly all k-mers
  // unique) the rest is one plain launch.  The result is the same either way.
  const u64 P = c->npos_pad;
  u64 nA = P / 32;
  u32* seenpos; u64 *arec, *extc; ALLOC(seenpos, u32, P + 8); ALLOC(arec, u64, nA + 8); ALLOC(extc, u64, 2);
  dev_memset(c, extc, 0, 16);
  int tk = t_begin(c);
  int anchored = decided;
  if (const char* e = getenv("SPB_ANCHOR")) anchored = atoi(e) ? 2 : -1;      // 1 = force on, 0 = force off (A/B measurements, tests)
  u64 done = std::min(P, CH);
  launch_extract_k(c, 0, done, slots, cap, seenpos, wbits, dbits, nullptr, nullptr);
  if (done < P && anchored >= 0) {
    u64 hi = std::min(P, done + CH);
    launch_extract_k(c, done, hi, slots, cap, seenpos, wbits, dbits, arec, extc);
    const u64 n_ext = dread(c, extc), n_dup = dread(c, extc + 1), chunk_n = hi - done;
    if (anchored == 0) anchored = (double)n_ext >= 0.25 * (double)chunk_n ? 1 : -1;
    done = hi;
    // partitioned dedup only pays when almost nothing is duplicated (every duplicate costs it a scattered write)
    if (anchored < 0 && allow_switch && c->n_inst >= switch_min && (double)n_dup < 0.02 * (double)chunk_n) {
      // mostly unique k-mers and a table far beyond L2: the partitioned "links" dedup is faster; start over with it
      dfree(c, arec); dfree(c, extc); dfree(c, seenpos); dfree(c, slots); dfree(c, wbits); dfree(c, dbits);
      return 1;
    }
  }
  if (anchored > 0) for (; done < P; done += CH) launch_extract_k(c, done, std::min(P, done + CH), slots, cap, seenpos, wbits, dbits, arec, nullptr);
  else launch_extract_k(c, done, P, slots, cap, seenpos, wbits, dbits, nullptr, nullptr);
  t_end(c, tk, "k_extract_insert_ms"); stat_set(c, "anchored", anchored > 0 ? 1 : 0);
  dfree(c, arec); dfree(c, extc);
  u64 nb = c->npos_pad / 256;
  SPB_LAUNCH(k_first_bits, nblk(nb, TPB), TPB, c->stream, wbits, dbits, c->fbits, nb);
  dev_memset(c, c->fbits + nw32, 0, 32);
  dfree(c, wbits); dfree(c, dbits);
  int r = dd_ids(c, 0, T, slots, cap, seenpos);
  dfree(c, slots); dfree(c, seenpos);
  return r;
}

int spb_build_dbg(spb_ctx* c, spb_counts* out) {
  if (!c) return SPB_EINVAL;
  if (c->state < 1) return fail(c, SPB_ESTATE, "spb_build_dbg before spb_load");
  if (c->state > 1) return fail(c, SPB_ESTATE, "spb_build_dbg called twice; call spb_load again");
  SPB_ON_DEVICE(c);
  stats_reset(c);
  int tm;
  // Default: single global table (two random HBM accesses per instance).  SPB_DEDUP=staged selects the partitioned
  // path (records -> buckets -> L2 scratch tables -> scatter back); measured slower on one GPU in round 1
  // (profiles/README.md), its stages are what the multi-GPU exchange is built from.
  const char* mode = getenv("SPB_DEDUP");          // global | links | staged  (default: adaptive global -> links)
  if (mode && (!strcmp(mode, "links") || !strcmp(mode, "smem") || !strcmp(mode, "part"))) {
    tm = t_begin(c);
    int rc = !strcmp(mode, "part") ? dedup_part_links(c) : 1;
    if (rc == 1 && strcmp(mode, "links")) { stats_reset(c); tm = t_begin(c); rc = dedup_smem_links(c); }
    if (rc == 1) { stats_reset(c); tm = t_begin(c); rc = dedup_staged_links(c); }
    if (rc) return rc;
    t_end(c, tm, "dedup_links_ms");
  } else if (!(mode && !strcmp(mode, "staged"))) {
    tm = t_begin(c);
    int rc = dedup_global_table(c, !mode);            // also reports k_extract_insert_ms (the kernel alone)
    if (rc == 1) {
      stat_set(c, "switched_to_links", 1);
      const auto drop_stats = [&]() { c->stat.erase(std::remove_if(c->stat.begin(), c->stat.end(), [](const spb_ctx::StatRec& r) { return r.name == "records_partition_ms" || r.name == "bucket_dedup_ms" || r.name == "smem_tables" || r.name == "k_roll_partition_ms" || r.name == "k_partition2_ms" || r.name == "k_bucket_dedup_ms" || r.name == "k_make_records_roll_ms" || r.name == "k_smem_dedup_ms"; }), c->stat.end()); };
      rc = (getenv("SPB_NO_SMEM_TABLES") || getenv("SPB_NO_PARTITION")) ? 1 : dedup_part_links(c);
      if (rc == 1) { drop_stats(); rc = getenv("SPB_NO_SMEM_TABLES") ? 1 : dedup_smem_links(c); }
      if (rc == 1) { drop_stats(); rc = dedup_staged_links(c); }
    }
    if (rc) return rc;
    t_end(c, tm, "dedup_global_ms");
  } else {
    // ---- S1: records + partition
    tm = t_begin(c);
    u32 bb = default_bucket_bits(c->n_inst, 1);
    if (const char* e = getenv("SPB_BUCKET_BITS")) bb = (u32)atoi(e);
    PRIM(dd_records(c, 0, c->T, bb));
    t_end(c, tm, "records_partition_ms");
    // ---- S2: L2-staged dedup per bucket
    tm = t_begin(c);
    u32* rep; ALLOC(rep, u32, c->dd_n + 1);
    u64 cc = c->dd_n;
    PRIM(dd_owner(c, c->dd_keys, c->dd_pos, &cc, 1, bb, 0, 1u << bb, rep));
    t_end(c, tm, "bucket_dedup_ms");
    // ---- S3 + ids
    tm = t_begin(c);
    PRIM(dd_scatter(c, rep));
    dfree(c, rep);
    PRIM(dd_ids(c, 0, c->T, nullptr));
    t_end(c, tm, "scatter_ids_ms");
  }
  tm = t_begin(c);
  PRIM(dd_graph(c));
  t_end(c, tm, "edges_ms");
  int r = check_err(c, "spb_build_dbg");
  if (r) return r;
  c->state = 2;
  if (out) *out = c->cnt;
  return SPB_OK;
}

// ------------------------------------------------------------------------------------------- multi-GPU stages
// One process per GPU, every rank holds the whole packed input (spb_load).  Positions are split into equal
// slices (one per rank); k-mers are hash-partitioned over the ranks (owner = top log2(n_owners) hash bits):
//   spb_mg_records -> [all-to-all of (hash, position) records] -> spb_mg_dedup (exact dedup in the owner's table)
//   -> [all-to-all of representatives back] -> spb_mg_scatter -> [all-gather of the first-appearance and
//   orientation bitmaps] -> spb_mg_ids -> [all-gather of the per-position vertex ids] -> spb_mg_graph.
// The exchanges are NCCL collectives issued by the host side (torch.distributed) on the device pointers below.
static bool pow2(u32 x) { return x && !(x & (x - 1)); }

int spb_mg_slice(spb_ctx* c, uint32_t n_ranks, uint64_t* slice_len) {
  if (!c || !slice_len || !n_ranks) return SPB_EINVAL;
  *slice_len = nblk(nblk(c->T, n_ranks), 1024) * 1024;
  return SPB_OK;
}

int spb_mg_records(spb_ctx* c, uint64_t pos_lo, uint64_t pos_hi, uint32_t n_owners, uint64_t* owner_counts, void** keys_dev, void** pos_dev) {
  if (!c) return SPB_EINVAL;
  if (c->state != 1) return fail(c, SPB_ESTATE, "spb_mg_records needs a freshly loaded input");
  if (!pow2(n_owners)) return fail(c, SPB_EINVAL, "n_owners must be a power of two");
  if (pos_lo % 256) return fail(c, SPB_EINVAL, "pos_lo must be a multiple of 256");
  SPB_ON_DEVICE(c);
  u32 bb = 0; while ((1u << bb) < n_owners) ++bb;
  uint64_t slice = 0; spb_mg_slice(c, n_owners, &slice);
  if (pos_hi > c->T) pos_hi = c->T;
  PRIM(dd_records(c, pos_lo, pos_hi, bb, slice * n_owners));
  for (u32 o = 0; o < n_owners; ++o) owner_counts[o] = c->dd_bstart_h[o + 1] - c->dd_bstart_h[o];
  *keys_dev = c->dd_keys; *pos_dev = c->dd_pos;
  return check_err(c, "spb_mg_records");
}

int spb_mg_dedup(spb_ctx* c, const uint64_t* keys, const uint32_t* pos, const uint64_t* chunk_counts, uint32_t n_chunks,
                 uint32_t n_owners, uint32_t owner, uint32_t* rep_out) {
  if (!c) return SPB_EINVAL;
  if (c->state != 1) return fail(c, SPB_ESTATE, "spb_mg_dedup needs a loaded input");
  if (!pow2(n_owners) || owner >= n_owners) return fail(c, SPB_EINVAL, "bad owner / n_owners");
  SPB_ON_DEVICE(c);
  u32 bb = 0; while ((1u << bb) < n_owners) ++bb;
  PRIM(dd_owner(c, (const u64*)keys, pos, (const u64*)chunk_counts, n_chunks, bb, owner, owner + 1, rep_out));
  return check_err(c, "spb_mg_dedup");
}

// Duplicate profile of the input from its first two 4 M-position chunks (chunk 0 into a scratch table, chunk 1 probed with
// the anchor + extend scheme): extended fraction and duplicate fraction of chunk 1.  The multi-GPU host side uses it (on
// rank 0, broadcast) to choose the return protocol: links when almost nothing is duplicated, representatives otherwise.
int spb_mg_probe(spb_ctx* c, double* ext_ratio, double* dup_ratio) {
  if (!c || !ext_ratio || !dup_ratio) return SPB_EINVAL;
  if (c->state != 1) return fail(c, SPB_ESTATE, "spb_mg_probe needs a freshly loaded input");
  SPB_ON_DEVICE(c);
  u64 CH = 1ull << 22;
  if (const char* e = getenv("SPB_CHUNK_BITS")) { int b = atoi(e); if (b >= 8 && b <= 32) CH = 1ull << b; }
  *ext_ratio = 0; *dup_ratio = 0;
  if (c->npos_pad <= CH) return SPB_OK;              // a single chunk: nothing to compare against
  u64 n_ext = 0, n_dup = 0, chunk_n = 0;
  PRIM(probe_sample_chunks(c, std::max<u64>(CH / 2, 256), &n_ext, &n_dup, &chunk_n));
  if (chunk_n) { *ext_ratio = (double)n_ext / (double)chunk_n; *dup_ratio = (double)n_dup / (double)chunk_n; }
  return check_err(c, "spb_mg_probe");
}

// links (from << 32 | to): sort by position, resolve chains, split by the rank that owns the position slice
static int finish_links(spb_ctx* c, u64* links, u64 nl, u32 n_owners, uint64_t* send_counts, void** links_dev) {
  u64* tmp; ALLOC(tmp, u64, nl + 1);
  PRIM(prim_sort_u64(c, links, tmp, nl));
  SPB_LAUNCH(k_links_resolve, nblk(nl, TPB), TPB, c->stream, links, nl, tmp);
  uint64_t slice = 0; spb_mg_slice(c, n_owners, &slice);
  u64* bounds; ALLOC(bounds, u64, n_owners + 2);
  SPB_LAUNCH(k_links_bounds, nblk(n_owners + 1, TPB), TPB, c->stream, tmp, nl, slice, n_owners, bounds);
  std::vector<u64> bh(n_owners + 1);
  dev_copy(c, bh.data(), bounds, (n_owners + 1) * 8ull); dev_sync(c);
  for (u32 r = 0; r < n_owners; ++r) send_counts[r] = bh[r + 1] - bh[r];
  dfree(c, bounds); dfree(c, links);
  dfree(c, c->mg_links);
  c->mg_links = tmp;
  *links_dev = tmp;
  return 0;
}

// Owner side, "links" protocol: dedup of the received records in L2-resident sub-bucket tables; returns only the
// exceptions as (position -> first position) links, sorted by position, with per-destination-rank counts.
int spb_mg_dedup_links(spb_ctx* c, uint64_t* keys_io, uint32_t* pos_io, uint64_t n_total, uint32_t n_owners, uint64_t* send_counts,
                       void** links_dev) {
  if (!c) return SPB_EINVAL;
  if (c->state != 1) return fail(c, SPB_ESTATE, "spb_mg_dedup_links needs a loaded input");
  if (!pow2(n_owners)) return fail(c, SPB_EINVAL, "n_owners must be a power of two");
  SPB_ON_DEVICE(c);
  u32 bbo = 0; while ((1u << bbo) < n_owners) ++bbo;
  u32 sb = 0; while (sb < 8 && (n_total >> sb) > 2200000ull) ++sb;          // sub-buckets of ~2 M records
  if (const char* e = getenv("SPB_BUCKET_BITS")) sb = (u32)atoi(e);
  u64* keys = (u64*)keys_io; u32* pos = pos_io;
  u64 *k2 = nullptr, *k2a = nullptr; u32 *p2 = nullptr, *p2a = nullptr;
  if (sb > 0 && n_total) {
    ALLOC(k2, u64, n_total); ALLOC(p2, u32, n_total);
    k2a = k2; p2a = p2;                                // the partition may swap the roles of the two buffer pairs
    PRIM(prim_partition_u64_u32(c, keys, pos, k2, p2, n_total, 64 - bbo - sb, 64 - bbo));
  }
  // (after the partition `keys` / `pos` point at the sorted copy, which may be the scratch pair)
  const u32 B = 1u << sb;
  u64 *coff, *sub; ALLOC(coff, u64, 2); ALLOC(sub, u64, B + 2);
  u64 co[2] = {0, n_total};
  memcpy(c->pinned, co, 16); dev_copy(c, coff, c->pinned, 16);
  // bucket = bits [64-bbo-sb, 64-bbo): shift the owner bits out by comparing on (key << bbo)
  std::vector<u64> bstart(B + 1, 0);
  if (n_total) {
    SPB_LAUNCH(k_bucket_bounds_shift, nblk(B + 1, TPB), TPB, c->stream, keys, n_total, bbo, sb, B, sub);
    dev_copy(c, bstart.data(), sub, (B + 1) * 8ull); dev_sync(c);
  }
  u64 nmax = 0;
  for (u32 b = 0; b < B; ++b) nmax = std::max(nmax, bstart[b + 1] - bstart[b]);
  u64 cap = 1024; while (cap < 2 * nmax) cap <<= 1;
  c->cnt.table_slots = cap;
  u64* scratch[4] = {nullptr, nullptr, nullptr, nullptr};
  u64* nlinks; ALLOC(nlinks, u64, 2); dev_memset(c, nlinks, 0, 16);
  u64 lcap = std::max<u64>(4096, n_total / 8);
  u64* links = nullptr; u64 nl = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    ALLOC(links, u64, lcap + 1);
    dev_memset(c, nlinks, 0, 16);
    {
      TwoStreams ts(c);
      for (int i = 0; i < ts.n; ++i) if (!scratch[i]) ALLOC(scratch[i], u64, cap);
      u32 step = 0;
      for (u32 b = 0; b < B; ++b) {
        u64 lo = bstart[b], n = bstart[b + 1] - lo;
        if (n == 0) continue;
        const int i = step % ts.n;
        ts.clear(c, i, scratch[i], cap * 8);
        SPB_LAUNCH(k_bkt_insert_list, nblk(n, TPB), TPB, ts.st[i], keys, pos, lo, n, c->F, c->R, c->T, c->k, scratch[i], cap - 1, links, nlinks, lcap,
                   (u64*)nullptr, 0);
        ++step;
      }
      ts.join(c);
    }
    nl = dread(c, nlinks);
    if (nl <= lcap) break;
    dfree(c, links); lcap = nl;
  }
  for (auto& x : scratch) dfree(c, x);
  dfree(c, nlinks); dfree(c, coff); dfree(c, sub);
  dfree(c, k2a); dfree(c, p2a);
  if (getenv("SPB_DEBUG")) fprintf(stderr, "[spb] dedup_links: n_total=%llu B=%u cap=%llu links=%llu\n", (unsigned long long)n_total, B, (unsigned long long)cap, (unsigned long long)nl);
  int r = finish_links(c, links, nl, n_owners, send_counts, links_dev);
  return r ? r : check_err(c, "spb_mg_dedup_links");
}

// Multi-GPU, "owner computes" protocol: no record travels.  Every rank walks ALL positions with the rolling hashes (a few
// ms) and keeps only the records whose hash it owns, sorts them into sub-buckets and deduplicates them in shared-memory
// tables (k_smem_dedup, list form).  Output as spb_mg_dedup_links: links (position -> first position) sorted by
// position, split by the rank that owns the position slice.  Also leaves the orientation bitmap complete on every rank.
int spb_mg_owned_links(spb_ctx* c, uint32_t rank, uint32_t n_owners, uint64_t* send_counts, void** links_dev) {
  if (!c) return SPB_EINVAL;
  if (c->state != 1) return fail(c, SPB_ESTATE, "spb_mg_owned_links needs a loaded input");
  if (!pow2(n_owners) || n_owners < 2 || rank >= n_owners) return fail(c, SPB_EINVAL, "n_owners must be a power of two >= 2, rank < n_owners");
  if (c->k < 17) return fail(c, SPB_EINVAL, "the owner-computes protocol needs k >= 17 (rolling hashes over 16-base windows): use the record protocols");
  SPB_ON_DEVICE(c);
  uint64_t slice = 0; spb_mg_slice(c, n_owners, &slice);
  int r0 = alloc_position_state(c, slice * n_owners); if (r0) return r0;
  const u64 T = c->T, n = nblk(T, TPB) * TPB;
  u32 bbo = 0; while ((1u << bbo) < n_owners) ++bbo;
  const RollK K = roll_constants(c->k);
  u64* count; ALLOC(count, u64, 2);
  u64 cap = n / n_owners + n / (8 * n_owners) + 65536, n_own = 0;
  u64 *keys = nullptr, *keys2 = nullptr; u32 *pos = nullptr, *pos2 = nullptr;
  for (int attempt = 0; attempt < 2; ++attempt) {
    ALLOC(keys, u64, cap); ALLOC(pos, u32, cap);
    dev_memset(c, count, 0, 16);
    SPB_LAUNCH(k_make_records_roll, nblk(n / SPB_RUN, 128), 128, c->stream, c->F, c->R, c->vbits, T, c->k, 0ull, T, n, keys, pos, c->obits, 0u, K,
               bbo, rank, count, cap);
    n_own = dread(c, count);
    if (n_own <= cap) break;
    dfree(c, keys); dfree(c, pos); cap = n_own;      // skewed hashes (masses of identical k-mers at one owner)
  }
  dfree(c, count);
  u64 avg = 7680;
  if (const char* e = getenv("SPB_SMEM_AVG")) avg = std::max(1, atoi(e));
  u32 nbits = 0; while ((n_own >> nbits) > avg && nbits < 24) ++nbits;
  if (nbits > 0 && n_own) {
    ALLOC(keys2, u64, n_own); ALLOC(pos2, u32, n_own);
    PRIM(prim_partition_u64_u32(c, keys, pos, keys2, pos2, n_own, 64 - (int)bbo - (int)nbits, 64 - (int)bbo));
    dfree(c, keys2); dfree(c, pos2);
  }
  const u64 NB = 1ull << nbits;
  u64* sub; ALLOC(sub, u64, NB + 2);
  std::vector<u64> bstart(NB + 1, 0);
  if (n_own) {
    SPB_LAUNCH(k_bucket_bounds_shift, nblk(NB + 1, TPB), TPB, c->stream, keys, n_own, bbo, nbits, (u32)NB, sub);
    dev_copy(c, bstart.data(), sub, (NB + 1) * 8ull); dev_sync(c);
  }
  u64 nmax = 0;
  for (u64 b = 0; b < NB; ++b) nmax = std::max(nmax, bstart[b + 1] - bstart[b]);
  u64* nlinks; ALLOC(nlinks, u64, 2);
  u64 lcap = std::max<u64>(4096, n_own / 8);
  u64* links = nullptr; u64 nl = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    ALLOC(links, u64, lcap + 1);
    dev_memset(c, nlinks, 0, 16);
    if (n_own == 0) break;
    if (nmax <= smem_limit()) {
      c->cnt.table_slots = SPB_ST_SLOTS;
      SPB_LAUNCH_PHASED_DSMEM(k_smem_dedup<true>, 2, NB, 512, SPB_ST_SLOTS * 4, c->stream, keys, pos, sub, c->F, c->R, T, c->k, bbo + nbits,
                              (u32*)nullptr, (u32*)nullptr, c->errflag, links, nlinks, lcap);
    } else {
      // a sub-bucket beyond the shared-memory table (masses of identical k-mers): groups of sub-buckets through tables in L2
      u32 gb = nbits > 6 ? 6 : nbits;                // <= 64 groups
      u64 gmax = 0;
      for (u64 g = 0; g < (1ull << gb); ++g) gmax = std::max(gmax, bstart[(g + 1) << (nbits - gb)] - bstart[g << (nbits - gb)]);
      u64 tcap = 1024; while (tcap < 2 * gmax) tcap <<= 1;
      c->cnt.table_slots = tcap;
      u64* scratch; ALLOC(scratch, u64, tcap);
      for (u64 g = 0; g < (1ull << gb); ++g) {
        const u64 lo = bstart[g << (nbits - gb)], ng = bstart[(g + 1) << (nbits - gb)] - lo;
        if (ng == 0) continue;
        dev_memset(c, scratch, 0xFF, tcap * 8);
        SPB_LAUNCH(k_bkt_insert_list, nblk(ng, TPB), TPB, c->stream, keys, pos, lo, ng, c->F, c->R, T, c->k, scratch, tcap - 1, links, nlinks, lcap,
                   (u64*)nullptr, 0);
      }
      dfree(c, scratch);
    }
    nl = dread(c, nlinks);
    if (nl <= lcap) break;
    dfree(c, links); lcap = nl;
  }
  dfree(c, nlinks); dfree(c, sub); dfree(c, keys); dfree(c, pos);
  c->mg_owned = true;
  int r = finish_links(c, links, nl, n_owners, send_counts, links_dev);
  return r ? r : check_err(c, "spb_mg_owned_links");
}

// Multi-GPU, "part" protocol (default for mostly-unique inputs): record exchange with the hand-written partition kernels.
//   spb_mg_part_records   position side: rolling hashes of the local slice fused with the partition by (owner, level-1
//                         bucket) -- the owner is simply the top bits of the level-1 bucket index, so the regions of one
//                         owner are contiguous: [owner][2^l1][cap1] records of 8 bytes + [owner][2^l1] counts.  The host
//                         exchanges both with equal-split all-to-alls (the few per cent of slack travel along).
//   spb_mg_part_dedup     owner side: level-2 partition of the n_src x 2^l1 received regions, one CTA per final bucket
//                         with a shared-memory table (list form), links sorted by position and split by slice owner.
// The geometry (l1, l2, cap1) is a function of (n_inst, n_owners, slice length) only, so every rank derives the same one.
struct PartGeom { u32 own_bits, l1, l2, l2a; u64 cap1, capA; };      // l2a > 0: two level-2 passes (l2a bits, then l2 - l2a)
static int part_geometry(spb_ctx* c, u32 n_owners, PartGeom* g) {
  u32 ob = 0; while ((1u << ob) < n_owners) ++ob;
  u64 avg = SPB_D_AVG;
  if (const char* e = getenv("SPB_PART_AVG")) avg = std::max(1, atoi(e));
  const u64 per_owner = nblk(c->n_inst, n_owners);
  u32 nbits = 0; while ((per_owner >> nbits) > avg) ++nbits;
  if (ob > 9) return 1;
  u32 l1 = std::min<u32>(9 - ob, nbits);
  if (l1 > (nbits + 1) / 2 + 1) l1 = (nbits + 1) / 2 + 1;         // keep both levels at a useful fan-out
  const u32 l2 = nbits - l1;
  u32 l2max = 9;                                                   // one level-2 pass handles 2^9 buckets (tests lower it to reach the two-pass form)
  if (const char* e = getenv("SPB_PART_L2MAX")) l2max = (u32)std::max(1, std::min(9, atoi(e)));
  if (l2 > 2 * l2max) return 1;
  g->l2a = l2 > l2max ? (l2 + 1) / 2 : 0;                          // more than that: the level-2 kernel runs twice
  const u64 perA = per_owner >> (l1 + g->l2a);
  g->capA = (perA + perA / 8 + 4096 + 1) & ~1ull;
  uint64_t slice = 0; spb_mg_slice(c, n_owners, &slice);
  const u64 n1 = nblk(slice, (u64)n_owners << l1);                // positions of a slice per (owner, level-1 bucket)
  g->own_bits = ob; g->l1 = l1; g->l2 = l2;
  g->cap1 = (n1 + n1 / 16 + 2048 + 1) & ~1ull;
  return 0;
}
int spb_mg_part_records(spb_ctx* c, uint64_t pos_lo, uint64_t pos_hi, uint32_t n_owners, uint64_t* region_records, uint32_t* n_regions_per_owner,
                        void** recs_dev, void** counts_dev, int* overflow) {
  if (!c || !region_records || !n_regions_per_owner || !recs_dev || !counts_dev || !overflow) return SPB_EINVAL;
  if (c->state != 1) return fail(c, SPB_ESTATE, "spb_mg_part_records needs a freshly loaded input");
  if (!pow2(n_owners)) return fail(c, SPB_EINVAL, "n_owners must be a power of two");
  if (pos_lo % 256) return fail(c, SPB_EINVAL, "pos_lo must be a multiple of 256");
  SPB_ON_DEVICE(c);
  PartGeom g;
  *overflow = 0;
  if (c->k < 17 || part_geometry(c, n_owners, &g)) { *overflow = 1; return SPB_OK; }      // tiny k: the caller falls back to the record protocols
  uint64_t slice = 0; spb_mg_slice(c, n_owners, &slice);
  int r0 = alloc_position_state(c, slice * n_owners); if (r0) return r0;
  if (pos_hi > c->T) pos_hi = c->T;
  if (pos_hi < pos_lo) pos_hi = pos_lo;
  const u64 NR = (u64)n_owners << g.l1;
  dfree(c, c->dd_keys); dfree(c, c->dd_bstart);
  u64* rec1; u32* cursor1;
  ALLOC(rec1, u64, NR * g.cap1 + 8192 + 2); ALLOC(cursor1, u32, NR + 8);
  dev_memset(c, cursor1, 0, (NR + 8) * 4);
  const u32 run = 256;
  { const int tq_ = t_begin(c); launch_roll_partition(c, pos_lo, pos_hi, run, nblk(pos_hi - pos_lo, run), 0, 0, g.own_bits + g.l1, 1, rec1, g.cap1, cursor1); t_end(c, tq_, "k_roll_partition_ms"); }
  const u32 e = dread(c, c->errflag);
  if (e & ERR_OVERFLOW) { dev_memset(c, c->errflag, 0, 4); *overflow = 1; }
  c->dd_keys = rec1; c->dd_bstart = (u64*)cursor1;               // owned by the context until spb_mg_apply_links
  c->mg_owned = true;                                            // (the host all-gathers the orientation bitmap right after this call)
  *region_records = g.cap1 << g.l1; *n_regions_per_owner = 1u << g.l1;
  *recs_dev = rec1; *counts_dev = cursor1;
  return (e & ~(u32)ERR_OVERFLOW) ? check_err(c, "spb_mg_part_records") : SPB_OK;
}
int spb_mg_part_dedup(spb_ctx* c, const uint64_t* recs, const uint32_t* counts, uint32_t n_src, uint32_t n_owners, uint64_t* send_counts,
                      void** links_dev, int* overflow) {
  if (!c || !send_counts || !links_dev || !overflow) return SPB_EINVAL;
  if (c->state != 1) return fail(c, SPB_ESTATE, "spb_mg_part_dedup needs a loaded input");
  SPB_ON_DEVICE(c);
  PartGeom g;
  *overflow = 0;
  if (part_geometry(c, n_owners, &g)) { *overflow = 1; return SPB_OK; }
  const u64 NB1 = 1ull << g.l1, NR1 = NB1 * n_src, NBF = 1ull << (g.l1 + g.l2), cap2 = SPB_D_LIMIT;
  u64* rec2; u32* cursor2;
  ALLOC(rec2, u64, NBF * cap2 + SPB_P2_CHUNK + 2); ALLOC(cursor2, u32, NBF + 8);
  dev_memset(c, cursor2, 0, (NBF + 8) * 4);
  const u32 cpr = (u32)nblk(g.cap1, SPB_P2_CHUNK);
#ifdef SPB_EMU
  const u64 grid2 = NR1 * cpr;
#else
  const u64 grid2 = std::min<u64>(NR1 * cpr, 2 * 148);
#endif
  if (!g.l2a) {
    const int tq_ = t_begin(c); SPB_LAUNCH_PHASED_DSMEM(k_partition2, 1, grid2, SPB_P2_T, 3 * SPB_P2_CHUNK * 8, c->stream, (const u64*)recs, g.cap1, counts, cpr, (u32)NR1, (u32)NB1, g.l2, rec2, cap2,
                            NBF * cap2, cursor2, c->errflag, 0u); t_end(c, tq_, "k_partition2_ms");
  } else {
    // very large inputs (more than 2^18 final buckets per owner's level-1 fan-out): level 2 in two passes through an
    // intermediate set of 2^(l1 + l2a) regions
    const u64 NBA = 1ull << (g.l1 + g.l2a);
    if (getenv("SPB_DEBUG")) fprintf(stderr, "[spb] part dedup: level 2 in two passes (l1=%u l2=%u+%u, %llu intermediate regions of %llu)\n", g.l1, g.l2a, g.l2 - g.l2a, (unsigned long long)NBA, (unsigned long long)g.capA);
    stat_set(c, "partition2_passes", 2);
    u64* recA; u32* cursorA;
    ALLOC(recA, u64, NBA * g.capA + SPB_P2_CHUNK + 2); ALLOC(cursorA, u32, NBA + 8);
    dev_memset(c, cursorA, 0, (NBA + 8) * 4);
    const int tq_ = t_begin(c);
    SPB_LAUNCH_PHASED_DSMEM(k_partition2, 1, grid2, SPB_P2_T, 3 * SPB_P2_CHUNK * 8, c->stream, (const u64*)recs, g.cap1, counts, cpr, (u32)NR1, (u32)NB1, g.l2a, recA, g.capA,
                            NBA * g.capA, cursorA, c->errflag, 0u);
    const u32 cprA = (u32)nblk(g.capA, SPB_P2_CHUNK);
#ifdef SPB_EMU
    const u64 gridB = NBA * cprA;
#else
    const u64 gridB = std::min<u64>(NBA * cprA, 2 * 148);
#endif
    SPB_LAUNCH_PHASED_DSMEM(k_partition2, 1, gridB, SPB_P2_T, 3 * SPB_P2_CHUNK * 8, c->stream, (const u64*)recA, g.capA, cursorA, cprA, (u32)NBA, (u32)NBA, g.l2 - g.l2a, rec2, cap2,
                            NBF * cap2, cursor2, c->errflag, g.l2a);
    t_end(c, tq_, "k_partition2_ms");
    dfree(c, recA); dfree(c, cursorA);
  }
  u64* nlinks; ALLOC(nlinks, u64, 2);
  u64 lcap = std::max<u64>(4096, c->n_inst / n_owners / 8);
  u64* links = nullptr; u64 nl = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    ALLOC(links, u64, lcap + 1);
    dev_memset(c, nlinks, 0, 16);
    { const int tq_ = t_begin(c); SPB_LAUNCH_PHASED_DSMEM((k_bucket_dedup<true, 1, false>), 3, NBF, SPB_D_T, SPB_D_SMEM_U64(1) * 8, c->stream, rec2, cap2, cursor2, (u32)NBF, g.l2, c->F, c->R, c->obits, c->T, c->k,
                            (u32*)nullptr, (u32*)nullptr, c->errflag, links, nlinks, lcap, 0xFFFFFFFFu); t_end(c, tq_, "k_bucket_dedup_ms"); }
    nl = dread(c, nlinks);
    if (nl <= lcap) break;
    dfree(c, links); lcap = nl;
  }
  dfree(c, nlinks); dfree(c, rec2); dfree(c, cursor2);
  c->cnt.table_slots = SPB_D_SLOTS;
  const u32 e = dread(c, c->errflag);
  if (e & ERR_OVERFLOW) { dev_memset(c, c->errflag, 0, 4); *overflow = 1; dfree(c, links); for (u32 r = 0; r < n_owners; ++r) send_counts[r] = 0; *links_dev = nullptr; return SPB_OK; }
  int r = finish_links(c, links, nl, n_owners, send_counts, links_dev);
  return r ? r : check_err(c, "spb_mg_part_dedup");
}

// Position side, "links" protocol: the links received for the local slice -> first-appearance bits of the slice.
int spb_mg_apply_links(spb_ctx* c, const uint64_t* links, uint64_t n, uint64_t pos_lo, uint64_t pos_hi) {
  if (!c) return SPB_EINVAL;
  SPB_ON_DEVICE(c);
  if (c->state != 1 || (!c->dd_pos && !c->mg_owned)) return fail(c, SPB_ESTATE, "spb_mg_apply_links without spb_mg_records / spb_mg_owned_links");
  const u64 P = c->npos_pad, nw32 = P / 32;
  ALLOC(c->mg_seen, u32, P + 8); ALLOC(c->mg_dup, u32, nw32 + 8);
  dev_memset(c, c->mg_dup, 0, (nw32 + 8) * 4);
  SPB_LAUNCH(k_links_apply, nblk(n, TPB), TPB, c->stream, (const u64*)links, n, c->mg_seen, c->mg_dup);
  if (getenv("SPB_DEBUG")) { std::vector<u64> h(std::min<u64>(n, 4)); dev_copy(c, h.data(), links, h.size() * 8); dev_sync(c); fprintf(stderr, "[spb] apply_links: n=%llu first=%llx %llx\n", (unsigned long long)n, n ? (unsigned long long)h[0] : 0ull, n > 1 ? (unsigned long long)h[1] : 0ull); }
  // first appearance = valid and not linked; only the words of the local slice are meaningful (all-gathered next)
  // (vbits covers the positions of the input only; the padded tail of the grown position arrays stays zero)
  SPB_LAUNCH(k_first_bits, nblk(nblk(c->T, 256), TPB), TPB, c->stream, c->vbits, c->mg_dup, c->fbits, nblk(c->T, 256));
  dfree(c, c->mg_dup); dfree(c, c->mg_links);
  dfree(c, c->dd_keys); dfree(c, c->dd_pos); dfree(c, c->dd_bstart);
  (void)pos_lo; (void)pos_hi;
  return check_err(c, "spb_mg_apply_links");
}

int spb_mg_scatter(spb_ctx* c, const uint32_t* rep) {
  if (!c) return SPB_EINVAL;
  SPB_ON_DEVICE(c);
  if (c->state != 1 || !c->dd_pos) return fail(c, SPB_ESTATE, "spb_mg_scatter without spb_mg_records");
  PRIM(dd_scatter(c, rep));
  return check_err(c, "spb_mg_scatter");
}

int spb_mg_buffers(spb_ctx* c, void** fbits_dev, void** obits_dev, void** sp_dev) {
  if (!c || !c->sp) return SPB_EINVAL;
  if (fbits_dev) *fbits_dev = c->fbits;
  if (obits_dev) *obits_dev = c->obits;
  if (sp_dev) *sp_dev = c->sp;
  return SPB_OK;
}

int spb_mg_ids(spb_ctx* c, uint64_t pos_lo, uint64_t pos_hi) {
  if (!c) return SPB_EINVAL;
  SPB_ON_DEVICE(c);
  if (c->state != 1 || !c->sp) return fail(c, SPB_ESTATE, "spb_mg_ids out of order");
  if (pos_hi > c->T) pos_hi = c->T;
  PRIM(dd_ids(c, pos_lo, pos_hi, nullptr, 1, c->mg_seen));       // links protocol: the representatives live in mg_seen
  dfree(c, c->mg_seen);
  return check_err(c, "spb_mg_ids");
}

int spb_mg_graph(spb_ctx* c, spb_counts* out) {
  if (!c) return SPB_EINVAL;
  SPB_ON_DEVICE(c);
  if (c->state != 1 || !c->v2pos) return fail(c, SPB_ESTATE, "spb_mg_graph out of order");
  stats_reset(c); stat_set(c, "multi_gpu", 1);
  int tm = t_begin(c);
  PRIM(dd_graph(c));
  t_end(c, tm, "edges_ms");
  int r = check_err(c, "spb_mg_graph");
  if (r) return r;
  c->state = 2;
  if (out) *out = c->cnt;
  return SPB_OK;
}

// =========================================================================================== compaction, in phases
// One GPU runs the phases back to back over everything (spb_compact).  The sliced multi-GPU tail runs the per-vertex and
// per-position phases on its own tiles / positions / NBPs and the contracted-graph phases (small) replicated, with the
// host side gathering the short lists in between (spb_mg_slice_*).

// K5/K6: init detection and pieces over the own tiles -> sorted lists of inits, piece starts, piece ends (local)
static int cp_classify(spb_ctx* c, u32** init_l, u32** plo_l, u32** phi_l, u32* nI_out, u32* nP_out) {
  const u32 nV = c->nV;
  GView g{c->vflags, c->Jk, c->Js, c->nJ, nV};
  // groups of 8 * SPB_CW = 32 vertices; the slice's tile range is aligned to them (spb_mg_slice_ids)
  const u32 nGall = (u32)(((u64)nV + 8 * SPB_CW - 1) / (8 * SPB_CW));
  const u32 nTall = (u32)(((u64)nV + 7) / 8);
  const auto grp = [&](u32 t) { return t >= nTall ? nGall : t / SPB_CW; };      // tile cuts are multiples of SPB_CW, or the end
  const u32 t_lo = c->sl.on ? grp(c->sl.t_lo) : 0u, t_hi = c->sl.on ? grp(c->sl.t_hi) : nGall;
  const u32 f_lo = c->sl.on ? (t_lo ? t_lo - 1 : 0u) : 0u, f_hi = c->sl.on ? std::min(t_hi + 1, nGall) : nGall;   // piece flags: one group of halo
  const u32 nT = t_hi - t_lo;
  u32 *cnt3, *off3; ALLOC(cnt3, u32, 3ull * nT + 1); ALLOC(off3, u32, 3ull * nT + 1);
  SPB_LAUNCH(k_classify_pieces, nblk(f_hi - f_lo, TPB), TPB, c->stream, g, c->vflags, cnt3, nT, f_lo, f_hi, t_lo, t_hi, c->errflag);
  u64 tot3 = 0;
  PRIM(prim_scan_u32(c, cnt3, off3, 3ull * nT, &tot3));
  u32 marks[2] = {0, 0};
  if (nT) {
    pin_read(c, c->pinned, off3 + nT, 4); pin_read(c, (u32*)c->pinned + 1, off3 + 2ull * nT, 4); dev_sync(c);
    memcpy(marks, c->pinned, 8);
  }
  const u32 nI = marks[0], nP = marks[1] - marks[0];
  if (!c->sl.on && (u64)nI + 2ull * nP != tot3) return fail(c, SPB_ECUDA, "internal: piece starts (%u) != piece ends (%llu)", nP, (unsigned long long)(tot3 - marks[1]));
  const u32 nE = (u32)(tot3 - marks[1]);               // piece ends in the own groups (a piece may cross a cut: starts != ends locally)
  u32 *il, *pl, *ph; ALLOC(il, u32, nI + 1); ALLOC(pl, u32, nP + 1); ALLOC(ph, u32, nE + 1);
  SPB_LAUNCH(k_compact3, nblk(nT, TPB), TPB, c->stream, c->vflags, nV, off3, nT, nI, nP, il, pl, ph, t_lo);
  dfree(c, cnt3); dfree(c, off3);
  *init_l = il; *plo_l = pl; *phi_l = ph; nI_out[0] = nI; nP_out[0] = nP; nP_out[1] = nE;
  return 0;
}

static void cp_free_contracted(spb_ctx* c) {
  dfree(c, c->pl_nbr); dfree(c, c->pr_nbr); dfree(c, c->pl_side); dfree(c, c->pr_side); dfree(c, c->nxt);
  for (int i = 0; i < 2; ++i) { dfree(c, c->rk_ptr[i]); dfree(c, c->rk_acc[i]); dfree(c, c->rk_root[i]); dfree(c, c->rk_mnv[i]); }
}

// K7-K9 on the complete lists (c->init_list, c->plo, c->phi; c->nI, c->nP): piece ends, darts, list ranking.  *again = 1:
// pure cycles were found and cut (pass 0 only) -- classification has to run again.
static int cp_contract(spb_ctx* c, int pass, int* again) {
  const u32 nV = c->nV;
  GView g{c->vflags, c->Jk, c->Js, c->nJ, nV};
  *again = 0;
  const u32 nP = c->nP, nD = 2 * nP;
  c->nD = nD;
  ALLOC(c->pl_nbr, u32, nP + 1); ALLOC(c->pr_nbr, u32, nP + 1); ALLOC(c->pl_side, u8, nP + 1); ALLOC(c->pr_side, u8, nP + 1);
  SPB_LAUNCH(k_piece_ends, nblk(nP, TPB), TPB, c->stream, g, c->plo, c->phi, nP, c->pl_nbr, c->pl_side, c->pr_nbr, c->pr_side, c->errflag);
  ALLOC(c->nxt, u32, nD + 1);
  SPB_LAUNCH(k_dart_next, nblk(nD, TPB), TPB, c->stream, c->vflags, c->plo, nP, c->pl_nbr, c->pl_side, c->pr_nbr, c->pr_side, c->nxt);
  for (int i = 0; i < 2; ++i) { ALLOC(c->rk_ptr[i], u32, nD + 1); ALLOC(c->rk_acc[i], u32, nD + 1); ALLOC(c->rk_root[i], u32, nD + 1); ALLOC(c->rk_mnv[i], u32, nD + 1); }
  int cur = 0;
  SPB_LAUNCH(k_rank_init, nblk(nD, TPB), TPB, c->stream, c->nxt, c->plo, c->phi, nD, c->rk_ptr[0], c->rk_acc[0], c->rk_root[0], c->rk_mnv[0]);
  u32 rounds = 1; while ((1ull << rounds) < (u64)nD + 1) ++rounds;
  ++rounds;
  for (u32 r = 0; r < rounds && nD; ++r) {
    SPB_LAUNCH(k_rank_jump, nblk(nD, TPB), TPB, c->stream, nD, c->rk_ptr[cur], c->rk_acc[cur], c->rk_root[cur], c->rk_mnv[cur], c->rk_ptr[cur ^ 1],
               c->rk_acc[cur ^ 1], c->rk_root[cur ^ 1], c->rk_mnv[cur ^ 1]);
    cur ^= 1;
  }
  c->rk_cur = cur;
  u32* ucount; ALLOC(ucount, u32, 4);
  dev_memset(c, ucount, 0, 16);
  SPB_LAUNCH(k_count_unresolved, nblk(nD, TPB), TPB, c->stream, c->rk_ptr[cur], nD, ucount);
  const u32 unresolved = dread(c, ucount);
  if (unresolved == 0) { dfree(c, ucount); return 0; }
  if (pass == 1) { dfree(c, ucount); return fail(c, SPB_ECUDA, "internal: %u darts unresolved after cycle cutting", unresolved); }
  // ---- K9: pure-cycle components: cut one edge per ring, then redo the classification
  PRIM(ensure_edges(c));                               // the edge rows still hold the edges that are cut for the walks
  SPB_LAUNCH(k_cycle_cut, nblk(nD, TPB), TPB, c->stream, g, c->vflags, c->Js, c->rk_ptr[cur], c->rk_mnv[cur], c->plo, nD, ucount + 1);
  c->ncyc = dread(c, ucount + 1);
  dfree(c, ucount);
  dfree(c, c->init_list); dfree(c, c->plo); dfree(c, c->phi);
  cp_free_contracted(c);
  *again = 1;
  return 0;
}

// K10: raw NBP table, lengths and output offsets (replicated; small)
static int cp_nbp_table(spb_ctx* c, u64* totv_out, u64* tots_out) {
  const u32 k = c->k; const u32 nV = c->nV;
  GView g{c->vflags, c->Jk, c->Js, c->nJ, nV};
  const u32 nP = c->nP, nD = c->nD, nI = c->nI;
  const int cur = c->rk_cur;
  u32* ideg; ALLOC(ideg, u32, nI + 1); ALLOC(c->ibase, u32, nI + 1);
  SPB_LAUNCH(k_init_degrees, nblk(nI, TPB), TPB, c->stream, g, c->init_list, nI, ideg);
  u64 nN64 = 0;
  PRIM(prim_scan_u32(c, ideg, c->ibase, nI, &nN64));
  dfree(c, ideg);
  const u32 nN = (u32)nN64; c->nN = nN;
  NbpTab& t = c->nt;
  ALLOC(t.a, u32, nN + 1); ALLOC(t.b, u32, nN + 1); ALLOC(t.L, u32, nN + 1); ALLOC(t.head, u32, nN + 1); ALLOC(t.last, u32, nN + 1);
  ALLOC(t.aside, u8, nN + 1); ALLOC(t.bside, u8, nN + 1); ALLOC(t.cyc, u8, nN + 1);
  u32* head_nbp; ALLOC(head_nbp, u32, nD + 1); dev_memset(c, head_nbp, 0xFF, (nD + 1) * 4ull);
  ALLOC(c->dart_nbp, u32, nD + 1);
  SPB_LAUNCH(k_nbp_heads, nblk(nI, TPB), TPB, c->stream, g, c->vflags, c->init_list, c->ibase, nI, c->plo, nP, t, head_nbp);
  SPB_LAUNCH(k_dart_finish, nblk(nD, TPB), TPB, c->stream, c->nxt, c->rk_root[cur], c->rk_acc[cur], c->plo, c->phi, c->pl_nbr, c->pl_side, c->pr_nbr, c->pr_side,
             head_nbp, nD, c->dart_nbp, t);
  dfree(c, head_nbp);
  u64 *vlen, *slen; ALLOC(vlen, u64, nN + 1); ALLOC(slen, u64, nN + 1);
  ALLOC(c->voff, u64, nN + 1); ALLOC(c->soff, u64, nN + 1);
  SPB_LAUNCH(k_nbp_lengths, nblk(nN, TPB), TPB, c->stream, t, nN, k, vlen, slen);
  u64 totv = 0, tots = 0;
  PRIM(prim_scan_u64(c, vlen, c->voff, nN, &totv));
  PRIM(prim_scan_u64(c, slen, c->soff, nN, &tots));
  dev_copy(c, c->voff + nN, &totv, 8); dev_copy(c, c->soff + nN, &tots, 8); dev_sync(c);
  dfree(c, vlen); dfree(c, slen);
  *totv_out = totv; *tots_out = tots;
  return 0;
}

// K11/K12: vertex lists + sequences of the own NBPs [n_lo, n_hi), vertex -> NBP map of the own vertices
static int cp_emit(spb_ctx* c, u64 totv, u64 tots) {
  const u64 T = c->T; const u32 k = c->k; const u32 nV = c->nV;
  const u32 nP = c->nP, nD = c->nD, nN = c->nN;
  const int cur = c->rk_cur;
  NbpTab& t = c->nt;
  spb_ctx::Slice& sl = c->sl;
  if (sl.on) {                                         // NBP ranges with (nearly) equal numbers of emitted vertices
    u32* cuts; ALLOC(cuts, u32, sl.G + 2);
    if (sl.G + 1 > 64) return fail(c, SPB_EINVAL, "more than 63 ranks");
    SPB_LAUNCH(k_nbp_cuts, 1, 64, c->stream, c->voff, nN, sl.G, cuts);
    pin_read(c, c->pinned, cuts + sl.rank, 8); dev_sync(c);
    u32 nc[2]; memcpy(nc, c->pinned, 8);
    sl.n_lo = nc[0]; sl.n_hi = nc[1];
    dfree(c, cuts);
    pin_read(c, c->pinned, c->voff + sl.n_lo, 8); pin_read(c, c->pinned + 1, c->voff + sl.n_hi, 8);
    pin_read(c, c->pinned + 2, c->soff + sl.n_lo, 8); pin_read(c, c->pinned + 3, c->soff + sl.n_hi, 8); dev_sync(c);
    sl.vert_lo = c->pinned[0]; sl.vert_n = c->pinned[1] - c->pinned[0]; sl.seq_lo = c->pinned[2]; sl.seq_n = c->pinned[3] - c->pinned[2];
  } else { sl.n_lo = 0; sl.n_hi = nN; sl.vert_lo = 0; sl.vert_n = totv; sl.seq_lo = 0; sl.seq_n = tots; }
  const u32 n_lo = sl.n_lo, n_hi = sl.n_hi;
  ALLOC(c->verts, u32, sl.vert_n + 1); ALLOC(c->seqs_base, u8, sl.seq_n + 16);
  c->seqs = c->seqs_base + (sl.seq_lo & 3);            // word-aligned stores of the sequence kernels keep the phase of the global offsets
  u32* verts = c->verts - sl.vert_lo; u8* seqs = c->seqs - sl.seq_lo;      // addressed with the global offsets
  u32* vnbp_w = nullptr;
  // The vertex -> NBP map costs 4 bytes per vertex to write.  On mostly-unique inputs the origin pass asks for a handful of
  // vertices only (windows of consecutive first appearances are decided by their flag bytes), which the piece list answers
  // by binary search; the array is then built lazily for the getters that want it (ensure_vnbp).
  const bool lazy_vnbp = !sl.on && (c->n_inst - nV) * 16 < (u64)nV && !getenv("SPB_EAGER_VNBP");
  if (!sl.on && !lazy_vnbp) { ALLOC(c->vnbp, u32, (u64)nV + 1); dev_memset(c, c->vnbp, 0xFF, ((u64)nV + 1) * 4); vnbp_w = c->vnbp; }
  const VPos vp = vpos_view(c);
  SPB_LAUNCH(k_emit_ends, nblk((u64)(n_hi - n_lo) * 32, TPB), TPB, c->stream, t, n_lo, n_hi, k, T, c->F, c->R, vp, c->voff, c->soff, verts, seqs);
  u32 *nch, *choff; ALLOC(nch, u32, nD + 1); ALLOC(choff, u32, nD + 1);
  SPB_LAUNCH(k_dart_chunks, nblk(nD, TPB), TPB, c->stream, c->plo, c->phi, nD, c->dart_nbp, n_lo, n_hi, nch);
  u64 nchunks = 0;
  PRIM(prim_scan_u32(c, nch, choff, nD, &nchunks));
  if (nchunks >= (1ull << 31)) return fail(c, SPB_EINVAL, "input too large for 32-bit dart offsets");
  ChunkParams* cpar; ALLOC(cpar, ChunkParams, nchunks + 1);
  SPB_LAUNCH(k_chunk_table, nblk(nchunks, TPB), TPB, c->stream, (u32)nchunks, choff, nD, c->plo, c->phi, c->dart_nbp, c->rk_acc[cur], k, T, vp,
             c->voff, c->soff, cpar);
  { const int tq_ = t_begin(c); SPB_LAUNCH(k_emit_darts, nchunks, 256, c->stream, cpar, c->F, c->R, verts, seqs, vnbp_w); t_end(c, tq_, "k_emit_darts_ms"); }
  dfree(c, cpar);
  SPB_LAUNCH(k_emit_short, std::min<u64>(nblk((u64)nD * 32, 256), 148 * 32), 256, c->stream, nD, c->plo, c->phi, c->dart_nbp, c->rk_acc[cur], k, T,
             c->F, c->R, vp, c->voff, c->soff, verts, seqs, vnbp_w, n_lo, n_hi);
  dfree(c, nch); dfree(c, choff);
  if (sl.on) {                                         // vertex -> NBP map of the own tiles (+ one group of 32: see k_classify_pieces)
    ALLOC(c->vnbp, u32, (u64)nV + 1);                  // addressed by the global vertex id; only [v_lo, v_hi) of it is touched
    const VNbp vn = vnbp_view(c);
    SPB_LAUNCH(k_fill_vnbp, nblk(nblk(vn.v_hi - vn.v_lo, 1024) * 32, TPB), TPB, c->stream, c->plo, c->phi, nP, c->dart_nbp, vn.v_lo, vn.v_hi, c->vnbp);
  }
  cp_free_contracted(c);
  return 0;
}

// K13, first half: (NBP, sequence) / (init, sequence) pairs of the own positions -> c->op, c->mp (unsorted), c->has01
static int cp_origin_pairs(spb_ctx* c) {
  const u64 T = c->T;
  const u32 nN = c->nN, nV = c->nV;
  u64* counts; ALLOC(counts, u64, 4);
  ALLOC(c->has01, u8, 2ull * c->n_seq + 8);
  u32 sb = 1; while ((1ull << sb) <= c->n_seq) ++sb;          // keys = (NBP or vertex id) << sb | sequence index
  (void)nN; (void)nV;
  const u64 w_lo = c->sl.on ? c->sl.pos_lo / 32 : 0, w_hi = c->sl.on ? (c->sl.pos_hi > c->sl.pos_lo ? (c->sl.pos_hi + 31) / 32 : w_lo) : (T + 31) / 32;
  const u64 npos = 32 * (w_hi - w_lo);
  u64 ocap = std::max<u64>(4096, npos / 4), mcap = std::max<u64>(4096, npos / 4);
  u64 *op = 0, *mp = 0; u64 no = 0, nm = 0;
  const VNbp vn = vnbp_view(c);
  for (int attempt = 0; attempt < 2; ++attempt) {
    ALLOC(op, u64, ocap + c->n_seq + 1); ALLOC(mp, u64, mcap + c->n_seq + 1);      // room for the quirk entries (one per sequence at most)
    dev_memset(c, counts, 0, 32); dev_memset(c, c->has01, 0, 2ull * c->n_seq + 8);
    { const int tq_ = t_begin(c); SPB_LAUNCH(k_origin_pairs, gs_grid(npos), TPB, c->stream, c->sp, T, w_lo, w_hi, c->fbits, c->blkoff, c->wpre, vn, c->vflags, c->seq_off, c->n_seq, op, counts, ocap, mp,
               counts + 1, mcap, c->has01, sb); t_end(c, tq_, "k_origin_pairs_ms"); }
    pin_read(c, c->pinned, counts, 16); dev_sync(c);
    no = c->pinned[0]; nm = c->pinned[1];
    if (no <= ocap && nm <= mcap) break;
    dfree(c, op); dfree(c, mp); ocap = std::max(ocap, no); mcap = std::max(mcap, nm);
  }
  dfree(c, counts);
  c->op = op; c->mp = mp; c->n_op = no; c->n_mp = nm; c->op_cap = ocap + c->n_seq; c->mp_cap = mcap + c->n_seq;
  return 0;
}

// K13, second half (replicated): all pairs -> leading-zero quirk, sort + unique, per-NBP origin lists.  op / mp have room
// for n_seq more entries; takes ownership of op, mp, has01.
static int cp_origins_finish(spb_ctx* c, u64* op, u64 no, u64 ocap, u64* mp, u64 nm, u64 mcap, u8* has01) {
  const u32 nN = c->nN, nV = c->nV;
  NbpTab& t = c->nt;
  u32 sb = 1; while ((1ull << sb) <= c->n_seq) ++sb;
  u32 ob = sb + 1; while (ob < 64 && (1ull << (ob - sb)) <= (u64)nN) ++ob;
  u32 mb = sb + 1; while (mb < 64 && (1ull << (mb - sb)) <= (u64)nV) ++mb;
  {
    u64* counts; ALLOC(counts, u64, 4);
    memcpy(c->pinned, &no, 8); memcpy(c->pinned + 1, &nm, 8);
    dev_copy(c, counts, c->pinned, 16);
    SPB_LAUNCH(k_origin_quirk, nblk(c->n_seq, TPB), TPB, c->stream, has01, c->n_seq, vnbp_view(c), c->vflags, op, counts, ocap, mp, counts + 1, mcap, sb);
    pin_read(c, c->pinned + 2, counts, 16); dev_sync(c);
    no = c->pinned[2]; nm = c->pinned[3];
    dfree(c, counts);
    if (no > ocap || nm > mcap) return fail(c, SPB_ECUDA, "internal: origin pair lists too small for the quirk entries");
  }
  dfree(c, has01);
  u64* tmp; u8* nopay = nullptr;
  ALLOC(tmp, u64, std::max(no, nm) + 1);
  PRIM(prim_sort_u64(c, op, tmp, no, ob));
  PRIM(prim_sort_u64(c, mp, tmp, nm, mb));
  dfree(c, tmp);
  PRIM(unique_sorted(c, op, nopay, no, &c->nO));
  PRIM(unique_sorted(c, mp, nopay, nm, &c->nM));
  if (!op) ALLOC(op, u64, 1);
  if (!mp) ALLOC(mp, u64, 1);
  u64* ocnt; ALLOC(ocnt, u64, nN + 1); ALLOC(c->ooff, u64, nN + 1);
  SPB_LAUNCH(k_nbp_origins, nblk(nN, TPB), TPB, c->stream, t, nN, c->dart_nbp, op, c->nO, mp, c->nM, sb, (const u64*)nullptr, ocnt, (u32*)nullptr);
  u64 toto = 0;
  PRIM(prim_scan_u64(c, ocnt, c->ooff, nN, &toto));
  dev_copy(c, c->ooff + nN, &toto, 8); dev_sync(c);
  ALLOC(c->origins, u32, toto + 1);
  SPB_LAUNCH(k_nbp_origins, nblk(nN, TPB), TPB, c->stream, t, nN, c->dart_nbp, op, c->nO, mp, c->nM, sb, c->ooff, ocnt, c->origins);
  dfree(c, ocnt); dfree(c, op); dfree(c, mp);
  c->cnt.n_origin_pairs = toto;
  return 0;
}

// K14: connected components (replicated)
static int cp_components(spb_ctx* c) {
  const u32 nI = c->nI, nN = c->nN, nP = c->nP;
  NbpTab& t = c->nt;
  ALLOC(c->parent, u32, nI + 1); ALLOC(c->label_of_root, u32, nI + 1);
  u32* cmin; ALLOC(cmin, u32, nI + 1); dev_memset(c, cmin, 0xFF, (nI + 1) * 4ull);
  u64 *rk, *rk2; ALLOC(rk, u64, nI + 1); ALLOC(rk2, u64, nI + 1);
  u32* nroots; ALLOC(nroots, u32, 4); dev_memset(c, nroots, 0, 16);
  SPB_LAUNCH(k_uf_init, nblk(nI, TPB), TPB, c->stream, c->parent, nI);
  SPB_LAUNCH(k_uf_union, nblk(nN, TPB), TPB, c->stream, t, nN, c->init_list, nI, c->parent);
  SPB_LAUNCH(k_comp_min_init, nblk(nI, TPB), TPB, c->stream, c->init_list, nI, c->parent, cmin);
  SPB_LAUNCH(k_comp_min_piece, nblk(nP, TPB), TPB, c->stream, c->plo, nP, c->dart_nbp, t, c->init_list, nI, c->parent, cmin);
  SPB_LAUNCH(k_comp_roots, nblk(nI, TPB), TPB, c->stream, c->parent, cmin, nI, rk, nroots);
  u32 nr = dread(c, nroots);
  PRIM(prim_sort_u64(c, rk, rk2, nr));
  SPB_LAUNCH(k_comp_labels, nblk(nr, TPB), TPB, c->stream, rk, nr, c->label_of_root);
  ALLOC(c->nbp_comp, int32_t, nN + 1);
  SPB_LAUNCH(k_nbp_comp, nblk(nN, TPB), TPB, c->stream, t, nN, c->init_list, nI, c->parent, c->label_of_root, c->nbp_comp);
  dfree(c, cmin); dfree(c, rk); dfree(c, rk2); dfree(c, nroots);
  c->cnt.n_comp = nr;
  return 0;
}

int spb_compact(spb_ctx* c, spb_counts* out) {
  if (!c) return SPB_EINVAL;
  if (c->state != 2) return fail(c, SPB_ESTATE, "spb_compact needs a freshly built DBG (state %d)", c->state);
  if (c->sl.on) return fail(c, SPB_ESTATE, "spb_compact: the DBG of this context is sliced over several ranks (spb_mg_slice_* / spb_mg_complete)");
  SPB_ON_DEVICE(c);
  int tm = t_begin(c);
  c->ncyc = 0;
  for (int pass = 0; pass < 2; ++pass) {
    u32 nI[1], nP[2];
    PRIM(cp_classify(c, &c->init_list, &c->plo, &c->phi, nI, nP));
    c->nI = nI[0]; c->nP = nP[0];
    int again = 0;
    PRIM(cp_contract(c, pass, &again));
    if (!again) break;
  }
  c->cnt.n_cycle_comp = c->ncyc;
  t_end(c, tm, "classify_rank_ms");
  tm = t_begin(c);
  u64 totv = 0, tots = 0;
  PRIM(cp_nbp_table(c, &totv, &tots));
  PRIM(cp_emit(c, totv, tots));
  t_end(c, tm, "emit_ms");
#ifndef SPB_EMU
  if (c->out_seq_host && tots <= c->out_seq_cap) {     // sequences to the caller's host buffer while the rest is computed
    if (!c->copy_stream) {
      CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&c->ev_emit, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->ev_d2h, cudaEventDisableTiming));
    }
    CK(cudaEventRecord(c->ev_emit, c->stream));
    CK(cudaStreamWaitEvent(c->copy_stream, c->ev_emit, 0));
    CK(cudaMemcpyAsync(c->out_seq_host, c->seqs, tots, cudaMemcpyDeviceToHost, c->copy_stream));
    CK(cudaEventRecord(c->ev_d2h, c->copy_stream));
    c->d2h_pending = true;
  }
#endif
  tm = t_begin(c);
  PRIM(ensure_edges(c));
  t_end(c, tm, "edge_rows_ms");
  tm = t_begin(c);
  PRIM(cp_origin_pairs(c));
  {
    u64 *op = c->op, *mp = c->mp; u8* h = c->has01;
    c->op = c->mp = nullptr; c->has01 = nullptr;
    PRIM(cp_origins_finish(c, op, c->n_op, c->op_cap, mp, c->n_mp, c->mp_cap, h));
  }
  t_end(c, tm, "origins_ms");
  tm = t_begin(c);
  PRIM(cp_components(c));
  t_end(c, tm, "components_ms");
  int r = check_err(c, "spb_compact");
  if (r) return r;
  c->cnt.n_inits = c->nI; c->cnt.n_pieces = c->nP; c->cnt.n_nbp = c->nN; c->cnt.nbp_vertices = totv; c->cnt.nbp_bases = tots;
  c->state = 3;
  if (out) *out = c->cnt;
  return SPB_OK;
}

// =========================================================================================== sliced multi-GPU tail
// After the dedup ("part" protocol with all links on every rank: fbits / obits / the links in mg_seen are complete
// everywhere) nothing of size N_inst is exchanged any more.  Rank r computes the vertex ids, junction entries and origin
// pairs of ITS positions, the flags / piece lists / edge rows / vertex -> NBP map of ITS vertices (first appearances of
// its positions, cut at multiples of 8) and the vertex lists and sequences of ITS NBPs (id ranges with equal shares of
// the output); the short lists (junction entries, inits, piece bounds, origin pairs) are all-gathered by the host side
// and the contracted graph (pieces, darts, list ranking, NBP table, origin sets, components) is computed by every rank.
// The few entries a sliced pass needs from another rank's range are recomputed from the replicated bitmaps (VPos, VidAt,
// VNbp in spb_kernels.cuh).  spb_mg_complete recomputes the sliced passes over the whole input on this rank (no
// communication) so that every single-GPU getter works; spb_mg_shard hands out the local parts.
static void free_graph_results(spb_ctx* c) {
  dfree(c, c->Jk); dfree(c, c->Js); dfree(c, c->edges); dfree(c, c->edge_off); c->edges_pending = false; dfree(c, c->seqlim);
  dfree(c, c->init_list); dfree(c, c->ibase); dfree(c, c->plo); dfree(c, c->phi); dfree(c, c->dart_nbp); dfree(c, c->vnbp);
  dfree(c, c->parent); dfree(c, c->label_of_root); dfree(c, c->nbp_comp);
  dfree(c, c->nt.a); dfree(c, c->nt.b); dfree(c, c->nt.L); dfree(c, c->nt.head); dfree(c, c->nt.last); dfree(c, c->nt.aside); dfree(c, c->nt.bside); dfree(c, c->nt.cyc);
  dfree(c, c->voff); dfree(c, c->soff); dfree(c, c->ooff); dfree(c, c->verts); dfree(c, c->seqs_base); c->seqs = nullptr; dfree(c, c->origins);
  dfree(c, c->op); dfree(c, c->mp); dfree(c, c->has01);
  cp_free_contracted(c);
}

int spb_mg_slice_ids(spb_ctx* c, uint32_t rank, uint32_t n_ranks) {
  if (!c || !n_ranks || rank >= n_ranks) return SPB_EINVAL;
  SPB_ON_DEVICE(c);
  if (c->state != 1 || !c->sp || !c->mg_seen) return fail(c, SPB_ESTATE, "spb_mg_slice_ids needs spb_mg_apply_links over all positions first");
  stats_reset(c); stat_set(c, "multi_gpu", 1);
  uint64_t S = 0; spb_mg_slice(c, n_ranks, &S);
  spb_ctx::Slice& sl = c->sl;
  sl = spb_ctx::Slice();
  sl.rank = rank; sl.G = n_ranks;
  sl.pos_lo = std::min<u64>((u64)rank * S, c->T); sl.pos_hi = std::min<u64>((u64)(rank + 1) * S, c->T);
  sl.ids_lo = sl.pos_lo ? sl.pos_lo - 1 : 0; sl.ids_hi = std::min<u64>(sl.pos_hi + 1, c->T);
  PRIM(dd_ids(c, sl.ids_lo, sl.ids_hi, nullptr, 1, c->mg_seen, 0, sl.pos_lo, sl.pos_hi));
  // vertices of the slice: ranks of its first and of the next slice's first position (S is a multiple of 256)
  u32 vf[2];
  for (int i = 0; i < 2; ++i) {
    const u64 p = i ? sl.pos_hi : sl.pos_lo;
    if (p >= c->T) { vf[i] = c->nV; continue; }
    pin_read(c, (u32*)c->pinned + i, c->blkoff + (p >> 8), 4);
  }
  dev_sync(c);
  for (int i = 0; i < 2; ++i) { const u64 p = i ? sl.pos_hi : sl.pos_lo; if (p < c->T) vf[i] = ((u32*)c->pinned)[i]; }
  sl.v_lo = vf[0]; sl.v_hi = vf[1];
  // own tiles of 8 vertices: cut below the first vertex of every slice (an empty slice at the end of the input owns nothing)
  const u32 nTall = (c->nV + 7) / 8;
  sl.t_lo = sl.pos_lo >= c->T ? nTall : (sl.v_lo >> 5) << 2; sl.t_hi = sl.pos_hi >= c->T ? nTall : (sl.v_hi >> 5) << 2;     // cut below multiples of 32 ids
  sl.on = true;
  return SPB_OK;                                       // (error flags are read once per step, by spb_mg_slice_graph / spb_mg_slice_finish: no host sync here)
}

int spb_mg_slice_junctions(spb_ctx* c, uint64_t* n, void** jk_dev, void** js_dev) {
  if (!c || !n || !jk_dev || !js_dev) return SPB_EINVAL;
  SPB_ON_DEVICE(c);
  if (c->state != 1 || !c->sl.on) return fail(c, SPB_ESTATE, "spb_mg_slice_junctions out of order");
  u64* Jk; u8* Js; u64 nj;
  dfree(c, c->Jk); dfree(c, c->Js);
  const u64 w_lo = c->sl.pos_lo / 32, w_hi = c->sl.pos_hi > c->sl.pos_lo ? (c->sl.pos_hi + 31) / 32 : w_lo;   // slices start at multiples of 1024
  PRIM(graph_junctions(c, w_lo, w_hi, &Jk, &Js, &nj));
  c->Jk = Jk; c->Js = Js;                              // kept until spb_mg_slice_graph replaces them by the complete list
  *n = nj; *jk_dev = Jk; *js_dev = Js;
  return SPB_OK;                                       // (error flags are read once per step, by spb_mg_slice_graph / spb_mg_slice_finish: no host sync here)
}

int spb_mg_slice_graph(spb_ctx* c, const uint64_t* jk_all, const uint8_t* js_all, uint64_t n_all, uint64_t* n_edges_local) {
  if (!c) return SPB_EINVAL;
  SPB_ON_DEVICE(c);
  if (c->state != 1 || !c->sl.on) return fail(c, SPB_ESTATE, "spb_mg_slice_graph out of order");
  u64* Jk; u8* Js;
  ALLOC(Jk, u64, n_all + 1); ALLOC(Js, u8, n_all + 1);
  dev_copy(c, Jk, jk_all, 8 * n_all); dev_copy(c, Js, js_all, n_all);
  dfree(c, c->Jk); dfree(c, c->Js);
  PRIM(graph_finish(c, Jk, Js, n_all));
  if (n_edges_local) *n_edges_local = c->sl.edge_n;
  int r = check_err(c, "spb_mg_slice_graph");
  if (r) return r;
  c->state = 2;
  c->ncyc = 0;
  return SPB_OK;
}

int spb_mg_slice_classify(spb_ctx* c, uint32_t* n_init, uint32_t* n_start, uint32_t* n_end, void** init_dev, void** plo_dev, void** phi_dev) {
  if (!c || !n_init || !n_start || !n_end) return SPB_EINVAL;
  SPB_ON_DEVICE(c);
  if (c->state != 2 || !c->sl.on) return fail(c, SPB_ESTATE, "spb_mg_slice_classify out of order");
  dfree(c, c->init_list); dfree(c, c->plo); dfree(c, c->phi);
  u32 nI[1], nP[2];
  PRIM(cp_classify(c, &c->init_list, &c->plo, &c->phi, nI, nP));      // local lists, replaced by the complete ones in spb_mg_slice_contract
  *n_init = nI[0]; *n_start = nP[0]; *n_end = nP[1];
  *init_dev = c->init_list; *plo_dev = c->plo; *phi_dev = c->phi;
  return SPB_OK;                                       // (error flags are read once per step, by spb_mg_slice_graph / spb_mg_slice_finish: no host sync here)
}

int spb_mg_slice_contract(spb_ctx* c, const uint32_t* init_all, uint32_t n_init, const uint32_t* plo_all, const uint32_t* phi_all, uint32_t n_piece,
                          uint64_t n_edges_total, int pass, int* again) {
  if (!c || !again) return SPB_EINVAL;
  SPB_ON_DEVICE(c);
  if (c->state != 2 || !c->sl.on) return fail(c, SPB_ESTATE, "spb_mg_slice_contract out of order");
  u32 *il, *pl, *ph;
  ALLOC(il, u32, n_init + 1); ALLOC(pl, u32, n_piece + 1); ALLOC(ph, u32, n_piece + 1);
  dev_copy(c, il, init_all, 4ull * n_init); dev_copy(c, pl, plo_all, 4ull * n_piece); dev_copy(c, ph, phi_all, 4ull * n_piece);
  dfree(c, c->init_list); dfree(c, c->plo); dfree(c, c->phi);
  c->init_list = il; c->plo = pl; c->phi = ph; c->nI = n_init; c->nP = n_piece;
  c->cnt.n_edges = n_edges_total;
  SPB_LAUNCH(k_mark_inits, nblk(n_init, TPB), TPB, c->stream, c->init_list, n_init, c->vflags);       // the init bit is read at arbitrary vertices
  PRIM(cp_contract(c, pass, again));
  return SPB_OK;                                       // (error flags are read once per step, by spb_mg_slice_graph / spb_mg_slice_finish: no host sync here)
}

int spb_mg_slice_emit(spb_ctx* c, uint64_t* n_op, void** op_dev, uint64_t* n_mp, void** mp_dev, void** has01_dev) {
  if (!c || !n_op || !n_mp) return SPB_EINVAL;
  SPB_ON_DEVICE(c);
  if (c->state != 2 || !c->sl.on || !c->nxt) return fail(c, SPB_ESTATE, "spb_mg_slice_emit out of order");
  c->cnt.n_cycle_comp = c->ncyc;
  u64 totv = 0, tots = 0;
  PRIM(cp_nbp_table(c, &totv, &tots));
  PRIM(cp_emit(c, totv, tots));
  c->cnt.n_inits = c->nI; c->cnt.n_pieces = c->nP; c->cnt.n_nbp = c->nN; c->cnt.nbp_vertices = totv; c->cnt.nbp_bases = tots;
  PRIM(cp_origin_pairs(c));
  PRIM(ensure_edges(c));
  *n_op = c->n_op; *n_mp = c->n_mp;
  *op_dev = c->op; *mp_dev = c->mp; *has01_dev = c->has01;
  return SPB_OK;                                       // (error flags are read once per step, by spb_mg_slice_graph / spb_mg_slice_finish: no host sync here)
}

int spb_mg_slice_finish(spb_ctx* c, const uint64_t* op_all, uint64_t n_op, const uint64_t* mp_all, uint64_t n_mp, const uint8_t* has01_any, spb_counts* out) {
  if (!c) return SPB_EINVAL;
  SPB_ON_DEVICE(c);
  if (c->state != 2 || !c->sl.on || !c->has01) return fail(c, SPB_ESTATE, "spb_mg_slice_finish out of order");
  u64 *op, *mp; u8* h;
  ALLOC(op, u64, n_op + c->n_seq + 1); ALLOC(mp, u64, n_mp + c->n_seq + 1); ALLOC(h, u8, 2ull * c->n_seq + 8);
  dev_copy(c, op, op_all, 8 * n_op); dev_copy(c, mp, mp_all, 8 * n_mp); dev_copy(c, h, has01_any, 2ull * c->n_seq);
  dfree(c, c->op); dfree(c, c->mp); dfree(c, c->has01);
  PRIM(cp_origins_finish(c, op, n_op, n_op + c->n_seq, mp, n_mp, n_mp + c->n_seq, h));
  PRIM(cp_components(c));
  int r = check_err(c, "spb_mg_slice_finish");
  if (r) return r;
  c->state = 3;
  if (out) *out = c->cnt;
  return SPB_OK;
}

// local parts of the sliced results: which = 0 per-position vertex ids (u32), 1 vertex2coords rows (2 x u32 per vertex),
// 2 edge rows (2 x u32), 3 NBP vertex lists (u32), 4 NBP sequences (u8).  first / count in elements of the complete array
// (rows for 1 and 2; the first edge row of a rank is the sum of the counts of the ranks below it).
int spb_mg_shard(spb_ctx* c, int which, void** dev, uint64_t* first, uint64_t* count) {
  if (!c || !dev || !first || !count) return SPB_EINVAL;
  SPB_ON_DEVICE(c);
  if (!c->sl.on || c->state < 2) return fail(c, SPB_ESTATE, "spb_mg_shard: no sliced results in this context");
  const spb_ctx::Slice& sl = c->sl;
  switch (which) {
    case 0: *dev = c->sp + sl.pos_lo; *first = sl.pos_lo; *count = sl.pos_hi - sl.pos_lo; return SPB_OK;
    case 1: {
      static_assert(sizeof(u32) == 4, "");
      u32* tmp; ALLOC(tmp, u32, 2ull * (sl.v_hi - sl.v_lo) + 2);
      SPB_LAUNCH(k_coords, nblk(sl.v_hi - sl.v_lo, TPB), TPB, c->stream, c->v2pos + sl.v_lo, sl.v_hi - sl.v_lo, c->seq_off, c->n_seq, tmp);
      dfree(c, c->shard_tmp); c->shard_tmp = tmp;
      *dev = tmp; *first = sl.v_lo; *count = sl.v_hi - sl.v_lo; return SPB_OK;
    }
    case 2: PRIM(ensure_edges(c)); *dev = c->edges; *first = 0; *count = sl.edge_n; return SPB_OK;
    case 3: if (c->state < 3) break; *dev = c->verts; *first = sl.vert_lo; *count = sl.vert_n; return SPB_OK;
    case 4: if (c->state < 3) break; *dev = c->seqs; *first = sl.seq_lo; *count = sl.seq_n; return SPB_OK;
    default: return SPB_EINVAL;
  }
  return fail(c, SPB_ESTATE, "spb_mg_shard: not compacted yet");
}

// local part of a sliced result to its place in the complete array at host_base (asynchronous on the context's stream when
// the host memory is page-locked, e.g. a shared buffer registered by every rank); *bytes = size of the part
int spb_mg_shard_to_host(spb_ctx* c, int which, void* host_base, uint64_t* bytes) {
  if (!c || !host_base) return SPB_EINVAL;
  void* dev = nullptr; uint64_t first = 0, count = 0;
  int r = spb_mg_shard(c, which, &dev, &first, &count);
  if (r) return r;
  const u64 w = which == 4 ? 1 : (which == 1 || which == 2) ? 8 : 4;
  if (count) dev_copy(c, (u8*)host_base + first * w, dev, count * w);
  if (bytes) *bytes = count * w;
  return SPB_OK;
}

// every sliced pass again, over the whole input, on this rank alone: afterwards the context is in the state a single GPU
// would have reached (all getters work).  No communication: the bitmaps and links are complete on every rank.
int spb_mg_complete(spb_ctx* c, spb_counts* out) {
  if (!c) return SPB_EINVAL;
  SPB_ON_DEVICE(c);
  if (!c->sl.on) { if (out) *out = c->cnt; return SPB_OK; }
  if (!c->mg_seen) return fail(c, SPB_ESTATE, "spb_mg_complete: the links are gone");
  const bool compacted = c->state == 3;
  free_graph_results(c);
  dfree(c, c->shard_tmp);
  c->sl = spb_ctx::Slice();
  c->state = 1;
  PRIM(dd_ids(c, 0, c->T, nullptr, 1, c->mg_seen));
  PRIM(dd_graph(c));
  int r = check_err(c, "spb_mg_complete");
  if (r) return r;
  c->state = 2;
  if (compacted) return spb_compact(c, out);
  if (out) *out = c->cnt;
  return SPB_OK;
}

// ------------------------------------------------------------------------------------------- getters
#define NEED(st) if (!c) return SPB_EINVAL; SPB_ON_DEVICE(c); if (c->state < (st)) return fail(c, SPB_ESTATE, "%s: results not available yet (state %d)", __func__, c->state)
// getters that read the per-position / per-vertex / per-NBP-vertex arrays: those are sliced over the ranks until spb_mg_complete
#define NEED_FULL(st) NEED(st); if (c->sl.on) return fail(c, SPB_ESTATE, "%s: the results of this context are sliced over the ranks (spb_mg_shard for the local part, spb_mg_complete for all of it)", __func__)

int spb_get_vertex2coords(spb_ctx* c, uint32_t* out) {
  NEED_FULL(2);
  u32* tmp; ALLOC(tmp, u32, 2ull * c->nV + 2);
  SPB_LAUNCH(k_coords, nblk(c->nV, TPB), TPB, c->stream, c->v2pos, c->nV, c->seq_off, c->n_seq, tmp);
  dev_copy(c, out, tmp, 8ull * c->nV); dev_sync(c);
  dfree(c, tmp);
  return check_err(c, __func__);
}
int spb_get_edges(spb_ctx* c, uint32_t* out) { NEED_FULL(2); PRIM(ensure_edges(c)); dev_copy(c, out, c->edges, 8ull * c->cnt.n_edges); dev_sync(c); return SPB_OK; }
int spb_get_instance_vertices(spb_ctx* c, uint32_t* out) { NEED_FULL(2); dev_copy(c, out, c->sp, 4ull * c->T); dev_sync(c); return SPB_OK; }
int spb_get_seqlimits(spb_ctx* c, uint32_t* out) { NEED(2); dev_copy(c, out, c->seqlim, 8ull * c->n_seq); dev_sync(c); return SPB_OK; }

static int build_intervals(spb_ctx* c) {
  if (c->iv) return 0;
  const u64 T = c->T;
  u8* flags; ALLOC(flags, u8, T + 8);
  SPB_LAUNCH(k_run_flags, nblk(T, TPB), TPB, c->stream, c->sp, T, flags);
  u32 nR = 0, nR2 = 0;
  PRIM(prim_count_flag(c, flags, 1, T, &nR));
  u32 *rs, *re; ALLOC(rs, u32, nR + 1); ALLOC(re, u32, nR + 1);
  PRIM(prim_select_flag(c, flags, 1, T, rs, &nR));
  PRIM(prim_select_flag(c, flags, 2, T, re, &nR2));
  dfree(c, flags);
  if (nR != nR2) return fail(c, SPB_ECUDA, "internal: run starts (%u) != run ends (%u)", nR, nR2);
  u64 *key, *val, *k2, *v2; ALLOC(key, u64, nR + 1); ALLOC(val, u64, nR + 1); ALLOC(k2, u64, nR + 1); ALLOC(v2, u64, nR + 1);
  SPB_LAUNCH(k_run_intervals, nblk(nR, TPB), TPB, c->stream, c->sp, rs, re, nR, c->seq_off, c->n_seq, key, val);
  dfree(c, rs); dfree(c, re);
  PRIM(prim_sort_u64_u64(c, key, val, k2, v2, nR));
  u64* runmax = k2;        // reuse the sort's spare key buffer
  PRIM(prim_segmax_u64(c, val, runmax, nR));
  u32 *head, *pos; ALLOC(head, u32, nR + 1); ALLOC(pos, u32, nR + 1);
  SPB_LAUNCH(k_interval_heads, nblk(nR, TPB), TPB, c->stream, key, runmax, nR, head);
  u64 nIv = 0;
  PRIM(prim_scan_u32(c, head, pos, nR, &nIv));
  u32* iv_seq; ALLOC(iv_seq, u32, nIv + 1); ALLOC(c->iv, u32, 2 * nIv + 2);
  SPB_LAUNCH(k_interval_write, nblk(nR, TPB), TPB, c->stream, key, runmax, nR, head, pos, c->iv, iv_seq);
  ALLOC(c->iv_off, u64, c->n_seq + 2);
  SPB_LAUNCH(k_interval_offsets, nblk(c->n_seq + 1, TPB), TPB, c->stream, iv_seq, (u32)nIv, c->n_seq, c->iv_off);
  dfree(c, key); dfree(c, val); dfree(c, k2); dfree(c, v2); dfree(c, head); dfree(c, pos); dfree(c, iv_seq);
  c->nIv = (u32)nIv; c->cnt.n_seq_intervals = nIv;
  dev_sync(c);
  return check_err(c, "seq_intervals");
}
int spb_get_seq_intervals(spb_ctx* c, uint64_t* off, uint32_t* intervals) {
  NEED_FULL(2);
  int r = build_intervals(c);
  if (r) return r;
  if (off) dev_copy(c, off, c->iv_off, 8ull * (c->n_seq + 1));
  if (intervals) dev_copy(c, intervals, c->iv, 8ull * c->nIv);
  dev_sync(c);
  return SPB_OK;
}
int spb_get_components(spb_ctx* c, int32_t* out) {
  NEED_FULL(3);
  int32_t* tmp; ALLOC(tmp, int32_t, (u64)c->nV + 1);
  PRIM(ensure_vnbp(c));
  SPB_LAUNCH(k_vertex_comp, nblk(c->nV, TPB), TPB, c->stream, c->vflags, c->nV, c->vnbp, c->init_list, c->nI, c->parent, c->label_of_root,
             c->nbp_comp, tmp);
  dev_copy(c, out, tmp, 4ull * c->nV); dev_sync(c);
  dfree(c, tmp);
  return check_err(c, __func__);
}
int spb_get_nbps(spb_ctx* c, uint64_t* off, uint32_t* verts, int32_t* comp, uint8_t* is_cycle) {
  NEED(3);
  if (off) dev_copy(c, off, c->voff, 8ull * (c->nN + 1));
  if (verts && c->sl.on) return fail(c, SPB_ESTATE, "spb_get_nbps: the vertex lists are sliced over the ranks (spb_mg_shard / spb_mg_complete)");
  if (verts) dev_copy(c, verts, c->verts, 4ull * c->cnt.nbp_vertices);
  if (comp) dev_copy(c, comp, c->nbp_comp, 4ull * c->nN);
  if (is_cycle) dev_copy(c, is_cycle, c->nt.cyc, (u64)c->nN);
  dev_sync(c);
  return SPB_OK;
}
int spb_get_nbp_seqs(spb_ctx* c, uint64_t* off, uint8_t* ascii) {
  NEED(3);
  if (off) dev_copy(c, off, c->soff, 8ull * (c->nN + 1));
  if (ascii && c->sl.on) return fail(c, SPB_ESTATE, "spb_get_nbp_seqs: the sequences are sliced over the ranks (spb_mg_shard / spb_mg_complete)");
#ifndef SPB_EMU
  if (ascii && ascii == c->out_seq_host && c->d2h_pending) {           // copied by spb_compact on the second stream: wait for it
    dev_sync(c);
    CK(cudaEventSynchronize(c->ev_d2h));
    c->d2h_pending = false;
    return SPB_OK;
  }
#endif
  if (ascii) dev_copy(c, ascii, c->seqs, c->cnt.nbp_bases);
  dev_sync(c);
  return SPB_OK;
}

// Page-locked host buffer that is to receive the NBP sequences (the bytes spb_get_nbp_seqs returns): when set, spb_compact
// starts the copy on a second stream as soon as the sequences are emitted, so that it overlaps with the origin sets,
// components and edge rows; spb_get_nbp_seqs(ctx, off, the same pointer) then only waits for it.  nullptr clears.
int spb_set_host_output(spb_ctx* c, uint8_t* nbp_seqs_host, uint64_t capacity) {
  if (!c) return SPB_EINVAL;
  SPB_ON_DEVICE(c);
#ifndef SPB_EMU
  if (c->d2h_pending) { cudaEventSynchronize(c->ev_d2h); c->d2h_pending = false; }
#endif
  c->out_seq_host = nbp_seqs_host; c->out_seq_cap = nbp_seqs_host ? capacity : 0;
  return SPB_OK;
}
int spb_get_nbp_origins(spb_ctx* c, uint64_t* off, uint32_t* seq_idx) {
  NEED(3);
  if (off) dev_copy(c, off, c->ooff, 8ull * (c->nN + 1));
  if (seq_idx) dev_copy(c, seq_idx, c->origins, 4ull * c->cnt.n_origin_pairs);
  dev_sync(c);
  return SPB_OK;
}
uint64_t spb_launch_count(void) { return g_spb_launches; }
int spb_get_counts(spb_ctx* c, spb_counts* out) {
  if (!c || !out) return SPB_EINVAL;
  *out = c->cnt;
  return SPB_OK;
}
static int build_nbp_graph(spb_ctx* c) {
  if (c->g_off) return 0;
  const u32 nN = c->nN;
  u64 *key, *key2; u32 *id, *id2;
  ALLOC(key, u64, nN + 1); ALLOC(key2, u64, nN + 1); ALLOC(id, u32, nN + 1); ALLOC(id2, u32, nN + 1);
  SPB_LAUNCH(k_nbpg_keys, nblk(nN, TPB), TPB, c->stream, c->nt, nN, key, id);
  PRIM(prim_partition_u64_u32(c, key, id, key2, id2, nN, 0, 34));
  dfree(c, key2); dfree(c, id2);
  u64* cnt; ALLOC(cnt, u64, nN + 1); ALLOC(c->g_off, u64, nN + 2); ALLOC(c->g_rev, u32, nN + 1);
  SPB_LAUNCH(k_nbpg_join, nblk(nN, TPB), TPB, c->stream, c->nt, nN, key, id, c->dart_nbp, (const u64*)nullptr, cnt, (u32*)nullptr, c->g_rev);
  u64 tot = 0;
  PRIM(prim_scan_u64(c, cnt, c->g_off, nN, &tot));
  dev_copy(c, c->g_off + nN, &tot, 8); dev_sync(c);
  ALLOC(c->g_succ, u32, tot + 1);
  SPB_LAUNCH(k_nbpg_join, nblk(nN, TPB), TPB, c->stream, c->nt, nN, key, id, c->dart_nbp, c->g_off, cnt, c->g_succ, (u32*)nullptr);
  dfree(c, key); dfree(c, id); dfree(c, cnt);
  c->g_edges = tot;
  dev_sync(c);
  return check_err(c, "nbp_graph");
}
int spb_get_nbp_graph(spb_ctx* c, uint64_t* n_edges, uint64_t* off, uint32_t* succ, uint32_t* rev) {
  NEED_FULL(3);
  int r = build_nbp_graph(c);
  if (r) return r;
  if (n_edges) *n_edges = c->g_edges;
  if (off) dev_copy(c, off, c->g_off, 8ull * (c->nN + 1));
  if (succ) dev_copy(c, succ, c->g_succ, 4ull * c->g_edges);
  if (rev) dev_copy(c, rev, c->g_rev, 4ull * c->nN);
  dev_sync(c);
  return SPB_OK;
}
// SURVEY 8(f) row 1, second half: which input sequences contain which raw NBP in their own orientation (reference
// get_seqPaths_NBPs, lib/Assembler.py:1121-1139).  Built once per compaction: sorted unique (NBP << sb | sequence) pairs
// and the per-NBP counts (what the reference accumulates into nv2rightOrientation, :448-453).
static int build_occurrences(spb_ctx* c) {
  if (c->occ_counts) return 0;
  int r = build_nbp_graph(c);                        // start keys are rebuilt below; the graph supplies the partner ids
  if (r) return r;
  const u32 nN = c->nN;
  const u64 T = c->T;
  u64 *key, *key2; u32 *id, *id2;
  ALLOC(key, u64, nN + 1); ALLOC(key2, u64, nN + 1); ALLOC(id, u32, nN + 1); ALLOC(id2, u32, nN + 1);
  SPB_LAUNCH(k_nbpg_keys, nblk(nN, TPB), TPB, c->stream, c->nt, nN, key, id);
  PRIM(prim_partition_u64_u32(c, key, id, key2, id2, nN, 0, 34));
  dfree(c, key2); dfree(c, id2);
  u64* cnt2; ALLOC(cnt2, u64, 2);
  u64 ccap = std::max<u64>(4096, 4ull * nN + 4ull * c->n_seq);
  u64* cand = nullptr; u64 nc = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    ALLOC(cand, u64, ccap);
    dev_memset(c, cnt2, 0, 16);
    SPB_LAUNCH(k_occ_candidates, gs_grid(T), TPB, c->stream, c->sp, T, c->vflags, c->v2pos, c->obits, c->seq_off, c->n_seq, key, id, nN, c->soff,
               cand, cnt2, ccap);
    nc = dread(c, cnt2);
    if (nc <= ccap) break;
    dfree(c, cand); ccap = nc;
  }
  dfree(c, key); dfree(c, id);
  u32 sb = 1; while ((1ull << sb) < c->n_seq) ++sb;
  u64* pairs; ALLOC(pairs, u64, nc + 1);
  dev_memset(c, cnt2, 0, 16);
#ifdef SPB_EMU
  SPB_LAUNCH(k_occ_verify, nblk(nc, TPB), TPB, c->stream, cand, nc, c->F, c->seqs, c->soff, c->seq_off, c->n_seq, sb, pairs, cnt2);
#else
  SPB_LAUNCH(k_occ_verify, nblk(nc * 32, TPB), TPB, c->stream, cand, nc, c->F, c->seqs, c->soff, c->seq_off, c->n_seq, sb, pairs, cnt2);
#endif
  u64 np = dread(c, cnt2);
  dfree(c, cand); dfree(c, cnt2);
  u32 nu = 0;
  if (np) {
    u64* tmp; ALLOC(tmp, u64, np + 1);
    PRIM(prim_sort_u64(c, pairs, tmp, np));
    dfree(c, tmp);
    u8* nopay = nullptr;
    PRIM(unique_sorted(c, pairs, nopay, np, &nu));
  }
  ALLOC(c->occ_counts, u32, nN + 1);
  PRIM(ensure_vnbp(c));
  SPB_LAUNCH(k_occ_counts, nblk(nN, TPB), TPB, c->stream, pairs, nu, sb, c->nt, nN, c->g_rev, c->vnbp, c->vflags, c->nV, c->occ_counts);
  // per-sequence view: (sequence << 32 | NBP), sorted; pairs of NBPs the reference's pre-filter rejects are dropped
  u64 *sk, *sk2; ALLOC(sk, u64, (u64)nu + 1); ALLOC(sk2, u64, (u64)nu + 1);
  SPB_LAUNCH(k_occ_rekey, nblk(nu, TPB), TPB, c->stream, pairs, nu, sb, c->occ_counts, sk);
  PRIM(prim_sort_u64(c, sk, sk2, nu));
  dfree(c, sk2); dfree(c, pairs);
  ALLOC(c->occ_off, u64, (u64)c->n_seq + 2);
  SPB_LAUNCH(k_occ_seq_offsets, nblk((u64)c->n_seq + 1, TPB), TPB, c->stream, sk, nu, c->n_seq, c->occ_off);
  ALLOC(c->occ_ids, u32, (u64)nu + 1);
  SPB_LAUNCH(k_occ_ids, nblk(nu, TPB), TPB, c->stream, sk, nu, c->occ_ids);
  dfree(c, sk);
  u64 last = 0;
  dev_copy(c, &last, c->occ_off + c->n_seq, 8); dev_sync(c);
  c->occ_pairs = last;                               // pairs that survive the filter
  return check_err(c, "nbp_occurrences");
}
int spb_get_nbp_orientation_counts(spb_ctx* c, uint32_t* counts) {
  NEED_FULL(3);
  if (!counts) return SPB_EINVAL;
  if (c->nN == 0) return SPB_OK;
  int r = build_occurrences(c);
  if (r) return r;
  dev_copy(c, counts, c->occ_counts, 4ull * c->nN); dev_sync(c);
  return SPB_OK;
}
int spb_get_seq_nbps(spb_ctx* c, uint64_t* n_pairs, uint64_t* off, uint32_t* nbp_ids) {
  NEED_FULL(3);
  if (c->nN == 0) { if (n_pairs) *n_pairs = 0; if (off) for (u32 s = 0; s <= c->n_seq; ++s) off[s] = 0; return SPB_OK; }
  int r = build_occurrences(c);
  if (r) return r;
  if (n_pairs) *n_pairs = c->occ_pairs;
  if (off) dev_copy(c, off, c->occ_off, 8ull * (c->n_seq + 1));
  if (nbp_ids) dev_copy(c, nbp_ids, c->occ_ids, 4ull * c->occ_pairs);
  dev_sync(c);
  return SPB_OK;
}
// bubble candidates = NBPs that share (first vertex, last vertex) with another NBP
// (reference lib/Assembler.py:728-734); group[i] = smallest NBP id of i's group, 0xFFFFFFFF when i is alone.
int spb_get_bubble_groups(spb_ctx* c, uint32_t* group) {
  NEED(3);
  if (!group) return SPB_EINVAL;
  const u32 nN = c->nN;
  if (nN == 0) return SPB_OK;
  u64 *key, *key2; u32 *id, *id2, *out;
  ALLOC(key, u64, nN + 1); ALLOC(key2, u64, nN + 1); ALLOC(id, u32, nN + 1); ALLOC(id2, u32, nN + 1); ALLOC(out, u32, nN + 1);
  SPB_LAUNCH(k_bubble_keys, nblk(nN, TPB), TPB, c->stream, c->nt, nN, key, id);
  PRIM(prim_partition_u64_u32(c, key, id, key2, id2, nN, 0, 64));
  dfree(c, key2); dfree(c, id2);
  SPB_LAUNCH(k_bubble_groups, nblk(nN, TPB), TPB, c->stream, key, id, nN, out);
  dev_copy(c, group, out, 4ull * nN); dev_sync(c);
  dfree(c, key); dfree(c, id); dfree(c, out);
  return check_err(c, "bubble_groups");
}
// SURVEY 8(f) row 2 (first half): sequences of host-supplied vertex paths -- reconstruct_sequence for the joined paths of
// the host tail (reference lib/Assembler.py:1038-1118, called at :911-921).  off uint64[n_paths + 1] and verts
// uint32[off[n_paths]] are host arrays; ascii receives off[n_paths] + n_paths * (k - 1) bytes, path i at
// off[i] + i * (k - 1).  SPB_EINVAL if a path is shorter than 2 or two consecutive vertices are not adjacent.
// host-supplied paths -> device vertex lists.  kind 0: vertex ids; kind 1: raw NBP ids (expanded on the device).  On
// return d_off[n_paths + 1] / d_verts hold the vertex paths on the device, *total their length.
static int paths_to_device(spb_ctx* c, int kind, uint32_t n_paths, const uint64_t* off, const uint32_t* items, u64** d_off_out, u32** d_verts_out, u64* total_out) {
  const u64 n_items = off[n_paths];
  u64* d_off; u32* d_items;
  ALLOC(d_off, u64, (u64)n_paths + 2); ALLOC(d_items, u32, n_items + 1);
  dev_copy(c, d_off, off, 8ull * (n_paths + 1)); dev_copy(c, d_items, items, 4ull * n_items);
  if (kind == 0) { *d_off_out = d_off; *d_verts_out = d_items; *total_out = n_items; return 0; }
  if (c->state < 3 || c->sl.on) return fail(c, SPB_ESTATE, "paths as NBP ids need the complete NBP vertex lists (spb_compact)");
  u64 *el_len, *el_off; ALLOC(el_len, u64, n_items + 1); ALLOC(el_off, u64, n_items + 1);
  SPB_LAUNCH(k_nbp_path_lengths, nblk(n_items, TPB), TPB, c->stream, d_off, n_paths, d_items, c->nN, c->voff, c->verts, el_len, c->errflag);
  u64 total = 0;
  PRIM(prim_scan_u64(c, el_len, el_off, n_items, &total));
  u32 e = dread(c, c->errflag);
  if (e & ERR_BADPATH) {
    dev_memset(c, c->errflag, 0, 4); dfree(c, d_off); dfree(c, d_items); dfree(c, el_len); dfree(c, el_off);
    return fail(c, SPB_EINVAL, "a path holds an unknown NBP id, or two consecutive NBPs that do not share their junction vertex");
  }
  u64* v_off; u32* v_verts; ALLOC(v_off, u64, (u64)n_paths + 2); ALLOC(v_verts, u32, total + 1);
  SPB_LAUNCH(k_nbp_path_write, std::min<u64>(nblk(n_items * 32, TPB), 148 * 16), TPB, c->stream, d_off, n_paths, d_items, c->voff, c->verts, el_off, v_off, v_verts);
  dev_copy(c, v_off + n_paths, &total, 8); dev_sync(c);
  dfree(c, d_off); dfree(c, d_items); dfree(c, el_len); dfree(c, el_off);
  *d_off_out = v_off; *d_verts_out = v_verts; *total_out = total;
  return 0;
}

// kind 0: paths are vertex lists; kind 1: lists of raw NBP ids.  vert_off (optional, host uint64[n_paths + 1]) receives the
// vertex offsets of the paths; ascii == nullptr: sizes only (path i needs vert_off[i + 1] - vert_off[i] + k - 1 bytes at
// vert_off[i] + i * (k - 1)).
int spb_path_sequences_ex(spb_ctx* c, int kind, uint32_t n_paths, const uint64_t* off, const uint32_t* items, uint64_t* vert_off, uint8_t* ascii) {
  NEED_FULL(2);
  if (!off || !items || (kind != 0 && kind != 1) || (!ascii && !vert_off)) return SPB_EINVAL;
  if (n_paths == 0) { if (vert_off) vert_off[0] = 0; return SPB_OK; }
  if (off[n_paths] == 0) return fail(c, SPB_EINVAL, "spb_path_sequences: empty paths");
  u64* d_off; u32* d_verts; u64 total = 0;
  PRIM(paths_to_device(c, kind, n_paths, off, items, &d_off, &d_verts, &total));
  if (vert_off) { dev_copy(c, vert_off, d_off, 8ull * (n_paths + 1)); dev_sync(c); }
  if (!ascii) { dfree(c, d_off); dfree(c, d_verts); return SPB_OK; }
  u8* d_out; ALLOC(d_out, u8, total + (u64)n_paths * (c->k - 1) + 1);
  GView g{c->vflags, c->Jk, c->Js, c->nJ, c->nV};
  SPB_LAUNCH(k_path_seqs, gs_grid(total), TPB, c->stream, g, c->F, c->R, c->T, c->k, c->v2pos, d_off, n_paths, d_verts, d_out, c->errflag);
  dev_copy(c, ascii, d_out, total + (u64)n_paths * (c->k - 1)); dev_sync(c);
  dfree(c, d_off); dfree(c, d_verts); dfree(c, d_out);
  u32 e = dread(c, c->errflag);
  if (e & ERR_BADPATH) { dev_memset(c, c->errflag, 0, 4); return fail(c, SPB_EINVAL, "spb_path_sequences: a path is shorter than 2 vertices or steps between vertices that are not adjacent"); }
  return check_err(c, "spb_path_sequences");
}
int spb_path_sequences(spb_ctx* c, uint32_t n_paths, const uint64_t* off, const uint32_t* verts, uint8_t* ascii) {
  if (!ascii) return SPB_EINVAL;
  return spb_path_sequences_ex(c, 0, n_paths, off, verts, nullptr, ascii);
}

static int build_init_seq_pairs(spb_ctx* c) {
  if (c->ivp) return 0;
  const u64 T = c->T;
  u32 sb = 1; while ((1ull << sb) <= c->n_seq) ++sb;
  u64* cnt; ALLOC(cnt, u64, 2);
  u8* has01; ALLOC(has01, u8, 2ull * c->n_seq + 8);
  u64 cap = std::max<u64>(4096, c->n_inst / 16);
  u64* pairs = nullptr; u64 np = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    ALLOC(pairs, u64, cap + 1);
    dev_memset(c, cnt, 0, 16); dev_memset(c, has01, 0, 2ull * c->n_seq + 8);
    SPB_LAUNCH(k_init_seq_pairs, gs_grid(T), TPB, c->stream, c->sp, T, c->vflags, c->seq_off, c->n_seq, sb, pairs, cnt, cap, has01);
    SPB_LAUNCH(k_init_seq_quirk, nblk(c->n_seq, TPB), TPB, c->stream, has01, c->n_seq, c->vflags, sb, pairs, cnt, cap);
    np = dread(c, cnt);
    if (np <= cap) break;
    dfree(c, pairs); cap = np;
  }
  dfree(c, cnt); dfree(c, has01);
  u32 nu = 0;
  if (np) {
    u64* tmp; ALLOC(tmp, u64, np + 1);
    PRIM(prim_sort_u64(c, pairs, tmp, np));
    dfree(c, tmp);
    u8* nopay = nullptr;
    PRIM(unique_sorted(c, pairs, nopay, np, &nu));
  }
  c->ivp = pairs; c->n_ivp = nu; c->ivp_sb = sb;
  return check_err(c, "init_seq_pairs");
}
// SURVEY 8(f) row 2 (second half): origins of host-supplied vertex paths that are concatenations of whole NBPs -- the
// goodOris of the joined paths (reference get_ori2cov_onevertex, lib/Assembler.py:1212-1248, + :941-945, without the
// `_Nsplit_` cut and without bubble_vertex2origins, which the caller unions in).  Result as CSR over the paths:
// out_off uint64[n_paths + 1], seq_idx uint32[<= cap] ascending per path; *n_pairs = pairs needed (call again with a
// larger `cap` when it exceeds it).
int spb_path_origins(spb_ctx* c, uint32_t n_paths, const uint64_t* off, const uint32_t* verts, uint64_t cap, uint64_t* n_pairs,
                     uint64_t* out_off, uint32_t* seq_idx) {
  return spb_path_origins_ex(c, 0, n_paths, off, verts, 0, nullptr, nullptr, nullptr, cap, n_pairs, out_off, seq_idx);
}
// kind as for spb_path_sequences_ex; bubble_vertex2origins as a CSR over ascending vertex ids (n_bub vertices, host arrays):
// bub_vert uint32[n_bub], bub_off uint64[n_bub + 1], bub_seq uint32[bub_off[n_bub]] (sequence indices)
int spb_path_origins_ex(spb_ctx* c, int kind, uint32_t n_paths, const uint64_t* off, const uint32_t* items, uint32_t n_bub, const uint32_t* bub_vert,
                        const uint64_t* bub_off, const uint32_t* bub_seq, uint64_t cap, uint64_t* n_pairs, uint64_t* out_off, uint32_t* seq_idx) {
  NEED_FULL(3);
  if (!off || !items || !n_pairs || !out_off || (cap && !seq_idx) || (kind != 0 && kind != 1) || (n_bub && (!bub_vert || !bub_off || !bub_seq))) return SPB_EINVAL;
  *n_pairs = 0;
  if (n_paths == 0) { out_off[0] = 0; return SPB_OK; }
  if (off[n_paths] == 0) return fail(c, SPB_EINVAL, "spb_path_origins: empty paths");
  for (u32 i = 1; i < n_bub; ++i) if (bub_vert[i] <= bub_vert[i - 1]) return fail(c, SPB_EINVAL, "spb_path_origins: bubble vertices must ascend");
  int r = build_init_seq_pairs(c);
  if (r) return r;
  PRIM(ensure_vnbp(c));
  u64* d_off; u32* d_verts; u64 *cnt, *pairs = nullptr; u64 total = 0;
  PRIM(paths_to_device(c, kind, n_paths, off, items, &d_off, &d_verts, &total));
  ALLOC(cnt, u64, 2);
  BubbleOri bub{nullptr, nullptr, nullptr, 0};
  u32 *d_bv = nullptr, *d_bs = nullptr; u64* d_bo = nullptr;
  if (n_bub) {
    const u64 nbs = bub_off[n_bub];
    for (u64 x = 0; x < nbs; ++x) if (bub_seq[x] >= c->n_seq) return fail(c, SPB_EINVAL, "spb_path_origins: bubble origin index out of range");
    ALLOC(d_bv, u32, n_bub + 1); ALLOC(d_bo, u64, (u64)n_bub + 2); ALLOC(d_bs, u32, nbs + 1);
    dev_copy(c, d_bv, bub_vert, 4ull * n_bub); dev_copy(c, d_bo, bub_off, 8ull * (n_bub + 1)); dev_copy(c, d_bs, bub_seq, 4ull * nbs);
    bub = BubbleOri{d_bv, d_bo, d_bs, n_bub};
  }
  u64 pcap = std::max<u64>(4096, 8ull * n_paths + cap + (n_bub ? bub_off[n_bub] : 0)), np = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    ALLOC(pairs, u64, pcap + 1);
    dev_memset(c, cnt, 0, 16);
    SPB_LAUNCH(k_path_origin_pairs, gs_grid(total), TPB, c->stream, d_off, n_paths, d_verts, c->nV, c->vflags, c->vnbp, c->ooff, c->origins,
               c->ivp, c->n_ivp, c->ivp_sb, bub, pairs, cnt, pcap, c->errflag);
    np = dread(c, cnt);
    if (np <= pcap) break;
    dfree(c, pairs); pcap = np;
  }
  dfree(c, d_off); dfree(c, d_verts); dfree(c, cnt); dfree(c, d_bv); dfree(c, d_bo); dfree(c, d_bs);
  u32 e = dread(c, c->errflag);
  if (e & ERR_BADPATH) {
    dev_memset(c, c->errflag, 0, 4); dfree(c, pairs);
    return fail(c, SPB_EINVAL, "spb_path_origins: a path is shorter than 2 vertices, holds an unknown vertex, or is not a concatenation of NBPs");
  }
  u32 nu = 0;
  if (np) {
    u64* tmp; ALLOC(tmp, u64, np + 1);
    PRIM(prim_sort_u64(c, pairs, tmp, np));
    dfree(c, tmp);
    u8* nopay = nullptr;
    PRIM(unique_sorted(c, pairs, nopay, np, &nu));
  }
  *n_pairs = nu;
  u64* d_out; ALLOC(d_out, u64, (u64)n_paths + 2);
  SPB_LAUNCH(k_path_origin_offsets, nblk((u64)n_paths + 1, TPB), TPB, c->stream, pairs, nu, n_paths, d_out);
  dev_copy(c, out_off, d_out, 8ull * (n_paths + 1));
  if (nu && nu <= cap) {
    u32* ids; ALLOC(ids, u32, (u64)nu + 1);
    SPB_LAUNCH(k_occ_ids, nblk(nu, TPB), TPB, c->stream, pairs, nu, ids);
    dev_copy(c, seq_idx, ids, 4ull * nu);
    dev_sync(c);
    dfree(c, ids);
  }
  dev_sync(c);
  dfree(c, d_out); dfree(c, pairs);
  return check_err(c, "spb_path_origins");
}
// ------------------------------------------------------------------------------------ FASTA reader
// Native restatement of the reference's read_fasta(fasta, ambigs='as_Ns', Ns='split', gap_size) (lib/utils.py:5-30, :52-59):
// the concatenated ASCII buffer + offsets that spb_load takes, without a pass through Python strings.
static void fasta_error(char* err, uint64_t cap, const std::string& m) { if (err && cap) snprintf(err, cap, "%s", m.c_str()); }
int spb_read_fasta(const char* path, uint32_t gap_size, spb_fasta* out, char* err, uint64_t err_cap) {
  if (!path || !out || gap_size == 0) return SPB_EINVAL;
  memset(out, 0, sizeof *out);
  FILE* f = fopen(path, "rb");
  if (!f) { fasta_error(err, err_cap, std::string("cannot open ") + path); return SPB_EINVAL; }
  std::string txt;
  { char buf[1 << 16]; size_t n; while ((n = fread(buf, 1, sizeof buf, f)) > 0) txt.append(buf, n); }
  fclose(f);
  {                                                                           // the reference opens the file in text mode: universal newlines
    size_t w = 0;
    for (size_t i = 0; i < txt.size(); ++i) {
      char ch = txt[i];
      if (ch == '\r') { ch = '\n'; if (i + 1 < txt.size() && txt[i + 1] == '\n') ++i; }
      txt[w++] = ch;
    }
    txt.resize(w);
  }
  auto is_ws = [](unsigned char ch) { return ch == ' ' || (ch >= 9 && ch <= 13) || (ch >= 28 && ch <= 31); };   // str.strip() / re \s on ASCII
  size_t b = 0, e = txt.size();
  while (b < e && is_ws((unsigned char)txt[b])) ++b;
  while (e > b && is_ws((unsigned char)txt[e - 1])) --e;
  while (b < e && txt[b] == '>') ++b;                                        // .lstrip('>')
  std::vector<std::string> names; std::vector<std::string> seqs;
  std::vector<std::string> keys;                                              // for the duplicate check (small inputs: linear is fine up to a few thousand; use a sorted copy)
  std::map<std::string, int> seen;
  const std::string gap(gap_size, 'N');
  size_t pos = b;
  while (pos <= e) {
    size_t nx = txt.find("\n>", pos);
    if (nx == std::string::npos || nx >= e) nx = e;
    const size_t nl = txt.find('\n', pos);
    if (nl == std::string::npos || nl >= nx) { fasta_error(err, err_cap, "record without a sequence line"); return SPB_EINVAL; }
    std::string name = txt.substr(pos, nl - pos);
    const size_t sp = name.find(' ');
    if (sp != std::string::npos) name.resize(sp);
    std::string seq; seq.reserve(nx - nl);
    for (size_t i = nl + 1; i < nx; ++i) {
      unsigned char ch = (unsigned char)txt[i];
      if (is_ws(ch) || ch == '.' || ch == '-') continue;
      if (ch >= 'a' && ch <= 'z') ch = (unsigned char)(ch - 32);
      if (ch == 'U') ch = 'T';
      switch (ch) { case 'R': case 'Y': case 'K': case 'M': case 'S': case 'W': case 'B': case 'D': case 'H': case 'V': ch = 'N'; default: break; }
      if (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T' && ch != 'N') {
        fasta_error(err, err_cap, "Illegal chars in seq \"" + name + "\"");
        return SPB_EINVAL;
      }
      seq.push_back((char)ch);
    }
    if (seen.count(name)) { fasta_error(err, err_cap, "Sequence \"" + name + "\" is duplicated in your input file"); return SPB_EINVAL; }
    auto put = [&](const std::string& n, std::string&& q) {
      auto it = seen.find(n);
      if (it != seen.end()) seqs[it->second] = std::move(q);                  // dict assignment keeps the first insertion position
      else { seen[n] = (int)names.size(); names.push_back(n); seqs.push_back(std::move(q)); }
    };
    if (seq.find(gap) == std::string::npos) put(name, std::move(seq));
    else {
      int idx = 0; size_t a = 0;
      for (;;) {
        size_t g = seq.find(gap, a);
        std::string piece = seq.substr(a, g == std::string::npos ? std::string::npos : g - a);
        size_t p0 = 0, p1 = piece.size();
        while (p0 < p1 && piece[p0] == 'N') ++p0;
        while (p1 > p0 && piece[p1 - 1] == 'N') --p1;
        if (p1 > p0) { put(name + "_Nsplit_" + std::to_string(idx), piece.substr(p0, p1 - p0)); ++idx; }
        if (g == std::string::npos) break;
        a = g + gap.size();
      }
    }
    if (nx >= e) break;
    pos = nx + 2;
  }
  uint64_t total = 0, nbytes = 0;
  for (auto& q : seqs) total += q.size();
  for (auto& n : names) nbytes += n.size() + 1;
  out->n_seq = (uint32_t)names.size(); out->n_bases = total; out->names_bytes = nbytes;
  out->ascii = (uint8_t*)malloc(total ? total : 1); out->seq_off = (uint64_t*)malloc(8 * (names.size() + 1)); out->names = (char*)malloc(nbytes ? nbytes : 1);
  if (!out->ascii || !out->seq_off || !out->names) { spb_free_fasta(out); return SPB_ENOMEM; }
  uint64_t o = 0, no = 0;
  for (size_t i = 0; i < names.size(); ++i) {
    out->seq_off[i] = o; memcpy(out->ascii + o, seqs[i].data(), seqs[i].size()); o += seqs[i].size();
    memcpy(out->names + no, names[i].data(), names[i].size()); no += names[i].size(); out->names[no++] = '\n';
  }
  out->seq_off[names.size()] = o;
  return SPB_OK;
}
// ------------------------------------------------------------------------------------------- writers
// copy `n` bytes that may live in device memory to the host, through a small pinned staging buffer
static bool fetch_bytes(const uint8_t* src, u64 n, std::vector<uint8_t>& host) {
  host.resize(n);
  if (!n) return true;
#ifdef SPB_EMU
  memcpy(host.data(), src, n);
  return true;
#else
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, src) != cudaSuccess || (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged)) {
    cudaGetLastError();                                // plain host memory (also: no device present at all)
    memcpy(host.data(), src, n);
    return true;
  }
  return cudaMemcpy(host.data(), src, n, cudaMemcpyDeviceToHost) == cudaSuccess;
#endif
}
// FASTA records ">{name}\n{seq}\n" in the order given -- what the reference's write_fasta (lib/utils.py:62-65) produces for
// a dict in insertion order, as used for NBPs.fasta / graph.fastg / *.core.fasta / *.aux.fasta (scripts/SuperPang.py:279-282).
// names: the n names back to back (no terminators) with name_off[n + 1]; seqs + seq_off[n + 1]: sequence bytes, in host
// OR device memory (e.g. the buffers of spb_get_nbp_seqs / spb_path_sequences left on the device); written in chunks of
// 64 MB so that a multi-gigabyte result never needs a second full copy on the host.
int spb_write_fasta(const char* path, uint64_t n, const char* names, const uint64_t* name_off, const uint8_t* seqs, const uint64_t* seq_off,
                    char* err, uint64_t err_cap) {
  if (!path || (n && (!names || !name_off || !seqs || !seq_off))) return SPB_EINVAL;
  FILE* f = fopen(path, "wb");
  if (!f) { fasta_error(err, err_cap, std::string("cannot open ") + path + " for writing"); return SPB_EINVAL; }
  std::vector<uint64_t> so;
  { std::vector<uint8_t> tmp; if (n && !fetch_bytes((const uint8_t*)seq_off, 8 * (n + 1), tmp)) { fclose(f); fasta_error(err, err_cap, "cannot read the sequence offsets"); return SPB_ECUDA; }
    so.resize(n + 1); if (n) memcpy(so.data(), tmp.data(), 8 * (n + 1)); }
  const u64 CH = 64ull << 20;
  std::vector<uint8_t> buf; std::string out;
  u64 i = 0;
  while (i < n) {                                       // records [i, j): as many as fit into one chunk (at least one)
    u64 j = i + 1;
    while (j < n && so[j + 1] - so[i] <= CH) ++j;
    if (!fetch_bytes(seqs + so[i], so[j] - so[i], buf)) { fclose(f); fasta_error(err, err_cap, "cannot read the sequences"); return SPB_ECUDA; }
    out.clear(); out.reserve((so[j] - so[i]) + (name_off[j] - name_off[i]) + 3 * (j - i));
    for (u64 r = i; r < j; ++r) {
      out.push_back('>'); out.append(names + name_off[r], name_off[r + 1] - name_off[r]); out.push_back('\n');
      out.append((const char*)buf.data() + (so[r] - so[i]), so[r + 1] - so[r]); out.push_back('\n');
    }
    if (fwrite(out.data(), 1, out.size(), f) != out.size()) { fclose(f); fasta_error(err, err_cap, std::string("short write to ") + path); return SPB_EINVAL; }
    i = j;
  }
  if (fclose(f) != 0) { fasta_error(err, err_cap, std::string("cannot close ") + path); return SPB_EINVAL; }
  return SPB_OK;
}
// NBP2origins.tsv / the edge table (scripts/SuperPang.py:284-294): header, then per node
//   "{node}\t{origins joined by ','}\t{bin of every origin joined by ','}\t{number of distinct bins}/{n_genomes}\n".
// Origins are indices into the table of source-sequence names (orig_names / orig_name_off[n_orig + 1]), in the order the
// caller wants them printed, as a CSR (origin_off[n + 1], origin_idx -- host or device memory, the form of
// spb_get_nbp_origins); orig_bin[n_orig] = index of the bin (genome) of every source sequence, bin names likewise.
int spb_write_origins_tsv(const char* path, const char* header, uint64_t n, const char* node_names, const uint64_t* node_off,
                          const uint64_t* origin_off, const uint32_t* origin_idx, uint32_t n_orig, const char* orig_names,
                          const uint64_t* orig_name_off, const uint32_t* orig_bin, const char* bin_names, const uint64_t* bin_name_off,
                          uint32_t n_bins, uint32_t n_genomes, char* err, uint64_t err_cap) {
  if (!path || !header || (n && (!node_names || !node_off || !origin_off))) return SPB_EINVAL;
  std::vector<uint8_t> t0, t1;
  if (n && !fetch_bytes((const uint8_t*)origin_off, 8 * (n + 1), t0)) { fasta_error(err, err_cap, "cannot read the origin offsets"); return SPB_ECUDA; }
  const uint64_t* oo = (const uint64_t*)t0.data();
  const u64 tot = n ? oo[n] : 0;
  if (tot && !fetch_bytes((const uint8_t*)origin_idx, 4 * tot, t1)) { fasta_error(err, err_cap, "cannot read the origin lists"); return SPB_ECUDA; }
  const uint32_t* oi = (const uint32_t*)t1.data();
  FILE* f = fopen(path, "wb");
  if (!f) { fasta_error(err, err_cap, std::string("cannot open ") + path + " for writing"); return SPB_EINVAL; }
  std::string out(header);
  std::vector<uint32_t> seen_mark(n_bins, 0xFFFFFFFFu);
  for (u64 r = 0; r < n; ++r) {
    out.append(node_names + node_off[r], node_off[r + 1] - node_off[r]); out.push_back('\t');
    for (u64 x = oo[r]; x < oo[r + 1]; ++x) {
      const u32 o = oi[x];
      if (o >= n_orig) { fclose(f); fasta_error(err, err_cap, "origin index out of range"); return SPB_EINVAL; }
      if (x > oo[r]) out.push_back(',');
      out.append(orig_names + orig_name_off[o], orig_name_off[o + 1] - orig_name_off[o]);
    }
    out.push_back('\t');
    u32 distinct = 0;
    for (u64 x = oo[r]; x < oo[r + 1]; ++x) {
      const u32 b = orig_bin[oi[x]];
      if (b >= n_bins) { fclose(f); fasta_error(err, err_cap, "bin index out of range"); return SPB_EINVAL; }
      if (x > oo[r]) out.push_back(',');
      out.append(bin_names + bin_name_off[b], bin_name_off[b + 1] - bin_name_off[b]);
      if (seen_mark[b] != (u32)r) { seen_mark[b] = (u32)r; ++distinct; }
    }
    out.push_back('\t'); out.append(std::to_string(distinct)); out.push_back('/'); out.append(std::to_string(n_genomes)); out.push_back('\n');
    if (out.size() > (32u << 20)) { if (fwrite(out.data(), 1, out.size(), f) != out.size()) { fclose(f); fasta_error(err, err_cap, "short write"); return SPB_EINVAL; } out.clear(); }
  }
  if (fwrite(out.data(), 1, out.size(), f) != out.size()) { fclose(f); fasta_error(err, err_cap, "short write"); return SPB_EINVAL; }
  if (fclose(f) != 0) { fasta_error(err, err_cap, std::string("cannot close ") + path); return SPB_EINVAL; }
  return SPB_OK;
}

void spb_free_fasta(spb_fasta* fa) {
  if (!fa) return;
  free(fa->ascii); free(fa->seq_off); free(fa->names);
  memset(fa, 0, sizeof *fa);
}
int spb_get_stage_stats(spb_ctx* c, char* buf, uint64_t cap) {
  if (!c || !buf || !cap) return SPB_EINVAL;
  const u64 syncs = c->n_syncs;
  dev_sync(c);
  std::vector<std::pair<std::string, double>> acc;                            // insertion order, repeated names summed
  acc.push_back({"host_syncs", (double)syncs});
  for (auto& r : c->stat) {
    double v = r.val;
#ifndef SPB_EMU
    if (r.ev >= 0) { float ms = 0; if (cudaEventElapsedTime(&ms, c->evpool[r.ev], c->evpool[r.ev + 1]) != cudaSuccess) { cudaGetLastError(); ms = 0; } v = ms; }
#endif
    bool found = false;
    for (auto& a : acc) if (a.first == r.name) { a.second += v; found = true; break; }
    if (!found) acc.push_back({r.name, v});
  }
  std::string js = "{";
  for (size_t i = 0; i < acc.size(); ++i) { char b[192]; snprintf(b, sizeof b, "%s\"%s\": %.4f", i ? ", " : "", acc[i].first.c_str(), acc[i].second); js += b; }
  js += "}";
  snprintf(buf, cap, "%s", js.c_str());
  return SPB_OK;
}

}  // extern "C"
