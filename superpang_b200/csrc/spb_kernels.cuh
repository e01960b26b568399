// superpang_b200 -- device kernels (sm_100a) for DBG build + NBP compaction.
//
// Data layout in HBM:
//   F, R      2-bit packed forward strand and reverse-complement strand of the CONCATENATED input,
//             32 bases per u64, first base in the top bits, so that integer compare of limbs ==
//             lexicographic compare of bases (A<C<G<T).  rc(F[a..a+k)) == R[T-a-k..T-a).
//   position  p = global base offset; a k-mer instance is identified by its start position.
//   slots     open-addressing table, u64 = tag(24) | first-known position(39) | orientation(1).
//   sp        u32 per position: vertex id of the instance (SPB_INVALID where no k-mer starts).
//   vflags    u8 per vertex (VF_* bits); implicit edge (v, v+1) is the VF_R bit of v.
//   J         sorted unique directed "junction" entries (src<<32|dst) + side byte; these are the
//             only edges that are not (v, v+1) between consecutive first appearances.
#pragma once
#include "spb_platform.h"

#define SPB_INVALID 0xFFFFFFFFu
#define SPB_TERM 0xFFFFFFFFu
#define SPB_EMPTY 0xFFFFFFFFFFFFFFFFull
#define SPB_POSMASK ((1ull << 39) - 1)

enum { VF_R = 1, VF_HASJ = 2, VF_SEQLIM = 4, VF_INIT = 8, VF_CYC = 16, VF_PSTART = 32, VF_PEND = 64, VF_IN2 = 128 };
enum { JS_SRC_R = 1, JS_DST_R = 2, JS_CUT = 4 };
enum { SIDE_L = 0, SIDE_R = 1 };
enum { ERR_BADCHAR = 1, ERR_DEGENERATE = 2, ERR_OVERFLOW = 4, ERR_BADPATH = 8 };

// ------------------------------------------------------------------------------------ bit helpers
__device__ __forceinline__ u32 get_bit(const u32* bits, u64 i) { return (bits[i >> 5] >> (i & 31)) & 1u; }

// One bit per thread of a full grid: warp ballot on the GPU (no atomics); the emulator ORs into a
// pre-zeroed array.  `i` must be the global thread id and every thread of the warp must call it.
__device__ __forceinline__ void store_bit(u32* bits, u64 i, bool flag) {
#ifdef SPB_EMU
  if (flag) bits[i >> 5] |= 1u << (i & 31);
#else
  u32 m = __ballot_sync(0xffffffffu, flag);
  if ((threadIdx.x & 31) == 0) bits[i >> 5] = m;
#endif
}

// append `n` items to a list guarded by a u64 counter; returns the first slot
__device__ __forceinline__ u64 list_reserve(u64* counter, u32 n) {
#ifdef SPB_EMU
  return atomicAdd(counter, (u64)n);
#else
  // warp-aggregated: one atomic per warp
  u32 active = __activemask();
  u32 lane = threadIdx.x & 31;
  // inclusive prefix over active lanes (n is 0..3: do it with ballots per bit)
  u32 b0 = __ballot_sync(active, n & 1), b1 = __ballot_sync(active, n & 2);
  u32 below = (1u << lane) - 1;
  u32 before = __popc(b0 & below) + 2 * __popc(b1 & below);
  u32 total = __popc(b0) + 2 * __popc(b1);
  int leader = __ffs(active) - 1;
  u64 base = 0;
  if ((int)lane == leader) base = atomicAdd(counter, (u64)total);
  base = __shfl_sync(active, base, leader);
  return base + before;
#endif
}

// append one item per flagged thread to a list guarded by a u64 counter (one atomic per warp); callable under
// divergence.  Returns the slot of the calling thread (meaningful only when flag is set).
__device__ __forceinline__ u64 warp_append(u64* counter, bool flag) {
#ifdef SPB_EMU
  return flag ? atomicAdd(counter, 1ull) : ~0ull;
#else
  u32 act = __activemask();
  u32 bal = __ballot_sync(act, flag);
  if (!flag) return ~0ull;
  u32 lane = threadIdx.x & 31;
  int leader = __ffs(bal) - 1;
  u64 base = 0;
  if ((int)lane == leader) base = atomicAdd(counter, (u64)__popc(bal));
  base = __shfl_sync(bal, base, leader);
  return base + __popc(bal & ((1u << lane) - 1));
#endif
}

// ------------------------------------------------------------------------------------ windows
// Packed strands are arrays of u32 words, 16 bases per word, first base in the top bits, so that
// unsigned compare of 32-bit limbs == lexicographic compare of bases (A<C<G<T).
// limb j of the window starting at base a:  funnelshift(w[a/16 + j], w[a/16 + j + 1], 2*(a%16)).
__device__ __forceinline__ u32 limb_at(const u32* __restrict__ w, u32 sh, u32 j) { return __funnelshift_l(__ldg(w + j + 1), __ldg(w + j), sh); }
__device__ __forceinline__ u32 nlimbs(u32 k) { return (2 * k + 31) / 32; }
__device__ __forceinline__ u32 tail_mask(u32 k) { u32 nb = 2 * k - 32 * (nlimbs(k) - 1); return nb == 32 ? ~0u : (~0u << (32 - nb)); }

// lexicographic compare of two k-base windows: -1 / 0 / +1
__device__ __forceinline__ int win_cmp(const u32* A, u64 a, const u32* B, u64 b, u32 k) {
  const u32 W = nlimbs(k);
  const u32* x = A + (a >> 4); const u32* y = B + (b >> 4);
  const u32 sx = (u32)(a & 15) * 2, sy = (u32)(b & 15) * 2;
  u32 x0 = __ldg(x), y0 = __ldg(y);
  for (u32 j = 0; j < W; ++j) {
    u32 x1 = __ldg(x + j + 1), y1 = __ldg(y + j + 1);
    u32 lx = __funnelshift_l(x1, x0, sx), ly = __funnelshift_l(y1, y0, sy);
    if (j == W - 1) { u32 m = tail_mask(k); lx &= m; ly &= m; }
    if (lx != ly) return lx > ly ? 1 : -1;
    x0 = x1; y0 = y1;
  }
  return 0;
}

// 64-bit hash of a k-base window: two 32-bit multiplicative chains over the limbs + fmix64.
// Only the distribution matters: equality is always verified on the full 2k bits.
__device__ __forceinline__ u64 hash_finish(u32 a, u32 b) {
  u64 h = ((u64)a << 32) | b;
  h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
  return h;
}
template <int W> __device__ __forceinline__ u64 win_hash_t(const u32* A, u64 a, u32 k) {
  const u32* x = A + (a >> 4); const u32 sh = (u32)(a & 15) * 2;
  u32 ha = 0x9E3779B9u ^ k, hb = 0x85EBCA6Bu;
  u32 x0 = __ldg(x);
#pragma unroll
  for (int j = 0; j < W; ++j) {
    u32 x1 = __ldg(x + j + 1);
    u32 l = __funnelshift_l(x1, x0, sh);
    if (j == W - 1) l &= tail_mask(k);
    ha = ha * 0x9E3779B1u + l; hb = (hb ^ l) * 0x85EBCA77u;
    x0 = x1;
  }
  return hash_finish(ha, hb);
}
__device__ __forceinline__ u64 win_hash_rt(const u32* A, u64 a, u32 k) {
  const u32 W = nlimbs(k);
  const u32* x = A + (a >> 4); const u32 sh = (u32)(a & 15) * 2;
  u32 ha = 0x9E3779B9u ^ k, hb = 0x85EBCA6Bu;
  u32 x0 = __ldg(x);
  for (u32 j = 0; j < W; ++j) {
    u32 x1 = __ldg(x + j + 1);
    u32 l = __funnelshift_l(x1, x0, sh);
    if (j == W - 1) l &= tail_mask(k);
    ha = ha * 0x9E3779B1u + l; hb = (hb ^ l) * 0x85EBCA77u;
    x0 = x1;
  }
  return hash_finish(ha, hb);
}
template <int W> __device__ __forceinline__ u64 win_hash(const u32* A, u64 a, u32 k) {
  if (W > 0) return win_hash_t<(W > 0 ? W : 1)>(A, a, k);
  return win_hash_rt(A, a, k);
}

__device__ __forceinline__ u32 base_at(const u32* S, u64 a) { return (S[a >> 4] >> (30 - 2 * (a & 15))) & 3u; }

// ------------------------------------------------------------------------------------ K0: pack
// One thread per packed word of F and of R (16 bases).  A=0 C=1 G=2 T=3.
__device__ __forceinline__ u32 ascii_code(u32 c, u32* bad) {
  if (c != 'A' && c != 'C' && c != 'G' && c != 'T') *bad = 1;
  u32 t = (c >> 1) & 3u;        // A0 C1 T2 G3
  return t ^ (t >> 1);          // A0 C1 G2 T3
}
__global__ void k_pack(const u8* __restrict__ ascii, u64 T, u32* __restrict__ F, u64 nw, u32* err) {
  u64 w = SPB_GTID();
  if (w >= nw) return;
  u32 f = 0, bad = 0;
  u64 p0 = w * 16;
  if (p0 + 16 <= T && ((size_t)(ascii + p0) & 15) == 0) {
    const u32* q = (const u32*)(ascii + p0);
    for (u32 j = 0; j < 4; ++j) {
      u32 c4 = __ldg(q + j);
      for (u32 i = 0; i < 4; ++i) f = (f << 2) | ascii_code((c4 >> (8 * i)) & 0xFF, &bad);
    }
  } else {
    for (u32 j = 0; j < 16; ++j) { u64 p = p0 + j; f = (f << 2) | (p < T ? ascii_code(__ldg(ascii + p), &bad) : 0u); }
  }
  F[w] = f;
  if (bad) atomicOr(err, (u32)ERR_BADCHAR);
}
// reverse-complement strand from the packed forward strand: R[t] = 3 - F[T-1-t]
__device__ __forceinline__ u32 revcomp16(u32 x) {       // reverse the sixteen 2-bit groups and complement them
  u32 y = __brev(x);
  y = ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
  return ~y;
}
__global__ void k_pack_rc(const u32* __restrict__ F, u64 T, u32* __restrict__ R, u64 nw) {
  u64 w = SPB_GTID();
  if (w >= nw) return;
  u64 r0 = w * 16;                                   // R bases [r0, r0+16) = complement of F bases [T-r0-16, T-r0) reversed
  u32 out;
  if (r0 + 16 <= T) {
    u64 a = T - r0 - 16;
    u32 x = __funnelshift_l(__ldg(F + (a >> 4) + 1), __ldg(F + (a >> 4)), (u32)(a & 15) * 2);
    out = revcomp16(x);
  } else {
    out = 0;
    for (u32 j = 0; j < 16; ++j) { u64 t = r0 + j; u32 c = t < T ? 3u - base_at(F, T - 1 - t) : 0u; out = (out << 2) | c; }
  }
  R[w] = out;
}

// valid-start bitmap: bit p set iff a k-mer of some sequence starts at p.  One thread per 32 positions.
__device__ __forceinline__ u32 upper_bound_u64(const u64* a, u32 n, u64 x) {  // first i with a[i] > x
  u32 lo = 0, hi = n;
  while (lo < hi) { u32 mid = (lo + hi) >> 1; if (__ldg(a + mid) <= x) lo = mid + 1; else hi = mid; }
  return lo;
}
__global__ void k_valid_bits(const u64* __restrict__ seq_off, u32 n_seq, u64 T, u32 k, u32* __restrict__ vbits, u64 nwords32) {
  u64 w = SPB_GTID();
  if (w >= nwords32) return;
  u64 p0 = w * 32;
  u32 m = 0;
  if (p0 < T) {
    u32 s = upper_bound_u64(seq_off, n_seq + 1, p0) - 1;     // sequence containing p0
    for (; s < n_seq; ++s) {
      u64 b = __ldg(seq_off + s), e = __ldg(seq_off + s + 1);
      if (b >= p0 + 32) break;
      if (e - b > k) {                                        // len > k (reference lib/Assembler.py:151)
        u64 lo = b > p0 ? b : p0, hi = e - k + 1;             // valid starts [b, e-k]
        if (hi > p0 + 32) hi = p0 + 32;
        if (hi > lo) { u32 n = (u32)(hi - lo); m |= (n == 32 ? ~0u : ((1u << n) - 1)) << (lo - p0); }
      }
    }
  }
  vbits[w] = m;
}

// ------------------------------------------------------------------------------------ K1: extract + dedup insert
// One thread per base position.  Canonical = lexicographic max of (k-mer, reverse complement)
// (reference lib/Assembler.py:989-991).  The table keeps, per distinct canonical k-mer, the
// SMALLEST position seen (atomicMin) -> first appearance in (sequence, position) order
// (reference lib/Assembler.py:186-195).  Equality is always verified on the full 2k bits.
//   wbits[p]  the thread installed its entry (CAS into an empty slot, or an atomicMin that lowered it)
//   dbits[q]  position q was installed earlier and then displaced by a smaller position
// first appearance  <=>  wbits & ~dbits   (every displacement is observed exactly once through the
// value atomicMin returns), so no second pass over the table is needed to find the firsts.
__device__ __forceinline__ bool orient_fwd(const u32* __restrict__ F, const u32* __restrict__ R, u64 T, u32 k, u64 p) {
  u64 ra = T - p - k;                               // almost always decided by the first limb
  const u32* x = F + (p >> 4); const u32* y = R + (ra >> 4);
  u32 lx = __funnelshift_l(__ldg(x + 1), __ldg(x), (u32)(p & 15) * 2), ly = __funnelshift_l(__ldg(y + 1), __ldg(y), (u32)(ra & 15) * 2);
  return (lx != ly) ? (lx > ly) : (win_cmp(F, p, R, ra, k) >= 0);
}
// hash + probe + exact verification + atomicMin; returns won, and for a non-winner the matched (smaller) position
// and that position's orientation bit
template <int W> __device__ __forceinline__ bool insert_full(const u32* __restrict__ F, const u32* __restrict__ R, u64 T, u32 k, u64* slots,
                                                             u64 capmask, u32* dbits, u64 p, bool o, u32* seen, u32* seen_o) {
  const u32* S = o ? F : R;
  u64 a = o ? p : T - p - k;
  u64 h = win_hash<W>(S, a, k);
  u64 idx = h & capmask, tag = h >> 40;
  u64 entry = (tag << 40) | (p << 1) | (o ? 1ull : 0ull);
  for (;;) {
    u64 cur = atomicCAS(slots + idx, SPB_EMPTY, entry);
    if (cur == SPB_EMPTY) return true;
    if ((cur >> 40) == tag) {
      u64 q = (cur >> 1) & SPB_POSMASK;
      bool oq = cur & 1;
      if (q == p || win_cmp(S, a, oq ? F : R, oq ? q : T - q - k, k) == 0) {
        if (entry < cur) {
          u64 old = atomicMin(slots + idx, entry);
          if (old > entry) {                         // we lowered the slot: old's position is no longer the first
            u64 dq = (old >> 1) & SPB_POSMASK;
            atomicOr(dbits + (dq >> 5), 1u << (dq & 31));
            return true;
          }
          cur = old;
        }
        *seen = (u32)((cur >> 1) & SPB_POSMASK); *seen_o = (u32)(cur & 1);
        return false;
      }
    }
    idx = (idx + 1) & capmask;
  }
}
__device__ __forceinline__ bool all_valid(const u32* __restrict__ vbits, u64 a, u32 n) {   // n <= 32 consecutive valid starts from a
  u64 two = (u64)__ldg(vbits + (a >> 5)) | ((u64)__ldg(vbits + (a >> 5) + 1) << 32);
  u64 m = ((n >= 64 ? ~0ull : ((1ull << n) - 1))) << (a & 31);
  return (two & m) == m;
}
enum { AR_OQ = 1, AR_OP = 2, AR_MATCHED = 4, AR_WON = 8, AR_VALID = 16 };

// Pass 1 ("anchors"): every 32nd position goes through the table.  arec[i] = matched position | flags << 32.
template <int W> __global__ void SPB_LAUNCH_BOUNDS(256)
k_anchor_insert(const u32* __restrict__ F, const u32* __restrict__ R, const u32* __restrict__ vbits, u64 T, u32 k,
                u64* slots, u64 capmask, u32* __restrict__ seenpos, u32* dbits, u64* __restrict__ arec, u64 pos_lo, u64 pos_hi) {
  u64 p = pos_lo + SPB_GTID() * 32;                  // pos_lo is a multiple of 32
  if (p >= pos_hi) return;
  u64 rec = 0;
  if (p < T && get_bit(vbits, p)) {
    bool o = orient_fwd(F, R, T, k, p);
    u32 seen = SPB_INVALID, so = 0;
    bool won = insert_full<W>(F, R, T, k, slots, capmask, dbits, p, o, &seen, &so);
    seenpos[p] = won ? SPB_INVALID : seen;
    rec = (u64)(won ? 0u : seen) | ((u64)(AR_VALID | (o ? AR_OP : 0) | (won ? AR_WON : AR_MATCHED) | (so ? AR_OQ : 0)) << 32);
  }
  arec[p >> 5] = rec;
}

// Pass 2: all other positions.  If the anchor p0 of the 32-block matched an earlier position q0 (same or opposite
// strand), position p0 + j is tested against q0 +- j by comparing only the j new bases: equal => its k-mer equals the
// k-mer at m = q0 +- j (exact, by induction on the shared k-1 bases), no hash, no table access.  In genome sets that
// share long stretches this removes almost every random HBM access; everything else takes the full path.
//   wbits[p]  the thread installed its entry (CAS into an empty slot, or an atomicMin that lowered it)
//   dbits[q]  position q was installed earlier and then displaced by a smaller position
// first appearance  <=>  wbits & ~dbits; a non-first instance leaves in seenpos[p] a smaller position with the same
// k-mer (k_assign_ids follows these links to the first appearance).
template <int W> __global__ void SPB_LAUNCH_BOUNDS(256)
k_extract_insert(const u32* __restrict__ F, const u32* __restrict__ R, const u32* __restrict__ vbits, u64 T, u32 k,
                 u64* slots, u64 capmask, u32* __restrict__ seenpos, u32* obits, u32* wbits, u32* dbits,
                 const u64* __restrict__ arec, u64 pos_lo, u64 pos_hi, u64* ext_count) {
  u64 p = pos_lo + SPB_GTID();                       // pos_lo and the grid are multiples of 256
  bool valid = p < pos_hi && p < T && get_bit(vbits, p);
  bool o = false, won = false, ext = false;
  if (valid) {
    const u32 j = (u32)(p & 31);
    u32 seen = SPB_INVALID, so = 0;
    if (!arec) {                                     // plain mode: every position takes the full path
      o = orient_fwd(F, R, T, k, p);
      won = insert_full<W>(F, R, T, k, slots, capmask, dbits, p, o, &seen, &so);
      seenpos[p] = won ? SPB_INVALID : seen;
    } else {
      const u64 rec = __ldg(arec + (p >> 5));
      const u32 fl = (u32)(rec >> 32);
      o = j ? orient_fwd(F, R, T, k, p) : (fl & AR_OP) != 0;
      if (j == 0) won = (fl & AR_WON) != 0;          // the anchor itself was handled by k_anchor_insert
      else {
        if ((fl & AR_VALID) && (fl & AR_MATCHED)) {
          const u64 p0 = p - j, q0 = (u32)rec;
          const u32 vm = j == 31 ? 0xFFFFFFFFu : ((1u << (j + 1)) - 1);
          if ((__ldg(vbits + (p >> 5)) & vm) == vm) {                    // p0 .. p are k-mer starts of one sequence
            bool opposite = ((fl & AR_OQ) != 0) != ((fl & AR_OP) != 0);
            if (!opposite) {
              if (all_valid(vbits, q0, j + 1) && win_cmp(F, p0 + k, F, q0 + k, j) == 0) { ext = true; seen = (u32)(q0 + j); }
            } else if (q0 >= j) {
              if (all_valid(vbits, q0 - j, j + 1) && win_cmp(F, p0 + k, R, T - q0, j) == 0) { ext = true; seen = (u32)(q0 - j); }
            }
          }
        }
        if (!ext) won = insert_full<W>(F, R, T, k, slots, capmask, dbits, p, o, &seen, &so);
        seenpos[p] = won ? SPB_INVALID : seen;
      }
    }
  }
  store_bit(obits, p, o);
  store_bit(wbits, p, won);
  if (ext_count) {                                   // [0] extended positions, [1] duplicates (valid, not installed)
    bool dup = valid && !won;
#ifdef SPB_EMU
    if (ext) atomicAdd(ext_count, 1ull);
    if (dup) atomicAdd(ext_count + 1, 1ull);
#else
    u32 b = __ballot_sync(0xffffffffu, ext), b2 = __ballot_sync(0xffffffffu, dup);
    if ((threadIdx.x & 31) == 0) { if (b) atomicAdd(ext_count, (u64)__popc(b)); if (b2) atomicAdd(ext_count + 1, (u64)__popc(b2)); }
#endif
  }
}

// ------------------------------------------------------------------------------------ staged dedup
// The global table above needs two random HBM row activations per instance.  The staged path keeps
// the table in L2 instead: (S1) every instance becomes a 12-byte record (64-bit hash with the
// orientation in bit 0, position), the records are radix-partitioned by the top hash bits into
// buckets of ~1 M records; (S2) each bucket is deduplicated in a small scratch table that lives in L2;
// (S3) the representatives are scattered back to position order bucket-interleaved ("i-major"), so
// that writes to one sector happen close together in time and merge in L2.  With several GPUs the
// buckets are the unit of ownership: S1 and S3 run on the GPU that owns the positions, S2 on the GPU
// that owns the bucket, with an all-to-all of records before and of representatives after S2.
__device__ __forceinline__ u64 mix64(u64 x) {
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
  return x;
}
// spread = 1: positions that start no k-mer get a pseudo-random key and the position SPB_INVALID (they spread evenly
// over the sub-buckets of the two-level path instead of piling up in the last one); spread = 0: key ~0.
template <int W> __global__ void SPB_LAUNCH_BOUNDS(256)
k_make_records(const u32* __restrict__ F, const u32* __restrict__ R, const u32* __restrict__ vbits, u64 T, u32 k,
               u64 pos_lo, u64 pos_hi, u64* __restrict__ keys, u32* __restrict__ pos, u32* obits, u32 spread) {
  u64 i = SPB_GTID();
  u64 p = pos_lo + i;
  bool valid = p < pos_hi && get_bit(vbits, p);
  bool o = false;
  u64 key = ~0ull;
  if (valid) {
    u64 ra = T - p - k;
    const u32* x = F + (p >> 4); const u32* y = R + (ra >> 4);
    u32 lx = __funnelshift_l(__ldg(x + 1), __ldg(x), (u32)(p & 15) * 2), ly = __funnelshift_l(__ldg(y + 1), __ldg(y), (u32)(ra & 15) * 2);
    o = (lx != ly) ? (lx > ly) : (win_cmp(F, p, R, ra, k) >= 0);
    u64 h = win_hash<W>(o ? F : R, o ? p : ra, k);
    key = (h & ~1ull) | (o ? 1ull : 0ull);
    if (key == ~0ull) key ^= 2ull;
  }
  if (spread && !valid) { keys[i] = mix64(p); pos[i] = SPB_INVALID; }
  else { keys[i] = key; pos[i] = (u32)p; }
  store_bit(obits, p, o);
}

// Rolling variant of k_make_records.  One thread walks SPB_RUN consecutive positions and keeps, for the k-mer at p and
// two odd multipliers B (32-bit arithmetic: one IMAD per update),
//   HF(p) = sum_i F[p+i] B^(k-1-i)            polynomial hash of the forward string
//   HR(p) = sum_j (3 - F[p+j]) B^j            the same polynomial of the reverse-complement string
// both updated in O(1) per position (HR with B^-1 mod 2^32), so a position costs a few dozen instructions instead of
// a hash over all 2k bits (~220 at k=301); the hash of the canonical string is the pair of the chosen orientation,
// finished by roll_finish.  Everything is read from the forward strand: the first 16 bases of the reverse-complement
// k-mer are revcomp16 of the last 16 bases of the forward one.  Records: as k_make_records (orientation in bit 0).
#ifndef SPB_RUN
#define SPB_RUN 256
#endif
struct RollK { u32 B[2], Binv[2], Bk1[2], negBk[2], negBinv[2], B4[2]; };   // B, B^-1, B^(k-1), -B^k, -B^-1, B^4  (mod 2^32)
__device__ __forceinline__ u64 roll_finish(u32 a, u32 b) {
  // (the rotations matter: a M + b is always even -- every coefficient M B1^e + B2^e is -- which would leave 2^31 values)
  u32 t = a * 0x9E3779B1u + __funnelshift_l(b, b, 15); t ^= t >> 16; t *= 0x85EBCA6Bu; t ^= t >> 13; t *= 0xC2B2AE35u; t ^= t >> 16;
  u32 u = b * 0xC2B2AE3Du + __funnelshift_l(a, a, 13); u ^= u >> 15; u *= 0x2C1B3C6Du; u ^= u >> 12; u *= 0x297A2D39u; u ^= u >> 15;
  return ((u64)t << 32) | u;
}
__global__ void SPB_LAUNCH_BOUNDS(128)
k_make_records_roll(const u32* __restrict__ F, const u32* __restrict__ R, const u32* __restrict__ vbits, u64 T, u32 k,
                    u64 pos_lo, u64 pos_hi, u64 n, u64* __restrict__ keys, u32* __restrict__ pos, u32* obits, u32 spread, RollK K,
                    u32 own_bits, u32 own_id, u64* own_count, u64 own_cap) {
  // own_bits > 0 ("owner computes", multi-GPU): only the records whose top own_bits hash bits equal own_id are kept,
  // compacted behind the counter own_count (order: by warp group, not by position); nothing else is written.
  const u64 g = SPB_GTID();
  const bool alive = g * SPB_RUN < n;
#ifdef SPB_EMU
  if (!alive) return;
#else
  // the 32 keys a lane produces per group of steps are transposed through shared memory, so that the warp writes 32
  // rows of 256 contiguous bytes instead of 32 scattered 8-byte pieces per store
  __shared__ u64 tile[4][32][33];
  const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#endif
  const u64 p0 = pos_lo + g * SPB_RUN;               // multiple of 256
  bool have = false, fresh = false;
  u32 f1 = 0, f2 = 0, r1 = 0, r2 = 0;                // HF, HR for the two multipliers
  u32 head = 0, tail = 0, wb = 0, wc = 0;            // 16 bases at p / ending at p+k-1; cached words of the two input streams
  for (u32 w = 0; w < SPB_RUN / 32; ++w) {
    const u64 pw = p0 + 32ull * w;
    const u32 vb = (alive && pw < T) ? __ldg(vbits + (pw >> 5)) : 0u;
    u32 ow = 0;
#ifndef SPB_EMU
    u32 mine = 0;                                    // owned records of this lane's group
#endif
    for (u32 j = 0; j < 32; ++j) {
      const u64 p = pw + j;
      const bool valid = ((vb >> j) & 1u) && p < pos_hi;
      u64 key = ~0ull;
      if (valid) {
        if (!have) {                                 // direct evaluation (run start, or first k-mer of a sequence)
          f1 = f2 = r1 = r2 = 0;
          u32 wd = 0, q1 = 1, q2 = 1;                // q = B^i
          for (u32 i = 0; i < k; ++i) {
            const u64 a = p + i;
            if (i == 0 || (a & 15) == 0) wd = __ldg(F + (a >> 4));
            const u32 x = (wd >> (30 - 2 * (u32)(a & 15))) & 3u;
            f1 = f1 * K.B[0] + x; f2 = f2 * K.B[1] + x;
            r1 += (3u - x) * q1; r2 += (3u - x) * q2;
            q1 *= K.B[0]; q2 *= K.B[1];
          }
          const u64 e = p + k - 16;
          head = __funnelshift_l(__ldg(F + (p >> 4) + 1), __ldg(F + (p >> 4)), (u32)(p & 15) * 2);
          tail = __funnelshift_l(__ldg(F + (e >> 4) + 1), __ldg(F + (e >> 4)), (u32)(e & 15) * 2);
          have = true; fresh = true;
        } else {                                     // roll from p-1
          const u32 out = head >> 30;
          const u64 ib = p + k - 1, ic = p + 15;
          if (fresh || (ib & 15) == 0) wb = __ldg(F + (ib >> 4));
          if (fresh || (ic & 15) == 0) wc = __ldg(F + (ic >> 4));
          fresh = false;
          const u32 in = (wb >> (30 - 2 * (u32)(ib & 15))) & 3u;
          head = (head << 2) | ((wc >> (30 - 2 * (u32)(ic & 15))) & 3u);
          tail = (tail << 2) | in;
          const u32 co = 3u - out, ci = 3u - in;
          f1 = f1 * K.B[0] + in + out * K.negBk[0];                              // (HF - out B^(k-1)) B + in
          f2 = f2 * K.B[1] + in + out * K.negBk[1];
          r1 = r1 * K.Binv[0] + (ci * K.Bk1[0] + co * K.negBinv[0]);             // (HR - co) / B + ci B^(k-1)
          r2 = r2 * K.Binv[1] + (ci * K.Bk1[1] + co * K.negBinv[1]);
        }
        const u32 lx = head, ly = revcomp16(tail);
        const bool o = (lx != ly) ? (lx > ly) : (win_cmp(F, p, R, T - p - k, k) >= 0);
        key = (roll_finish(o ? f1 : r1, o ? f2 : r2) & ~1ull) | (o ? 1ull : 0ull);
        if (key == ~0ull) key ^= 2ull;
        ow |= (o ? 1u : 0u) << j;
      } else {
        have = false;
        if (spread) key = mix64(p);
      }
#ifdef SPB_EMU
      if (own_bits) {
        if (valid && (u32)(key >> (64 - own_bits)) == own_id) { const u64 i = atomicAdd(own_count, 1ull); if (i < own_cap) { keys[i] = key; pos[i] = (u32)p; } }
      } else { keys[p - pos_lo] = key; pos[p - pos_lo] = (spread && !valid) ? SPB_INVALID : (u32)p; }
#else
      tile[wid][lane][j] = key;
      if (own_bits && valid && (u32)(key >> (64 - own_bits)) == own_id) ++mine;
#endif
    }
    if (alive) obits[pw >> 5] = ow;
#ifndef SPB_EMU
    __syncwarp();
    const u64 g0 = g - lane;                         // first run of the warp
    const u64 nruns = n / SPB_RUN;
    const u32 rows = g0 >= nruns ? 0u : (u32)min((u64)32, nruns - g0);
    u32 vm = vb;                                     // valid steps of this lane's group (for the spread marking)
    if (pw + 32 > pos_hi) vm = pw >= pos_hi ? 0u : (vm & ((1u << (u32)(pos_hi - pw)) - 1u));
    u32 pv = (u32)(pos_lo + g0 * SPB_RUN + 32u * w + lane);
    if (own_bits) {                                  // compaction: one counter update per warp and group
      u32 incl = mine;
#pragma unroll
      for (u32 d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
      const u32 total = __shfl_sync(0xffffffffu, incl, 31);
      u64 base = 0;
      if (lane == 0 && total) base = atomicAdd(own_count, (u64)total);
      base = __shfl_sync(0xffffffffu, base, 0);
      for (u32 r = 0; r < rows; ++r) {
        const u32 vr = __shfl_sync(0xffffffffu, vm, r);
        const u64 off = base + __shfl_sync(0xffffffffu, incl - mine, r);
        const u64 key = tile[wid][r][lane];
        const bool own = ((vr >> lane) & 1u) && (u32)(key >> (64 - own_bits)) == own_id;
        const u32 bal = __ballot_sync(0xffffffffu, own);
        const u64 i = off + __popc(bal & ((1u << lane) - 1u));
        if (own && i < own_cap) { keys[i] = key; pos[i] = pv + r * SPB_RUN; }
      }
      __syncwarp();
      continue;
    }
    u64* kq = keys + (g0 * SPB_RUN + 32u * w + lane);
    u32* pq = pos + (g0 * SPB_RUN + 32u * w + lane);
    for (u32 r = 0; r < rows; ++r) {                 // row r = the 32 keys of lane r
      const u32 vr = __shfl_sync(0xffffffffu, vm, r);
      kq[(u64)r * SPB_RUN] = tile[wid][r][lane];
      pq[(u64)r * SPB_RUN] = (spread && !((vr >> lane) & 1u)) ? SPB_INVALID : pv + r * SPB_RUN;
    }
    __syncwarp();
#endif
  }
}

// Single-pass radix partition fused with the record generation (replaces k_make_records + a library radix pass on the
// single-GPU links path).  A block handles a tile of 2048 positions: hashes go to shared memory with a per-bucket
// histogram; one atomicAdd per (block, bucket) on the global bucket cursors reserves the output run; the records
// are then written into their runs.  Buckets are fixed-capacity regions (hashes are uniform; an overflow is flagged
// and the caller falls back to the exact two-pass partition).
#define SPB_PT 8
template <int W> __global__ void SPB_LAUNCH_BOUNDS(256)
k_partition_records(const u32* __restrict__ F, const u32* __restrict__ R, const u32* __restrict__ vbits, u64 T, u32 k, u32 bb,
                    u64* __restrict__ keys, u32* __restrict__ pos, u64 bucket_cap, u32* cursor, u32* obits, u32* err) {
  SPB_SHARED u64 skey[256 * SPB_PT];
  SPB_SHARED u32 hist[256], base[256];
  const u32 B = 1u << bb;
  const u64 tile = (u64)blockIdx.x * (256 * SPB_PT);
  SPB_PHASE_BEGIN(0)
    hist[threadIdx.x] = 0;
  SPB_PHASE_END
  SPB_PHASE_BEGIN(1)
    for (u32 j = 0; j < SPB_PT; ++j) {
      u64 p = tile + j * 256 + threadIdx.x;
      bool valid = p < T && get_bit(vbits, p);
      bool o = false;
      u64 key = ~0ull;
      if (valid) {
        o = orient_fwd(F, R, T, k, p);
        u64 h = win_hash<W>(o ? F : R, o ? p : T - p - k, k);
        key = (h & ~1ull) | (o ? 1ull : 0ull);
        if (key == ~0ull) key ^= 2ull;
        atomicAdd(&hist[bb ? (u32)(key >> (64 - bb)) : 0u], 1u);
      }
      skey[j * 256 + threadIdx.x] = key;
      store_bit(obits, p, o);
    }
  SPB_PHASE_END
  SPB_PHASE_BEGIN(2)
    if (threadIdx.x < B) {
      u32 cnt = hist[threadIdx.x];
      base[threadIdx.x] = cnt ? atomicAdd(cursor + threadIdx.x, cnt) : 0u;
      hist[threadIdx.x] = 0;
    }
  SPB_PHASE_END
  SPB_PHASE_BEGIN(3)
    for (u32 j = 0; j < SPB_PT; ++j) {
      u64 key = skey[j * 256 + threadIdx.x];
      if (key == ~0ull) continue;
      u32 b = bb ? (u32)(key >> (64 - bb)) : 0u;
      u64 idx = (u64)base[b] + atomicAdd(&hist[b], 1u);
      if (idx < bucket_cap) { keys[(u64)b * bucket_cap + idx] = key; pos[(u64)b * bucket_cap + idx] = (u32)(tile + j * 256 + threadIdx.x); }
      else atomicOr(err, (u32)ERR_OVERFLOW);
    }
  SPB_PHASE_END
}

// sub[r * (B + 1) + b] = first record of chunk r whose bucket (top `bb` bits of the key) is >= b_lo + b
__global__ void k_bucket_bounds(const u64* __restrict__ keys, const u64* __restrict__ chunk_off, u32 nchunks, u32 bb, u32 b_lo, u32 B,
                                u64* __restrict__ sub) {
  u64 g = SPB_GTID();
  if (g >= (u64)nchunks * (B + 1)) return;
  u32 r = (u32)(g / (B + 1)), b = (u32)(g % (B + 1));
  u64 lo = chunk_off[r], hi = chunk_off[r + 1];
  u64 bucket = (u64)b_lo + b;
  if (bb == 0) { sub[g] = b ? hi - chunk_off[r] : 0; return; }
  u64 a = lo, e = hi;
  while (a < e) { u64 mid = (a + e) >> 1; if ((__ldg(keys + mid) >> (64 - bb)) < bucket) a = mid + 1; else e = mid; }
  sub[g] = a - lo;
}

// sub[b] = first record whose sub-bucket (the `sb` hash bits below the `bbo` owner bits) is >= b; records sorted by those bits
__global__ void k_bucket_bounds_shift(const u64* __restrict__ keys, u64 n, u32 bbo, u32 sb, u32 B, u64* __restrict__ sub) {
  u64 b = SPB_GTID();
  if (b > B) return;
  if (sb == 0) { sub[b] = b ? n : 0; return; }
  u64 lo = 0, hi = n;
  while (lo < hi) { u64 mid = (lo + hi) >> 1; if (((__ldg(keys + mid) << bbo) >> (64 - sb)) < b) lo = mid + 1; else hi = mid; }
  sub[b] = lo;
}
struct BktView {
  const u64* keys; const u32* pos; u32* sidx;      // records (all chunks concatenated) and per-record slot / representative
  const u64* chunk_off; const u64* sub;            // chunk offsets [nchunks+1], sub-range table [nchunks][B+1]
  u32 nchunks, B, b;                               // b = bucket index relative to b_lo
};
__device__ __forceinline__ u64 bkt_record(const BktView& v, u64 t) {   // t-th record of bucket v.b -> index into the record arrays
  for (u32 r = 0; r < v.nchunks; ++r) {
    u64 s0 = v.sub[(u64)r * (v.B + 1) + v.b], s1 = v.sub[(u64)r * (v.B + 1) + v.b + 1];
    if (t < s1 - s0) return v.chunk_off[r] + s0 + t;
    t -= s1 - s0;
  }
  return ~0ull;
}
__global__ void SPB_LAUNCH_BOUNDS(256)
k_bkt_insert(BktView v, u64 n_b, const u32* __restrict__ F, const u32* __restrict__ R, u64 T, u32 k, u64* scratch, u64 capmask, u32* dbits) {
  u64 t = SPB_GTID();
  if (t >= n_b) return;
  u64 rec = bkt_record(v, t);
  u64 key = v.keys[rec];
  if (key == ~0ull) { v.sidx[rec] = SPB_INVALID; return; }
  u64 p = v.pos[rec];
  bool o = key & 1;
  u64 tag = (key >> 1) & 0x7FFFFFull, idx = (key >> 24) & capmask;
  u64 entry = (tag << 40) | (p << 1) | (o ? 1ull : 0ull);
  const u32* S = o ? F : R;
  u64 a = o ? p : T - p - k;
  u32 seen = (u32)p;                                // winners are their own candidate representative
  for (;;) {
    u64 cur = atomicCAS(scratch + idx, SPB_EMPTY, entry);
    if (cur == SPB_EMPTY) break;
    if ((cur >> 40) == tag) {
      u64 q = (cur >> 1) & SPB_POSMASK;
      bool oq = cur & 1;
      if (q == p || win_cmp(S, a, oq ? F : R, oq ? q : T - q - k, k) == 0) {
        if (entry < cur) {
          u64 old = atomicMin(scratch + idx, entry);
          u64 oq2 = (old >> 1) & SPB_POSMASK;
          if (old > entry) atomicOr(dbits + (oq2 >> 5), 1u << (oq2 & 31));    // oq2 is no longer the first position
          else q = oq2;
        }
        if (q > p) q = p;
        seen = (u32)q;
        break;
      }
    }
    idx = (idx + 1) & capmask;
  }
  v.sidx[rec] = seen;                                // exact unless `seen` is displaced later (k_bkt_final re-probes then)
}
// representative of every record of the bucket; the same launch clears the other scratch table
__global__ void k_bkt_final(BktView v, u64 n_b, const u32* __restrict__ F, const u32* __restrict__ R, u64 T, u32 k,
                            const u64* __restrict__ scratch, u64 capmask, const u32* __restrict__ dbits, u64* clear, u64 clear_n) {
  u64 t = SPB_GTID();
  if (t < n_b) {
    u64 rec = bkt_record(v, t);
    u32 r = v.sidx[rec];
    if (r != SPB_INVALID && get_bit(dbits, r)) {     // candidate was displaced after we saw it: exact re-probe by key
      u64 key = v.keys[rec], p = v.pos[rec];
      bool o = key & 1;
      u64 tag = (key >> 1) & 0x7FFFFFull, idx = (key >> 24) & capmask;
      const u32* S = o ? F : R;
      u64 a = o ? p : T - p - k;
      for (;;) {
        u64 cur = scratch[idx];
        if (cur == SPB_EMPTY) break;
        if ((cur >> 40) == tag) {
          u64 q = (cur >> 1) & SPB_POSMASK;
          bool oq = cur & 1;
          if (q == p || win_cmp(S, a, oq ? F : R, oq ? q : T - q - k, k) == 0) { v.sidx[rec] = (u32)q; break; }
        }
        idx = (idx + 1) & capmask;
      }
    }
  }
  u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 j = t; j < clear_n; j += stride) clear[j] = SPB_EMPTY;
}
// "links" variant of the per-bucket dedup: nothing is written per record.  Only the exceptions leave a trace, as a
// link "position -> smaller position with the same k-mer": (a) an instance that found its k-mer already installed,
// (b) an installed position that a smaller one displaced (written by the displacer).  Every position without a
// link is a first appearance, and every chain of links ends at one (k_assign_ids follows them).  With mostly unique
// k-mers almost no scattered write happens; the table of the bucket lives in L2.  The same launch clears the other
// scratch table for the next bucket.
__global__ void SPB_LAUNCH_BOUNDS(256)
k_bkt_insert_links(const u64* __restrict__ keys, const u32* __restrict__ pos, u64 lo, u64 n, const u32* __restrict__ F,
                   const u32* __restrict__ R, u64 T, u32 k, u64* scratch, u64 capmask, u32* seenpos, u32* dupbits,
                   u64* clear, u64 clear_n) {
  u64 t = SPB_GTID();
  if (t < n) {
    u64 key = keys[lo + t];
    if (key != ~0ull) {
      u64 p = pos[lo + t];
      bool o = key & 1;
      u64 tag = (key >> 1) & 0x7FFFFFull, idx = (key >> 24) & capmask;
      u64 entry = (tag << 40) | (p << 1) | (o ? 1ull : 0ull);
      const u32* S = o ? F : R;
      u64 a = o ? p : T - p - k;
      for (;;) {
        u64 cur = atomicCAS(scratch + idx, SPB_EMPTY, entry);
        if (cur == SPB_EMPTY) break;
        if ((cur >> 40) == tag) {
          u64 q = (cur >> 1) & SPB_POSMASK;
          bool oq = cur & 1;
          if (q == p || win_cmp(S, a, oq ? F : R, oq ? q : T - q - k, k) == 0) {
            u64 from = p, to = q;                                   // default: we link to the installed position
            if (entry < cur) {
              u64 old = atomicMin(scratch + idx, entry);
              u64 oq2 = (old >> 1) & SPB_POSMASK;
              if (old > entry) { from = oq2; to = p; }               // we displaced oq2: it links to us
              else to = oq2;
            }
            seenpos[from] = (u32)to;
            atomicOr(dupbits + (from >> 5), 1u << (from & 31));
            break;
          }
        }
        idx = (idx + 1) & capmask;
      }
    }
  }
  u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 j = t; j < clear_n; j += stride) clear[j] = SPB_EMPTY;
}

// ------------------------------------------------------------------------------------ shared-memory tables
// Two-level form of the "links" dedup: the records are radix-sorted on the top `nbits` hash bits into sub-buckets of
// ~7.6 K records, and one CTA deduplicates a sub-bucket in a hash table in shared memory -- no global atomics at all,
// the records stream through once.  slot = tag (18 hash bits) << 14 | index of the record inside the sub-bucket; a
// tag match is verified on the full 2k bits, and the earlier POSITION keeps the slot (CAS take-over), so the records
// of a sub-bucket may arrive in any order.  Exceptions leave links exactly as in k_bkt_insert_links.  Records are
// loaded four at a time per thread (independent loads in flight) before their inserts.
#define SPB_ST_SLOTS 16384u          // 64 KB of shared memory: three CTAs per SM
#define SPB_ST_LIMIT 12288u          // records per sub-bucket the table accepts (load <= 0.75); more -> caller falls back
#define SPB_ST_EMPTY 0xFFFFFFFFu
#ifndef SPB_ST_BATCH
#define SPB_ST_BATCH 4
#endif
#ifndef SPB_ST_MINB
#define SPB_ST_MINB 3
#endif
template <class V> __device__ __forceinline__ V st_pick(const V (&a)[SPB_ST_BATCH], u32 u) {   // a[u] without local memory
  V r = a[0];
#pragma unroll
  for (u32 j = 1; j < SPB_ST_BATCH; ++j) if (u == j) r = a[j];
  return r;
}
template <bool LIST> __global__ void SPB_LAUNCH_BOUNDS2(512, SPB_ST_MINB)
k_smem_dedup(const u64* __restrict__ keys, const u32* __restrict__ pos, const u64* __restrict__ sub, const u32* __restrict__ F,
             const u32* __restrict__ R, u64 T, u32 k, u32 nbits, u32* __restrict__ seenpos, u32* dupbits, u32* err,
             u64* __restrict__ links, u64* nlinks, u64 lcap) {
  // LIST (multi-GPU owner): the exceptions are appended to a list (from << 32 | to) instead of being written in place,
  // and the first appearance is decided by comparing positions (CAS take-over): the records arrive in any order.
  // !LIST (single GPU): the radix sort is stable, so index order is position order and atomicMin on the slot suffices.
  SPB_DYN_SHARED(u32, tab, SPB_ST_SLOTS);
  const u64 seg0 = sub[blockIdx.x];
  const u64 n64 = sub[blockIdx.x + 1] - seg0;
  if (n64 == 0) return;                              // uniform over the block
  if (n64 > SPB_ST_LIMIT) { if (threadIdx.x == 0) atomicOr(err, (u32)ERR_OVERFLOW); return; }
  const u32 n = (u32)n64;
  SPB_PHASE_BEGIN(0)
    for (u32 s = threadIdx.x * 4; s < SPB_ST_SLOTS; s += blockDim.x * 4) { tab[s] = SPB_ST_EMPTY; tab[s + 1] = SPB_ST_EMPTY; tab[s + 2] = SPB_ST_EMPTY; tab[s + 3] = SPB_ST_EMPTY; }
  SPB_PHASE_END
  SPB_PHASE_BEGIN(1)
    for (u32 base = 0; base < n; base += blockDim.x * SPB_ST_BATCH) {
      u64 kk[SPB_ST_BATCH]; u32 pp[SPB_ST_BATCH];
#pragma unroll
      for (u32 u = 0; u < SPB_ST_BATCH; ++u) {
        const u32 i = base + u * blockDim.x + threadIdx.x;
        pp[u] = i < n ? pos[seg0 + i] : SPB_INVALID;
        kk[u] = i < n ? keys[seg0 + i] : 0ull;
      }
      // A lane takes its next record as soon as the current one is placed, so a warp iterates max-over-lanes of the
      // SUM of the probe lengths of a batch, not the sum of the maxima (linear probing has a long tail).  The loop
      // exit is warp-uniform, which keeps the lanes converged at the atomic.
      u32 u = 0, s = 0, tag = 0, v = 0;
      u64 key = 0, p = 0;
      bool active = false;
      for (;;) {
        if (!active && u < SPB_ST_BATCH) {
          key = st_pick(kk, u); p = st_pick(pp, u);
          const u32 i = base + u * blockDim.x + threadIdx.x;
          ++u;
          if ((u32)p != SPB_INVALID) {                 // else: past the end, or a position that starts no k-mer
            const u64 rem = key << nbits;              // hash bits below the sub-bucket bits: 14 -> home slot, next 18 -> tag
            s = (u32)(rem >> 50);
            tag = (u32)(rem >> 32) & 0x3FFFFu; v = (tag << 14) | i;
            active = true;
          }
        }
        if (active) {
          const u32 old = atomicCAS(&tab[s], SPB_ST_EMPTY, v);
          if (old == SPB_ST_EMPTY) active = false;
          else {
            bool same = false;
            if ((old >> 14) == tag) {
              const u64 rq = seg0 + (old & 0x3FFFu);
              const u64 q = pos[rq];
              const bool o = key & 1, oq = keys[rq] & 1;
              if (win_cmp(o ? F : R, o ? p : T - p - k, oq ? F : R, oq ? q : T - q - k, k) == 0) {
                u64 from = p, to = q;                 // default: we link to the installed record
                if (LIST) {
                  if (p < q) {                        // ours is the earlier position: take the slot over
                    u32 cur = old; u64 pc = q;
                    for (;;) {
                      const u32 prev = atomicCAS(&tab[s], cur, v);
                      if (prev == cur) { from = pc; to = p; break; }     // we displaced it: it links to us
                      cur = prev; pc = pos[seg0 + (cur & 0x3FFFu)];      // another instance of the k-mer got in between
                      if (pc < p) { to = pc; break; }
                    }
                  }
                  const u64 slot = warp_append(nlinks, true);
                  if (slot < lcap) links[slot] = (from << 32) | to;
                } else {
                  if (v < old) {
                    const u32 prev = atomicMin(&tab[s], v);
                    const u64 pq = pos[seg0 + (prev & 0x3FFFu)];
                    if (prev > v) { from = pq; to = p; }                 // we displaced it: it links to us
                    else to = pq;
                  }
                  seenpos[from] = (u32)to;
                  atomicOr(dupbits + (from >> 5), 1u << (from & 31));
                }
                same = true;
              }
            }
            if (same) active = false; else s = (s + ((tag & (SPB_ST_SLOTS - 1)) | 1u)) & (SPB_ST_SLOTS - 1);   // double hashing: no clusters
          }
        }
        if (__all_sync(0xffffffffu, !active && u >= SPB_ST_BATCH)) break;
      }
    }
  SPB_PHASE_END
}

// Owner side of the multi-GPU "links" protocol: like k_bkt_insert_links, but the links are appended to a list
// (from << 32 | to) instead of being written in place (the positions belong to other ranks).
__global__ void SPB_LAUNCH_BOUNDS(256)
k_bkt_insert_list(const u64* __restrict__ keys, const u32* __restrict__ pos, u64 lo, u64 n, const u32* __restrict__ F,
                  const u32* __restrict__ R, u64 T, u32 k, u64* scratch, u64 capmask, u64* __restrict__ links, u64* nlinks, u64 lcap,
                  u64* clear, u64 clear_n) {
  u64 t = SPB_GTID();
  bool emit = false;
  u64 link = 0;
  if (t < n) {
    u64 key = keys[lo + t];
    if (key != ~0ull) {
      u64 p = pos[lo + t];
      bool o = key & 1;
      u64 tag = (key >> 1) & 0x7FFFFFull, idx = (key >> 24) & capmask;
      u64 entry = (tag << 40) | (p << 1) | (o ? 1ull : 0ull);
      const u32* S = o ? F : R;
      u64 a = o ? p : T - p - k;
      for (;;) {
        u64 cur = atomicCAS(scratch + idx, SPB_EMPTY, entry);
        if (cur == SPB_EMPTY) break;
        if ((cur >> 40) == tag) {
          u64 q = (cur >> 1) & SPB_POSMASK;
          bool oq = cur & 1;
          if (q == p || win_cmp(S, a, oq ? F : R, oq ? q : T - q - k, k) == 0) {
            u64 from = p, to = q;
            if (entry < cur) {
              u64 old = atomicMin(scratch + idx, entry);
              u64 oq2 = (old >> 1) & SPB_POSMASK;
              if (old > entry) { from = oq2; to = p; } else to = oq2;
            }
            emit = true; link = (from << 32) | to;
            break;
          }
        }
        idx = (idx + 1) & capmask;
      }
    }
  }
  u64 slot = warp_append(nlinks, emit);
  if (emit && slot < lcap) links[slot] = link;
  u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 j = t; j < clear_n; j += stride) clear[j] = SPB_EMPTY;
}
// links sorted by `from`: make every `to` a first appearance (a `to` that is itself the `from` of another link is
// followed; all positions of a chain carry the same k-mer and therefore live at this owner)
__global__ void k_links_resolve(const u64* __restrict__ links, u64 n, u64* __restrict__ out) {
  u64 i = SPB_GTID();
  if (i >= n) return;
  u64 l = links[i];
  u32 to = (u32)l;
  for (;;) {
    u64 lo = 0, hi = n;                               // lower_bound on the `from` part
    while (lo < hi) { u64 mid = (lo + hi) >> 1; if ((links[mid] >> 32) < to) lo = mid + 1; else hi = mid; }
    if (lo < n && (links[lo] >> 32) == to) to = (u32)links[lo]; else break;
  }
  out[i] = (l & 0xFFFFFFFF00000000ull) | to;
}
__global__ void k_links_bounds(const u64* __restrict__ links, u64 n, u64 slice, u32 n_ranks, u64* __restrict__ bounds) {
  u64 r = SPB_GTID();
  if (r > n_ranks) return;
  u64 key = (r * slice) << 32;
  if (r == n_ranks) { bounds[r] = n; return; }
  u64 lo = 0, hi = n;
  while (lo < hi) { u64 mid = (lo + hi) >> 1; if (links[mid] < key) lo = mid + 1; else hi = mid; }
  bounds[r] = lo;
}
// position side: apply the links received for the local slice
__global__ void k_links_apply(const u64* __restrict__ links, u64 n, u32* __restrict__ seenpos, u32* dupbits) {
  u64 i = SPB_GTID();
  if (i >= n) return;
  u32 from = (u32)(links[i] >> 32);
  seenpos[from] = (u32)links[i];
  atomicOr(dupbits + (from >> 5), 1u << (from & 31));
}

// S3: bucket-interleaved scatter back to position order
__global__ void k_unpartition(const u32* __restrict__ pos, const u32* __restrict__ rep, const u64* __restrict__ bstart, u32 B, u32* __restrict__ sp, u32* fbits) {
  u64 g = SPB_GTID();
  u32 b = (u32)(g % B);
  u64 rec = bstart[b] + g / B;
  if (rec >= bstart[b + 1]) return;
  u32 r = rep[rec];
  if (r == SPB_INVALID) return;
  u32 p = pos[rec];
  if (r == p) atomicOr(fbits + (p >> 5), 1u << (p & 31));
  else sp[p] = r;
}
__global__ void k_block_counts(const u32* __restrict__ fbits, u32* __restrict__ cnt, u8* __restrict__ wpre, u64 nblk) {
  u64 b = SPB_GTID();
  if (b >= nblk) return;
  u32 c = 0;
  for (u32 j = 0; j < 8; ++j) { wpre[b * 8 + j] = (u8)c; c += __popc(fbits[b * 8 + j]); }
  cnt[b] = c;
}

#define SPB_JU 4                                     // bitmap words per warp iteration (independent loads in flight)
// The per-position passes (ids, junction entries, origin pairs) walk the first-appearance bitmap in windows of 32 words
// (1024 positions, one word per lane).  A window in which every position is a first appearance -- the normal case on
// mostly-unique inputs -- is handled with ONE memory round trip per window (the passes are latency-bound, not
// bandwidth-bound); any other window goes through the general code, SPB_JU words at a time.
#define SPB_WIN 32
__device__ __forceinline__ bool window_all_first(const u32* __restrict__ fbits, u64 wb, u64 nwords, u32 lane) {
#ifdef SPB_EMU
  (void)lane;
  if (wb + SPB_WIN > nwords) return false;
  for (u32 i = 0; i < SPB_WIN; ++i) if (fbits[wb + i] != 0xFFFFFFFFu) return false;
  return true;
#else
  const u32 fw = wb + lane < nwords ? __ldg(fbits + wb + lane) : 0u;
  return __all_sync(0xffffffffu, fw == 0xFFFFFFFFu);
#endif
}
// first-appearance bitmap fbits = wbits & ~dbits, and per 256-position block counts
__global__ void k_first_bits(const u32* __restrict__ wbits, const u32* __restrict__ dbits, u32* __restrict__ fbits, u64 nblk) {
  u64 b = SPB_GTID();
  if (b >= nblk) return;
  for (u32 j = 0; j < 8; ++j) fbits[b * 8 + j] = wbits[b * 8 + j] & ~dbits[b * 8 + j];
}

// rank of position x among first appearances: block offset (256 positions) + prefix of the words of the block
// + bits below x in its own word
__device__ __forceinline__ u32 rank_first(const u32* fbits, const u32* blkoff, const u8* wpre, u32 x) {
  return __ldg(blkoff + (x >> 8)) + __ldg(wpre + (x >> 5)) + __popc(__ldg(fbits + (x >> 5)) & ((1u << (x & 31)) - 1));
}

// Multi-GPU, sliced tail: a rank fills the per-position / per-vertex arrays only for its own slice.
// The few entries a kernel needs from outside that slice are recomputed from the replicated bitmaps and rank tables:
//   VPos   first position of vertex v        = select(fbits, v): block by binary search in blkoff, word by wpre, bit in the word
//   VidAt  vertex id of the instance at p    = rank_first(first position of its k-mer) (links in `seen`)
// Single GPU: lo = 0, hi = everything -> plain array reads.
struct VPos { const u32* v2pos; u32 vlo, vhi; const u32* fbits; const u32* blkoff; const u8* wpre; u32 nb; };
__device__ __forceinline__ u32 vpos_at(const VPos& q, u32 v) {
  if (v >= q.vlo && v < q.vhi) return q.v2pos[v];
  u32 lo = 0, hi = q.nb;                             // last block b with blkoff[b] <= v (empty blocks share their offset with the next one)
  while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (__ldg(q.blkoff + mid) <= v) lo = mid + 1; else hi = mid; }
  const u32 b = lo - 1;
  const u32 r = v - __ldg(q.blkoff + b);
  u32 j = 7;
  while (j > 0 && __ldg(q.wpre + 8ull * b + j) > r) --j;
  u32 w = __ldg(q.fbits + 8ull * b + j);
  for (u32 i = r - __ldg(q.wpre + 8ull * b + j); i > 0; --i) w &= w - 1;
#ifdef SPB_EMU
  return (u32)(256ull * b + 32 * j + (u32)__builtin_ctz(w));
#else
  return (u32)(256ull * b + 32 * j + (u32)(__ffs((int)w) - 1));
#endif
}
struct VidAt { const u32* sp; u64 lo, hi; const u32* fbits; const u32* seen; const u32* blkoff; const u8* wpre; };
__device__ __forceinline__ u32 vid_at(const VidAt& q, u64 p) {     // p must start a k-mer
  if (p >= q.lo && p < q.hi) return q.sp[p];
  const u32 r = get_bit(q.fbits, p) ? (u32)p : q.seen[p];
  return rank_first(q.fbits, q.blkoff, q.wpre, r);
}

// exact lookup of the first position of the k-mer at p in the finished table (rarely needed, see k_assign_ids)
template <int W> __device__ __forceinline__ u32 lookup_rep(const u32* __restrict__ F, const u32* __restrict__ R, u64 T, u32 k,
                                                           const u64* __restrict__ slots, u64 capmask, u64 p) {
  u64 ra = T - p - k;
  bool o = win_cmp(F, p, R, ra, k) >= 0;
  const u32* S = o ? F : R;
  u64 a = o ? p : ra;
  u64 h = win_hash<W>(S, a, k);
  u64 idx = h & capmask, tag = h >> 40;
  for (;;) {
    u64 cur = __ldcg(slots + idx);
    if (cur == SPB_EMPTY) return (u32)p;             // cannot happen: every instance was inserted
    if ((cur >> 40) == tag) {
      u64 q = (cur >> 1) & SPB_POSMASK;
      bool oq = cur & 1;
      if (q == p || win_cmp(S, a, oq ? F : R, oq ? q : T - q - k, k) == 0) return (u32)q;
    }
    idx = (idx + 1) & capmask;
  }
}

// K3: vertex ids (= rank of the representative among first appearances), vertex -> first position,
// implicit-edge bit.  On entry sp[p] holds, for an instance that is not a first appearance, the position it
// matched during the insert (global-table path; exact unless that position was displaced later -> table lookup)
// or its exact representative (staged / multi-GPU path, slots == nullptr).  On exit sp[p] = vertex id.
template <int W> __global__ void
k_assign_ids(const u32* sp_in, u32* sp, const u32* __restrict__ vbits, u64 T, u64 lo, u64 hi, const u32* __restrict__ fbits,
             const u32* __restrict__ blkoff, const u8* __restrict__ wpre, const u64* __restrict__ slots, u64 capmask, const u32* __restrict__ F,
             const u32* __restrict__ R, u32 k, u32* __restrict__ v2pos, u8* __restrict__ vflags, int follow, u32* err, u64 tlo, u64 thi) {
  // [tlo, thi): positions whose vertex -> position entries are written (everything on one GPU; the slice plus its halo on
  // the sliced multi-GPU tail); the flag byte of every vertex is written everywhere.
  // One warp per SPB_JU x 32 positions (SPB_JU words of the bitmaps), grid-stride.  The dependent loads (bitmap word ->
  // link -> bitmap word at the link target -> rank tables) are issued for all SPB_JU words before they are consumed.
  //   sp_in[p] = a smaller position with the same k-mer (global-table path: follow the links down to a first
  //   appearance; a displaced winner on the way ends in an exact table lookup), a pure link (links path: follow) or
  //   the exact representative (multi-GPU path)
  const u64 nwords = (T + 31) / 32;
  const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
  const u32 lane = threadIdx.x & 31;
  for (u64 wb = (SPB_GTID() >> 5) * SPB_WIN; wb < nwords; wb += nwarps * SPB_WIN) {
   {
    // ---- window of 1024 consecutive first appearances: vertex vb + i at position g0 + i.  The rank of the window and the
    // bit that follows it are loaded together with the bitmap words (independent addresses: one round trip).
    const u32 vb_w = __ldg(blkoff + (wb >> 3)) + __ldg(wpre + wb);
    const u32 nextbit = wb + SPB_WIN <= nwords ? (__ldg(fbits + wb + SPB_WIN) & 1u) : 0u;
    if (window_all_first(fbits, wb, nwords, lane)) {
      const u64 g0 = wb * 32, g1 = g0 + 32 * SPB_WIN;
      if (g0 >= lo && g1 <= hi) {
#pragma unroll 8
        for (u32 j = 0; j < SPB_WIN; ++j) sp[g0 + 32 * j + lane] = vb_w + 32 * j + lane;
      } else if (g0 < hi && g1 > lo) {
        for (u32 j = 0; j < SPB_WIN; ++j) { const u64 p = g0 + 32 * j + lane; if (p >= lo && p < hi) sp[p] = vb_w + 32 * j + lane; }
      }
      if (g0 >= tlo && g1 <= thi) {
#pragma unroll 8
        for (u32 j = 0; j < SPB_WIN; ++j) v2pos[vb_w + 32 * j + lane] = (u32)(g0 + 32 * j + lane);
      } else if (g0 < thi && g1 > tlo) {
        for (u32 j = 0; j < SPB_WIN; ++j) { const u64 p = g0 + 32 * j + lane; if (p >= tlo && p < thi) v2pos[vb_w + 32 * j + lane] = (u32)p; }
      }
      // flag bytes [vb, vb + 1024): VF_R everywhere except possibly the last one; aligned 4-byte stores for the middle
      const u32 ve = vb_w + 32 * SPB_WIN;
      const u8 lastf = nextbit ? VF_R : 0;
      const u32 a0 = (vb_w + 3) & ~3u, a1 = ve & ~3u;
#pragma unroll 8
      for (u32 j = 0; j < SPB_WIN / 4; ++j) {
        const u32 wv = a0 + 4 * (32 * j + lane);
        if (wv + 4 <= a1) *(u32*)(vflags + wv) = (wv + 4 == ve) ? (0x00010101u | ((u32)lastf << 24)) : 0x01010101u;
      }
      if (lane < a0 - vb_w) vflags[vb_w + lane] = VF_R;
      if (lane < ve - a1) vflags[a1 + lane] = (a1 + lane == ve - 1) ? lastf : (u8)VF_R;
      continue;
    }
   }
   for (u64 wd0 = wb; wd0 < wb + SPB_WIN && wd0 < nwords; wd0 += SPB_JU) {
    u32 f[SPB_JU + 1], r[SPB_JU], fr[SPB_JU];
    bool first[SPB_JU], mine[SPB_JU], act[SPB_JU];
#pragma unroll
    for (u32 j = 0; j <= SPB_JU; ++j) f[j] = wd0 + j <= nwords ? __ldg(fbits + wd0 + j) : 0u;
    bool allfirst = wd0 + SPB_JU <= nwords;
#pragma unroll
    for (u32 j = 0; j < SPB_JU; ++j) allfirst = allfirst && f[j] == 0xFFFFFFFFu;
    if (allfirst) {                                  // warp-uniform: 128 first appearances in a row, ids are consecutive
      const u64 g0 = wd0 * 32, g1 = g0 + 32 * SPB_JU;
      const u32 vb = __ldg(blkoff + (wd0 >> 3)) + __ldg(wpre + wd0);     // vertex of position g0; position g0 + i is vertex vb + i
      // (all tests on the group's range are warp-uniform: nothing but the flag bytes is written for a group outside the
      // caller's position ranges, and a group inside them needs no per-lane test)
      if (g0 >= lo && g1 <= hi) {
#pragma unroll
        for (u32 j = 0; j < SPB_JU; ++j) sp[g0 + 32 * j + lane] = vb + 32 * j + lane;
      } else if (g0 < hi && g1 > lo) {
#pragma unroll
        for (u32 j = 0; j < SPB_JU; ++j) { const u64 p = g0 + 32 * j + lane; if (p >= lo && p < hi) sp[p] = vb + 32 * j + lane; }
      }
      if (g0 >= tlo && g1 <= thi) {
#pragma unroll
        for (u32 j = 0; j < SPB_JU; ++j) v2pos[vb + 32 * j + lane] = (u32)(g0 + 32 * j + lane);
      } else if (g0 < thi && g1 > tlo) {
#pragma unroll
        for (u32 j = 0; j < SPB_JU; ++j) { const u64 p = g0 + 32 * j + lane; if (p >= tlo && p < thi) v2pos[vb + 32 * j + lane] = (u32)p; }
      }
      // flag bytes of the 32 * SPB_JU consecutive vertices [vb, vb + 128): all VF_R except possibly the last one.  One
      // aligned 4-byte store per lane for the aligned middle part (full sectors), byte stores for the ragged ends.
      {
        const u32 ve = vb + 32 * SPB_JU;
        const u8 lastf = (f[SPB_JU] & 1u) ? VF_R : 0;
        const u32 a0 = (vb + 3) & ~3u, a1 = ve & ~3u;             // aligned words [a0, a1)
        const u32 wv = a0 + 4 * lane;
        if (wv + 4 <= a1) *(u32*)(vflags + wv) = (wv + 4 == ve) ? (0x00010101u | ((u32)lastf << 24)) : 0x01010101u;
        if (lane < a0 - vb) vflags[vb + lane] = VF_R;             // head (a0 - vb < 4 <= 128: never the last vertex)
        if (lane < ve - a1) vflags[a1 + lane] = (a1 + lane == ve - 1) ? lastf : (u8)VF_R;
      }
      continue;
    }
#pragma unroll
    for (u32 j = 0; j < SPB_JU; ++j) {               // stage 1: the link of every repeated instance
      const u64 wd = wd0 + j, p = wd * 32 + lane;
      const bool inside = wd < nwords && p < T;
      mine[j] = inside && p >= lo && p < hi;         // multi-GPU: sp is filled for the local slice only, the vertex table everywhere
      const bool vb = inside && ((__ldg(vbits + wd) >> lane) & 1u);
      first[j] = (f[j] >> lane) & 1u;
      act[j] = vb && (first[j] || mine[j]);
      if (inside && !vb && mine[j]) sp[p] = SPB_INVALID;
      r[j] = (u32)p;
      if (act[j] && !first[j]) r[j] = sp_in[p];
    }
#pragma unroll
    for (u32 j = 0; j < SPB_JU; ++j) {               // stage 2: bitmap word at the link target
      fr[j] = f[j];
      if (act[j] && !first[j] && r[j] != SPB_INVALID) fr[j] = __ldg(fbits + (r[j] >> 5));
    }
    u32 bo[SPB_JU], wp[SPB_JU];
#pragma unroll
    for (u32 j = 0; j < SPB_JU; ++j) {               // stage 3: finish the (rare) longer chains; rank tables
      bo[j] = wp[j] = 0;
      if (!act[j]) continue;
      if (!first[j]) {
        u32 x = r[j], fx = fr[j];
        if (slots) {
          for (u32 hop = 0;; ++hop) {
            if (x == SPB_INVALID || hop > 256) { x = lookup_rep<W>(F, R, T, k, slots, capmask, wd0 * 32 + 32ull * j + lane); fx = __ldg(fbits + (x >> 5)); break; }
            if ((fx >> (x & 31)) & 1u) break;
            x = sp_in[x];
            if (x != SPB_INVALID) fx = __ldg(fbits + (x >> 5));
          }
        } else if (follow) {                         // links without a table: chains strictly descend and end at a first appearance
          while (!((fx >> (x & 31)) & 1u)) {
            const u32 nx = sp_in[x];
            if (nx >= x) { atomicOr(err, (u32)ERR_OVERFLOW); break; }
            x = nx; fx = __ldg(fbits + (x >> 5));
          }
        }
        r[j] = x; fr[j] = fx;
      }
      bo[j] = __ldg(blkoff + (r[j] >> 8)); wp[j] = __ldg(wpre + (r[j] >> 5));
    }
#pragma unroll
    for (u32 j = 0; j < SPB_JU; ++j) {               // stage 4: ids
      if (!act[j]) continue;
      const u64 p = (wd0 + j) * 32 + lane;
      const u32 vid = bo[j] + wp[j] + __popc(fr[j] & ((1u << (r[j] & 31)) - 1));
      if (mine[j]) sp[p] = vid;
      if (first[j]) {
        if (p >= tlo && p < thi) v2pos[vid] = (u32)p;
        const u32 nextfirst = lane < 31 ? (f[j] >> (lane + 1)) & 1u : f[j + 1] & 1u;
        vflags[vid] = nextfirst ? VF_R : 0;          // (p, p+1) both first appearances => edge (vid, vid+1)
      }
    }
   }
  }
}

// K4: junction candidates.  The instance pair (p, p+1) yields edge {sp[p], sp[p+1]} (reference
// lib/Assembler.py:199-201).  It is dropped when it coincides with an implicit edge (v, v+1).
__global__ void k_junction_emit(const u32* __restrict__ sp, u64 T, u64 w_lo, u64 w_hi, const u32* __restrict__ fbits, const u8* __restrict__ vflags, VPos vp,
                                const u32* __restrict__ obits, const u32* __restrict__ vobits, u64* __restrict__ Jk, u8* __restrict__ Js,
                                u64* counter, u64 cap, u32* err) {
  // [w_lo, w_hi): bitmap words whose pairs (p, p + 1) are ours (sp must be filled up to position 32 * w_hi)
  // One warp per SPB_JU x 32 positions (SPB_JU words of the first-appearance bitmap), grid-stride.  A word whose 32
  // pairs (p, p+1) are all pairs of consecutive first appearances holds only implicit edges and is skipped at once.
  // The three dependent loads of a pair (bitmap -> vertex ids -> flags / orientation bits) are issued for all SPB_JU
  // words before any of them is consumed: the kernel is bound by that latency chain, not by bandwidth.
  const u64 nwords = (T + 31) / 32;
  const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
  const u32 lane = threadIdx.x & 31;
  for (u64 wb = w_lo + (SPB_GTID() >> 5) * SPB_WIN; wb < w_hi; wb += nwarps * SPB_WIN) {
   {
    // window of 1024 first appearances followed by one more: nothing but implicit edges
    const u32 nextbit = wb + SPB_WIN <= nwords ? (__ldg(fbits + wb + SPB_WIN) & 1u) : 0u;
    if (wb + SPB_WIN <= w_hi && window_all_first(fbits, wb, nwords, lane) && nextbit) continue;
   }
   for (u64 wd0 = wb; wd0 < wb + SPB_WIN && wd0 < w_hi; wd0 += SPB_JU) {
    u32 f[SPB_JU + 1];
#pragma unroll
    for (u32 j = 0; j <= SPB_JU; ++j) f[j] = wd0 + j <= nwords ? __ldg(fbits + wd0 + j) : 0u;
    u32 u[SPB_JU], w[SPB_JU], ob[SPB_JU], on[SPB_JU];
    bool live[SPB_JU];
    bool any = false;
#pragma unroll
    for (u32 j = 0; j < SPB_JU; ++j) any = any || (wd0 + j < w_hi && (f[j] & ((f[j] >> 1) | (f[j + 1] << 31))) != 0xFFFFFFFFu);
    if (!any) continue;                              // warp-uniform: nothing but consecutive first appearances here
#pragma unroll
    for (u32 j = 0; j < SPB_JU; ++j) {               // stage 1: vertex ids of the pairs that are not two consecutive firsts
      const u64 p = (wd0 + j) * 32 + lane;
      const bool word = wd0 + j < w_hi && (f[j] & ((f[j] >> 1) | (f[j + 1] << 31))) != 0xFFFFFFFFu;     // warp-uniform
      const u32 fp = (f[j] >> lane) & 1u, fq = lane < 31 ? (f[j] >> (lane + 1)) & 1u : f[j + 1] & 1u;
      live[j] = word && p + 1 < T && !(fp && fq);
      u[j] = live[j] ? sp[p] : SPB_INVALID; w[j] = live[j] ? sp[p + 1] : SPB_INVALID;
      ob[j] = word ? __ldg(obits + wd0 + j) : 0u; on[j] = word ? __ldg(obits + wd0 + j + 1) : 0u;
    }
    u8 fl[SPB_JU]; u32 ou[SPB_JU], ow[SPB_JU];
#pragma unroll
    for (u32 j = 0; j < SPB_JU; ++j) {               // stage 2: flags + stored orientations of candidate implicit edges
      fl[j] = 0; ou[j] = ow[j] = 0;
      live[j] = live[j] && u[j] != SPB_INVALID && w[j] != SPB_INVALID;
      if (live[j] && (w[j] == u[j] + 1 || u[j] == w[j] + 1)) {
        fl[j] = vflags[w[j] == u[j] + 1 ? u[j] : w[j]];
        ou[j] = vobits ? get_bit(vobits, u[j]) : get_bit(obits, vpos_at(vp, u[j]));
        ow[j] = vobits ? get_bit(vobits, w[j]) : get_bit(obits, vpos_at(vp, w[j]));
      }
    }
#pragma unroll
    for (u32 j = 0; j < SPB_JU; ++j) {               // stage 3: check implicit duplicates, emit junction edges
      const u32 op = (ob[j] >> lane) & 1u, oq = lane < 31 ? (ob[j] >> (lane + 1)) & 1u : on[j] & 1u;    // orientation at p, p+1
      u32 n = 0;
      if (live[j]) {
        const bool up = w[j] == u[j] + 1, down = u[j] == w[j] + 1;
        if ((up || down) && (fl[j] & VF_R)) {
          // a later instance of an implicit edge must attach on the same sides as the first one (both instances on
          // the stored strand when walked upwards, both on the other strand when walked downwards); otherwise the
          // edge is orientation-ambiguous, like a junction edge seen with two different payloads (k_unique_heads)
          const u32 su = op ^ ou[j], sw = oq ^ ow[j];
          if (su != sw || su != (down ? 1u : 0u)) atomicOr(err, (u32)ERR_DEGENERATE);
        } else n = (u[j] == w[j]) ? 1 : 2;
      }
#ifndef SPB_EMU
      if (__ballot_sync(0xffffffffu, n != 0) == 0) continue;
#endif
      u64 slot = list_reserve(counter, n);
      if (n == 0 || slot + n > cap) continue;
      u32 su = op ^ get_bit(obits, vpos_at(vp, u[j]));                   // instance strand vs. vertex' stored k-mer
      u32 sw = oq ^ get_bit(obits, vpos_at(vp, w[j]));
      u32 side_u = su ? SIDE_L : SIDE_R;                           // where w attaches on u
      u32 side_w = sw ? SIDE_R : SIDE_L;                           // where u attaches on w
      Jk[slot] = ((u64)u[j] << 32) | w[j]; Js[slot] = (u8)(side_u | (side_w << 1));
      if (n == 2) { Jk[slot + 1] = ((u64)w[j] << 32) | u[j]; Js[slot + 1] = (u8)(side_w | (side_u << 1)); }
    }
   }
  }
}

// orientation bit of every vertex' stored k-mer (= of its first appearance); v2pos ascends with the vertex id, so the
// gathers are nearly sequential.  One warp per 32 vertices, grid-stride.
__global__ void k_vertex_obits(const u32* __restrict__ v2pos, const u32* __restrict__ obits, u32 nV, u32* __restrict__ vobits) {
  const u64 nwords = ((u64)nV + 31) / 32;
  const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
  const u32 lane = threadIdx.x & 31;
  for (u64 wd = SPB_GTID() >> 5; wd < nwords; wd += nwarps) {
    const u64 v = wd * 32 + lane;
    const bool b = v < nV && get_bit(obits, v2pos[v]);
#ifdef SPB_EMU
    if (b) vobits[wd] |= 1u << lane;                 // the caller zeroes the array
#else
    const u32 m = __ballot_sync(0xffffffffu, b);
    if (lane == 0) vobits[wd] = m;
#endif
  }
}

// heads of runs of equal keys in a sorted u64 array; optionally check that the payload agrees
__global__ void k_unique_heads(const u64* __restrict__ keys, const u8* __restrict__ pay, u64 n, u32* __restrict__ head, u32* err) {
  u64 i = SPB_GTID();
  if (i >= n) return;
  bool h = (i == 0) || keys[i] != keys[i - 1];
  head[i] = h ? 1u : 0u;
  // the same edge reached with different attachment sides is orientation-ambiguous (the reference asserts when it walks
  // there); self-loops are exempt: their sides are never used (lib/Assembler.py:1013, :334)
  if (!h && pay && pay[i] != pay[i - 1] && (u32)(keys[i] >> 32) != (u32)keys[i]) atomicOr(err, (u32)ERR_DEGENERATE);
}
__global__ void k_unique_scatter(const u64* __restrict__ keys, const u8* __restrict__ pay, u64 n, const u32* __restrict__ head,
                                 const u32* __restrict__ pos, u64* __restrict__ okeys, u8* __restrict__ opay) {
  u64 i = SPB_GTID();
  if (i >= n || !head[i]) return;
  okeys[pos[i]] = keys[i];
  if (pay) opay[pos[i]] = pay[i];
}

// ------------------------------------------------------------------------------------ graph view
struct GView {
  const u8* vflags; const u64* Jk; const u8* Js; u32 nJ; u32 nV;
};
__device__ __forceinline__ u32 lower_bound_key(const u64* a, u32 n, u64 x) {  // first i with a[i] >= x
  u32 lo = 0, hi = n;
  while (lo < hi) { u32 mid = (lo + hi) >> 1; if (__ldg(a + mid) < x) lo = mid + 1; else hi = mid; }
  return lo;
}
__device__ __forceinline__ u32 upper_bound_u32(const u32* a, u32 n, u32 x) {  // first i with a[i] > x
  u32 lo = 0, hi = n;
  while (lo < hi) { u32 mid = (lo + hi) >> 1; if (__ldg(a + mid) <= x) lo = mid + 1; else hi = mid; }
  return lo;
}
// f(nbr, side_of_nbr_on_v, side_of_v_on_nbr) for every distinct non-self neighbour of v, in the
// fixed order: v-1 (implicit), v+1 (implicit), junction neighbours ascending.
template <class Fn> __device__ __forceinline__ void for_each_nbr(const GView& g, u32 v, Fn f) {
  u8 fl = g.vflags[v];
  if (v > 0 && (g.vflags[v - 1] & VF_R)) f(v - 1, (u32)SIDE_L, (u32)SIDE_R);
  if (fl & VF_R) f(v + 1, (u32)SIDE_R, (u32)SIDE_L);
  if (fl & VF_HASJ) {
    u32 i = lower_bound_key(g.Jk, g.nJ, (u64)v << 32);
    for (; i < g.nJ && (u32)(g.Jk[i] >> 32) == v; ++i) {
      u32 d = (u32)g.Jk[i]; u32 s = g.Js[i];
      if (d != v && !(s & JS_CUT)) f(d, s & 1u, (s >> 1) & 1u);
    }
  }
}

__global__ void k_mark_hasj(const u64* __restrict__ Jk, u32 nJ, u8* vflags) {
  u64 i = SPB_GTID();
  if (i >= nJ) return;
  u32 s = (u32)(Jk[i] >> 32);
  if (i > 0 && (u32)(Jk[i - 1] >> 32) == s) return;
  atomicOr((u32*)vflags + (s >> 2), (u32)VF_HASJ << ((s & 3) * 8));
}
__global__ void k_mark_seqlim(const u64* __restrict__ seq_off, u32 n_seq, u32 k, VidAt va, u8* vflags, u32* seqlim) {
  u64 s = SPB_GTID();
  if (s >= n_seq) return;
  u64 b = seq_off[s], e = seq_off[s + 1];
  u32 v0 = SPB_INVALID, v1 = SPB_INVALID;
  if (e - b > k) {
    v0 = vid_at(va, b); v1 = vid_at(va, e - k);
    atomicOr((u32*)vflags + (v0 >> 2), (u32)VF_SEQLIM << ((v0 & 3) * 8));
    atomicOr((u32*)vflags + (v1 >> 2), (u32)VF_SEQLIM << ((v1 & 3) * 8));
  }
  seqlim[2 * s] = v0; seqlim[2 * s + 1] = v1;
}

// K5: init detection (reference lib/Assembler.py:995-1035): init iff #distinct non-self
// neighbours != 2, or the vertex is a sequence limit and both neighbours attach to the same side.
// Eight vertices (one u64 of flag bytes) per thread.  A vertex without junction entries has only the
// implicit neighbours v-1 (left) and v+1 (right): init <=> not both present; the general rule runs
// only for the few vertices that carry junction entries.
__device__ __forceinline__ bool classify_slow(const GView& g, u32 v, u8 f, u32* err) {
  u32 cnt = 0, s0 = 0, s1 = 0;
  for_each_nbr(g, v, [&](u32, u32 side, u32) { if (cnt == 0) s0 = side; else if (cnt == 1) s1 = side; ++cnt; });
  bool init = cnt != 2;
  if (!init && s0 == s1) {
    init = true;
    if (!(f & VF_SEQLIM) && err) atomicOr(err, (u32)ERR_DEGENERATE);  // the reference would assert (lib/Assembler.py:1104)
  }
  return init;
}
#define SPB_M8 0x0101010101010101ull

// K5 + K6 fused, 32 vertices (4 flag words = one 32-byte sector) per thread: init bits, then pieces = maximal runs of
// consecutive-id extender vertices linked by implicit edges, and the per-thread counts of inits / piece starts / piece
// ends for the compaction below.  The init bits of the two vertices next to the thread's range are recomputed from the
// neighbouring words (they depend on the stable bits R / HASJ / SEQLIM / CYC and on J only), so ONE pass over the flag bytes
// replaces the classification pass + the piece pass, and the count arrays are 4 times shorter than with one word per thread.
#define SPB_CW 4                                     // flag words per thread
__device__ __forceinline__ u64 classify_word(const GView& g, u64 v0, u64 w, u64 prev_byte, u32* err, bool report) {
  const u64 nvalid = g.nV - v0 < 8 ? g.nV - v0 : 8;
  const u64 vm = nvalid == 8 ? ~0ull : ((1ull << (8 * nvalid)) - 1);
  const u64 r = w & SPB_M8, rprev = (r << 8) | (prev_byte & VF_R);
  const u64 hasj = (w >> 1) & SPB_M8;
  const u64 init = ~(r & rprev) & SPB_M8 & ~hasj;    // no junction entries: init <=> not both implicit neighbours
  u64 out = (w & ~(SPB_M8 * (u64)(VF_INIT | VF_PSTART | VF_PEND | VF_IN2))) | ((init & vm) << 3);
  u64 slow = hasj & vm;
  while (slow) {                                      // the few vertices with junction entries: general rule
#ifdef SPB_EMU
    const u32 i = (u32)(__builtin_ctzll(slow) >> 3);
#else
    const u32 i = (u32)((__ffsll((long long)slow) - 1) >> 3);
#endif
    slow &= slow - 1;
    if (classify_slow(g, (u32)(v0 + i), (u8)(w >> (8 * i)), report ? err : nullptr)) out |= (u64)VF_INIT << (8 * i);
  }
  return out;
}
// Flags are written for the groups [f_lo, f_hi) of 32 vertices; the counts cover the groups [t_lo, t_hi) inside them,
// nT = t_hi - t_lo (the sliced multi-GPU tail also flags one group on either side of its own range: the per-position passes
// of the slice read the piece flags of vertices up to 31 ids beyond the 32-aligned cut).
__global__ void k_classify_pieces(GView g, u8* vflags, u32* __restrict__ cnt3, u32 nT, u32 f_lo, u32 f_hi, u32 t_lo, u32 t_hi, u32* err) {
  const u64 t = SPB_GTID() + f_lo;
  const u64 v0 = t * 8 * SPB_CW;
  if (t >= f_hi || v0 >= g.nV) return;
  const u64 nVw = ((u64)g.nV + 7) / 8;               // flag words that hold vertices (the array is padded with zero bytes)
  const u64 w0 = v0 >> 3;
  // words w0 - 1 .. w0 + SPB_CW (neighbours included), classified from their stable bits
  u64 c[SPB_CW + 2];
#pragma unroll
  for (int i = -1; i <= SPB_CW; ++i) {
    const long long wi = (long long)w0 + i;
    u64 out = 0;
    if (wi >= 0 && (u64)wi < nVw) {
      const u64 w = *(const u64*)(g.vflags + 8 * wi);
      const u64 pb = wi ? g.vflags[8 * wi - 1] : 0;
      out = classify_word(g, 8ull * wi, w, pb, err, i >= 0 && i < SPB_CW);
    }
    c[i + 1] = out;
  }
  u32 nI = 0, nS = 0, nE = 0;
#pragma unroll
  for (int i = 0; i < SPB_CW; ++i) {
    const u64 wi = w0 + i;
    if (wi >= nVw) break;
    const u64 w = c[i + 1], pf = c[i] >> 56, nf = c[i + 2] & 0xFF;
    const u64 vb = 8 * wi;
    const u64 nvalid = g.nV - vb < 8 ? g.nV - vb : 8;
    const u64 vm = nvalid == 8 ? ~0ull : ((1ull << (8 * nvalid)) - 1);
    const u64 init = (w >> 3) & SPB_M8 & vm, r = w & SPB_M8 & vm;
    const u64 ext = ~init & SPB_M8 & vm;
    const u64 p_r = (r << 8) | (pf & 1), p_init = (init << 8) | ((pf >> 3) & 1);
    const u64 n_init = (init >> 8) | (((nf >> 3) & 1) << 56);
    const u64 start = ext & ~(p_r & ~p_init);           // no extender predecessor through an implicit edge
    const u64 end = ext & ~(r & ~n_init);               // no extender successor through an implicit edge
    *(u64*)(vflags + vb) = w | (start << 5) | (end << 6);
    nI += (u32)__popcll(init); nS += (u32)__popcll(start); nE += (u32)__popcll(end);
  }
  if (t < t_lo || t >= t_hi) return;
  const u64 i3 = t - t_lo;
  cnt3[i3] = nI; cnt3[nT + i3] = nS; cnt3[2 * (u64)nT + i3] = nE;
}
__global__ void k_compact3(const u8* __restrict__ vflags, u32 nV, const u32* __restrict__ off3, u32 nT, u32 totI, u32 totS,
                           u32* __restrict__ init_list, u32* __restrict__ plo, u32* __restrict__ phi, u32 t_lo) {
  const u64 i3 = SPB_GTID();
  const u64 v0 = (i3 + t_lo) * 8 * SPB_CW;
  if (i3 >= nT || v0 >= nV) return;
  u64 w[SPB_CW]; u64 any = 0;
#pragma unroll
  for (u32 i = 0; i < SPB_CW; ++i) { w[i] = v0 + 8 * i < nV ? *(const u64*)(vflags + v0 + 8 * i) : 0ull; any |= w[i]; }
  if (!(any & (SPB_M8 * (u64)(VF_INIT | VF_PSTART | VF_PEND)))) return;     // nothing to list in this group (the usual case): no offsets needed
  u32 oI = off3[i3], oS = off3[nT + i3] - totI, oE = off3[2 * (u64)nT + i3] - totI - totS;
  for (u32 x = 0; x < 8 * SPB_CW && v0 + x < nV; ++x) {
    const u32 f = (u32)(w[x >> 3] >> (8 * (x & 7))) & 0xFF;
    if (f & VF_INIT) init_list[oI++] = (u32)(v0 + x);
    if (f & VF_PSTART) plo[oS++] = (u32)(v0 + x);
    if (f & VF_PEND) phi[oE++] = (u32)(v0 + x);
  }
}

// sliced multi-GPU tail: the init vertices found by the other ranks (the init bit is read at arbitrary vertices below)
__global__ void k_mark_inits(const u32* __restrict__ init_list, u32 nI, u8* vflags) {
  u64 i = SPB_GTID();
  if (i >= nI) return;
  const u32 v = init_list[i];
  atomicOr((u32*)vflags + (v >> 2), (u32)VF_INIT << ((v & 3) * 8));
}
__global__ void k_init_degrees(GView g, const u32* __restrict__ init_list, u32 nI, u32* __restrict__ ideg) {
  u64 i = SPB_GTID();
  if (i >= nI) return;
  u32 c = 0;
  for_each_nbr(g, init_list[i], [&](u32, u32, u32) { ++c; });
  ideg[i] = c;
}

// outside neighbours of the two ends of every piece: the neighbour on the LEFT of lo, on the RIGHT of hi
__global__ void k_piece_ends(GView g, const u32* __restrict__ plo, const u32* __restrict__ phi, u32 nP,
                             u32* __restrict__ pl_nbr, u8* __restrict__ pl_side, u32* __restrict__ pr_nbr, u8* __restrict__ pr_side, u32* err) {
  u64 P = SPB_GTID();
  if (P >= nP) return;
  u32 ln = SPB_INVALID, rn = SPB_INVALID, ls = 0, rs = 0;
  for_each_nbr(g, plo[P], [&](u32 n, u32 side, u32 oside) { if (side == SIDE_L) { ln = n; ls = oside; } });
  for_each_nbr(g, phi[P], [&](u32 n, u32 side, u32 oside) { if (side == SIDE_R) { rn = n; rs = oside; } });
  if (ln == SPB_INVALID || rn == SPB_INVALID) atomicOr(err, (u32)ERR_DEGENERATE);
  pl_nbr[P] = ln; pl_side[P] = (u8)ls; pr_nbr[P] = rn; pr_side[P] = (u8)rs;
}

__device__ __forceinline__ u32 piece_of(const u32* plo, u32 nP, u32 v) { return upper_bound_u32(plo, nP, v) - 1; }

// K7: darts.  d = 2P + dir; dir 0 = "up" (enter at lo from its left, leave hi to its right, vertices in
// forward orientation), dir 1 = "down".  nxt[d] = following dart or SPB_TERM when the walk reaches an init.
__global__ void k_dart_next(const u8* __restrict__ vflags, const u32* __restrict__ plo, u32 nP,
                            const u32* __restrict__ pl_nbr, const u8* __restrict__ pl_side, const u32* __restrict__ pr_nbr,
                            const u8* __restrict__ pr_side, u32* __restrict__ nxt) {
  u64 d = SPB_GTID();
  if (d >= 2ull * nP) return;
  u32 P = (u32)(d >> 1); bool up = !(d & 1);
  u32 y = up ? pr_nbr[P] : pl_nbr[P];
  u32 sy = up ? pr_side[P] : pl_side[P];          // side of y on which we arrive
  if (y == SPB_INVALID || (vflags[y] & VF_INIT)) { nxt[d] = SPB_TERM; return; }
  u32 Q = piece_of(plo, nP, y);
  nxt[d] = 2 * Q + (sy == SIDE_L ? 0u : 1u);       // arriving on y's left => traverse y forward
}

// list ranking over the predecessor links (synchronous pointer jumping, ping-pong buffers).
//   ptr[d]  = an ancestor dart, or SPB_TERM once the head of the list is known
//   acc[d]  = #vertices from that ancestor's first vertex to d's first vertex; absolute rank when resolved
//   root[d] = head dart (valid when resolved);  mnv[d] = min vertex id seen (used to break pure cycles)
__global__ void k_rank_init(const u32* __restrict__ nxt, const u32* __restrict__ plo, const u32* __restrict__ phi, u32 nD,
                            u32* __restrict__ ptr, u32* __restrict__ acc, u32* __restrict__ root, u32* __restrict__ mnv) {
  u64 d = SPB_GTID();
  if (d >= nD) return;
  u32 b = nxt[d ^ 1];
  if (b == SPB_TERM) { ptr[d] = SPB_TERM; acc[d] = 1; root[d] = (u32)d; }
  else { u32 pb = b ^ 1; ptr[d] = pb; acc[d] = phi[pb >> 1] - plo[pb >> 1] + 1; root[d] = SPB_TERM; }
  mnv[d] = plo[d >> 1];
}
__global__ void k_rank_jump(u32 nD, const u32* __restrict__ ptr, const u32* __restrict__ acc, const u32* __restrict__ root,
                            const u32* __restrict__ mnv, u32* __restrict__ optr, u32* __restrict__ oacc, u32* __restrict__ oroot,
                            u32* __restrict__ omnv) {
  u64 d = SPB_GTID();
  if (d >= nD) return;
  u32 b = ptr[d];
  if (b == SPB_TERM) { optr[d] = SPB_TERM; oacc[d] = acc[d]; oroot[d] = root[d]; omnv[d] = mnv[d]; return; }
  optr[d] = ptr[b]; oacc[d] = acc[d] + acc[b]; oroot[d] = root[b];
  u32 m = mnv[d], mb = mnv[b];
  omnv[d] = m < mb ? m : mb;
}
__global__ void k_count_unresolved(const u32* __restrict__ ptr, u32 nD, u32* cnt) {
  u64 d = SPB_GTID();
  if (d >= nD) return;
  if (ptr[d] != SPB_TERM) atomicAdd(cnt, 1u);
}

// Pure-cycle components (reference lib/Assembler.py:304-325): a = lowest vertex id of the ring,
// n0 = its lower-id neighbour; both become inits and the edge (a, n0) is cut.
__global__ void k_cycle_cut(GView g, u8* vflags, u8* Js, const u32* __restrict__ ptr, const u32* __restrict__ mnv,
                            const u32* __restrict__ plo, u32 nD, u32* ncyc) {
  u64 d = SPB_GTID();
  if (d >= nD || (d & 1)) return;              // one representative: the up-dart of the piece that starts at the minimum
  if (ptr[d] == SPB_TERM) return;
  u32 a = plo[d >> 1];
  if (mnv[d] != a) return;
  u32 n0 = SPB_INVALID;
  for_each_nbr(g, a, [&](u32 n, u32, u32) { if (n < n0) n0 = n; });
  atomicAdd(ncyc, 1u);
  // mark both ends as cycle inits
  atomicOr((u32*)vflags + (a >> 2), (u32)(VF_CYC) << ((a & 3) * 8));
  atomicOr((u32*)vflags + (n0 >> 2), (u32)(VF_CYC) << ((n0 & 3) * 8));
  if (n0 == a + 1 && (g.vflags[a] & VF_R)) {
    // implicit edge: clear the R bit (atomic AND via CAS-free trick: only this thread touches bit R of a)
    u32* wp = (u32*)vflags + (a >> 2);
    u32 bit = (u32)VF_R << ((a & 3) * 8);
    u32 old = *wp;
    while (true) { u32 prev = atomicCAS(wp, old, old & ~bit); if (prev == old) break; old = prev; }
  } else {
    u32 i = lower_bound_key(g.Jk, g.nJ, ((u64)a << 32) | n0); Js[i] |= JS_CUT;
    u32 j = lower_bound_key(g.Jk, g.nJ, ((u64)n0 << 32) | a); Js[j] |= JS_CUT;
  }
}

// K10: NBP table.  One raw NBP per (init a, distinct neighbour u) (reference lib/Assembler.py:327-339).
struct NbpTab {
  u32* a; u8* aside; u32* b; u8* bside; u32* L; u32* head; u32* last; u8* cyc;
};
__global__ void k_nbp_heads(GView g, u8* vflags, const u32* __restrict__ init_list, const u32* __restrict__ ibase, u32 nI,
                            const u32* __restrict__ plo, u32 nP, NbpTab t, u32* __restrict__ head_nbp) {
  u64 i = SPB_GTID();
  if (i >= nI) return;
  u32 a = init_list[i];
  u32 id = ibase[i];
  u8 cyc = (g.vflags[a] & VF_CYC) ? 1 : 0;
  bool two = false;
  for_each_nbr(g, a, [&](u32 u, u32 side, u32 oside) {
    t.a[id] = a; t.aside[id] = (u8)side; t.cyc[id] = cyc;
    if (g.vflags[u] & VF_INIT) {
      t.L[id] = 2; t.b[id] = u; t.bside[id] = (u8)oside; t.head[id] = SPB_TERM; t.last[id] = SPB_TERM; two = true;
    } else {
      u32 Q = piece_of(plo, nP, u);
      u32 d = 2 * Q + (oside == SIDE_L ? 0u : 1u);
      head_nbp[d] = id; t.head[id] = d;
    }
    ++id;
  });
  if (two || cyc) atomicOr((u32*)vflags + (a >> 2), (u32)VF_IN2 << ((a & 3) * 8));
}
__global__ void k_dart_finish(const u32* __restrict__ nxt, const u32* __restrict__ root, const u32* __restrict__ acc,
                              const u32* __restrict__ plo, const u32* __restrict__ phi, const u32* __restrict__ pl_nbr,
                              const u8* __restrict__ pl_side, const u32* __restrict__ pr_nbr, const u8* __restrict__ pr_side,
                              const u32* __restrict__ head_nbp, u32 nD, u32* __restrict__ dart_nbp, NbpTab t) {
  u64 d = SPB_GTID();
  if (d >= nD) return;
  if (root[d] == SPB_TERM) { dart_nbp[d] = SPB_INVALID; return; }   // unresolved (cannot happen after cycle cutting)
  u32 id = head_nbp[root[d]];
  dart_nbp[d] = id;
  if (nxt[d] == SPB_TERM) {
    u32 P = (u32)(d >> 1); bool up = !(d & 1);
    t.L[id] = acc[d] + (phi[P] - plo[P] + 1) + 1;
    t.b[id] = up ? pr_nbr[P] : pl_nbr[P];
    t.bside[id] = up ? pr_side[P] : pl_side[P];
    t.last[id] = (u32)d;
  }
}
// per-NBP lengths (vertices / bases) as u64 for the offset scans; cycle NBPs repeat their first vertex
__global__ void k_nbp_lengths(NbpTab t, u32 nN, u32 k, u64* __restrict__ vlen, u64* __restrict__ slen) {
  u64 i = SPB_GTID();
  if (i >= nN) return;
  u64 L = (u64)t.L[i] + t.cyc[i];
  vlen[i] = L; slen[i] = L + k - 1;
}

__device__ __forceinline__ u8 code2ascii(u32 c) { return (u8)("ACGT"[c]); }

// K11: NBP head / tail: first vertex with its k oriented bases, last vertex with its last base.
// (verts / seq: buffers of the NBPs [n_lo, n_hi) only, addressed with the global offsets minus those of NBP n_lo)
__global__ void k_emit_ends(NbpTab t, u32 n_lo, u32 n_hi, u32 k, u64 T, const u32* __restrict__ F, const u32* __restrict__ R,
                            VPos vp, const u64* __restrict__ voff, const u64* __restrict__ soff,
                            u32* __restrict__ verts, u8* __restrict__ seq) {
  u64 id = (SPB_GTID() >> 5) + n_lo;           // one warp per NBP
  u32 lane = threadIdx.x & 31;
  if (id >= n_hi) return;
  u32 a = t.a[id], b = t.b[id];
  u64 vo = voff[id], so = soff[id];
  u32 L = t.L[id];
  bool afwd = t.aside[id] == SIDE_R;          // second vertex on a's right => a read forward
  bool bfwd = t.bside[id] == SIDE_L;          // predecessor on b's left  => b read forward
  u64 pa = vpos_at(vp, a), pb = vpos_at(vp, b);
  const u32* S = afwd ? F : R;
  u64 s0 = afwd ? pa : T - pa - k;
  for (u32 j = lane; j < k; j += 32) seq[so + j] = code2ascii(base_at(S, s0 + j));
  if (lane == 0) {
    verts[vo] = a; verts[vo + L - 1] = b;
    seq[so + k - 1 + (L - 1)] = code2ascii(bfwd ? base_at(F, pb + k - 1) : base_at(R, T - 1 - pb));
    if (t.cyc[id]) {                            // closing vertex of a broken cycle (reference lib/Assembler.py:341-343)
      verts[vo + L] = a;
      seq[so + k - 1 + L] = code2ascii(afwd ? base_at(F, pa + k - 1) : base_at(R, T - 1 - pa));
    }
  }
}

// K12: the bulk of every NBP: the vertices of its darts.  Work is cut into chunks of SPB_CHUNK ranks.
#define SPB_CHUNK 4096
#define SPB_SHORT_DART 512
// long darts are cut into chunks (k_emit_darts), short ones are emitted element-parallel (k_emit_short)
__global__ void k_dart_chunks(const u32* __restrict__ plo, const u32* __restrict__ phi, u32 nD, const u32* __restrict__ dart_nbp, u32 n_lo, u32 n_hi,
                              u32* __restrict__ nch) {
  u64 d = SPB_GTID();
  if (d >= nD) return;
  u32 len = phi[d >> 1] - plo[d >> 1] + 1;
  const u32 id = dart_nbp[d];
  bool is_short = len < SPB_SHORT_DART || id < n_lo || id >= n_hi;      // darts of other ranks' NBPs: no chunks here
  nch[d] = is_short ? 0 : (len + SPB_CHUNK - 1) / SPB_CHUNK;
}
// short darts: one warp per dart (grid-stride over all darts, long ones are skipped); the per-dart parameters are
// loaded once per warp and the lanes write consecutive ranks
__global__ void SPB_LAUNCH_BOUNDS(256)
k_emit_short(u32 nD, const u32* __restrict__ plo, const u32* __restrict__ phi,
             const u32* __restrict__ dart_nbp, const u32* __restrict__ acc, u32 k, u64 T, const u32* __restrict__ F,
             const u32* __restrict__ R, VPos vp, const u64* __restrict__ voff,
             const u64* __restrict__ soff, u32* __restrict__ verts, u8* __restrict__ seq, u32* __restrict__ vnbp, u32 n_lo, u32 n_hi) {
  const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
  const u32 lane = threadIdx.x & 31;
  for (u64 d = SPB_GTID() >> 5; d < nD; d += nwarps) {
    u32 P = (u32)(d >> 1); bool up = !(d & 1);
    u32 lo = plo[P], hi = phi[P], len = hi - lo + 1;
    if (len >= SPB_SHORT_DART) continue;
    u32 id = dart_nbp[d], un = dart_nbp[d ^ 1];
    if (id < n_lo || id >= n_hi) continue;
    if (id < un) un = id;
    u64 r0 = acc[d];
    u64 vo = voff[id] + r0, so = soff[id] + (k - 1) + r0;
    u64 src = up ? (u64)vpos_at(vp, lo) + k - 1 : T - 1 - (u64)vpos_at(vp, hi);
    const u32* S = up ? F : R;
    for (u32 t = lane; t < len; t += 32) {
      u32 v = up ? lo + t : hi - t;
      verts[vo + t] = v;
      seq[so + t] = code2ascii(base_at(S, src + t));
      if (up && vnbp) vnbp[v] = un;
    }
  }
}
struct ChunkParams { u64 vo, so, src; u32 lo, hi, un, t0, n; u32 up; };
__device__ __forceinline__ ChunkParams chunk_params(u32 c, const u32* choff, u32 nD, const u32* plo, const u32* phi, const u32* dart_nbp,
                                                    const u32* acc, u32 k, u64 T, const VPos& vp, const u64* voff, const u64* soff) {
  ChunkParams q;
  u32 d = upper_bound_u32(choff, nD, c) - 1;
  u32 P = d >> 1; q.up = !(d & 1);
  q.lo = plo[P]; q.hi = phi[P];
  u32 len = q.hi - q.lo + 1;
  u32 id = dart_nbp[d];
  u32 un = dart_nbp[d ^ 1]; q.un = id < un ? id : un;
  q.vo = voff[id] + acc[d]; q.so = soff[id] + (k - 1) + acc[d];
  q.src = q.up ? (u64)vpos_at(vp, q.lo) + k - 1 : T - 1 - (u64)vpos_at(vp, q.hi);
  q.t0 = (c - choff[d]) * SPB_CHUNK;
  q.n = len - q.t0 < SPB_CHUNK ? len - q.t0 : SPB_CHUNK;
  return q;
}
// parameters of every chunk, one thread per chunk: the binary search and the dependent table loads of all chunks run in
// parallel here instead of once per block at the head of k_emit_darts (where they were a serial latency chain)
__global__ void k_chunk_table(u32 nchunks, const u32* __restrict__ choff, u32 nD, const u32* __restrict__ plo, const u32* __restrict__ phi,
                              const u32* __restrict__ dart_nbp, const u32* __restrict__ acc, u32 k, u64 T, VPos vp,
                              const u64* __restrict__ voff, const u64* __restrict__ soff, ChunkParams* __restrict__ out) {
  u64 c = SPB_GTID();
  if (c >= nchunks) return;
  out[c] = chunk_params((u32)c, choff, nD, plo, phi, dart_nbp, acc, k, T, vp, voff, soff);
}
__global__ void SPB_LAUNCH_BOUNDS(256)
k_emit_darts(const ChunkParams* __restrict__ cp, const u32* __restrict__ F, const u32* __restrict__ R,
             u32* __restrict__ verts, u8* __restrict__ seq, u32* __restrict__ vnbp) {
  const ChunkParams q = cp[blockIdx.x];
  const u32* S = q.up ? F : R;
  for (u32 t = q.t0 + threadIdx.x; t < q.t0 + q.n; t += blockDim.x) {
    u32 v = q.up ? q.lo + t : q.hi - t;
    verts[q.vo + t] = v;
    if (q.up && vnbp) vnbp[v] = q.un;
  }
  // sequence bytes: one aligned 4-byte word of the output per thread (4 bases = 8 bits of the strand)
  u64 D0 = q.so + q.t0, D1 = D0 + q.n;
  for (u64 wq = (D0 & ~3ull) + 4ull * threadIdx.x; wq < D1; wq += 4ull * blockDim.x) {
    u64 b0 = wq < D0 ? D0 : wq, b1 = wq + 4 < D1 ? wq + 4 : D1;
    u64 i = q.src + (b0 - q.so);                     // strand base index of the first byte we write
    u32 bits = __funnelshift_l(__ldg(S + (i >> 4) + 1), __ldg(S + (i >> 4)), (u32)(i & 15) * 2) >> 24;   // 4 bases
    u32 word = 0;
    for (u32 j = 0; j < 4; ++j) word |= ((0x54474341u >> (8 * ((bits >> (6 - 2 * j)) & 3u))) & 0xFFu) << (8 * j);
    if (b0 == wq && b1 == wq + 4) *(u32*)(seq + wq) = word;
    else for (u64 qq = b0; qq < b1; ++qq) seq[qq] = (u8)(word >> (8 * (qq - b0)));
  }
}

// Sliced multi-GPU tail: NBPs are emitted by the rank that owns their id range, so the vertex -> (undirected) NBP map is
// not a by-product of the emission there; every rank fills it for its own vertex range from the sorted piece list.  One
// warp per 1024 vertices: first piece by binary search, then the pieces that intersect the tile in order.
__global__ void k_fill_vnbp(const u32* __restrict__ plo, const u32* __restrict__ phi, u32 nP, const u32* __restrict__ dart_nbp, u32 v_lo, u32 v_hi,
                            u32* __restrict__ vnbp) {
  const u64 wid = SPB_GTID() >> 5;
  const u32 lane = threadIdx.x & 31;
  const u64 a = (u64)v_lo + 1024ull * wid;
  if (a >= v_hi) return;
  const u32 b = (u32)(a + 1024 < v_hi ? a + 1024 : v_hi);
  for (u32 v = (u32)a + lane; v < b; v += 32) vnbp[v] = SPB_INVALID;
  u32 P = upper_bound_u32(plo, nP, (u32)a);          // pieces that start at or before a: the last one may reach into the tile
  P = P ? P - 1 : 0;
  for (; P < nP && plo[P] < b; ++P) {
    const u32 lo = plo[P] > (u32)a ? plo[P] : (u32)a, hi = phi[P] < b - 1 ? phi[P] : b - 1;
    if (hi < lo) continue;
    u32 un = dart_nbp[2 * P], dn = dart_nbp[2 * P + 1];
    if (dn < un) un = dn;
    const u32 first = lo + ((lane + 32u - ((lo - (u32)a) & 31u)) & 31u);   // every lane keeps to the vertices it cleared above
    for (u32 v = first; v <= hi; v += 32) vnbp[v] = un;
  }
}
// vertex -> (undirected) NBP id: array inside [v_lo, v_hi), piece lookup outside (repeated instances whose vertex
// belongs to another rank's range)
struct VNbp { const u32* vnbp; u32 v_lo, v_hi; const u32* plo; const u32* phi; u32 nP; const u32* dart_nbp; };
__device__ __forceinline__ u32 vnbp_at(const VNbp& q, u32 v) {
  if (v >= q.v_lo && v < q.v_hi) return q.vnbp[v];
  u32 P = upper_bound_u32(q.plo, q.nP, v);
  if (!P) return SPB_INVALID;
  --P;
  if (v > q.phi[P]) return SPB_INVALID;               // not inside a piece: an init
  const u32 un = q.dart_nbp[2 * P], dn = q.dart_nbp[2 * P + 1];
  return dn < un ? dn : un;
}
// first NBP of every rank: NBP ranges with (nearly) equal numbers of emitted vertices; cuts[r] = first id with voff >= r * total / G
__global__ void k_nbp_cuts(const u64* __restrict__ voff, u32 nN, u32 G, u32* __restrict__ cuts) {
  u64 r = SPB_GTID();
  if (r > G) return;
  if (r == G) { cuts[r] = nN; return; }
  const u64 want = (voff[nN] / G) * r;
  u32 lo = 0, hi = nN;
  while (lo < hi) { u32 mid = (lo + hi) >> 1; if (voff[mid] < want) lo = mid + 1; else hi = mid; }
  cuts[r] = lo;
}

// do any of the n flag bytes from vflags[v0] on carry one of the bits of `mask8`?  (warp-cooperative: aligned 4-byte loads;
// the up to three bytes before v0 that share its word are looked at too -- a false positive only costs the general path)
__device__ __forceinline__ bool window_flags_any(const u8* __restrict__ vflags, u32 v0, u32 n, u32 mask8, u32 lane) {
#ifdef SPB_EMU
  (void)lane;
  for (u32 i = 0; i < n; ++i) if (vflags[v0 + i] & mask8) return true;
  return false;
#else
  const u32 m = mask8 * 0x01010101u;
  const u32 b = v0 & ~3u, e = v0 + n;
  bool hit = false;
  for (u32 w = b + 4 * lane; w < e; w += 128) hit |= (__ldg((const u32*)(vflags + w)) & m) != 0;
  return __any_sync(0xffffffffu, hit);
#endif
}

// K13: origins.  (undirected NBP, sequence) pairs where the NBP changes along a sequence; (init vertex,
// sequence) membership pairs for the inits of 2-vertex NBPs (reference lib/Assembler.py:1212-1262).
__global__ void k_origin_pairs(const u32* __restrict__ sp, u64 T, u64 w_lo, u64 w_hi, const u32* __restrict__ fbits, const u32* __restrict__ blkoff,
                               const u8* __restrict__ wpre, VNbp vn, const u8* __restrict__ vflags,
                               const u64* __restrict__ seq_off, u32 n_seq, u64* __restrict__ opairs, u64* ocount, u64 ocap,
                               u64* __restrict__ mpairs, u64* mcount, u64 mcap, u8* __restrict__ has01, u32 sb) {
  // one warp per SPB_JU x 32 positions, grid-stride.  Where p-1 and p are consecutive first appearances (vertices v-1, v
  // joined by an implicit edge) the NBP changes at p only if v starts a piece, so one flag byte decides and the two vnbp
  // gathers are skipped.  The dependent loads (bitmap -> vertex ids -> flags / NBP ids) are issued for all SPB_JU words
  // before they are consumed (latency-bound kernel).
  const u64 nwords = (T + 31) / 32;
  const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
  const u32 lane = threadIdx.x & 31;
  for (u64 wb = w_lo + (SPB_GTID() >> 5) * SPB_WIN; wb < w_hi; wb += nwarps * SPB_WIN) {   // [w_lo, w_hi): our bitmap words (sp filled from position 32 * w_lo - 1)
   {
    // window of 1024 first appearances preceded by one more: vertices vb .. vb + 1023 in a row, every position follows its
    // predecessor over an implicit edge, so the NBP changes only where a vertex is an init or starts a piece -- if no flag
    // byte of the window says so (pieces are long), there is nothing to emit here
    const u32 vb_w = __ldg(blkoff + (wb >> 3)) + __ldg(wpre + wb);
    const u32 prevbit = wb ? (__ldg(fbits + wb - 1) >> 31) : 0u;
    if (wb + SPB_WIN <= w_hi && prevbit && vb_w > 1 && window_all_first(fbits, wb, nwords, lane) &&
        !window_flags_any(vflags, vb_w, 32 * SPB_WIN, VF_INIT | VF_PSTART | VF_IN2, lane)) continue;
   }
   for (u64 wd0 = wb; wd0 < wb + SPB_WIN && wd0 < w_hi; wd0 += SPB_JU) {
    u32 v[SPB_JU], pv[SPB_JU];
    bool pair_first[SPB_JU];
#pragma unroll
    for (u32 j = 0; j < SPB_JU; ++j) {               // stage 1: vertex ids at p and p-1
      const u64 wd = wd0 + j, p = wd * 32 + lane;
      v[j] = pv[j] = SPB_INVALID; pair_first[j] = false;
      if (p < T && wd < w_hi) {
        const u32 f = __ldg(fbits + wd);
        const u32 fprev = wd ? (__ldg(fbits + wd - 1) >> 31) : 0u;
        pair_first[j] = ((f & ((f << 1) | fprev)) >> lane) & 1u;       // p-1 and p both first appearances
        v[j] = sp[p];
        if (p > 0 && !pair_first[j]) pv[j] = sp[p - 1];
      }
    }
    u8 fl[SPB_JU]; u32 c[SPB_JU], prev[SPB_JU];
#pragma unroll
    for (u32 j = 0; j < SPB_JU; ++j) {               // stage 2: flags; NBP ids where the cheap test cannot decide
      fl[j] = 0; c[j] = prev[j] = SPB_INVALID;
      if (v[j] != SPB_INVALID) {
        fl[j] = vflags[v[j]];
        if (!pair_first[j]) { c[j] = vnbp_at(vn, v[j]); if (pv[j] != SPB_INVALID) prev[j] = vnbp_at(vn, pv[j]); }
      }
    }
#pragma unroll
    for (u32 j = 0; j < SPB_JU; ++j) {               // stage 3
      const u64 p = (wd0 + j) * 32 + lane;
      const u32 vv = v[j];
      if (vv == SPB_INVALID) continue;
      bool pfirst = pv[j] == SPB_INVALID;
      if (pair_first[j]) {
        if (vv > 1 && !(fl[j] & (VF_INIT | VF_PSTART))) continue;
        c[j] = vnbp_at(vn, vv);                             // rare: a piece starts here
        const u32 q = sp[p - 1];                     // p > 0 and valid, since p-1 is a first appearance
        prev[j] = vnbp_at(vn, q); pfirst = false;
      }
      const bool in2 = (fl[j] & VF_IN2) != 0;
      const bool emit_o = (c[j] != SPB_INVALID) && (pfirst || c[j] != prev[j]);
      if (!emit_o && !in2 && vv > 1) continue;
      u32 s = upper_bound_u64(seq_off, n_seq + 1, p) - 1;
      if (vv <= 1) has01[2 * s + vv] = 1;
      u64 i = warp_append(ocount, emit_o); if (emit_o && i < ocap) opairs[i] = ((u64)c[j] << sb) | s;
      u64 jj = warp_append(mcount, in2);   if (in2 && jj < mcap) mpairs[jj] = ((u64)vv << sb) | s;
    }
   }
  }
}
// compress_vertices leading-zero quirk (reference lib/vtools.pyx:29-45): a sequence whose smallest
// vertex id is 1 is recorded as containing vertex 0 as well.
__global__ void k_origin_quirk(const u8* __restrict__ has01, u32 n_seq, VNbp vn, const u8* __restrict__ vflags,
                               u64* __restrict__ opairs, u64* ocount, u64 ocap, u64* __restrict__ mpairs, u64* mcount, u64 mcap, u32 sb) {
  u64 s = SPB_GTID();
  if (s >= n_seq) return;
  if (!(has01[2 * s + 1] && !has01[2 * s])) return;
  u32 c = vnbp_at(vn, 0);
  if (c != SPB_INVALID) { u64 i = atomicAdd(ocount, 1ull); if (i < ocap) opairs[i] = ((u64)c << sb) | s; }
  if (vflags[0] & VF_IN2) { u64 i = atomicAdd(mcount, 1ull); if (i < mcap) mpairs[i] = (u64)s; }
}

__device__ __forceinline__ void key_range(const u64* keys, u32 n, u32 id, u32 sb, u32* lo, u32* hi) {   // entries with key >> sb == id
  *lo = lower_bound_key(keys, n, (u64)id << sb);
  *hi = lower_bound_key(keys, n, ((u64)id + 1) << sb);
}
// merge-count / merge-write of two sorted lists of (id << sb | seq) keys: union (mode 0) or intersection (mode 1) of the seq parts
__device__ __forceinline__ u32 merge_lists(const u64* A, u32 a0, u32 a1, const u64* B, u32 b0, u32 b1, int mode, u32 sb, u32* out) {
  const u64 m = (1ull << sb) - 1;
  u32 n = 0;
  while (a0 < a1 || b0 < b1) {
    u32 x = a0 < a1 ? (u32)(A[a0] & m) : 0xFFFFFFFFu, y = b0 < b1 ? (u32)(B[b0] & m) : 0xFFFFFFFFu;
    if (a0 < a1 && b0 < b1 && x == y) { if (out) out[n] = x; ++n; ++a0; ++b0; }
    else if (b0 >= b1 || (a0 < a1 && x < y)) { if (mode == 0) { if (out) out[n] = x; ++n; } ++a0; }
    else { if (mode == 0) { if (out) out[n] = y; ++n; } ++b0; }
  }
  return n;
}
// per-NBP origin lists.  L >= 3: sequences touching an interior vertex; cycle NBPs also count their
// last init (it is interior once the first vertex is appended); L == 2: sequences containing both.
__global__ void k_nbp_origins(NbpTab t, u32 nN, const u32* __restrict__ dart_nbp, const u64* __restrict__ opairs, u32 nO,
                              const u64* __restrict__ mpairs, u32 nM, u32 sb, const u64* __restrict__ ooff, u64* __restrict__ ocnt,
                              u32* __restrict__ out) {
  u64 id = SPB_GTID();
  if (id >= nN) return;
  u32* dst = out ? out + ooff[id] : nullptr;
  u32 n;
  if (t.head[id] == SPB_TERM) {
    u32 a0, a1, b0, b1;
    key_range(mpairs, nM, t.a[id], sb, &a0, &a1); key_range(mpairs, nM, t.b[id], sb, &b0, &b1);
    n = merge_lists(mpairs, a0, a1, mpairs, b0, b1, 1, sb, dst);
  } else {
    u32 un = dart_nbp[t.head[id]], rv = dart_nbp[t.last[id] ^ 1];
    if (rv < un) un = rv;
    u32 a0, a1, b0 = 0, b1 = 0;
    key_range(opairs, nO, un, sb, &a0, &a1);
    if (t.cyc[id]) key_range(mpairs, nM, t.b[id], sb, &b0, &b1);
    n = merge_lists(opairs, a0, a1, mpairs, b0, b1, 0, sb, dst);
  }
  if (!out) ocnt[id] = n;
}

// K14: connected components = union-find over inits, one union per NBP (a, b).
__device__ __forceinline__ u32 uf_find(u32* parent, u32 x) {
  while (true) { u32 p = parent[x]; if (p == x) return x; u32 gp = parent[p]; if (gp != p) parent[x] = gp; x = p; }
}
__device__ __forceinline__ u32 lower_bound_u32(const u32* a, u32 n, u32 x) {
  u32 lo = 0, hi = n;
  while (lo < hi) { u32 mid = (lo + hi) >> 1; if (__ldg(a + mid) < x) lo = mid + 1; else hi = mid; }
  return lo;
}
__global__ void k_uf_init(u32* parent, u32 n) { u64 i = SPB_GTID(); if (i < n) parent[i] = (u32)i; }
__global__ void k_uf_union(NbpTab t, u32 nN, const u32* __restrict__ init_list, u32 nI, u32* parent) {
  u64 id = SPB_GTID();
  if (id >= nN) return;
  u32 x = lower_bound_u32(init_list, nI, t.a[id]), y = lower_bound_u32(init_list, nI, t.b[id]);
  while (true) {
    x = uf_find(parent, x); y = uf_find(parent, y);
    if (x == y) break;
    if (x > y) { u32 z = x; x = y; y = z; }
    if (atomicCAS(parent + y, y, x) == y) break;     // hook the larger root under the smaller
  }
}
// component minimum vertex (may be an extender): inits contribute themselves, pieces their lowest vertex
__global__ void k_comp_min_init(const u32* __restrict__ init_list, u32 nI, u32* parent, u32* cmin) {
  u64 i = SPB_GTID();
  if (i >= nI) return;
  u32 r = uf_find(parent, (u32)i);
  if (init_list[i] < cmin[r]) atomicMin(cmin + r, init_list[i]);   // one giant component: do not hammer a single address
}
__global__ void k_comp_min_piece(const u32* __restrict__ plo, u32 nP, const u32* __restrict__ dart_nbp, NbpTab t,
                                 const u32* __restrict__ init_list, u32 nI, u32* parent, u32* cmin) {
  u64 P = SPB_GTID();
  if (P >= nP) return;
  u32 a = t.a[dart_nbp[2 * P]];
  u32 r = uf_find(parent, lower_bound_u32(init_list, nI, a));
  if (plo[P] < cmin[r]) atomicMin(cmin + r, plo[P]);
}
__global__ void k_comp_roots(u32* parent, const u32* __restrict__ cmin, u32 nI, u64* __restrict__ rootkeys, u32* nroots) {
  u64 i = SPB_GTID();
  if (i >= nI) return;
  if (parent[i] == (u32)i) { u32 j = atomicAdd(nroots, 1u); rootkeys[j] = ((u64)cmin[i] << 32) | (u32)i; }
}
__global__ void k_comp_labels(const u64* __restrict__ rootkeys, u32 nroots, u32* __restrict__ label_of_root) {
  u64 j = SPB_GTID();
  if (j >= nroots) return;
  label_of_root[(u32)rootkeys[j]] = (u32)j;
}
__global__ void k_nbp_comp(NbpTab t, u32 nN, const u32* __restrict__ init_list, u32 nI, u32* parent,
                           const u32* __restrict__ label_of_root, int32_t* __restrict__ nbp_comp) {
  u64 id = SPB_GTID();
  if (id >= nN) return;
  nbp_comp[id] = (int32_t)label_of_root[uf_find(parent, lower_bound_u32(init_list, nI, t.a[id]))];
}
__global__ void k_vertex_comp(const u8* __restrict__ vflags, u32 nV, const u32* __restrict__ vnbp, const u32* __restrict__ init_list,
                              u32 nI, u32* parent, const u32* __restrict__ label_of_root, const int32_t* __restrict__ nbp_comp,
                              int32_t* __restrict__ out) {
  u64 v = SPB_GTID();
  if (v >= nV) return;
  if (vflags[v] & VF_INIT) out[v] = (int32_t)label_of_root[uf_find(parent, lower_bound_u32(init_list, nI, (u32)v))];
  else out[v] = nbp_comp[vnbp[v]];
}

// ------------------------------------------------------------------------------------ NBP graph
// The reference links NBP1 -> NBP2 when the last k bases of NBP1 equal the first k bases of NBP2
// (lib/Assembler.py:365-384).  A k-mer string is a (vertex, orientation) pair (odd k: no palindromes), so the join
// key is (first vertex, read forward?) for the start of an NBP and (last vertex, read forward?) for its end.
__device__ __forceinline__ u64 nbp_start_key(const NbpTab& t, u32 i) { return ((u64)t.a[i] << 1) | (t.aside[i] == SIDE_R ? 1u : 0u); }
__device__ __forceinline__ u64 nbp_end_key(const NbpTab& t, u32 i) {
  if (t.cyc[i]) return nbp_start_key(t, i);                      // a broken cycle ends with its first vertex again
  return ((u64)t.b[i] << 1) | (t.bside[i] == SIDE_L ? 1u : 0u);
}
__global__ void k_nbpg_keys(NbpTab t, u32 nN, u64* __restrict__ key, u32* __restrict__ id) {
  u64 i = SPB_GTID();
  if (i >= nN) return;
  key[i] = nbp_start_key(t, (u32)i); id[i] = (u32)i;
}
// successors of every NBP (count pass when succ == nullptr) and the id of its reverse-orientation partner
__global__ void k_nbpg_join(NbpTab t, u32 nN, const u64* __restrict__ skey, const u32* __restrict__ sid, const u32* __restrict__ dart_nbp,
                            const u64* __restrict__ off, u64* __restrict__ cnt, u32* __restrict__ succ, u32* __restrict__ rev) {
  u64 i = SPB_GTID();
  if (i >= nN) return;
  u64 ek = nbp_end_key(t, (u32)i);
  u32 lo = lower_bound_key(skey, nN, ek), hi = lower_bound_key(skey, nN, ek + 1);
  if (!succ) { cnt[i] = hi - lo; }
  else { u64 o = off[i]; for (u32 j = lo; j < hi; ++j) succ[o + (j - lo)] = sid[j]; }
  if (rev && !succ) {
    u32 r = SPB_INVALID;
    if (t.head[i] != SPB_TERM) r = dart_nbp[t.last[i] ^ 1];       // the reversed last dart heads the partner
    else {                                                          // 2-vertex NBP (a, b): its partner is (b, a)
      u32 l2 = lower_bound_key(skey, nN, (u64)t.b[i] << 1), h2 = lower_bound_key(skey, nN, ((u64)t.b[i] + 1) << 1);
      for (u32 j = l2; j < h2; ++j) { u32 c = sid[j]; if (t.head[c] == SPB_TERM && t.b[c] == t.a[i]) { r = c; break; } }
    }
    rev[i] = r;
  }
}

// ------------------------------------------------------------------------------------ NBP orientation counts
// For every raw NBP: in how many input sequences does its string occur in the sequence's own orientation?  (reference
// get_seqPaths_NBPs, lib/Assembler.py:1121-1139, consumed as nv2rightOrientation at :448-453.)  An occurrence starts where
// the sequence reads the NBP's first k-mer in the NBP's orientation: candidates come from a lookup of (vertex, read
// forward?) in the sorted start keys of the NBP graph, and every candidate is then compared base by base against the
// emitted NBP sequence, so the count is exact whatever the walk does in between.
__global__ void k_occ_candidates(const u32* __restrict__ sp, u64 T, const u8* __restrict__ vflags, const u32* __restrict__ v2pos,
                                 const u32* __restrict__ obits, const u64* __restrict__ seq_off, u32 n_seq, const u64* __restrict__ skey,
                                 const u32* __restrict__ sid, u32 nN, const u64* __restrict__ soff, u64* __restrict__ cand, u64* ncand, u64 cap) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 p = SPB_GTID(); p < T; p += stride) {
    const u32 u = sp[p];
    if (u == SPB_INVALID || !(vflags[u] & VF_INIT)) continue;      // NBPs start at inits (incl. the cut vertex of a cycle)
    const u32 su = get_bit(obits, p) ^ get_bit(obits, v2pos[u]);   // instance strand vs. stored k-mer
    const u64 key = ((u64)u << 1) | (su ? 0u : 1u);                // (vertex, read forward?)  == nbp_start_key
    u32 lo = lower_bound_key(skey, nN, key);
    if (lo >= nN || skey[lo] != key) continue;
    const u32 s = upper_bound_u64(seq_off, n_seq + 1, p) - 1;
    const u64 end = seq_off[s + 1];
    for (; lo < nN && skey[lo] == key; ++lo) {
      const u32 x = sid[lo];
      if (p + (soff[x + 1] - soff[x]) > end) continue;             // does not fit into the rest of the sequence
      const u64 i = atomicAdd(ncand, 1ull);
      if (i < cap) cand[i] = (p << 32) | x;
    }
  }
}
// one warp per candidate: exact comparison of the NBP sequence (ASCII) with the packed forward strand at p
__global__ void k_occ_verify(const u64* __restrict__ cand, u64 ncand, const u32* __restrict__ F, const u8* __restrict__ seqs,
                             const u64* __restrict__ soff, const u64* __restrict__ seq_off, u32 n_seq, u32 sb, u64* __restrict__ pairs, u64* npairs) {
#ifdef SPB_EMU
  const u64 i = SPB_GTID();
  if (i >= ncand) return;
  const u64 p = cand[i] >> 32; const u32 x = (u32)cand[i];
  const u64 len = soff[x + 1] - soff[x];
  for (u64 j = 0; j < len; ++j) if (seqs[soff[x] + j] != code2ascii(base_at(F, p + j))) return;
  const u32 s = upper_bound_u64(seq_off, n_seq + 1, p) - 1;
  pairs[atomicAdd(npairs, 1ull)] = ((u64)x << sb) | s;
#else
  const u64 i = SPB_GTID() >> 5;
  if (i >= ncand) return;                            // warp-uniform
  const u32 lane = threadIdx.x & 31;
  const u64 p = cand[i] >> 32; const u32 x = (u32)cand[i];
  const u64 len = soff[x + 1] - soff[x], so = soff[x];
  bool ok = true;
  for (u64 j0 = 0; j0 < len && ok; j0 += 32) {
    const u64 j = j0 + lane;
    const bool bad = j < len && seqs[so + j] != code2ascii(base_at(F, p + j));
    ok = __ballot_sync(0xffffffffu, bad) == 0;
  }
  if (ok && lane == 0) {
    const u32 s = upper_bound_u64(seq_off, n_seq + 1, p) - 1;
    pairs[atomicAdd(npairs, 1ull)] = ((u64)x << sb) | s;
  }
#endif
}
// count of distinct sequences per NBP from the sorted unique pairs; the reference's compress_vertices quirk
// (lib/vtools.pyx:29-45: a vertex list whose smallest vertex is 1 is recorded as containing 0 too) makes its
// pre-filter `vertex_overlap(cNBP, seqPath) == refLen` fail for every NBP that holds vertex 1 but not vertex 0.
__device__ __forceinline__ bool nbp_holds(const NbpTab& t, const u32* rev, const u32* vnbp, const u8* vflags, u32 x, u32 v) {
  if (t.a[x] == v || t.b[x] == v) return true;
  if (vflags[v] & VF_INIT) return false;
  const u32 r = rev[x];
  return vnbp[v] == (r < x ? r : x);
}
// (NBP << sb | sequence) -> (sequence << 32 | NBP) for the per-sequence view; pairs of NBPs the quirk excludes get ~0
__global__ void k_occ_rekey(const u64* __restrict__ pairs, u32 npairs, u32 sb, const u32* __restrict__ counts, u64* __restrict__ out) {
  const u64 i = SPB_GTID();
  if (i >= npairs) return;
  const u64 x = pairs[i] >> sb, s = pairs[i] & ((1ull << sb) - 1);
  out[i] = counts[x] ? ((s << 32) | x) : ~0ull;
}
__global__ void k_occ_seq_offsets(const u64* __restrict__ keys, u32 n, u32 n_seq, u64* __restrict__ off) {
  const u64 s = SPB_GTID();
  if (s > n_seq) return;
  off[s] = lower_bound_key(keys, n, s << 32);
}
__global__ void k_occ_ids(const u64* __restrict__ keys, u32 n, u32* __restrict__ ids) {
  const u64 i = SPB_GTID();
  if (i < n) ids[i] = (u32)keys[i];
}
__global__ void k_occ_counts(const u64* __restrict__ pairs, u32 npairs, u32 sb, NbpTab t, u32 nN, const u32* __restrict__ rev,
                             const u32* __restrict__ vnbp, const u8* __restrict__ vflags, u32 nV, u32* __restrict__ out) {
  const u64 x = SPB_GTID();
  if (x >= nN) return;
  u32 c = lower_bound_key(pairs, npairs, (x + 1) << sb) - lower_bound_key(pairs, npairs, x << sb);
  if (nV > 1 && nbp_holds(t, rev, vnbp, vflags, (u32)x, 1u) && !nbp_holds(t, rev, vnbp, vflags, (u32)x, 0u)) c = 0;
  out[x] = c;
}

// ------------------------------------------------------------------------------------ bubble candidates
// The reference groups the NBPs by (first vertex, last vertex) and aligns the members of every group with more than
// one path (lib/Assembler.py:728-734).  Key = first << 32 | last; a broken cycle ends with its first vertex.
__global__ void k_bubble_keys(NbpTab t, u32 nN, u64* __restrict__ key, u32* __restrict__ id) {
  u64 i = SPB_GTID();
  if (i >= nN) return;
  key[i] = ((u64)t.a[i] << 32) | (t.cyc[i] ? t.a[i] : t.b[i]); id[i] = (u32)i;
}
// group[i] = smallest NBP id of the group of i when the group has more than one member, else SPB_INVALID
__global__ void k_bubble_groups(const u64* __restrict__ skey, const u32* __restrict__ sid, u32 nN, u32* __restrict__ group) {
  u64 j = SPB_GTID();
  if (j >= nN) return;
  const u64 k0 = skey[j];
  const u32 lo = lower_bound_key(skey, nN, k0), hi = lower_bound_key(skey, nN, k0 + 1);
  u32 m = SPB_INVALID;
  if (hi - lo > 1) for (u32 q = lo; q < hi; ++q) m = sid[q] < m ? sid[q] : m;
  group[sid[j]] = m;
}

// ------------------------------------------------------------------------------------ sequences of arbitrary paths
// reconstruct_sequence (reference lib/Assembler.py:1038-1118) for vertex paths the host hands in (joined NBPs): the
// first k-mer plus the last base of every following one, each vertex read in the orientation in which it overlaps its
// predecessor.  The orientation of a vertex in a path is local: it is read as stored (= as at its first appearance,
// vertex2kmer :1063-1070) iff its predecessor attaches on its left side, so every path element is independent.
// side of v on which nbr attaches (SIDE_L / SIDE_R), 2 = not adjacent.  Graph structure first; edges the cycle breaking
// removed from it are recovered from the k-1 base overlaps of the stored k-mers.
__device__ __forceinline__ u32 attach_side(const GView& g, const u32* __restrict__ F, const u32* __restrict__ R, u64 T, u32 k,
                                           const u32* __restrict__ v2pos, u32 v, u32 nbr) {
  if (nbr == v + 1 && (g.vflags[v] & VF_R)) return SIDE_R;
  if (nbr + 1 == v && (g.vflags[nbr] & VF_R)) return SIDE_L;
  if (g.vflags[v] & VF_HASJ) {
    const u32 i = lower_bound_key(g.Jk, g.nJ, ((u64)v << 32) | nbr);
    if (i < g.nJ && g.Jk[i] == (((u64)v << 32) | nbr)) return g.Js[i] & 1u;
  }
  const u64 pa = v2pos[v], pb = v2pos[nbr];          // stored k-mers: A = F[pa, pa+k), B = F[pb, pb+k), rc(B) = R[T-pb-k, T-pb)
  if (win_cmp(F, pa + 1, F, pb, k - 1) == 0 || win_cmp(F, pa + 1, R, T - pb - k, k - 1) == 0) return SIDE_R;     // A[1:] == B[:-1] / rc(B)[:-1]
  if (win_cmp(F, pb + 1, F, pa, k - 1) == 0 || win_cmp(R, T - pb - k + 1, F, pa, k - 1) == 0) return SIDE_L;     // B[1:] / rc(B)[1:] == A[:-1]
  return 2u;
}
__global__ void k_path_seqs(GView g, const u32* __restrict__ F, const u32* __restrict__ R, u64 T, u32 k, const u32* __restrict__ v2pos,
                            const u64* __restrict__ off, u32 n_paths, const u32* __restrict__ verts, u8* __restrict__ out, u32* err) {
  const u64 total = off[n_paths];
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 e = SPB_GTID(); e < total; e += stride) {
    const u32 path = upper_bound_u64(off, n_paths + 1, e) - 1;
    const u64 i = e - off[path], len = off[path + 1] - off[path];
    const u32 v = verts[e];
    if (len < 2 || v >= g.nV) { atomicOr(err, (u32)ERR_BADPATH); continue; }
    const u64 p = v2pos[v];
    if (i == 0) {                                    // first k-mer: forward iff the second vertex attaches on its right side
      const u32 w = verts[e + 1];
      const u32 side = w < g.nV && w != v ? attach_side(g, F, R, T, k, v2pos, v, w) : 2u;
      if (side > 1) { atomicOr(err, (u32)ERR_BADPATH); continue; }
      const u64 o = e + (u64)path * (k - 1);
      for (u32 j = 0; j < k; ++j) out[o + j] = code2ascii(side == SIDE_R ? base_at(F, p + j) : base_at(R, T - p - k + j));
    } else {                                         // last base of the k-mer, read forward iff the predecessor attaches on the left
      const u32 w = verts[e - 1];
      const u32 side = w < g.nV && w != v ? attach_side(g, F, R, T, k, v2pos, v, w) : 2u;
      if (side > 1) { atomicOr(err, (u32)ERR_BADPATH); continue; }
      out[e + (u64)(path + 1) * (k - 1)] = code2ascii(side == SIDE_L ? base_at(F, p + k - 1) : 3u - base_at(F, p));
    }
  }
}

// ------------------------------------------------------------------------------------ origins of arbitrary paths
// get_ori2cov_onevertex + goodOris (reference lib/Assembler.py:1212-1248, :941-946) for host-supplied paths that are
// concatenations of whole NBPs (the joined paths of the tail): a path of more than two vertices originates from every
// sequence that holds one of its INTERIOR vertices.  An interior extender contributes the origin list of its NBP (already
// computed: the union over the NBP's interior), an interior init contributes the sequences that hold it -- (init vertex,
// sequence) pairs collected once from the per-position vertex ids, incl. the leading-zero quirk for vertex 0.
__global__ void k_init_seq_pairs(const u32* __restrict__ sp, u64 T, const u8* __restrict__ vflags, const u64* __restrict__ seq_off, u32 n_seq,
                                 u32 sb, u64* __restrict__ pairs, u64* npairs, u64 cap, u8* __restrict__ has01) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 p = SPB_GTID(); p < T; p += stride) {
    const u32 v = sp[p];
    if (v == SPB_INVALID) continue;
    const bool ini = (vflags[v] & VF_INIT) != 0;
    if (!ini && v > 1) continue;
    const u32 s = upper_bound_u64(seq_off, n_seq + 1, p) - 1;
    if (v <= 1) has01[2 * s + v] = 1;
    if (ini) { const u64 i = atomicAdd(npairs, 1ull); if (i < cap) pairs[i] = ((u64)v << sb) | s; }
  }
}
__global__ void k_init_seq_quirk(const u8* __restrict__ has01, u32 n_seq, const u8* __restrict__ vflags, u32 sb, u64* __restrict__ pairs,
                                 u64* npairs, u64 cap) {
  const u64 s = SPB_GTID();
  if (s >= n_seq) return;
  if (has01[2 * s + 1] && !has01[2 * s] && (vflags[0] & VF_INIT)) { const u64 i = atomicAdd(npairs, 1ull); if (i < cap) pairs[i] = s; }   // vertex 0
}
// Paths given as lists of raw NBP ids (consecutive NBPs share one vertex: the reference joins them as path + succ[1:],
// lib/Assembler.py:810) -> vertex lists.  el_len[e] = vertices element e contributes; after the scan one warp per element
// copies them.  A pair of consecutive NBPs that does not share its junction vertex flags ERR_BADPATH.
__global__ void k_nbp_path_lengths(const u64* __restrict__ poff, u32 n_paths, const u32* __restrict__ ids, u32 nN, const u64* __restrict__ voff,
                                   const u32* __restrict__ verts, u64* __restrict__ el_len, u32* err) {
  const u64 total = poff[n_paths];
  const u64 e = SPB_GTID();
  if (e >= total) return;
  const u32 id = ids[e];
  if (id >= nN) { atomicOr(err, (u32)ERR_BADPATH); el_len[e] = 0; return; }
  const u32 path = upper_bound_u64(poff, n_paths + 1, e) - 1;
  const bool first = e == poff[path];
  const u64 len = voff[id + 1] - voff[id];
  if (!first) {
    const u32 prev = ids[e - 1];
    if (prev >= nN || verts[voff[prev + 1] - 1] != verts[voff[id]]) atomicOr(err, (u32)ERR_BADPATH);
  }
  el_len[e] = first ? len : len - 1;
}
__global__ void k_nbp_path_write(const u64* __restrict__ poff, u32 n_paths, const u32* __restrict__ ids, const u64* __restrict__ voff,
                                 const u32* __restrict__ verts, const u64* __restrict__ el_off, u64* __restrict__ out_off, u32* __restrict__ out_verts) {
  const u64 total = poff[n_paths];
  const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
  const u32 lane = threadIdx.x & 31;
  for (u64 e = SPB_GTID() >> 5; e < total; e += nwarps) {
    const u32 id = ids[e];
    const u32 path = upper_bound_u64(poff, n_paths + 1, e) - 1;
    const bool first = e == poff[path];
    if (first && lane == 0) out_off[path] = el_off[e];
    const u64 src = voff[id] + (first ? 0 : 1), n = voff[id + 1] - src;
    for (u64 t = lane; t < n; t += 32) out_verts[el_off[e] + t] = verts[src + t];
  }
}

// bubble_vertex2origins (reference lib/Assembler.py:922, :1236): extra (vertex -> sequences) lists left behind by popped
// bubbles, as a CSR over ascending vertex ids; every vertex a path is asked about contributes them as well
struct BubbleOri { const u32* vert; const u64* off; const u32* seq; u32 n; };
__device__ __forceinline__ bool bubble_range(const BubbleOri& b, u32 v, u64* lo, u64* hi) {
  if (!b.n) return false;
  u32 l = 0, h = b.n;
  while (l < h) { const u32 m = (l + h) >> 1; if (b.vert[m] < v) l = m + 1; else h = m; }
  if (l >= b.n || b.vert[l] != v) return false;
  *lo = b.off[l]; *hi = b.off[l + 1];
  return true;
}
__device__ __forceinline__ bool bubble_has(const BubbleOri& b, u32 v, u32 s) {
  u64 lo, hi;
  if (!bubble_range(b, v, &lo, &hi)) return false;
  for (u64 x = lo; x < hi; ++x) if (b.seq[x] == s) return true;
  return false;
}
__device__ __forceinline__ bool ivp_has(const u64* ivp, u32 n_ivp, u32 sb, u32 v, u32 s) {
  const u32 i = lower_bound_key(ivp, n_ivp, ((u64)v << sb) | s);
  return i < n_ivp && ivp[i] == (((u64)v << sb) | s);
}
__global__ void k_path_origin_pairs(const u64* __restrict__ off, u32 n_paths, const u32* __restrict__ verts, u32 nV, const u8* __restrict__ vflags,
                                    const u32* __restrict__ vnbp, const u64* __restrict__ ooff, const u32* __restrict__ origins,
                                    const u64* __restrict__ ivp, u32 n_ivp, u32 sb, BubbleOri bub, u64* __restrict__ out, u64* nout, u64 cap, u32* err) {
  const u64 total = off[n_paths];
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 e = SPB_GTID(); e < total; e += stride) {
    const u32 path = upper_bound_u64(off, n_paths + 1, e) - 1;
    const u64 i = e - off[path], len = off[path + 1] - off[path];
    const u32 v = verts[e];
    if (len < 2 || v >= nV) { atomicOr(err, (u32)ERR_BADPATH); continue; }
    if (len == 2) {                                  // both vertices must be held: intersection of the two init lists
      if (i) continue;
      const u32 w = verts[e + 1];
      if (w >= nV || !(vflags[v] & VF_INIT) || !(vflags[w] & VF_INIT)) { atomicOr(err, (u32)ERR_BADPATH); continue; }
      u32 a = lower_bound_key(ivp, n_ivp, (u64)v << sb), ae = lower_bound_key(ivp, n_ivp, ((u64)v + 1) << sb);
      u32 b = lower_bound_key(ivp, n_ivp, (u64)w << sb), be = lower_bound_key(ivp, n_ivp, ((u64)w + 1) << sb);
      const u64 m = (1ull << sb) - 1;
      if (!bub.n) {
        while (a < ae && b < be) {
          const u64 x = ivp[a] & m, y = ivp[b] & m;
          if (x < y) ++a; else if (y < x) ++b;
          else { const u64 j = atomicAdd(nout, 1ull); if (j < cap) out[j] = ((u64)path << 32) | x; ++a; ++b; }
        }
      } else {
        // with bubble lists: a sequence counts for a vertex when the vertex is in it OR a popped bubble says so; it is kept
        // when that holds for both vertices (the sort + unique of the caller removes what is found twice)
        u64 lo = 0, hi = 0;
        for (u32 x = a; x < ae; ++x) { const u32 q = (u32)(ivp[x] & m); if (ivp_has(ivp, n_ivp, sb, w, q) || bubble_has(bub, w, q)) { const u64 j = atomicAdd(nout, 1ull); if (j < cap) out[j] = ((u64)path << 32) | q; } }
        if (bubble_range(bub, v, &lo, &hi)) for (u64 x = lo; x < hi; ++x) { const u32 q = bub.seq[x]; if (ivp_has(ivp, n_ivp, sb, w, q) || bubble_has(bub, w, q)) { const u64 j = atomicAdd(nout, 1ull); if (j < cap) out[j] = ((u64)path << 32) | q; } }
      }
      continue;
    }
    if (i == 0 || i == len - 1) continue;            // the end vertices overlap other paths (:1224-1227)
    {
      u64 lo = 0, hi = 0;
      if (bubble_range(bub, v, &lo, &hi)) for (u64 x = lo; x < hi; ++x) { const u64 j = atomicAdd(nout, 1ull); if (j < cap) out[j] = ((u64)path << 32) | bub.seq[x]; }
    }
    if (vflags[v] & VF_INIT) {
      for (u32 a = lower_bound_key(ivp, n_ivp, (u64)v << sb), ae = lower_bound_key(ivp, n_ivp, ((u64)v + 1) << sb); a < ae; ++a) {
        const u64 j = atomicAdd(nout, 1ull);
        if (j < cap) out[j] = ((u64)path << 32) | (ivp[a] & ((1ull << sb) - 1));
      }
    } else {
      const u32 x = vnbp[v];
      if (x == SPB_INVALID) { atomicOr(err, (u32)ERR_BADPATH); continue; }
      const u32 pv = verts[e - 1];
      if (i > 1 && pv < nV && !(vflags[pv] & VF_INIT) && vnbp[pv] == x) continue;      // same NBP as the element before: listed once
      for (u64 a = ooff[x]; a < ooff[x + 1]; ++a) {
        const u64 j = atomicAdd(nout, 1ull);
        if (j < cap) out[j] = ((u64)path << 32) | origins[a];
      }
    }
  }
}
__global__ void k_path_origin_offsets(const u64* __restrict__ keys, u32 n, u32 n_paths, u64* __restrict__ off) {
  const u64 p = SPB_GTID();
  if (p > n_paths) return;
  off[p] = lower_bound_key(keys, n, p << 32);
}

// ------------------------------------------------------------------------------------ getters
__global__ void k_coords(const u32* __restrict__ v2pos, u32 nV, const u64* __restrict__ seq_off, u32 n_seq, u32* __restrict__ out) {
  u64 v = SPB_GTID();
  if (v >= nV) return;
  u64 p = v2pos[v];
  u32 s = upper_bound_u64(seq_off, n_seq + 1, p) - 1;
  out[2 * v] = s; out[2 * v + 1] = (u32)(p - seq_off[s]);
}

// explicit edge list (reference lib/Assembler.py:247-249): per vertex v the rows (v, x), x >= v:
// self-loop first, then the implicit (v, v+1), then junction neighbours above v.  8 vertices per thread.
__global__ void k_edge_counts8(GView g, u32* __restrict__ cnt, u32 t_lo, u32 t_hi) {   // tiles [t_lo, t_hi) of 8 vertices; cnt[t - t_lo]
  u64 t = SPB_GTID() + t_lo;
  u64 v0 = t * 8;
  if (t >= t_hi || v0 >= g.nV) return;
  u64 w = *(const u64*)(g.vflags + v0);
  u32 c = 0;
  for (u32 i = 0; i < 8 && v0 + i < g.nV; ++i) {
    u32 fl = (u32)(w >> (8 * i)) & 0xFF;
    if (fl & VF_R) ++c;
    if (fl & VF_HASJ) {
      u32 v = (u32)(v0 + i);
      u32 j = lower_bound_key(g.Jk, g.nJ, ((u64)v << 32) | v);
      for (; j < g.nJ && (u32)(g.Jk[j] >> 32) == v; ++j) ++c;
    }
  }
  cnt[t - t_lo] = c;
}
// rows of vertex v start at off8[v / 8 - t_lo] + (rows of the lower vertices of its group); one thread per vertex so that
// the common case -- only the implicit (v, v+1) row -- is written as consecutive 8-byte stores
__device__ __forceinline__ u32 edge_rows_of(const GView& g, u32 v, u32 fl) {
  u32 c = (fl & VF_R) ? 1 : 0;
  if (fl & VF_HASJ) {
    u32 j = lower_bound_key(g.Jk, g.nJ, ((u64)v << 32) | v);
    for (; j < g.nJ && (u32)(g.Jk[j] >> 32) == v; ++j) ++c;
  }
  return c;
}
__global__ void k_edge_write(GView g, const u32* __restrict__ off8, u32* __restrict__ edges, u32 t_lo, u32 t_hi) {
  // One warp per window of 1024 vertices of the tiles [t_lo, t_hi), grid-stride.  A window in which every vertex carries the
  // implicit edge and no junction entries (the usual case) is written without per-vertex offsets: row i is
  // (v0 + i, v0 + i + 1) at off8[first tile] + i, two rows per 16-byte store.  Any other window: one vertex per lane and step,
  // rows of vertex v start at off8[v / 8 - t_lo] + (rows of the lower vertices of its tile).
  uint2* rows = (uint2*)edges;
  const u64 v_first = 8ull * t_lo;
  const u64 v_end = 8ull * t_hi < g.nV ? 8ull * t_hi : g.nV;
  if (v_end <= v_first) return;
  const u64 n_win = (v_end - v_first + 1023) / 1024;
  const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
  const u32 lane = threadIdx.x & 31;
  for (u64 wi = SPB_GTID() >> 5; wi < n_win; wi += nwarps) {
    const u64 v0 = v_first + 1024 * wi;
#ifndef SPB_EMU
    if (v0 + 1024 <= v_end) {
      const u64 need = SPB_M8 * (u64)(VF_R | VF_HASJ), want = SPB_M8 * (u64)VF_R;
      const u64* fp = reinterpret_cast<const u64*>(g.vflags + v0 + 32 * lane);                     // 32 flag bytes per lane (v0 is a multiple of 8)
      const u64 f0 = fp[0], f1 = fp[1], f2 = fp[2], f3 = fp[3];
      const bool plain = ((f0 & need) == want) && ((f1 & need) == want) && ((f2 & need) == want) && ((f3 & need) == want);
      if (__all_sync(0xffffffffu, plain)) {
        const u64 o = off8[(v0 >> 3) - t_lo];
        uint2* dst = rows + o;
        if ((o & 1) == 0) {                                                                       // 16-byte aligned: two rows per store
#pragma unroll 4
          for (u32 j = 0; j < 16; ++j) {
            const u32 i = 2 * (32 * j + lane);
            const u32 v = (u32)v0 + i;
            *reinterpret_cast<uint4*>(dst + i) = make_uint4(v, v + 1, v + 1, v + 2);
          }
        } else {
#pragma unroll 4
          for (u32 j = 0; j < 32; ++j) { const u32 i = 32 * j + lane; dst[i] = make_uint2((u32)v0 + i, (u32)v0 + i + 1); }
        }
        continue;
      }
    }
#endif
    for (u32 j = 0; j < 32; ++j) {
      const u64 v = v0 + 32 * j + lane;
      if (v >= v_end) break;
      const u64 w = *(const u64*)(g.vflags + (v & ~7ull));
      u64 o = off8[(v >> 3) - t_lo];
      const u32 i = (u32)(v & 7);
      const u64 below = i ? (w & ((1ull << (8 * i)) - 1)) : 0;
      if (!(below & (0x0101010101010101ull * VF_HASJ))) o += __popcll(below & (0x0101010101010101ull * VF_R));
      else for (u32 x = 0; x < i; ++x) o += edge_rows_of(g, (u32)(v - i + x), (u32)(w >> (8 * x)) & 0xFF);
      const u32 fl = (u32)(w >> (8 * i)) & 0xFF;
      bool r = (fl & VF_R) != 0;
      if (fl & VF_HASJ) {
        u32 x = lower_bound_key(g.Jk, g.nJ, ((u64)v << 32) | v);
        for (; x < g.nJ && (u32)(g.Jk[x] >> 32) == (u32)v; ++x) {
          u32 d = (u32)g.Jk[x];
          if (r && d > (u32)v + 1) { rows[o++] = make_uint2((u32)v, (u32)v + 1); r = false; }
          rows[o++] = make_uint2((u32)v, d);
        }
      }
      if (r) rows[o] = make_uint2((u32)v, (u32)v + 1);
    }
  }
}

// seqPaths (reference lib/vtools.pyx:18-49).  Along a sequence the vertex ids form monotone runs with
// step +1 or -1; a run boundary lies before p when p opens its sequence, the step is not +-1, or
// the direction changes.  Run starts and run ends are compacted separately (order preserving), so
// the i-th start pairs with the i-th end.
__device__ __forceinline__ int run_step(const u32* sp, u64 p) {   // step from p-1 to p: +1, -1, or 0 (= boundary)
  if (p == 0) return 0;
  u32 a = sp[p - 1], b = sp[p];
  if (a == SPB_INVALID || b == SPB_INVALID) return 0;
  return (b == a + 1) ? 1 : (a == b + 1) ? -1 : 0;
}
__device__ __forceinline__ bool run_boundary(const u32* sp, u64 p) {  // true if a run starts at p (sp[p] valid)
  int s = run_step(sp, p);
  if (s == 0) return true;
  int s1 = run_step(sp, p - 1);
  return s1 != 0 && s1 != s;
}
__global__ void k_run_flags(const u32* __restrict__ sp, u64 T, u8* __restrict__ flags) {
  u64 p = SPB_GTID();
  if (p >= T) return;
  u8 f = 0;
  if (sp[p] != SPB_INVALID) {
    if (run_boundary(sp, p)) f |= 1;
    if (p + 1 >= T || sp[p + 1] == SPB_INVALID || run_boundary(sp, p + 1)) f |= 2;
  }
  flags[p] = f;
}
__global__ void k_run_intervals(const u32* __restrict__ sp, const u32* __restrict__ rs, const u32* __restrict__ re, u32 nR,
                                const u64* __restrict__ seq_off, u32 n_seq, u64* __restrict__ key, u64* __restrict__ val) {
  u64 i = SPB_GTID();
  if (i >= nR) return;
  u32 a = sp[rs[i]], b = sp[re[i]];
  u32 lo = a < b ? a : b, hi = a < b ? b : a;
  u64 s = upper_bound_u64(seq_off, n_seq + 1, rs[i]) - 1;
  key[i] = (s << 32) | lo; val[i] = (s << 32) | hi;
}
// after sorting by (seq, lo) and a segmented running max of hi: interval heads
__global__ void k_interval_heads(const u64* __restrict__ key, const u64* __restrict__ runmax, u32 nR, u32* __restrict__ head) {
  u64 i = SPB_GTID();
  if (i >= nR) return;
  bool h = (i == 0) || (key[i] >> 32) != (key[i - 1] >> 32) || (u32)key[i] > (u32)runmax[i - 1] + 1;
  head[i] = h ? 1u : 0u;
}
__global__ void k_interval_write(const u64* __restrict__ key, const u64* __restrict__ runmax, u32 nR, const u32* __restrict__ head,
                                 const u32* __restrict__ pos, u32* __restrict__ iv, u32* __restrict__ iv_seq) {
  u64 i = SPB_GTID();
  if (i >= nR) return;
  if (head[i]) {
    u32 lo = (u32)key[i];
    bool first_of_seq = (i == 0) || (key[i] >> 32) != (key[i - 1] >> 32);
    if (first_of_seq && lo <= 1) lo = 0;          // compress_vertices `last = cs = 0` quirk (lib/vtools.pyx:29-38)
    iv[2 * (u64)pos[i]] = lo; iv_seq[pos[i]] = (u32)(key[i] >> 32);
  }
  bool last = (i + 1 == nR) || head[i + 1];
  if (last) iv[2 * (u64)(head[i] ? pos[i] : pos[i] - 1) + 1] = (u32)runmax[i];   // pos = exclusive scan of head
}
__global__ void k_interval_offsets(const u32* __restrict__ iv_seq, u32 nIv, u32 n_seq, u64* __restrict__ off) {
  // off[s] = first interval row of sequence s (rows are grouped by sequence, ascending)
  u64 s = SPB_GTID();
  if (s > n_seq) return;
  off[s] = lower_bound_u32(iv_seq, nIv, (u32)s);
}
