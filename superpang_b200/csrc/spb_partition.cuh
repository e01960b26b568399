// superpang_b200 -- two-level partitioned dedup with shared-memory tables, hand-written end to end (round 2).
//
// Replaces "k_make_records_roll -> library radix sort (2 passes over 12-byte records) -> k_smem_dedup" on the mostly-unique
// path.  Records are 8 bytes: [32 hash bits | 32-bit position]; the bucket a record sits in
// carries the hash bits above those 32, so no bit is stored twice:
//
//   k_roll_partition   rolling canonical hashes (as k_make_records_roll) FUSED with the level-1 partition: a CTA stages
//                      the records of T1 x GS positions in shared memory, counting-sorts them by the top L1 hash bits
//                      (shared-memory histogram -> ranks), reserves one run per bucket with ONE global atomic per
//                      (CTA, bucket, group) and copies the staged records out as contiguous runs.  Buckets are
//                      fixed-capacity regions rec1[b * cap1 ..) with a cursor each (hashes are uniform; an overflow is
//                      flagged and the host falls back to the sort-based path).
//   k_partition2       the same counting sort over 8192-record chunks of a level-1 region, by the next L2 hash bits,
//                      into the final regions rec2[(b1 << L2 | b2) * cap2 ..).
//   k_bucket_dedup     one CTA per final bucket: the bucket (<= 40 KB) is brought into shared memory by ONE bulk
//                      asynchronous copy (cp.async.bulk, completion on an mbarrier) while the CTA clears its 8192-slot
//                      table; inserts then touch shared memory only (slot = tag | record index, CAS-first, double
//                      hashing, lane-persistent probing); a tag + hash match is verified on the full 2k bits against the
//                      packed genome; the EARLIER position keeps the slot (CAS take-over), so records may arrive in any
//                      order.  Exceptions (repeated k-mers) leave links (position -> smaller position with the same
//                      k-mer) in place (single GPU) or in a list (multi-GPU owner).
//
// Algorithmic bytes per instance: 8 written (P1) + 8 read + 8 written (P2) + 8 read (dedup) = 32, against 80 for the
// library-sort path (12 written, 4 + 2 x 24 in the sort, 12 read).
//
// The emulator build (SPB_EMU, test infrastructure) runs simplified sequential forms of P1 / P2 (same records into the
// same regions, in a different order -- the dedup is order-independent) and the phased form of the dedup kernel.
#pragma once

#define SPB_PART_MAXB 512            // buckets per level <= 2^9 (shared-memory histogram)
#define SPB_P2_T 512
#define SPB_P2_PT 8                  // records per thread: 4096-record chunks
#define SPB_P2_CHUNK (SPB_P2_T * SPB_P2_PT)
#define SPB_D_T 512
#define SPB_D_SLOTS 8192u            // 32 KB table
#define SPB_D_LIMIT 4800u            // records per final bucket (37.5 KB): table load <= 0.59; more -> overflow -> fallback
#define SPB_D_DEFER 2560u            // records that may wait for the second pass (u16 indices)
#define SPB_D_AVG 3840u              // target mean (load 0.47)
#define SPB_D_EMPTY 0xFFFFFFFFu

#ifndef SPB_EMU
__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory"); }
// bulk asynchronous copy global -> shared (TMA engine, no registers, no per-thread loads); completes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, u32 bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
  asm volatile("{\n.reg .pred p;\nSPB_WAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra SPB_DONE_%=;\nbra SPB_WAIT_%=;\nSPB_DONE_%=:\n}"
               ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
#endif

struct RollSt { u32 f1, f2, r1, r2, head, rct, wb, wc; bool have, fresh; };   // rct = revcomp16 of the last 16 bases

// canonical rolling hash of the k-mer at p (64 bits, equal for a k-mer and its reverse complement); o = canonical is forward.
// The caller resets st.have = false whenever the previous position did not start a k-mer.
__device__ __forceinline__ u64 roll_key(RollSt& s, const u32* __restrict__ F, const u32* __restrict__ R, u64 T, u32 k, const RollK& K, u64 p, bool& o) {
  if (!s.have) {                                     // direct evaluation (run start, or first k-mer of a sequence)
    u32 f1 = 0, f2 = 0, r1 = 0, r2 = 0, wd = 0, q1 = 1, q2 = 1;
    for (u32 i = 0; i < k; ++i) {
      const u64 a = p + i;
      if (i == 0 || (a & 15) == 0) wd = __ldg(F + (a >> 4));
      const u32 x = (wd >> (30 - 2 * (u32)(a & 15))) & 3u;
      f1 = f1 * K.B[0] + x; f2 = f2 * K.B[1] + x;
      r1 += (3u - x) * q1; r2 += (3u - x) * q2;
      q1 *= K.B[0]; q2 *= K.B[1];
    }
    s.f1 = f1; s.f2 = f2; s.r1 = r1; s.r2 = r2;
    const u64 e = p + k - 16;
    s.head = __funnelshift_l(__ldg(F + (p >> 4) + 1), __ldg(F + (p >> 4)), (u32)(p & 15) * 2);
    s.rct = revcomp16(__funnelshift_l(__ldg(F + (e >> 4) + 1), __ldg(F + (e >> 4)), (u32)(e & 15) * 2));
    s.have = true; s.fresh = true;
  } else {                                           // roll from p - 1
    const u32 out = s.head >> 30;
    const u64 ib = p + k - 1, ic = p + 15;
    if (s.fresh || (ib & 15) == 0) s.wb = __ldg(F + (ib >> 4));
    if (s.fresh || (ic & 15) == 0) s.wc = __ldg(F + (ic >> 4));
    s.fresh = false;
    const u32 in = (s.wb >> (30 - 2 * (u32)(ib & 15))) & 3u;
    const u32 co = 3u - out, ci = 3u - in;
    s.head = (s.head << 2) | ((s.wc >> (30 - 2 * (u32)(ic & 15))) & 3u);
    s.rct = (s.rct >> 2) | (ci << 30);               // the reverse complement gains the complement of the new base in front
    s.f1 = s.f1 * K.B[0] + in + out * K.negBk[0];
    s.f2 = s.f2 * K.B[1] + in + out * K.negBk[1];
    s.r1 = s.r1 * K.Binv[0] + (ci * K.Bk1[0] + co * K.negBinv[0]);
    s.r2 = s.r2 * K.Binv[1] + (ci * K.Bk1[1] + co * K.negBinv[1]);
  }
  const u32 lx = s.head, ly = s.rct;
  o = (lx != ly) ? (lx > ly) : (win_cmp(F, p, R, T - p - k, k) >= 0);
  return roll_finish(o ? s.f1 : s.r1, o ? s.f2 : s.r2);
}

// Out-of-line form for the rare positions that cannot take the fast path below (run starts, sequence ends): everything
// by value, so that the state stays in registers at the call sites.
struct RollRes { RollSt s; u64 h; u32 o; };
__device__ __noinline__ RollRes roll_key_slow(RollSt s, const u32* __restrict__ F, const u32* __restrict__ R, u64 T, u32 k, RollK K, u64 p) {
  RollRes r; bool o;
  r.h = roll_key(s, F, R, T, k, K, p, o);
  r.s = s; r.o = o ? 1u : 0u;
  return r;
}

// Run starts.  A thread's run starts at a 16-aligned position p0 >= 16; instead of a direct evaluation base by base
// (k steps of ~20 instructions) the state of the window at p0 - 1 is assembled from per-byte tables (4 bases per step):
// window = base p0-1 followed by the k-1 bases from p0 on, which start on a word boundary.  tab[byte] = {sum_i x_i B^(3-i)
// for the two multipliers (forward hash of the 4 bases), sum_j (3 - x_j) B^j (reverse-complement hash)}.
struct RollTab { u32 f1, f2, r1, r2; };
__device__ __forceinline__ RollTab roll_tab_entry(u32 byte, const RollK& K) {
  RollTab t; t.f1 = t.f2 = t.r1 = t.r2 = 0;
  u32 q1 = 1, q2 = 1;
  for (u32 i = 0; i < 4; ++i) {
    const u32 x = (byte >> (6 - 2 * i)) & 3u;
    t.f1 = t.f1 * K.B[0] + x; t.f2 = t.f2 * K.B[1] + x;
    t.r1 += (3u - x) * q1; t.r2 += (3u - x) * q2;
    q1 *= K.B[0]; q2 *= K.B[1];
  }
  return t;
}
// state of the window [p0 - 1, p0 - 1 + k): afterwards the k-mer at p0 is one rolled step away (s.have = true)
__device__ __forceinline__ void roll_init_before(RollSt& s, const u32* __restrict__ F, u64 p0, u32 k, const RollK& K, const RollTab* tab) {
  const u32* x = F + (p0 >> 4);
  u32 f1 = 0, f2 = 0, r1 = 0, r2 = 0, q1 = 1, q2 = 1;      // over the k - 1 bases from p0 on
  const u32 nb = (k - 1) >> 2;                             // whole bytes
  u32 wd = 0;
  for (u32 m = 0; m < nb; ++m) {
    if ((m & 3) == 0) wd = __ldg(x + (m >> 2));
    const RollTab t = tab[(wd >> (24 - 8 * (m & 3))) & 0xFFu];
    f1 = f1 * K.B4[0] + t.f1; f2 = f2 * K.B4[1] + t.f2;
    r1 += t.r1 * q1; r2 += t.r2 * q2;
    q1 *= K.B4[0]; q2 *= K.B4[1];
  }
  for (u32 i = nb * 4; i < k - 1; ++i) {
    if ((i & 15) == 0) wd = __ldg(x + (i >> 4));
    const u32 b = (wd >> (30 - 2 * (i & 15))) & 3u;
    f1 = f1 * K.B[0] + b; f2 = f2 * K.B[1] + b;
    r1 += (3u - b) * q1; r2 += (3u - b) * q2;
    q1 *= K.B[0]; q2 *= K.B[1];
  }
  const u32 xm = __ldg(x - 1) & 3u;                        // base p0 - 1 in front: q = B^(k-1) here
  s.f1 = xm * q1 + f1; s.f2 = xm * q2 + f2;
  s.r1 = (3u - xm) + K.B[0] * r1; s.r2 = (3u - xm) + K.B[1] * r2;
  s.head = __funnelshift_l(__ldg(x), __ldg(x - 1), 30);
  const u64 e = p0 - 1 + k - 16;
  s.rct = revcomp16(__funnelshift_l(__ldg(F + (e >> 4) + 1), __ldg(F + (e >> 4)), (u32)(e & 15) * 2));
  s.have = true; s.fresh = true;
}

// Fast form of roll_key for 8 consecutive positions [pblk + q0, + 8) of a 32-aligned block whose k-mers are all valid and
// whose predecessor's state is in `s`: the words of the two input streams are preloaded (Blk32), and with the loops
// unrolled every shift amount is a compile-time constant.
struct Blk32 { u32 hw[3], tw[2]; };                  // bases [pblk, +48) and [pblk + k - 1, +32)
__device__ __forceinline__ void blk32_load(Blk32& b, const u32* __restrict__ F, u64 pblk, u32 k) {
  const u32* x = F + (pblk >> 4);
  b.hw[0] = __ldg(x); b.hw[1] = __ldg(x + 1); b.hw[2] = __ldg(x + 2);
  const u64 e = pblk + k - 1;
  const u32* y = F + (e >> 4);
  const u32 sh = (u32)(e & 15) * 2;
  const u32 y0 = __ldg(y), y1 = __ldg(y + 1), y2 = __ldg(y + 2);
  b.tw[0] = __funnelshift_l(y1, y0, sh); b.tw[1] = __funnelshift_l(y2, y1, sh);
}
__device__ __forceinline__ void roll_fast8(RollSt& s, const Blk32& b, const RollK& K, const u32* __restrict__ F, const u32* __restrict__ R, u64 T, u32 k,
                                           u64 pblk, const u32 q0, u64 (&h)[8], u32& ow) {
#pragma unroll
  for (u32 j = 0; j < 8; ++j) {
    const u32 q = q0 + j, w = q >> 4, q16 = q & 15;
    const u32 out = s.head >> 30;
    const u32 in = (b.tw[w] >> (30 - 2 * q16)) & 3u;
    const u32 co = out ^ 3u, ci = in ^ 3u;
    s.head = q16 ? __funnelshift_l(b.hw[w + 1], b.hw[w], 2 * q16) : b.hw[w];
    s.rct = __funnelshift_r(s.rct, ci, 2);           // (ci : rct) >> 2
    s.f1 = s.f1 * K.B[0] + in + out * K.negBk[0];
    s.f2 = s.f2 * K.B[1] + in + out * K.negBk[1];
    s.r1 = s.r1 * K.Binv[0] + (ci * K.Bk1[0] + co * K.negBinv[0]);
    s.r2 = s.r2 * K.Binv[1] + (ci * K.Bk1[1] + co * K.negBinv[1]);
    const bool o = (s.head != s.rct) ? (s.head > s.rct) : (win_cmp(F, pblk + q, R, T - (pblk + q) - k, k) >= 0);
    h[j] = roll_finish(o ? s.f1 : s.r1, o ? s.f2 : s.r2);
    ow |= (o ? 1u : 0u) << q;
  }
  s.fresh = true;                                    // the cached words of the generic path are stale now
}

// The hash bits are consumed from the top: own_bits (owner GPU) | l1bits (level-1 bucket) | 32 bits kept in the record
// (of which the level-2 partition consumes the first l2bits).
__device__ __forceinline__ u32 top_bits(u64 x, u32 nbits) { return nbits ? (u32)(x >> (64 - nbits)) : 0u; }
__device__ __forceinline__ u32 top_bits32(u32 hi, u32 nbits) { return (hi >> 1) >> (31 - nbits); }   // nbits = 0 .. 31, no select

// ------------------------------------------------------------------------------------ P1: rolling hashes + level-1 partition
// Thread g of the grid walks the positions [pos_lo + g * run, + run), run a multiple of 32; pos_lo a multiple of 256;
// nothing is written for positions >= pos_end (a multiple of 32).  MG (multi-GPU, owner computes): only the k-mers whose
// top own_bits hash bits equal own_id are kept.  A CTA stages the records of T1 x 8 positions per group.  Every bucket
// has `copies` regions (region index = copy << l1bits | bucket), CTA x appends to copy x % copies: all CTAs hammering
// the same 2^l1bits cursors made the atomic round trip (in the critical path of every group) several microseconds.
#define SPB_P1_GS 8
template <u32 T1, u32 MINB, bool MG> __global__ void SPB_LAUNCH_BOUNDS2(T1, MINB)
k_roll_partition(const u32* __restrict__ F, const u32* __restrict__ R, const u32* __restrict__ vbits, u64 T, u32 k, u64 pos_lo, u64 pos_hi, u64 pos_end,
                 u32 run, u64 n_runs, u32 own_bits, u32 own_id, u32 l1bits, u32 copies, u64* __restrict__ rec1, u64 cap1, u64 trash, u32* cursor1,
                 u32* __restrict__ obits, RollK K, u32* err) {
  // trash: index of T1 x 8 spare records behind the regions -- the runs of an overflowing bucket go there (and the
  // overflow is flagged), so that the copy-out loop needs no test
  constexpr u32 GS = SPB_P1_GS;
  const u64 g = SPB_GTID();
  const bool alive = g < n_runs;
  const u64 p0 = pos_lo + g * run;
  RollSt st; st.have = false; st.fresh = false; st.f1 = st.f2 = st.r1 = st.r2 = st.head = st.rct = st.wb = st.wc = 0;
#ifdef SPB_EMU
  if (!alive) return;
  if (p0 >= 16 && p0 < T) {
    RollTab tab[256];
    for (u32 b = 0; b < 256; ++b) tab[b] = roll_tab_entry(b, K);
    roll_init_before(st, F, p0, k, K, tab);
  }
  for (u32 w = 0; w < run / 32; ++w) {
    const u64 pw = p0 + 32ull * w;
    if (pw >= pos_end) break;
    u32 vb = pw < T ? vbits[pw >> 5] : 0u, ow = 0;
    if (pw + 32 > pos_hi) vb = pw >= pos_hi ? 0u : (vb & ((1u << (u32)(pos_hi - pw)) - 1u));
    Blk32 blk;
    if (vb == 0xFFFFFFFFu) blk32_load(blk, F, pw, k);
    for (u32 half = 0; half < 4; ++half) {
      u64 h[8]; u32 vm = (vb >> (8 * half)) & 0xFFu;
      if (vm == 0xFFu && st.have && vb == 0xFFFFFFFFu) roll_fast8(st, blk, K, F, R, T, k, pw, 8 * half, h, ow);
      else for (u32 j = 0; j < 8; ++j) {
        if ((vm >> j) & 1u) { bool o; h[j] = roll_key(st, F, R, T, k, K, pw + 8 * half + j, o); ow |= (o ? 1u : 0u) << (8 * half + j); }
        else st.have = false;
      }
      for (u32 j = 0; j < 8; ++j) if ((vm >> j) & 1u) {
        if (MG && top_bits(h[j], own_bits) != own_id) continue;
        const u64 hh = MG ? h[j] << own_bits : h[j];
        const u32 b = (((u32)(g / blockDim.x) % copies) << l1bits) + top_bits(hh, l1bits);
        const u32 i = atomicAdd(cursor1 + b, 1u);
        if (i < cap1) rec1[(u64)b * cap1 + i] = ((hh << l1bits) & 0xFFFFFFFF00000000ull) | (u32)(pw + 8 * half + j); else atomicOr(err, (u32)ERR_OVERFLOW);
        (void)trash;
      }
    }
    obits[pw >> 5] = ow;
  }
#else
  constexpr u32 NS = T1 * GS;                        // records staged per group (<= 8192)
  constexpr u32 PER = (SPB_PART_MAXB + T1 - 1) / T1; // buckets per thread in the scan
  static_assert((NS & (NS - 1)) == 0 && NS <= 8192, "staging size");
  extern __shared__ __align__(128) u64 sorted[];     // NS staged records: [32 hash bits | bucket << 16 | index inside the group]
  __shared__ u64 gdst[SPB_PART_MAXB];                // per bucket: global index of its run minus its offset in `sorted`
  __shared__ u32 hist[SPB_PART_MAXB], base[SPB_PART_MAXB];
  __shared__ u32 wsum[32];
  const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const u32 NB = 1u << l1bits;
  const u32 cta_p0 = (u32)(pos_lo + (u64)blockIdx.x * T1 * run);
  const u32 reg0 = (blockIdx.x % copies) << l1bits;   // this CTA's copy of the bucket regions (fewer CTAs per cursor)
  __shared__ RollTab rtab[256];
  for (u32 b = tid; b < SPB_PART_MAXB; b += T1) hist[b] = 0;
  for (u32 b = tid; b < 256; b += T1) rtab[b] = roll_tab_entry(b, K);
  __syncthreads();
  if (alive && p0 >= 16 && p0 < T) roll_init_before(st, F, p0, k, K, rtab);
  for (u32 w = 0; w < run / 32; ++w) {
    const u64 pw = p0 + 32ull * w;
    u32 vb = (alive && pw < T) ? __ldg(vbits + (pw >> 5)) : 0u, ow = 0;
    if (pw + 32 > pos_hi) vb = pw >= pos_hi ? 0u : (vb & ((1u << (u32)(pos_hi - pw)) - 1u));
    Blk32 blk;
    if (vb == 0xFFFFFFFFu) blk32_load(blk, F, pw, k);
#pragma unroll
    for (u32 half = 0; half < 4; ++half) {
      const u32 vm = (vb >> (8 * half)) & 0xFFu;
      u32 ehi[GS], elo[GS], rk[GS], km = 0;          // km: records kept (valid, and owned when MG)
      {
        u64 h[GS];
        // ---- A: hashes; staged entry = [32 hash bits | bucket | index inside the group]; rank from the shared histogram
        if (vm == 0xFFu && st.have && vb == 0xFFFFFFFFu) roll_fast8(st, blk, K, F, R, T, k, pw, 8 * half, h, ow);
        else {
#pragma unroll
          for (u32 j = 0; j < GS; ++j) {
            h[j] = 0;
            if ((vm >> j) & 1u) { const RollRes r = roll_key_slow(st, F, R, T, k, K, pw + 8 * half + j); st = r.s; h[j] = r.h; ow |= r.o << (8 * half + j); }
            else st.have = false;
          }
        }
#pragma unroll
        for (u32 j = 0; j < GS; ++j) {
          u32 hi = (u32)(h[j] >> 32), lo = (u32)h[j];
          bool keep = (vm >> j) & 1u;
          if (MG) { keep = keep && top_bits32(hi, own_bits) == own_id; hi = __funnelshift_l(lo, hi, own_bits); lo <<= own_bits; }
          const u32 b = top_bits32(hi, l1bits);
          ehi[j] = __funnelshift_l(lo, hi, l1bits);
          elo[j] = (b << 16) | (tid * GS + j);
          rk[j] = 0;
          if (keep) { rk[j] = atomicAdd(&hist[b], 1u); km |= 1u << j; }
        }
      }
      __syncthreads();
      // ---- B: one global atomic per non-empty bucket reserves its run; exclusive scan of the histogram
      u32 cnt[PER], gb[PER], sum = 0;
#pragma unroll
      for (u32 i = 0; i < PER; ++i) {
        const u32 b = tid * PER + i;
        cnt[i] = (b < NB) ? hist[b] : 0u; gb[i] = 0;
        if (cnt[i]) gb[i] = atomicAdd(cursor1 + reg0 + b, cnt[i]);   // consumed after phase C: the round trip overlaps with it
        sum += cnt[i];
      }
      u32 incl = sum;
#pragma unroll
      for (u32 d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
      if (lane == 31) wsum[wid] = incl;
      __syncthreads();
      u32 off = incl - sum;
      for (u32 x = 0; x < wid; ++x) off += wsum[x];
#pragma unroll
      for (u32 i = 0; i < PER; ++i) {
        const u32 b = tid * PER + i;
        if (b < SPB_PART_MAXB) {
          base[b] = off;
          off += cnt[i];
          hist[b] = 0;
        }
      }
      __syncthreads();
      // ---- C: records to their sorted place in shared memory
#pragma unroll
      for (u32 j = 0; j < GS; ++j)
        if ((km >> j) & 1u) sorted[base[elo[j] >> 16] + rk[j]] = ((u64)ehi[j] << 32) | elo[j];
#pragma unroll
      for (u32 i = 0; i < PER; ++i) {
        const u32 b = tid * PER + i;
        if (b < SPB_PART_MAXB) {
          u64 gd = (u64)(reg0 + b) * cap1 + gb[i] - base[b];
          if (cnt[i] && (u64)gb[i] + cnt[i] > cap1) { gd = trash; atomicOr(err, (u32)ERR_OVERFLOW); }   // trash + i stays inside the spare records
          gdst[b] = gd;
        }
      }
      __syncthreads();
      // ---- D: contiguous runs out (consecutive threads write consecutive records of a bucket)
      u32 tot = 0;
      for (u32 x = 0; x < T1 / 32; ++x) tot += wsum[x];
      const u32 pgrp = cta_p0 + 32u * w + 8u * half;
#pragma unroll
      for (u32 it = 0; it < GS; ++it) {
        const u32 i = it * T1 + tid;
        if (i < tot) {
          const uint2 e = *reinterpret_cast<const uint2*>(sorted + i);
          const u32 loc = e.x & 0xFFFFu;
          const u64 gd = gdst[e.x >> 16];
          rec1[gd + i] = ((u64)e.y << 32) | (pgrp + (loc / GS) * run + (loc % GS));
        }
      }
      __syncthreads();
    }
    if (alive && pw < pos_end) obits[pw >> 5] = ow;
  }
#endif
}

// ------------------------------------------------------------------------------------ P2: level-2 partition
// Input regions r = 0 .. n_regions-1: rec_in[r * cap_in ..), cnt_in[r] records; parent bucket of region r = r % parent_mod
// (single GPU: parent_mod = number of level-1 buckets; multi-GPU owner: the regions of the G sources follow each other).
// Output bucket = parent << l2bits | (top l2bits of the record), regions of cap2 records with cursors.
// GPU form: persistent CTAs (two per SM).  Chunk c + 1 is brought into the second staging buffer by a bulk asynchronous
// copy (cp.async.bulk, completion on an mbarrier) while chunk c is partitioned, and the round trip of the cursor atomics
// overlaps with the scatter into `sorted2`.
__global__ void SPB_LAUNCH_BOUNDS2(SPB_P2_T, 2)
k_partition2(const u64* __restrict__ rec_in, u64 cap_in, const u32* __restrict__ cnt_in, u32 chunks_per_region, u32 n_regions, u32 parent_mod, u32 l2bits,
             u64* __restrict__ rec2, u64 cap2, u64 trash, u32* cursor2, u32* err, u32 skip2) {
  // skip2: stored hash bits already consumed by an earlier level-2 pass (very large inputs run this kernel twice)
  // trash: index of SPB_P2_CHUNK spare records behind the regions (runs of an overflowing bucket; the overflow is flagged)
#ifdef SPB_EMU
  const u32 region = blockIdx.x / chunks_per_region, chunk = blockIdx.x % chunks_per_region;
  const u64 n = (u64)cnt_in[region] < cap_in ? (u64)cnt_in[region] : cap_in;
  const u64 start = (u64)chunk * SPB_P2_CHUNK;
  if (start >= n) return;
  const u32 m = n - start < (u64)SPB_P2_CHUNK ? (u32)(n - start) : (u32)SPB_P2_CHUNK;
  const u64* src = rec_in + (u64)region * cap_in + start;
  const u64 out0 = (u64)(region % parent_mod) << l2bits;
  for (u32 j = 0; j < SPB_P2_PT; ++j) {
    const u32 i = j * SPB_P2_T + threadIdx.x;
    if (i >= m) continue;
    const u64 r = src[i];
    const u64 ob = out0 + top_bits(r << skip2, l2bits);
    const u32 s = atomicAdd(cursor2 + ob, 1u);
    if (s < cap2) rec2[ob * cap2 + s] = r; else atomicOr(err, (u32)ERR_OVERFLOW);
  }
  (void)trash; (void)n_regions;
#else
  extern __shared__ __align__(128) u64 p2smem[];     // stage[2][CHUNK] | sorted2[CHUNK]
  u64* stage = p2smem;
  u64* sorted2 = p2smem + 2 * SPB_P2_CHUNK;
  __shared__ u64 gdst[SPB_PART_MAXB];
  __shared__ u32 hist[SPB_PART_MAXB], base[SPB_PART_MAXB];
  __shared__ u32 wsum[32];
  __shared__ __align__(8) u64 bar[2];
  __shared__ u32 s_m[2];                             // records of the chunk in each staging buffer (0: nothing was copied)
  __shared__ u32 s_region[2];
  const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const u32 NB = 1u << l2bits;
  const u64 total = (u64)n_regions * chunks_per_region;
  hist[tid] = 0;
  // producer (thread 0): copy of chunk id `c` into buffer `buf`
  auto region_count = [&](u64 c) -> u32 { return c < total ? __ldg(cnt_in + (u32)(c / chunks_per_region)) : 0u; };
  auto issue = [&](u64 c, u32 buf, u32 cnt_region) {  // cnt_region = region_count(c), loaded one iteration ahead
    u32 m = 0, region = 0;
    if (c < total) {
      region = (u32)(c / chunks_per_region);
      const u64 start = (c % chunks_per_region) * SPB_P2_CHUNK;
      const u64 n = (u64)cnt_region < cap_in ? (u64)cnt_region : cap_in;
      if (start < n) {
        m = n - start < (u64)SPB_P2_CHUNK ? (u32)(n - start) : (u32)SPB_P2_CHUNK;
        const u32 bytes = (m * 8u + 15u) & ~15u;
        mbar_expect_tx(bar + buf, bytes);
        bulk_g2s(stage + buf * SPB_P2_CHUNK, rec_in + (u64)region * cap_in + start, bytes, bar + buf);
      }
    }
    s_m[buf] = m; s_region[buf] = region;
  };
  if (tid == 0) {
    mbar_init(bar, 1); mbar_init(bar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    issue(blockIdx.x, 0, region_count(blockIdx.x));
  }
  u32 cnt_pf = tid == 0 ? region_count((u64)blockIdx.x + gridDim.x) : 0u;
  __syncthreads();
  u32 par = 0;                                       // bit b: parity of the next completion of barrier b
  u32 it = 0;
  for (u64 c = blockIdx.x; c < total; c += gridDim.x, ++it) {
    const u32 buf = it & 1;
    if (tid == 0) {                                  // the other buffer was released by the barrier that ended the previous iteration
      issue(c + gridDim.x, buf ^ 1, cnt_pf);
      cnt_pf = region_count(c + 2ull * gridDim.x);   // consumed one iteration later
    }
    const u32 m = s_m[buf];
    if (m) {                                         // uniform
      mbar_wait(bar + buf, (par >> buf) & 1u); par ^= 1u << buf;
      const u64 out0 = (u64)(s_region[buf] % parent_mod) << l2bits;
      const u64* st = stage + buf * SPB_P2_CHUNK;
      uint2 rec[SPB_P2_PT]; u32 rk[SPB_P2_PT];
#pragma unroll
      for (u32 j = 0; j < SPB_P2_PT; ++j) { const u32 i = j * SPB_P2_T + tid; rec[j] = i < m ? *reinterpret_cast<const uint2*>(st + i) : make_uint2(0u, 0u); }
#pragma unroll
      for (u32 j = 0; j < SPB_P2_PT; ++j) { const u32 i = j * SPB_P2_T + tid; rk[j] = i < m ? atomicAdd(&hist[top_bits32(rec[j].y << skip2, l2bits)], 1u) : 0u; }
      __syncthreads();
      const u32 b = tid;                             // SPB_P2_T == SPB_PART_MAXB: one bucket per thread
      const u32 cnt = b < NB ? hist[b] : 0u;
      u32 gb = 0;
      if (cnt) gb = atomicAdd(cursor2 + out0 + b, cnt);   // consumed after the scatter below
      u32 incl = cnt;
#pragma unroll
      for (u32 d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
      if (lane == 31) wsum[wid] = incl;
      __syncthreads();
      u32 off = incl - cnt;
      for (u32 w = 0; w < wid; ++w) off += wsum[w];
      base[b] = off;
      hist[b] = 0;
      __syncthreads();
#pragma unroll
      for (u32 j = 0; j < SPB_P2_PT; ++j) { const u32 i = j * SPB_P2_T + tid; if (i < m) *reinterpret_cast<uint2*>(sorted2 + base[top_bits32(rec[j].y << skip2, l2bits)] + rk[j]) = rec[j]; }
      u64 gd = (out0 + b) * cap2 + gb - off;
      if (cnt && (u64)gb + cnt > cap2) { gd = trash; atomicOr(err, (u32)ERR_OVERFLOW); }
      gdst[b] = gd;
      __syncthreads();
#pragma unroll
      for (u32 j = 0; j < SPB_P2_PT; ++j) {
        const u32 i = j * SPB_P2_T + tid;
        if (i < m) {
          const uint2 r = *reinterpret_cast<const uint2*>(sorted2 + i);
          *reinterpret_cast<uint2*>(rec2 + gdst[top_bits32(r.y << skip2, l2bits)] + i) = r;
        }
      }
    }
    __syncthreads();                                 // end of the iteration: every buffer of this chunk may be reused
  }
#endif
}

// ------------------------------------------------------------------------------------ dedup: one CTA per final bucket

// One probe of record `rec` (index i of the bucket, table value v = tag | i) at slot s: installs it, or recognises an
// earlier/later instance of the same k-mer (exact 2k-bit verification) and leaves a link; returns false when the slot
// holds another k-mer (the caller moves on to the next slot of the probe sequence).
// (second half of dd_probe, also used after the warp-cooperative verification) the k-mers at p (ours, table value v) and
// q (installed in slot s as `old`) are equal: the earlier position keeps the slot, the other one gets a link
template <bool LIST> __device__ __forceinline__ void
dd_link(u32* tab, const u64* recs, u32 s, u32 v, u32 old, u64 p, u64 q, u32* __restrict__ seenpos, u32* dupbits, u64* __restrict__ links, u64* nlinks, u64 lcap) {
  u64 from = p, to = q;                                // default: we link to the installed record
  if (p < q) {                                         // ours is the earlier position: take the slot over
    u32 cur = old; u64 pc = q;
    for (;;) {
      const u32 prev = atomicCAS(&tab[s], cur, v);
      if (prev == cur) { from = pc; to = p; break; }   // we displaced it: it links to us
      cur = prev; pc = (u32)recs[cur & 0x1FFFu];       // another instance of the k-mer got in between
      if (pc < p) { to = pc; break; }
    }
  }
  if (LIST) {
    const u64 slot = warp_append(nlinks, true);
    if (slot < lcap) links[slot] = (from << 32) | to;
  } else {
    seenpos[from] = (u32)to;
    atomicOr(dupbits + (from >> 5), 1u << (from & 31));
  }
}
template <bool LIST, bool OPT> __device__ __forceinline__ bool
dd_probe(u32* tab, const u64* recs, u32 s, u32 tag, u32 v, u64 rec, const u32* __restrict__ F, const u32* __restrict__ R, const u32* __restrict__ obits, u64 T, u32 k,
         u32* __restrict__ seenpos, u32* dupbits, u64* __restrict__ links, u64* nlinks, u64 lcap, u32 hmask) {
  const u32 old = atomicCAS(&tab[s], SPB_D_EMPTY, v);
  if (old == SPB_D_EMPTY) return true;
  if ((old >> 13) != tag) return false;
  const u64 rq = recs[old & 0x1FFFu];
  if ((((u32)(rq >> 32) ^ (u32)(rec >> 32)) & hmask) != 0) return false;         // all stored hash bits agree: verify the 2k bits
  const u64 p = (u32)rec, q = (u32)rq;
  if (OPT) {                                           // optimistic form: linked now, verified by k_verify_links afterwards
    dd_link<LIST>(tab, recs, s, v, old, p, q, seenpos, dupbits, links, nlinks, lcap);
    return true;
  }
  const bool o = get_bit(obits, p), oq = get_bit(obits, q);   // canonical orientations, left by k_roll_partition
  if (win_cmp(o ? F : R, o ? p : T - p - k, oq ? F : R, oq ? q : T - q - k, k) != 0) return false;
  dd_link<LIST>(tab, recs, s, v, old, p, q, seenpos, dupbits, links, nlinks, lcap);
  return true;
}
#ifndef SPB_EMU
// all 32 lanes: are the canonical k-mers at p and q equal?  (one limb per lane: one round trip instead of a serial loop)
__device__ __forceinline__ bool warp_same_kmer(const u32* __restrict__ F, const u32* __restrict__ R, const u32* __restrict__ obits, u64 T, u32 k, u64 p, u64 q) {
  const u32 lane = threadIdx.x & 31;
  const bool o = get_bit(obits, p), oq = get_bit(obits, q);
  const u32* A = o ? F : R; const u64 a = o ? p : T - p - k;
  const u32* B = oq ? F : R; const u64 b = oq ? q : T - q - k;
  const u32 W = nlimbs(k);
  bool neq = false;
  for (u32 j = lane; j < W; j += 32) {
    u32 la = limb_at(A + (a >> 4), (u32)(a & 15) * 2, j), lb = limb_at(B + (b >> 4), (u32)(b & 15) * 2, j);
    if (j == W - 1) { const u32 m = tail_mask(k); la &= m; lb &= m; }
    neq |= la != lb;
  }
  return !__any_sync(0xffffffffu, neq);
}
#endif
// slot / tag / step of a record: of the 32 stored hash bits the level-2 partition consumed the first skip_bits (equal for
// the whole bucket); the next 13 are the home slot, the rest the tag
__device__ __forceinline__ void dd_hash(u64 rec, u32 skip_bits, u32 i, u32& s, u32& tag, u32& v, u32& step, u32 hmask) {
  const u32 h = ((u32)(rec >> 32) & hmask) << skip_bits;
  s = h >> 19; tag = h & 0x7FFFFu; v = (tag << 13) | i;
  step = ((h >> 5) & (SPB_D_SLOTS - 1)) | 1u;          // double hashing: no clusters
}

// LIST = false: links written in place (seenpos[from] = to, dupbits); LIST = true: appended to `links` (from << 32 | to).
// Two passes over the bucket: (1) every record tries its home slot once -- no loop, no divergence; about three quarters
// are placed; the others go to a list in shared memory; (2) the listed records probe on, lane-persistent (a lane takes
// its next record as soon as the current one is done), so that the warps of pass 2 are full of records that need work.
// GPU form: persistent CTAs (two per SM); the records of the next bucket arrive in the second buffer by a bulk
// asynchronous copy (completion on an mbarrier) while the current bucket is inserted.
// NBUF = 2: persistent, two CTAs per SM (double-buffered records); NBUF = 1: one bucket per CTA, three CTAs per SM (the
// copy of a bucket overlaps with the work of the two other resident CTAs only).
#define SPB_D_SMEM_U64(NBUF) (SPB_D_SLOTS / 2 + (NBUF) * SPB_D_LIMIT + SPB_D_DEFER / 4 + 8)
// OPT (in-place form, one GPU): a record whose stored hash bits (and bucket: 9 + 32 = 41 bits at config 3) equal those of the installed one
// is linked WITHOUT looking at the genome; k_verify_links checks every link afterwards in one streaming pass (a link whose
// predecessor position is linked to the neighbouring target costs one base compare) and k_repair_* re-homes the few
// positions whose known bits collided with another k-mer (5.7e4 of 5e8 records at config 3).  hmask: stored hash bits in use (all; tests clear some to provoke
// collisions).
template <bool LIST, int NBUF, bool OPT> __global__ void SPB_LAUNCH_BOUNDS2(SPB_D_T, (NBUF == 1 ? 3 : 2))
k_bucket_dedup(const u64* __restrict__ recs_g, u64 cap, const u32* __restrict__ cnt, u32 n_buckets, u32 skip_bits, const u32* __restrict__ F, const u32* __restrict__ R,
               const u32* __restrict__ obits, u64 T, u32 k, u32* __restrict__ seenpos, u32* dupbits, u32* err, u64* __restrict__ links, u64* nlinks, u64 lcap, u32 hmask) {
  SPB_DYN_SHARED(u64, dsm, SPB_D_SMEM_U64(2));
  u32* tab = (u32*)dsm;
  u64* recs0 = dsm + SPB_D_SLOTS / 2;
  unsigned short* defer = (unsigned short*)(recs0 + NBUF * SPB_D_LIMIT);
  u64* bar = recs0 + NBUF * SPB_D_LIMIT + SPB_D_DEFER / 4;      // two mbarriers
  u32* ctl = (u32*)(bar + 2);                                // [0] ndefer, [1], [2] records in buffer 0 / 1
#ifdef SPB_EMU
  const u32 n = cnt[blockIdx.x];
  if (n == 0) return;                                // uniform over the block
  if (n > SPB_D_LIMIT || n > cap) { if (threadIdx.x == 0) atomicOr(err, (u32)ERR_OVERFLOW); return; }
  const u64* src = recs_g + (u64)blockIdx.x * cap;
  u64* recs = recs0;
  u32* ndefer = ctl;
  SPB_PHASE_BEGIN(0)
    if (threadIdx.x == 0) { for (u32 i = 0; i < n; ++i) recs[i] = src[i]; *ndefer = 0; }
    for (u32 s = threadIdx.x * 4; s < SPB_D_SLOTS; s += blockDim.x * 4) { tab[s] = SPB_D_EMPTY; tab[s + 1] = SPB_D_EMPTY; tab[s + 2] = SPB_D_EMPTY; tab[s + 3] = SPB_D_EMPTY; }
  SPB_PHASE_END
  SPB_PHASE_BEGIN(1)
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) {
      const u64 rec = recs[i];
      u32 s, tag, v, step;
      dd_hash(rec, skip_bits, i, s, tag, v, step, hmask);
      if (atomicCAS(&tab[s], SPB_D_EMPTY, v) != SPB_D_EMPTY) {
        const u32 d = atomicAdd(ndefer, 1u);
        if (d < SPB_D_DEFER) defer[d] = (unsigned short)i;
        else for (;;) {                                // list full (a bucket of repeats): finish this record here
          if (dd_probe<LIST, OPT>(tab, recs, s, tag, v, rec, F, R, obits, T, k, seenpos, dupbits, links, nlinks, lcap, hmask)) break;
          s = (s + step) & (SPB_D_SLOTS - 1);
        }
      }
    }
  SPB_PHASE_END
  SPB_PHASE_BEGIN(2)
    const u32 nd = *ndefer < SPB_D_DEFER ? *ndefer : SPB_D_DEFER;
    for (u32 d = threadIdx.x; d < nd; d += blockDim.x) {
      const u32 i = defer[d];
      const u64 rec = recs[i];
      u32 s, tag, v, step;
      dd_hash(rec, skip_bits, i, s, tag, v, step, hmask);
      while (!dd_probe<LIST, OPT>(tab, recs, s, tag, v, rec, F, R, obits, T, k, seenpos, dupbits, links, nlinks, lcap, hmask)) s = (s + step) & (SPB_D_SLOTS - 1);
    }
  SPB_PHASE_END
  (void)n_buckets;
#else
  const u32 tid = threadIdx.x;
  // producer (thread 0): records of bucket b into buffer buf; nb = cnt[b], loaded one iteration ahead
  auto issue = [&](u32 b, u32 buf, u32 nb) {
    u32 m = 0;
    if (b < n_buckets) {
      if (nb > SPB_D_LIMIT || nb > cap) atomicOr(err, (u32)ERR_OVERFLOW);
      else if (nb) {
        m = nb;
        const u32 bytes = ((m * 8u) + 15u) & ~15u;     // regions are 16-byte aligned and padded: reading one record past n is harmless
        mbar_expect_tx(bar + buf, bytes);
        bulk_g2s(recs0 + buf * SPB_D_LIMIT, recs_g + (u64)b * cap, bytes, bar + buf);
      }
    }
    ctl[1 + buf] = m;
  };
  auto bucket_count = [&](u64 b) -> u32 { return b < n_buckets ? __ldg(cnt + b) : 0u; };
  if (tid == 0) {
    mbar_init(bar, 1); mbar_init(bar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    issue(blockIdx.x, 0, bucket_count(blockIdx.x));
  }
  u32 cnt_pf = (NBUF == 2 && tid == 0) ? bucket_count((u64)blockIdx.x + gridDim.x) : 0u;
  u32 par = 0, it = 0;
  for (u32 b = blockIdx.x; b < n_buckets; b += gridDim.x, ++it) {
    const u32 buf = NBUF == 1 ? 0u : (it & 1);
    if (tid == 0) {                                  // the other buffer was released by the barrier that ended the previous iteration
      if (NBUF == 2) {
        issue(b + gridDim.x, buf ^ 1, cnt_pf);
        cnt_pf = bucket_count((u64)b + 2ull * gridDim.x);
      } else if (it) issue(b, 0, bucket_count(b));
      ctl[0] = 0;
    }
    {
      uint4* t4 = reinterpret_cast<uint4*>(tab);
      const uint4 e4 = make_uint4(SPB_D_EMPTY, SPB_D_EMPTY, SPB_D_EMPTY, SPB_D_EMPTY);
#pragma unroll
      for (u32 s = 0; s < SPB_D_SLOTS / 4 / SPB_D_T; ++s) t4[s * SPB_D_T + tid] = e4;
    }
    __syncthreads();
    const u32 n = ctl[1 + buf];
    if (n) {                                         // uniform
      const u64* recs = recs0 + buf * SPB_D_LIMIT;
      mbar_wait(bar + buf, (par >> buf) & 1u); par ^= 1u << buf;
      // ---- pass 1: home slot, once
      const u32 lane = tid & 31;
      for (u32 i0 = 0; i0 < n; i0 += SPB_D_T) {
        const u32 i = i0 + tid;
        u64 rec = 0; u32 s = 0, tag = 0, v = 0, step = 1;
        bool fail = false;
        if (i < n) {
          rec = recs[i];
          dd_hash(rec, skip_bits, i, s, tag, v, step, hmask);
          fail = atomicCAS(&tab[s], SPB_D_EMPTY, v) != SPB_D_EMPTY;
        }
        const u32 bal = __ballot_sync(0xffffffffu, fail);          // one list reservation per warp
        if (bal) {
          u32 d0 = 0;
          if (lane == (u32)__ffs(bal) - 1) d0 = atomicAdd(ctl, (u32)__popc(bal));
          d0 = __shfl_sync(0xffffffffu, d0, __ffs(bal) - 1);
          if (fail) {
            const u32 d = d0 + __popc(bal & ((1u << lane) - 1u));
            if (d < SPB_D_DEFER) defer[d] = (unsigned short)i;
            else for (;;) {                            // list full (a bucket of repeats): finish this record here
              if (dd_probe<LIST, OPT>(tab, recs, s, tag, v, rec, F, R, obits, T, k, seenpos, dupbits, links, nlinks, lcap, hmask)) break;
              s = (s + step) & (SPB_D_SLOTS - 1);
            }
          }
        }
      }
      __syncthreads();
      // ---- pass 2: the deferred records probe on.  A probe that meets a record with the same 32 hash bits asks for the
      // exact comparison, which the warp does together (one limb per lane), so that the rare repeats do not stall it.
      const u32 nd = ctl[0] < SPB_D_DEFER ? ctl[0] : SPB_D_DEFER;
      u32 d = tid, s = 0, tag = 0, v = 0, step = 1;
      u64 rec = 0;
      bool active = false;
      for (;;) {
        if (!active && d < nd) {
          const u32 i = defer[d];
          rec = recs[i];
          dd_hash(rec, skip_bits, i, s, tag, v, step, hmask);
          d += SPB_D_T;
          active = true;
        }
        bool cand = false; u32 old = 0; u64 rq = 0;
        if (active) {
          old = atomicCAS(&tab[s], SPB_D_EMPTY, v);
          if (old == SPB_D_EMPTY) active = false;
          else {
            if ((old >> 13) == tag) { rq = recs[old & 0x1FFFu]; cand = (((u32)(rq >> 32) ^ (u32)(rec >> 32)) & hmask) == 0; }
            if (!cand) s = (s + step) & (SPB_D_SLOTS - 1);
          }
        }
        if (OPT) {                                   // optimistic: every candidate is linked at once, no genome access
          if (cand) { dd_link<LIST>(tab, recs, s, v, old, (u32)rec, (u32)rq, seenpos, dupbits, links, nlinks, lcap); active = false; }
        }
        u32 cm = OPT ? 0u : __ballot_sync(0xffffffffu, cand);
        while (cm) {
          const u32 l = __ffs(cm) - 1; cm &= cm - 1;
          const u64 pp = __shfl_sync(0xffffffffu, (u32)rec, l), qq = __shfl_sync(0xffffffffu, (u32)rq, l);
          const bool same = warp_same_kmer(F, R, obits, T, k, pp, qq);
          if (lane == l) {
            if (same) { dd_link<LIST>(tab, recs, s, v, old, pp, qq, seenpos, dupbits, links, nlinks, lcap); active = false; }
            else s = (s + step) & (SPB_D_SLOTS - 1);
          }
        }
        if (__all_sync(0xffffffffu, !active && d >= nd)) break;
      }
    }
    __syncthreads();                                 // end of the iteration: table, list and this buffer may be reused
  }
#endif
}


// ------------------------------------------------------------------------------------ optimistic links: verification + repair
// canonical k-mers at p and q equal?  (exact: 2k bits of the packed genome, orientation bits from k_roll_partition)
__device__ __forceinline__ bool kmer_equal(const u32* __restrict__ F, const u32* __restrict__ R, const u32* __restrict__ obits, u64 T, u32 k, u64 p, u64 q) {
  const bool o = get_bit(obits, p), oq = get_bit(obits, q);
  return win_cmp(o ? F : R, o ? p : T - p - k, oq ? F : R, oq ? q : T - q - k, k) == 0;
}
// Can the link p -> q be checked against the link of p - 1?  Yes if p - 1 is linked to the position next to q on the matching
// side with the same strand relation: the two k-mers then overlap their (verified) predecessors in k - 1 bases and only the
// new base has to agree -- exact by induction along the run, whose first link is always compared in full.
__device__ __forceinline__ bool link_extends(const u32* __restrict__ seenpos, const u32* __restrict__ dupbits, const u32* __restrict__ obits, u64 p, u64 q, bool* same_out) {
  const bool same = get_bit(obits, p) == get_bit(obits, q);
  *same_out = same;
  if (p == 0 || !get_bit(dupbits, p - 1)) return false;
  const u64 q1 = seenpos[p - 1];
  if ((get_bit(obits, p - 1) == get_bit(obits, q1)) != same) return false;
  return same ? (q1 + 1 == q) : (q1 == q + 1);
}
// one linked position: check against the predecessor's link where possible, in full otherwise; a failed link is flagged
// (failbits) and listed
__device__ __forceinline__ void verify_one(u64 p, const u32* __restrict__ seenpos, const u32* __restrict__ dupbits, const u32* __restrict__ F, const u32* __restrict__ R,
                                           const u32* __restrict__ obits, u64 T, u32 k, u32* failbits, u32* __restrict__ flist, u64* nfail, u64 fcap) {
  const u64 q = seenpos[p];
  bool same;
  const bool ext = link_extends(seenpos, dupbits, obits, p, q, &same);
  bool ok;
  if (ext) ok = same ? base_at(F, p + k - 1) == base_at(F, q + k - 1) : base_at(F, p + k - 1) == (3u - base_at(F, q));
  else ok = kmer_equal(F, R, obits, T, k, p, q);
  if (!ok) {
    atomicOr(failbits + (p >> 5), 1u << (p & 31));
    const u64 i = atomicAdd(nfail, 1ull);
    if (i < fcap) flist[i] = (u32)p;
  }
}
// every linked position (the bitmap of linked positions is sparse on mostly-unique inputs: 32 words per warp iteration, only
// the non-zero ones are looked at)
__global__ void k_verify_links(const u32* __restrict__ seenpos, const u32* __restrict__ dupbits, u64 nwords, const u32* __restrict__ F, const u32* __restrict__ R,
                               const u32* __restrict__ obits, u64 T, u32 k, u32* failbits, u32* __restrict__ flist, u64* nfail, u64 fcap) {
  const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
  const u32 lane = threadIdx.x & 31;
#ifdef SPB_EMU
  for (u64 wd = SPB_GTID() >> 5; wd < nwords; wd += nwarps)
    if ((dupbits[wd] >> lane) & 1u) verify_one(wd * 32 + lane, seenpos, dupbits, F, R, obits, T, k, failbits, flist, nfail, fcap);
#else
  for (u64 wb = (SPB_GTID() >> 5) * 32; wb < nwords; wb += nwarps * 32) {
    const u32 d = wb + lane < nwords ? __ldg(dupbits + wb + lane) : 0u;
    u32 nz = __ballot_sync(0xffffffffu, d != 0);
    while (nz) {
      const u32 j = (u32)__ffs((int)nz) - 1;
      nz &= nz - 1;
      const u32 dj = __shfl_sync(0xffffffffu, d, j);
      if ((dj >> lane) & 1u) verify_one((wb + j) * 32 + lane, seenpos, dupbits, F, R, obits, T, k, failbits, flist, nfail, fcap);
    }
  }
#endif
}
// A link that was checked against its predecessor's link is only as good as that one: from every failed position walk on
// while the next position leaned on the one before it, comparing in full; the first that holds on its own ends the walk.
// One thread per failure of the first pass (new failures are appended and handled by the thread that finds them).
__global__ void k_verify_followups(u32 nf0, const u32* __restrict__ seenpos, const u32* __restrict__ dupbits, const u32* __restrict__ F, const u32* __restrict__ R,
                                   const u32* __restrict__ obits, u64 T, u32 k, u32* failbits, u32* flist, u64* nfail, u64 fcap) {
  const u64 i = SPB_GTID();
  if (i >= nf0) return;
  for (u64 t = (u64)flist[i] + 1;; ++t) {
    if (!get_bit(dupbits, t)) break;
    const u64 q = seenpos[t];
    bool same;
    if (!link_extends(seenpos, dupbits, obits, t, q, &same)) break;        // compared in full already
    if (get_bit(failbits, t)) break;                                        // somebody else's walk starts here
    if (kmer_equal(F, R, obits, T, k, t, q)) break;
    atomicOr(failbits + (t >> 5), 1u << (t & 31));
    const u64 j = atomicAdd(nfail, 1ull);
    if (j < fcap) flist[j] = (u32)t;
  }
}
// Repair of the failed links (distinct k-mers whose known hash bits agree: 5.7e4 of 5e8 records at config 3).  The records of one
// hash class were all linked towards the class' slot holders (always the smallest position seen so far), so the true first
// position of a failed one is either a holder further down its target's chain or another failed position of the class.
__global__ void k_repair_roots(const u32* __restrict__ flist, u32 nf, const u32* __restrict__ seenpos, const u32* __restrict__ dupbits, u64* __restrict__ keys) {
  const u64 i = SPB_GTID();
  if (i >= nf) return;
  const u32 p = flist[i];
  u32 x = seenpos[p];
  while (get_bit(dupbits, x)) x = seenpos[x];
  keys[i] = ((u64)x << 32) | p;                        // (root of the hash class, position): sorted by the caller
}
__global__ void k_repair_find(const u64* __restrict__ keys, u32 nf, const u32* __restrict__ seenpos, const u32* __restrict__ dupbits,
                              const u32* __restrict__ F, const u32* __restrict__ R, const u32* __restrict__ obits, u64 T, u32 k, u32* __restrict__ newt) {
  const u64 i = SPB_GTID();
  if (i >= nf) return;
  const u32 p = (u32)keys[i], root = (u32)(keys[i] >> 32);
  u32 best = p;
  for (u32 x = seenpos[p];; x = seenpos[x]) {          // holders from the target down to the root: positions descend
    if (x < best && kmer_equal(F, R, obits, T, k, p, x)) best = x;
    if (!get_bit(dupbits, x)) break;
  }
  for (u64 j = i; j-- > 0 && (u32)(keys[j] >> 32) == root;) {             // the failed positions of the same class below p (sorted)
    const u32 f = (u32)keys[j];
    if (f < best && kmer_equal(F, R, obits, T, k, p, f)) best = f;
  }
  newt[i] = best;
}
__global__ void k_repair_apply(const u64* __restrict__ keys, u32 nf, const u32* __restrict__ newt, u32* __restrict__ seenpos, u32* dupbits) {
  const u64 i = SPB_GTID();
  if (i >= nf) return;
  const u32 p = (u32)keys[i];
  if (newt[i] != p) { seenpos[p] = newt[i]; return; }
  u32* wp = dupbits + (p >> 5);                        // its own k-mer's first position: no link
  const u32 bit = 1u << (p & 31);
  u32 old = *wp;
  while (true) { const u32 prev = atomicCAS(wp, old, old & ~bit); if (prev == old) break; old = prev; }
}
