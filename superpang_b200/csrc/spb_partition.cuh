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
#define SPB_P2_PT 16                 // records per thread: 8192-record chunks
#define SPB_P2_CHUNK (SPB_P2_T * SPB_P2_PT)
#define SPB_D_T 512
#define SPB_D_SLOTS 8192u            // 32 KB table
#define SPB_D_LIMIT 5120u            // records per final bucket (40 KB): table load <= 0.625; more -> overflow -> fallback
#define SPB_D_AVG 3840u              // target mean (load 0.47)
#define SPB_D_EMPTY 0xFFFFFFFFu

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

// The hash bits are consumed from the top: own_bits (owner GPU) | l1bits (level-1 bucket) | 32 bits kept in the record
// (of which the level-2 partition consumes the first l2bits).
__device__ __forceinline__ u32 top_bits(u64 x, u32 nbits) { return nbits ? (u32)(x >> (64 - nbits)) : 0u; }

// ------------------------------------------------------------------------------------ P1: rolling hashes + level-1 partition
// Thread g of the grid walks the positions [pos_lo + g * run, + run), run a multiple of 32 and of GS; pos_lo a multiple
// of 256.  own_bits > 0 (multi-GPU, owner computes): only the k-mers whose top own_bits hash bits equal own_id are kept.
template <u32 T1, u32 GS, u32 MINB> __global__ void SPB_LAUNCH_BOUNDS2(T1, MINB)
k_roll_partition(const u32* __restrict__ F, const u32* __restrict__ R, const u32* __restrict__ vbits, u64 T, u32 k, u64 pos_lo, u64 pos_hi,
                 u32 run, u64 n_runs, u32 own_bits, u32 own_id, u32 l1bits, u64* __restrict__ rec1, u64 cap1, u32* cursor1,
                 u32* __restrict__ obits, RollK K, u32* err) {
  const u64 g = SPB_GTID();
  const bool alive = g < n_runs;
  const u64 p0 = pos_lo + g * run;
  RollSt st; st.have = false; st.fresh = false; st.f1 = st.f2 = st.r1 = st.r2 = st.head = st.rct = st.wb = st.wc = 0;
#ifdef SPB_EMU
  if (!alive) return;
  u32 vb = 0, ow = 0;
  for (u32 j = 0; j < run; ++j) {
    const u64 p = p0 + j;
    if ((p & 31) == 0) { vb = p < T ? vbits[p >> 5] : 0u; ow = 0; }
    const bool valid = ((vb >> (p & 31)) & 1u) && p < pos_hi;
    if (valid) {
      bool o;
      const u64 h = roll_key(st, F, R, T, k, K, p, o);
      ow |= (o ? 1u : 0u) << (p & 31);
      if (top_bits(h, own_bits) == own_id) {
        const u64 hh = h << own_bits;
        const u32 b = top_bits(hh, l1bits);
        const u32 i = atomicAdd(cursor1 + b, 1u);
        if (i < cap1) rec1[(u64)b * cap1 + i] = ((hh << l1bits) & 0xFFFFFFFF00000000ull) | (u32)p; else atomicOr(err, (u32)ERR_OVERFLOW);
      }
    } else st.have = false;
    if ((p & 31) == 31) obits[p >> 5] = ow;
  }
#else
  constexpr u32 NS = T1 * GS;                        // records staged per group
  constexpr u32 LOCMASK = NS - 1;                    // NS is a power of two <= 8192
  constexpr u32 PER = (SPB_PART_MAXB + T1 - 1) / T1; // buckets per thread in the scan
  static_assert((NS & (NS - 1)) == 0 && NS <= 8192, "staging size");
  extern __shared__ __align__(128) u64 sorted[];     // NS staged records
  __shared__ u32 hist[SPB_PART_MAXB], base[SPB_PART_MAXB], gbase[SPB_PART_MAXB];
  __shared__ u32 wsum[32];
  const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const u32 NB = 1u << l1bits;
  const u64 cta_p0 = pos_lo + (u64)blockIdx.x * T1 * run;
  for (u32 b = tid; b < SPB_PART_MAXB; b += T1) hist[b] = 0;
  __syncthreads();
  u32 vb = 0, ow = 0;
  for (u32 grp = 0; grp < run / GS; ++grp) {
    const u64 pg = p0 + (u64)grp * GS;
    if (((grp * GS) & 31) == 0) { vb = (alive && pg < T) ? __ldg(vbits + (pg >> 5)) : 0u; ow = 0; }
    u64 ent[GS]; u32 rk[GS]; u32 vm = 0;
    // ---- A: staged entry = [hash bits | index inside the group]; rank inside its bucket from the shared histogram
#pragma unroll
    for (u32 j = 0; j < GS; ++j) {
      const u64 p = pg + j;
      const bool valid = ((vb >> (p & 31)) & 1u) && p < pos_hi;
      ent[j] = 0; rk[j] = 0;
      if (valid) {
        bool o;
        const u64 h = roll_key(st, F, R, T, k, K, p, o);
        ow |= (o ? 1u : 0u) << (p & 31);
        if (top_bits(h, own_bits) == own_id) {
          ent[j] = ((h << own_bits) & ~(u64)LOCMASK) | (tid * GS + j);
          rk[j] = atomicAdd(&hist[top_bits(ent[j], l1bits)], 1u);
          vm |= 1u << j;
        }
      } else st.have = false;
    }
    if (alive && (((grp + 1) * GS) & 31) == 0) obits[pg >> 5] = ow;
    __syncthreads();
    // ---- B: one global atomic per non-empty bucket reserves its run; exclusive scan of the histogram
    u32 cnt[PER], gb[PER], sum = 0;
#pragma unroll
    for (u32 i = 0; i < PER; ++i) {
      const u32 b = tid * PER + i;
      cnt[i] = (b < NB) ? hist[b] : 0u; gb[i] = 0;
      if (cnt[i]) gb[i] = atomicAdd(cursor1 + b, cnt[i]);
      sum += cnt[i];
    }
    u32 incl = sum;
#pragma unroll
    for (u32 d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    u32 off = incl - sum;
    for (u32 w = 0; w < wid; ++w) off += wsum[w];
#pragma unroll
    for (u32 i = 0; i < PER; ++i) {
      const u32 b = tid * PER + i;
      if (b < SPB_PART_MAXB) {
        base[b] = off; off += cnt[i];
        if (cnt[i] && (u64)gb[i] + cnt[i] > cap1) { gb[i] = 0xFFFFFFFFu; atomicOr(err, (u32)ERR_OVERFLOW); }
        gbase[b] = gb[i];
        hist[b] = 0;
      }
    }
    __syncthreads();
    // ---- C: records to their sorted place in shared memory
#pragma unroll
    for (u32 j = 0; j < GS; ++j)
      if ((vm >> j) & 1u) sorted[base[top_bits(ent[j], l1bits)] + rk[j]] = ent[j];
    __syncthreads();
    // ---- D: contiguous runs out (consecutive threads write consecutive records of a bucket)
    u32 tot = 0;
    for (u32 w = 0; w < T1 / 32; ++w) tot += wsum[w];
    for (u32 i = tid; i < tot; i += T1) {
      const u64 e = sorted[i];
      const u32 b = top_bits(e, l1bits), gbv = gbase[b], loc = (u32)e & LOCMASK;
      const u64 p = cta_p0 + (u64)(loc / GS) * run + grp * GS + (loc % GS);
      if (gbv != 0xFFFFFFFFu) rec1[(u64)b * cap1 + gbv + (i - base[b])] = ((e << l1bits) & 0xFFFFFFFF00000000ull) | (u32)p;
    }
    __syncthreads();
  }
#endif
}

// ------------------------------------------------------------------------------------ P2: level-2 partition
// Input regions r = 0 .. n_regions-1: rec_in[r * cap_in ..), cnt_in[r] records; parent bucket of region r = r % parent_mod
// (single GPU: parent_mod = number of level-1 buckets; multi-GPU owner: the regions of the G sources follow each other).
// Output bucket = parent << l2bits | (top l2bits of the record), regions of cap2 records with cursors.
__global__ void SPB_LAUNCH_BOUNDS2(SPB_P2_T, 2)
k_partition2(const u64* __restrict__ rec_in, u64 cap_in, const u32* __restrict__ cnt_in, u32 chunks_per_region, u32 parent_mod, u32 l2bits,
             u64* __restrict__ rec2, u64 cap2, u32* cursor2, u32* err) {
  const u32 region = blockIdx.x / chunks_per_region, chunk = blockIdx.x % chunks_per_region;
  const u64 n = (u64)cnt_in[region] < cap_in ? (u64)cnt_in[region] : cap_in;
  const u64 start = (u64)chunk * SPB_P2_CHUNK;
  if (start >= n) return;
  const u32 m = n - start < (u64)SPB_P2_CHUNK ? (u32)(n - start) : (u32)SPB_P2_CHUNK;
  const u64* src = rec_in + (u64)region * cap_in + start;
  const u64 out0 = (u64)(region % parent_mod) << l2bits;
#ifdef SPB_EMU
  for (u32 j = 0; j < SPB_P2_PT; ++j) {
    const u32 i = j * SPB_P2_T + threadIdx.x;
    if (i >= m) continue;
    const u64 r = src[i];
    const u64 ob = out0 + top_bits(r, l2bits);
    const u32 s = atomicAdd(cursor2 + ob, 1u);
    if (s < cap2) rec2[ob * cap2 + s] = r; else atomicOr(err, (u32)ERR_OVERFLOW);
  }
#else
  extern __shared__ __align__(128) u64 sorted2[];    // SPB_P2_CHUNK records
  __shared__ u32 hist[SPB_PART_MAXB], base[SPB_PART_MAXB], gbase[SPB_PART_MAXB];
  __shared__ u32 wsum[32];
  const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const u32 NB = 1u << l2bits;
  for (u32 b = tid; b < SPB_PART_MAXB; b += SPB_P2_T) hist[b] = 0;
  __syncthreads();
  u64 rec[SPB_P2_PT]; u32 rk[SPB_P2_PT];
#pragma unroll
  for (u32 j = 0; j < SPB_P2_PT; ++j) { const u32 i = j * SPB_P2_T + tid; rec[j] = i < m ? __ldcs(src + i) : 0ull; }
#pragma unroll
  for (u32 j = 0; j < SPB_P2_PT; ++j) { const u32 i = j * SPB_P2_T + tid; rk[j] = i < m ? atomicAdd(&hist[top_bits(rec[j], l2bits)], 1u) : 0u; }
  __syncthreads();
  const u32 b = tid;                                 // SPB_P2_T == SPB_PART_MAXB: one bucket per thread
  const u32 cnt = b < NB ? hist[b] : 0u;
  u32 gb = 0;
  if (cnt) gb = atomicAdd(cursor2 + out0 + b, cnt);
  u32 incl = cnt;
#pragma unroll
  for (u32 d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  u32 off = incl - cnt;
  for (u32 w = 0; w < wid; ++w) off += wsum[w];
  base[b] = off;
  if (cnt && (u64)gb + cnt > cap2) { gb = 0xFFFFFFFFu; atomicOr(err, (u32)ERR_OVERFLOW); }
  gbase[b] = gb;
  __syncthreads();
#pragma unroll
  for (u32 j = 0; j < SPB_P2_PT; ++j) { const u32 i = j * SPB_P2_T + tid; if (i < m) sorted2[base[top_bits(rec[j], l2bits)] + rk[j]] = rec[j]; }
  __syncthreads();
  for (u32 i = tid; i < m; i += SPB_P2_T) {
    const u64 r = sorted2[i];
    const u32 bb = top_bits(r, l2bits), gbv = gbase[bb];
    if (gbv != 0xFFFFFFFFu) rec2[(out0 + bb) * cap2 + gbv + (i - base[bb])] = r;
  }
#endif
}

// ------------------------------------------------------------------------------------ dedup: one CTA per final bucket
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

// skip_bits = hash bits of the record already consumed by the level-2 partition (they are the same for the whole bucket).
// LIST = false: links written in place (seenpos[from] = to, dupbits); LIST = true: appended to `links` (from << 32 | to).
template <bool LIST> __global__ void SPB_LAUNCH_BOUNDS2(SPB_D_T, 3)
k_bucket_dedup(const u64* __restrict__ recs_g, u64 cap, const u32* __restrict__ cnt, u32 skip_bits, const u32* __restrict__ F, const u32* __restrict__ R,
               u64 T, u32 k, u32* __restrict__ seenpos, u32* dupbits, u32* err, u64* __restrict__ links, u64* nlinks, u64 lcap) {
  SPB_DYN_SHARED(u64, dsm, SPB_D_SLOTS / 2 + SPB_D_LIMIT + 2);
  u32* tab = (u32*)dsm;
  u64* recs = dsm + SPB_D_SLOTS / 2;
  u64* bar = dsm + SPB_D_SLOTS / 2 + SPB_D_LIMIT;
  const u32 n = cnt[blockIdx.x];
  if (n == 0) return;                                // uniform over the block
  if (n > SPB_D_LIMIT || n > cap) { if (threadIdx.x == 0) atomicOr(err, (u32)ERR_OVERFLOW); return; }
  const u64* src = recs_g + (u64)blockIdx.x * cap;
  SPB_PHASE_BEGIN(0)
#ifdef SPB_EMU
    if (threadIdx.x == 0) for (u32 i = 0; i < n; ++i) recs[i] = src[i];
#else
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const u32 bytes = ((n * 8u) + 15u) & ~15u;       // regions are 16-byte aligned and padded: reading one record past n is harmless
      mbar_expect_tx(bar, bytes);
      bulk_g2s(recs, src, bytes, bar);
    }
#endif
    for (u32 s = threadIdx.x * 4; s < SPB_D_SLOTS; s += blockDim.x * 4) { tab[s] = SPB_D_EMPTY; tab[s + 1] = SPB_D_EMPTY; tab[s + 2] = SPB_D_EMPTY; tab[s + 3] = SPB_D_EMPTY; }
  SPB_PHASE_END
  SPB_PHASE_BEGIN(1)
#ifndef SPB_EMU
    mbar_wait(bar, 0);
#endif
    u32 i = threadIdx.x, s = 0, tag = 0, v = 0, step = 1;
    u64 rec = 0;
    bool active = false;
    for (;;) {
      if (!active && i < n) {                          // a lane takes its next record as soon as the current one is placed
        rec = recs[i];
        const u32 h = (u32)(rec >> 32) << skip_bits;   // remaining hash bits: 13 -> home slot, the rest -> tag
        s = h >> 19; tag = h & 0x7FFFFu; v = (tag << 13) | i;
        step = ((h >> 5) & (SPB_D_SLOTS - 1)) | 1u;    // double hashing: no clusters
        i += blockDim.x;
        active = true;
      }
      if (active) {
        const u32 old = atomicCAS(&tab[s], SPB_D_EMPTY, v);
        if (old == SPB_D_EMPTY) active = false;
        else {
          bool same = false;
          if ((old >> 13) == tag) {
            const u64 rq = recs[old & 0x1FFFu];
            if ((rq >> 32) == (rec >> 32)) {           // all 32 stored hash bits agree: verify the 2k bits
              const u64 p = (u32)rec, q = (u32)rq;
              const bool o = orient_fwd(F, R, T, k, p), oq = orient_fwd(F, R, T, k, q);
              if (win_cmp(o ? F : R, o ? p : T - p - k, oq ? F : R, oq ? q : T - q - k, k) == 0) {
                u64 from = p, to = q;                  // default: we link to the installed record
                if (p < q) {                           // ours is the earlier position: take the slot over
                  u32 cur = old; u64 pc = q;
                  for (;;) {
                    const u32 prev = atomicCAS(&tab[s], cur, v);
                    if (prev == cur) { from = pc; to = p; break; }       // we displaced it: it links to us
                    cur = prev; pc = (u32)recs[cur & 0x1FFFu];           // another instance of the k-mer got in between
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
                same = true;
              }
            }
          }
          if (same) active = false; else s = (s + step) & (SPB_D_SLOTS - 1);
        }
      }
      if (__all_sync(0xffffffffu, !active && i >= n)) break;
    }
  SPB_PHASE_END
}
