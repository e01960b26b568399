/*
 * superpang_b200 -- C-ABI of the B200-native DBG-build + NBP-compaction path.
 *
 * The reference (fpusan/SuperPang v1.3.1) has NO plugin / FFI interface for this path: the only
 * seam is the Python class constructed at `src/superpang/scripts/SuperPang.py:199`
 * (`Assembler(fasta, ksize, threads, diskdb, debug, gap_size).run(...)`).  Each entry point below
 * therefore cites the *reference code it replaces* (paths relative to `src/superpang/`), and
 * INTEGRATION.md shows the `ctypes` binding a maintainer adds to `lib/Assembler.py`.
 *
 * Conventions
 *   - plain pointers + sizes, no torch / C++ types; every function returns 0 (SPB_OK) or a
 *     negative error code and leaves a message retrievable with spb_last_error();
 *   - all work is enqueued on the CUDA stream handed to spb_create() (e.g. torch's current stream);
 *   - result getters copy into CALLER-OWNED buffers that may be device memory (torch tensors) or
 *     host memory (numpy arrays) -- cudaMemcpyDefault resolves the direction through UVA;
 *   - there is no CPU fallback: without a CUDA device spb_create() fails with SPB_ECUDA.
 *
 * Vertex ids are the reference's ids (order of first appearance, `lib/Assembler.py:186-195`),
 * sequences are indexed by their position in the reference's `seqDict` (`lib/Assembler.py:94-96`).
 */
#ifndef SUPERPANG_B200_H
#define SUPERPANG_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPB_OK            0
#define SPB_EINVAL       -1   /* even k, k < 3, empty input, non-ACGT base, T >= 2^32-64 ...      */
#define SPB_ENOMEM       -2
#define SPB_EDEGENERATE  -3   /* inputs on which the reference's own asserts fire     */
#define SPB_ECUDA        -4
#define SPB_ESTATE       -5   /* call order violated (e.g. spb_compact before spb_build_dbg)      */

#define SPB_NONE 0xFFFFFFFFu  /* "no vertex" marker in per-sequence tables                        */

typedef struct spb_ctx spb_ctx;

typedef struct spb_counts {
  uint64_t n_seq;            /* sequences handed to spb_load (incl. those with len <= k)           */
  uint64_t n_bases;          /* total bases T                                                      */
  uint64_t n_inst;           /* k-mer instances: sum over len > k of (len - k + 1)                 */
  uint64_t n_vertices;       /* distinct canonical k-mers                                          */
  uint64_t n_edges;          /* unique undirected edges incl. self-loops (np.unique rows)          */
  uint64_t n_junction;       /* directed junction entries kept after the implicit-edge filter      */
  uint64_t n_inits;          /* NBP-init vertices                                                  */
  uint64_t n_pieces;         /* maximal runs of consecutive-id extender vertices                   */
  uint64_t n_comp;           /* connected components of the DBG                                    */
  uint64_t n_cycle_comp;     /* components that are pure cycles (broken at one edge)               */
  uint64_t n_nbp;            /* raw NBPs, both orientations (after the terminal-duplicate filter)  */
  uint64_t nbp_vertices;     /* sum of NBP lengths in vertices                                     */
  uint64_t nbp_bases;        /* sum of NBP sequence lengths (k - 1 + L each)                       */
  uint64_t n_origin_pairs;   /* sum over NBPs of #origin sequences                                 */
  uint64_t n_seq_intervals;  /* total rows of all seqPaths interval lists                          */
  uint64_t table_slots;      /* capacity of the k-mer hash table                                   */
} spb_counts;

/* Create / destroy a context bound to one CUDA device and one stream (NULL = legacy default
 * stream).  Replaces the reference's fork-pool "runtime" (`lib/Assembler.py:37-74`, `:135-148`). */
int  spb_create(spb_ctx** out, int device, void* cuda_stream);
void spb_destroy(spb_ctx* ctx);
const char* spb_last_error(const spb_ctx* ctx);

/* Hand over the N-split input sequences (what `read_fasta(..., Ns='split')` returned,
 * `lib/utils.py:5-30`, `lib/Assembler.py:94`): upper-case ASCII A/C/G/T only, concatenated, with
 * seq_off[n_seq + 1] base offsets.  Sequences with len <= k must be passed too so that `ref`
 * numbering matches (`lib/Assembler.py:150-153`).  `ascii` may be host (pinned or pageable) or
 * device memory.  2-bit packs both strands into HBM (replaces `Compressor.compress`,
 * `lib/Compressor.pyx:32-58`, and `reverse_complement`, `lib/cutils.pyx:16-28`). */
int spb_load(spb_ctx* ctx, const uint8_t* ascii, const uint64_t* seq_off, uint32_t n_seq, uint32_t k);

/* DBG build: canonical k-mer extraction, exact dedup with reference-identical vertex numbering,
 * per-instance vertex ids, (k-1)-overlap edges.  Replaces `Assembler.__init__`,
 * `lib/Assembler.py:150-252` (seq2hashes `:978-992`, the dict loop `:186-217`, np.unique `:247-249`). */
int spb_build_dbg(spb_ctx* ctx, spb_counts* out);

/* NBP compaction: init detection, chain contraction + list ranking, cycle breaking, NBP vertex
 * lists, NBP sequences, NBP->origin sets, connected components.  Replaces the head of
 * `Assembler.run`, `lib/Assembler.py:269-363` (is_NBP_init `:995-1035`, reconstruct_sequence
 * `:1038-1059`) and the origin lookup `get_ori2cov_onevertex` + goodOris (`:1212-1262`, `:944-946`). */
int spb_compact(spb_ctx* ctx, spb_counts* out);

/* ---- multi-GPU building blocks (one process per GPU; every rank loads the whole input) ---------
 * K-mers are hash-partitioned over the ranks; the host side (torch.distributed / NCCL) performs the
 * collectives between the stages on the device pointers these functions hand out:
 *   spb_mg_records   records (64-bit hash | orientation, position) of the positions [pos_lo, pos_hi), sorted by
 *                    owner; owner_counts[n_owners] = split sizes for the all-to-all
 *   spb_mg_dedup     owner side: exact dedup of the received records; rep_out[i] = first position of record i
 *   spb_mg_scatter   position side: representatives (returned by the reverse all-to-all, same order as sent)
 *                    -> first-appearance bits + per-position representative for the local slice
 *   spb_mg_buffers   device pointers of the first-appearance bitmap, orientation bitmap and per-position array
 *                    (all-gathered in place by the host, slice = spb_mg_slice positions per rank)
 *   spb_mg_ids       vertex ids for the local slice + the vertex table (after the bitmap all-gather)
 *   spb_mg_graph     junction edges / explicit edge list (after the all-gather of the vertex ids);
 *                    leaves the context in the same state as spb_build_dbg, spb_compact follows unchanged.
 * Replaces the reference's single-process dict loop, `lib/Assembler.py:186-217`, for n_ranks > 1. */
int spb_mg_slice(spb_ctx* ctx, uint32_t n_ranks, uint64_t* slice_len);
/* Duplicate profile of the first two 4 M-position chunks (fractions of chunk 1 that extend an anchor match / that are
 * duplicates); the host side picks the return protocol from it: links if almost nothing repeats, representatives otherwise. */
int spb_mg_probe(spb_ctx* ctx, double* ext_ratio, double* dup_ratio);
int spb_mg_records(spb_ctx* ctx, uint64_t pos_lo, uint64_t pos_hi, uint32_t n_owners, uint64_t* owner_counts,
                   void** keys_dev, void** pos_dev);
int spb_mg_dedup(spb_ctx* ctx, const uint64_t* keys, const uint32_t* pos, const uint64_t* chunk_counts,
                 uint32_t n_chunks, uint32_t n_owners, uint32_t owner, uint32_t* rep_out);
int spb_mg_scatter(spb_ctx* ctx, const uint32_t* rep);
/* "links" protocol (default): instead of one representative per record the owner returns only the exceptions --
 * (position << 32 | first position) for every instance that is not a first appearance -- sorted by position, with
 * send_counts[n_owners] per destination rank (position slices); the received records may be reordered in place.
 * spb_mg_apply_links consumes the links a rank receives for its slice (replaces spb_mg_dedup + spb_mg_scatter). */
int spb_mg_dedup_links(spb_ctx* ctx, uint64_t* keys_io, uint32_t* pos_io, uint64_t n_total, uint32_t n_owners,
                       uint64_t* send_counts, void** links_dev);
int spb_mg_apply_links(spb_ctx* ctx, const uint64_t* links, uint64_t n, uint64_t pos_lo, uint64_t pos_hi);
/* "Owner computes" protocol (default when the probe finds few duplicates): no record travels.  Rank `rank` of
 * `n_owners` walks all positions itself, keeps the k-mers whose hash it owns and deduplicates them; the output has
 * the form of spb_mg_dedup_links (links sorted by position, send_counts[r] = links for the position slice of rank r)
 * and is consumed by spb_mg_apply_links after the all-to-all.  Replaces spb_mg_records + the record exchange +
 * spb_mg_dedup_links; the orientation bitmap is complete on every rank afterwards (no all-gather needed for it).
 * Same reference lines as spb_build_dbg (lib/Assembler.py:148-201). */
int spb_mg_owned_links(spb_ctx* ctx, uint32_t rank, uint32_t n_owners, uint64_t* send_counts, void** links_dev);
/* "Part" protocol (default for mostly-unique inputs; round 2): the record exchange on the hand-written partition kernels.
 * spb_mg_part_records turns the local slice into 8-byte records partitioned by (owner rank, level-1 bucket): recs_dev =
 * [n_owners][*n_regions_per_owner regions] of fixed capacity (*region_records records per owner in total), counts_dev =
 * [n_owners][*n_regions_per_owner] uint32 -- both exchanged by equal-split all-to-alls.  The host then all-gathers the
 * orientation bitmap (spb_mg_buffers) and calls spb_mg_part_dedup on the received regions ([n_src] blocks in source-rank
 * order), which returns links in the form of spb_mg_dedup_links.  *overflow = 1 (from either call, on any rank) means the
 * input does not suit the fixed-capacity regions (masses of identical k-mers): every rank must then fall back to
 * spb_mg_records / spb_mg_dedup_links.  Replaces, on this path, the reference's single-process dict loop
 * (lib/Assembler.py:186-217) across GPUs. */
int spb_mg_part_records(spb_ctx* ctx, uint64_t pos_lo, uint64_t pos_hi, uint32_t n_owners, uint64_t* region_records,
                        uint32_t* n_regions_per_owner, void** recs_dev, void** counts_dev, int* overflow);
int spb_mg_part_dedup(spb_ctx* ctx, const uint64_t* recs, const uint32_t* counts, uint32_t n_src, uint32_t n_owners,
                      uint64_t* send_counts, void** links_dev, int* overflow);
int spb_mg_buffers(spb_ctx* ctx, void** fbits_dev, void** obits_dev, void** sp_dev);
int spb_mg_ids(spb_ctx* ctx, uint64_t pos_lo, uint64_t pos_hi);
int spb_mg_graph(spb_ctx* ctx, spb_counts* out);

/* ---- sliced multi-GPU tail (round 2): graph build + NBP compaction partitioned over the ranks --------------------------
 * Follows the "part" dedup when every rank holds all links (spb_mg_apply_links over [0, T)): fbits / obits / links are
 * complete everywhere and nothing of size N_inst is exchanged any more.  Rank r computes
 *   - vertex ids, junction entries and origin pairs of ITS positions (slice r of spb_mg_slice),
 *   - implicit-edge flags of all vertices (from the replicated bitmap), init / piece flags, the sorted init / piece-bound
 *     lists, explicit edge rows and the vertex -> NBP map of ITS vertices (first appearances of its positions, cut at
 *     multiples of 8),
 *   - vertex lists and sequences of ITS raw NBPs (id ranges with equal shares of the output),
 * the host side all-gathers the short lists in between (NCCL: junction entries, inits, piece bounds, origin pairs; the
 * exchange of the edge endpoints and of the compaction state that the reference does in one process,
 * lib/Assembler.py:199-249, :269-363), and every rank computes the contracted graph (pieces, darts, list ranking, NBP
 * table), the origin sets and the components -- all of size O(#NBPs) -- itself.  Call order:
 *   spb_mg_slice_ids -> spb_mg_slice_junctions -> [all-gather J] -> spb_mg_slice_graph -> { spb_mg_slice_classify ->
 *   [all-gather lists, sum of edge counts] -> spb_mg_slice_contract } (again while *again) -> spb_mg_slice_emit ->
 *   [all-gather pairs, OR of has01] -> spb_mg_slice_finish.
 * All list pointers are device pointers; the library copies what it is given.  Results stay sliced: spb_mg_shard hands
 * out the local parts (which: 0 per-position vertex ids, 1 vertex2coords rows, 2 edge rows, 3 NBP vertex lists, 4 NBP
 * sequences; first / count in elements (rows) of the complete array, edge rows counted from this rank's first row);
 * counts, offsets, components, origins and the NBP table are complete on every rank.  spb_mg_complete recomputes the
 * sliced passes over the whole input on the calling rank (no communication), after which every getter works as on one
 * GPU; getters that need sliced arrays return SPB_ESTATE before that. */
int spb_mg_slice_ids(spb_ctx* ctx, uint32_t rank, uint32_t n_ranks);
int spb_mg_slice_junctions(spb_ctx* ctx, uint64_t* n, void** jk_dev, void** js_dev);
int spb_mg_slice_graph(spb_ctx* ctx, const uint64_t* jk_all, const uint8_t* js_all, uint64_t n_all, uint64_t* n_edges_local);
int spb_mg_slice_classify(spb_ctx* ctx, uint32_t* n_init, uint32_t* n_start, uint32_t* n_end, void** init_dev, void** plo_dev, void** phi_dev);
int spb_mg_slice_contract(spb_ctx* ctx, const uint32_t* init_all, uint32_t n_init, const uint32_t* plo_all, const uint32_t* phi_all,
                          uint32_t n_piece, uint64_t n_edges_total, int pass, int* again);
int spb_mg_slice_emit(spb_ctx* ctx, uint64_t* n_op, void** op_dev, uint64_t* n_mp, void** mp_dev, void** has01_dev);
int spb_mg_slice_finish(spb_ctx* ctx, const uint64_t* op_all, uint64_t n_op, const uint64_t* mp_all, uint64_t n_mp,
                        const uint8_t* has01_any, spb_counts* out);
int spb_mg_shard(spb_ctx* ctx, int which, void** dev, uint64_t* first, uint64_t* count);
/* spb_mg_shard + copy of the local part to its place in the complete array at host_base (asynchronous on the context's
 * stream when the host memory is page-locked, e.g. shared memory registered by every rank); *bytes = size of the part. */
int spb_mg_shard_to_host(spb_ctx* ctx, int which, void* host_base, uint64_t* bytes);
int spb_mg_complete(spb_ctx* ctx, spb_counts* out);

/* Page-locked host buffer that is to receive the NBP sequences (the bytes spb_get_nbp_seqs returns; what the CLI writes
 * to NBPs.fasta, scripts/SuperPang.py:257-268).  When set, spb_compact starts the device-to-host copy on a second stream as
 * soon as the sequences are emitted, overlapping it with the origin sets, components and edge rows; a later
 * spb_get_nbp_seqs(ctx, off, the same pointer) only waits for that copy.  capacity in bytes; nullptr clears. */
int spb_set_host_output(spb_ctx* ctx, uint8_t* nbp_seqs_host, uint64_t capacity);

/* ---- getters (sizes come from spb_counts; buffers are caller-owned, host or device) ---------- */

/* uint32[n_vertices][2] = (ref, idx) of each vertex's first appearance (`lib/Assembler.py:194`). */
int spb_get_vertex2coords(spb_ctx* ctx, uint32_t* out);
/* uint32[n_edges][2], rows (min,max) in lexicographic order (`lib/Assembler.py:247-249`). */
int spb_get_edges(spb_ctx* ctx, uint32_t* out);
/* uint32[n_bases]: vertex id of the k-mer instance starting at every base position, SPB_NONE where
 * no k-mer starts (the per-sequence `sp` arrays of `lib/Assembler.py:183-197`, concatenated). */
int spb_get_instance_vertices(spb_ctx* ctx, uint32_t* out);
/* uint32[n_seq][2]: first / last vertex of each sequence (`seqLimitsPerSeq`, `lib/Assembler.py:206-211`);
 * SPB_NONE for sequences with len <= k. */
int spb_get_seqlimits(spb_ctx* ctx, uint32_t* out);
/* seqPaths (`compress_vertices`, `lib/vtools.pyx:18-49`, incl. its leading-zero quirk):
 * off uint64[n_seq + 1] row offsets, intervals uint32[n_seq_intervals][2]. */
int spb_get_seq_intervals(spb_ctx* ctx, uint64_t* off, uint32_t* intervals);
/* int32[n_vertices]: connected-component label, numbered by lowest vertex id
 * (`label_components`, `lib/Assembler.py:271`). */
int spb_get_components(spb_ctx* ctx, int32_t* out);
/* NBP vertex lists: off uint64[n_nbp + 1], verts uint32[nbp_vertices], comp int32[n_nbp],
 * is_cycle uint8[n_nbp] (`lib/Assembler.py:327-354`).  Any pointer may be NULL to skip it. */
int spb_get_nbps(spb_ctx* ctx, uint64_t* off, uint32_t* verts, int32_t* comp, uint8_t* is_cycle);
/* NBP sequences: off uint64[n_nbp + 1], ascii uint8[nbp_bases] (`lib/Assembler.py:357`). */
int spb_get_nbp_seqs(spb_ctx* ctx, uint64_t* off, uint8_t* ascii);
/* NBP origins: off uint64[n_nbp + 1], seq_idx uint32[n_origin_pairs] (sorted per NBP; indices into
 * the loaded sequence list; the caller cuts names at `_Nsplit_`, `lib/Assembler.py:946`). */
int spb_get_nbp_origins(spb_ctx* ctx, uint64_t* off, uint32_t* seq_idx);
/* NBP graph: successors of every raw NBP = the NBPs whose first k bases equal its last k bases
 * (`lib/Assembler.py:365-384`), as CSR (off uint64[n_nbp + 1], succ uint32[*n_edges]), and rev uint32[n_nbp] = the id
 * of the reverse-orientation partner of every NBP (the pairs of `lib/Assembler.py:405-455`).  Call once with NULL
 * buffers to learn *n_edges. */
int spb_get_nbp_graph(spb_ctx* ctx, uint64_t* n_edges, uint64_t* off, uint32_t* succ, uint32_t* rev);
/* Orientation counts: counts uint32[n_nbp], counts[i] = number of loaded sequences whose string
 * contains the sequence of raw NBP i, i.e. that walk the NBP in its own orientation -- what the reference accumulates
 * into nv2rightOrientation from get_seqPaths_NBPs (`lib/Assembler.py:1121-1139`, `:448-453`), including the effect of
 * the compress_vertices leading-zero quirk on its vertex_overlap pre-filter (`lib/vtools.pyx:29-45`). */
int spb_get_nbp_orientation_counts(spb_ctx* ctx, uint32_t* counts);
/* The same relation per sequence, as CSR: off uint64[n_seq + 1], nbp_ids uint32[*n_pairs] (ascending per sequence) = the
 * raw NBPs that sequence s contains in its own orientation -- the sets `get_seqPaths_NBPs` returns
 * (`lib/Assembler.py:1121-1139`), which the forward / reverse separation consumes (`:448`, `:480-520`).  Call once with
 * NULL buffers to learn *n_pairs. */
int spb_get_seq_nbps(spb_ctx* ctx, uint64_t* n_pairs, uint64_t* off, uint32_t* nbp_ids);
/* Bubble candidates: group uint32[n_nbp]; NBPs that share (first vertex, last vertex) with at least
 * one other NBP carry the smallest NBP id of their group, all others 0xFFFFFFFF -- the `path_limits` grouping of
 * `lib/Assembler.py:728-734` (the minimap2 alignment of the members, `:749-758`, stays on the host). */
int spb_get_bubble_groups(spb_ctx* ctx, uint32_t* group);
/* Sequences of host-supplied vertex paths: `reconstruct_sequence` (`lib/Assembler.py:1038-1118`) for
 * the joined paths of the host tail (`:911-921`).  off uint64[n_paths + 1], verts uint32[off[n_paths]] (host);
 * ascii uint8[off[n_paths] + n_paths * (k - 1)], path i written at off[i] + i * (k - 1).  Needs spb_build_dbg only.
 * SPB_EINVAL when a path has fewer than 2 vertices or steps between non-adjacent vertices (the reference asserts). */
int spb_path_sequences(spb_ctx* ctx, uint32_t n_paths, const uint64_t* off, const uint32_t* verts, uint8_t* ascii);
/* Origins of host-supplied vertex paths that are concatenations of whole NBPs: the goodOris of the
 * joined paths -- `get_ori2cov_onevertex` (`lib/Assembler.py:1212-1248`) + `:941-945`, before the `_Nsplit_` cut and
 * without `bubble_vertex2origins` (the caller unions those in).  off / verts as for spb_path_sequences.  Result as CSR:
 * out_off uint64[n_paths + 1], seq_idx uint32[cap] (ascending sequence indices per path); *n_pairs = number of pairs;
 * when it exceeds cap only out_off is filled: call again with cap >= *n_pairs.  Needs spb_compact. */
int spb_path_origins(spb_ctx* ctx, uint32_t n_paths, const uint64_t* off, const uint32_t* verts, uint64_t cap, uint64_t* n_pairs,
                     uint64_t* out_off, uint32_t* seq_idx);
/* Extended forms (completed in round 2).  kind 0: paths are vertex lists (as above); kind 1: paths are lists
 * of RAW NBP IDS (indices into spb_get_nbps; consecutive NBPs share their junction vertex, the way the reference joins them,
 * `lib/Assembler.py:810`) and are expanded to vertex lists on the device -- no vertex ever crosses PCIe.
 * spb_path_sequences_ex: vert_off (optional, host uint64[n_paths + 1]) receives the vertex offsets of the paths; ascii ==
 * NULL asks for the sizes only; path i is written at vert_off[i] + i * (k - 1).
 * spb_path_origins_ex: additionally takes `bubble_vertex2origins` (`lib/Assembler.py:922`, `:1236`) as a CSR over ascending
 * vertex ids -- bub_vert uint32[n_bub], bub_off uint64[n_bub + 1], bub_seq uint32 (sequence indices), host arrays -- whose
 * lists are united with the origins of every vertex the rule looks at, as the reference does. */
int spb_path_sequences_ex(spb_ctx* ctx, int kind, uint32_t n_paths, const uint64_t* off, const uint32_t* items, uint64_t* vert_off, uint8_t* ascii);
int spb_path_origins_ex(spb_ctx* ctx, int kind, uint32_t n_paths, const uint64_t* off, const uint32_t* items, uint32_t n_bub,
                        const uint32_t* bub_vert, const uint64_t* bub_off, const uint32_t* bub_seq, uint64_t cap, uint64_t* n_pairs,
                        uint64_t* out_off, uint32_t* seq_idx);
/* FASTA reader + N-splitter: native form of the reference's `read_fasta(fasta, ambigs='as_Ns',
 * Ns='split', gap_size)` (`lib/utils.py:5-30`, `:52-59`): whitespace / '.' / '-' removed, upper case, U -> T, IUPAC
 * ambiguity codes -> N, any other character is an error, records split at runs of gap_size Ns into `{name}_Nsplit_{i}`,
 * duplicated names are an error.  Fills the buffers spb_load takes; release with spb_free_fasta.  names: n_seq names,
 * each followed by '\n'.  err (optional) receives the message of a refused input. */
typedef struct spb_fasta {
  uint8_t* ascii; uint64_t* seq_off; char* names;
  uint32_t n_seq; uint64_t n_bases; uint64_t names_bytes;
} spb_fasta;
int spb_read_fasta(const char* path, uint32_t gap_size, spb_fasta* out, char* err, uint64_t err_cap);
void spb_free_fasta(spb_fasta* fa);

/* ---- output writers (writer half) -------------------------------------------------------------------
 * spb_write_fasta: ">{name}\n{seq}\n" per record in the order given -- byte for byte what the reference's write_fasta
 * (lib/utils.py:62-65) writes for a dict, i.e. the format of NBPs.fasta, graph.fastg, *.core.fasta and *.aux.fasta
 * (scripts/SuperPang.py:279-282).  names: the n names back to back with name_off[n + 1] (host); seqs / seq_off[n + 1] may be
 * HOST or DEVICE memory (the CSR of spb_get_nbp_seqs or spb_path_sequences left on the device): the sequences are streamed
 * to the file in 64 MB chunks.
 * spb_write_origins_tsv: NBP2origins.tsv and the edge table (scripts/SuperPang.py:284-294): `header`, then per node
 * "{node}\t{origins ','-joined}\t{their bins ','-joined}\t{distinct bins}/{n_genomes}\n"; origins as a CSR of indices
 * into the source-sequence name table (origin_off / origin_idx: host or device, the form of spb_get_nbp_origins). */
int spb_write_fasta(const char* path, uint64_t n, const char* names, const uint64_t* name_off, const uint8_t* seqs, const uint64_t* seq_off,
                    char* err, uint64_t err_cap);
int spb_write_origins_tsv(const char* path, const char* header, uint64_t n, const char* node_names, const uint64_t* node_off,
                          const uint64_t* origin_off, const uint32_t* origin_idx, uint32_t n_orig, const char* orig_names,
                          const uint64_t* orig_name_off, const uint32_t* orig_bin, const char* bin_names, const uint64_t* bin_name_off,
                          uint32_t n_bins, uint32_t n_genomes, char* err, uint64_t err_cap);
/* Current counters (n_seq_intervals is filled by the first spb_get_seq_intervals call, which may
 * be made with NULL buffers just to size them). */
int spb_get_counts(spb_ctx* ctx, spb_counts* out);
/* Per-stage device times of the last spb_build_dbg + spb_compact (ms, CUDA events) as a JSON
 * string written into buf (truncated to cap). */
int spb_get_stage_stats(spb_ctx* ctx, char* buf, uint64_t cap);

/* Number of this library's own kernel launches since it was loaded (CUB launches not counted). */
uint64_t spb_launch_count(void);
/* Library / build identification ("cuda sm_100a" for the product; the CPU kernel-logic emulator
 * used by the not-gpu test tier answers "emu"). */
const char* spb_build_kind(void);

#ifdef __cplusplus
}
#endif
#endif
