"""
ctypes binding of the C-ABI declared in `include/superpang_b200.h`.

The product library is `superpang_b200/libsuperpang_b200.so`, built by nvcc for sm_100a
(`__graft_entry__.build()`).  There is NO CPU fallback: if the library is missing, or no CUDA device
is present, construction fails loudly.  (`tests/emu/` builds a host emulation of the kernels for the
CPU-only test tier and injects it through the `lib=` argument; the package itself never loads it.)
"""
import ctypes as C
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsuperpang_b200.so")

SPB_NONE = 0xFFFFFFFF
_ERR = {-1: "SPB_EINVAL", -2: "SPB_ENOMEM", -3: "SPB_EDEGENERATE", -4: "SPB_ECUDA", -5: "SPB_ESTATE"}

EXPORTS = (
    "spb_create", "spb_destroy", "spb_last_error", "spb_load", "spb_build_dbg", "spb_compact",
    "spb_get_vertex2coords", "spb_get_edges", "spb_get_instance_vertices", "spb_get_seqlimits",
    "spb_get_seq_intervals", "spb_get_components", "spb_get_nbps", "spb_get_nbp_seqs",
    "spb_get_nbp_origins", "spb_get_counts", "spb_get_stage_stats", "spb_build_kind", "spb_launch_count",
    "spb_get_nbp_graph", "spb_get_nbp_orientation_counts", "spb_get_seq_nbps", "spb_get_bubble_groups", "spb_path_sequences", "spb_path_origins", "spb_read_fasta", "spb_free_fasta", "spb_mg_slice", "spb_mg_probe", "spb_mg_records", "spb_mg_dedup", "spb_mg_scatter", "spb_mg_dedup_links", "spb_mg_apply_links", "spb_mg_owned_links", "spb_mg_part_records", "spb_mg_part_dedup", "spb_mg_buffers", "spb_mg_ids", "spb_mg_graph",
    "spb_mg_slice_ids", "spb_mg_slice_junctions", "spb_mg_slice_graph", "spb_mg_slice_classify", "spb_mg_slice_contract", "spb_mg_slice_emit",
    "spb_mg_slice_finish", "spb_mg_shard", "spb_mg_complete", "spb_set_host_output", "spb_mg_shard_to_host", "spb_write_fasta", "spb_write_origins_tsv", "spb_path_sequences_ex", "spb_path_origins_ex",
)


class SpbCounts(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "n_seq", "n_bases", "n_inst", "n_vertices", "n_edges", "n_junction", "n_inits", "n_pieces",
        "n_comp", "n_cycle_comp", "n_nbp", "nbp_vertices", "nbp_bases", "n_origin_pairs",
        "n_seq_intervals", "table_slots")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class SpbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{_ERR.get(code, code)}: {msg}")
        self.code = code


class DegenerateInputError(SpbError):
    """Inputs on which the reference's own assertions fire (even k / orientation-ambiguous repeats)."""


class SpbFasta(C.Structure):
    _fields_ = [("ascii", C.POINTER(C.c_uint8)), ("seq_off", C.POINTER(C.c_uint64)), ("names", C.POINTER(C.c_char)),
                ("n_seq", C.c_uint32), ("n_bases", C.c_uint64), ("names_bytes", C.c_uint64)]


def read_fasta_native(path, gap_size=1, lib=None):
    """(names, ascii uint8[T], seq_off uint64[n + 1]) through the library's FASTA reader (reference read_fasta with
    ambigs='as_Ns', Ns='split'; lib/utils.py:5-30).  Raises ValueError for what the reference refuses."""
    lib = lib if lib is not None else load_library()
    fa, err = SpbFasta(), C.create_string_buffer(512)
    rc = lib.spb_read_fasta(os.fsencode(path), int(gap_size), C.byref(fa), err, 512)
    if rc != 0:
        raise ValueError(err.value.decode() or f"spb_read_fasta failed ({rc})")
    try:
        buf = np.ctypeslib.as_array(fa.ascii, shape=(max(int(fa.n_bases), 1),))[:int(fa.n_bases)].copy()
        off = np.ctypeslib.as_array(fa.seq_off, shape=(fa.n_seq + 1,)).copy()
        names = C.string_at(fa.names, int(fa.names_bytes)).decode().split("\n")[:-1] if fa.names_bytes else []
    finally:
        lib.spb_free_fasta(C.byref(fa))
    return names, buf, off


def _name_table(names):
    """(bytes blob, uint64 offsets[n + 1]) of a list of str."""
    enc = [n.encode() for n in names]
    off = np.zeros(len(enc) + 1, dtype=np.uint64)
    np.cumsum([len(e) for e in enc], out=off[1:])
    return b"".join(enc), off


def write_fasta_native(path, names, seqs, seq_off, lib=None):
    """FASTA file of `names` with the sequences of a CSR (seqs uint8, seq_off uint64[n + 1]) that may live in HOST or DEVICE
    memory (numpy arrays, torch tensors or raw device pointers) -- spb_write_fasta; byte-identical to the reference's
    write_fasta (lib/utils.py:62-65) on the same records."""
    lib = lib if lib is not None else load_library()
    blob, noff = _name_table(names)
    err = C.create_string_buffer(512)
    rc = lib.spb_write_fasta(os.fsencode(path), len(names), blob, noff.ctypes.data, _ptr(seqs), _ptr(seq_off), err, 512)
    if rc != 0:
        raise OSError(err.value.decode() or f"spb_write_fasta failed ({rc})")


def write_origins_tsv_native(path, header, node_names, origin_off, origin_idx, orig_names, orig_bin, bin_names, n_genomes, lib=None):
    """NBP2origins.tsv-style table (scripts/SuperPang.py:284-294) -- spb_write_origins_tsv.  origin_off / origin_idx: CSR of
    indices into orig_names (host or device memory); orig_bin[i] = index into bin_names of source sequence i."""
    lib = lib if lib is not None else load_library()
    nb, noff = _name_table(node_names)
    ob, ooff = _name_table(orig_names)
    bb, boff = _name_table(bin_names)
    orig_bin = np.ascontiguousarray(orig_bin, dtype=np.uint32)
    err = C.create_string_buffer(512)
    rc = lib.spb_write_origins_tsv(os.fsencode(path), header.encode(), len(node_names), nb, noff.ctypes.data, _ptr(origin_off), _ptr(origin_idx),
                                   len(orig_names), ob, ooff.ctypes.data, orig_bin.ctypes.data, bb, boff.ctypes.data, len(bin_names), int(n_genomes), err, 512)
    if rc != 0:
        raise OSError(err.value.decode() or f"spb_write_origins_tsv failed ({rc})")


def load_library(path=None):
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            f"superpang_b200: CUDA library not found at {path}. Build it with "
            f"`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
            f"There is no CPU fallback.")
    lib = C.CDLL(path)
    _declare(lib)
    return lib


def _declare(lib):
    vp, u64p, u32p, u8p, i32p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    lib.spb_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
    lib.spb_destroy.argtypes = [vp]
    lib.spb_destroy.restype = None
    lib.spb_last_error.argtypes = [vp]
    lib.spb_last_error.restype = C.c_char_p
    lib.spb_load.argtypes = [vp, u8p, u64p, C.c_uint32, C.c_uint32]
    lib.spb_build_dbg.argtypes = [vp, C.POINTER(SpbCounts)]
    lib.spb_compact.argtypes = [vp, C.POINTER(SpbCounts)]
    lib.spb_get_counts.argtypes = [vp, C.POINTER(SpbCounts)]
    lib.spb_get_vertex2coords.argtypes = [vp, u32p]
    lib.spb_get_edges.argtypes = [vp, u32p]
    lib.spb_get_instance_vertices.argtypes = [vp, u32p]
    lib.spb_get_seqlimits.argtypes = [vp, u32p]
    lib.spb_get_seq_intervals.argtypes = [vp, u64p, u32p]
    lib.spb_get_components.argtypes = [vp, i32p]
    lib.spb_get_nbps.argtypes = [vp, u64p, u32p, i32p, u8p]
    lib.spb_get_nbp_seqs.argtypes = [vp, u64p, u8p]
    lib.spb_get_nbp_origins.argtypes = [vp, u64p, u32p]
    lib.spb_get_stage_stats.argtypes = [vp, C.c_char_p, C.c_uint64]
    lib.spb_build_kind.restype = C.c_char_p
    lib.spb_launch_count.restype = C.c_uint64
    lib.spb_get_nbp_graph.argtypes = [vp, C.POINTER(C.c_uint64), u64p, u32p, u32p]
    lib.spb_get_nbp_orientation_counts.argtypes = [vp, u32p]
    lib.spb_get_bubble_groups.argtypes = [vp, u32p]
    lib.spb_read_fasta.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(SpbFasta), C.c_char_p, C.c_uint64]
    lib.spb_free_fasta.argtypes = [C.POINTER(SpbFasta)]
    lib.spb_free_fasta.restype = None
    lib.spb_path_sequences.argtypes = [vp, C.c_uint32, u64p, u32p, u8p]
    lib.spb_path_origins.argtypes = [vp, C.c_uint32, u64p, u32p, C.c_uint64, C.POINTER(C.c_uint64), u64p, u32p]
    lib.spb_get_seq_nbps.argtypes = [vp, C.POINTER(C.c_uint64), u64p, u32p]
    lib.spb_mg_slice.argtypes = [vp, C.c_uint32, C.POINTER(C.c_uint64)]
    lib.spb_mg_probe.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.spb_mg_records.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.spb_mg_dedup.argtypes = [vp, u64p, u32p, C.POINTER(C.c_uint64), C.c_uint32, C.c_uint32, C.c_uint32, u32p]
    lib.spb_mg_scatter.argtypes = [vp, u32p]
    lib.spb_mg_dedup_links.argtypes = [vp, u64p, u32p, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_void_p)]
    lib.spb_mg_apply_links.argtypes = [vp, u64p, C.c_uint64, C.c_uint64, C.c_uint64]
    lib.spb_mg_owned_links.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_void_p)]
    lib.spb_mg_part_records.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_int)]
    lib.spb_mg_part_dedup.argtypes = [vp, u64p, u32p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    lib.spb_mg_buffers.argtypes = [vp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.spb_mg_ids.argtypes = [vp, C.c_uint64, C.c_uint64]
    lib.spb_mg_graph.argtypes = [vp, C.POINTER(SpbCounts)]
    pv, pu64, pu32 = C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    lib.spb_mg_slice_ids.argtypes = [vp, C.c_uint32, C.c_uint32]
    lib.spb_mg_slice_junctions.argtypes = [vp, pu64, pv, pv]
    lib.spb_mg_slice_graph.argtypes = [vp, u64p, u8p, C.c_uint64, pu64]
    lib.spb_mg_slice_classify.argtypes = [vp, pu32, pu32, pu32, pv, pv, pv]
    lib.spb_mg_slice_contract.argtypes = [vp, u32p, C.c_uint32, u32p, u32p, C.c_uint32, C.c_uint64, C.c_int, C.POINTER(C.c_int)]
    lib.spb_mg_slice_emit.argtypes = [vp, pu64, pv, pu64, pv, pv]
    lib.spb_mg_slice_finish.argtypes = [vp, u64p, C.c_uint64, u64p, C.c_uint64, u8p, C.POINTER(SpbCounts)]
    lib.spb_mg_shard.argtypes = [vp, C.c_int, pv, pu64, pu64]
    lib.spb_mg_complete.argtypes = [vp, C.POINTER(SpbCounts)]
    lib.spb_set_host_output.argtypes = [vp, u8p, C.c_uint64]
    lib.spb_mg_shard_to_host.argtypes = [vp, C.c_int, vp, pu64]
    lib.spb_path_sequences_ex.argtypes = [vp, C.c_int, C.c_uint32, u64p, u32p, u64p, u8p]
    lib.spb_path_origins_ex.argtypes = [vp, C.c_int, C.c_uint32, u64p, u32p, C.c_uint32, u32p, u64p, u32p, C.c_uint64, C.POINTER(C.c_uint64), u64p, u32p]
    lib.spb_write_fasta.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, u64p, u8p, u64p, C.c_char_p, C.c_uint64]
    lib.spb_write_origins_tsv.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_char_p, u64p, u64p, u32p, C.c_uint32, C.c_char_p, u64p, u32p,
                                          C.c_char_p, u64p, C.c_uint32, C.c_uint32, C.c_char_p, C.c_uint64]
    return lib


def _ptr(a):
    """Raw address of a numpy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()          # torch tensor (host pinned or device)


def _wrap(ptr, n, dtype, cuda, device):
    """Zero-copy torch view of `n` elements of library-owned memory (device memory for the CUDA build)."""
    import torch
    dev = torch.device("cuda", device) if cuda else torch.device("cpu")
    if n == 0 or not ptr:
        return torch.empty(0, dtype=dtype, device=dev)
    typestr, npdt, size = {torch.int64: ("<i8", np.int64, 8), torch.int32: ("<i4", np.int32, 4), torch.uint8: ("|u1", np.uint8, 1)}[dtype]
    if cuda:
        class _Ptr:
            pass
        o = _Ptr()
        o.__cuda_array_interface__ = dict(shape=(int(n),), typestr=typestr, data=(int(ptr), False), version=3, strides=None)
        return torch.as_tensor(o, device=dev)
    buf = (C.c_byte * (int(n) * size)).from_address(int(ptr))
    return torch.from_numpy(np.frombuffer(buf, dtype=npdt))


def _exchange(out, inp, out_splits, in_splits, rank, group):
    """Variable-size all-to-all.  The chunk a rank keeps for itself is a plain device copy; the other chunks go to
    NCCL as grouped send/recv on views of the caller's buffers (no staging copies).  Backends without the list form
    (gloo, CPU tests) use all_to_all_single on the whole buffers."""
    import torch
    import torch.distributed as dist
    if out.is_cuda:
        outs = list(torch.split(out, out_splits))
        ins = list(torch.split(inp, in_splits))
        if in_splits[rank]:
            outs[rank].copy_(ins[rank])
        empty_o, empty_i = out[:0], inp[:0]
        outs[rank], ins[rank] = empty_o, empty_i
        dist.all_to_all(outs, ins, group=group)
    else:
        dist.all_to_all_single(out, inp, out_splits, in_splits, group=group)


def _gather_var(parts, counts_all, rank, world, group):
    """All-gather of variable-length device lists.  parts: tensors of this rank (one per list, any integer dtype);
    counts_all[r][i] = length of list i on rank r (already known everywhere).  The lists of a rank travel as ONE padded
    byte block (one collective); returns, per list, the concatenation over the ranks in rank order."""
    import torch
    import torch.distributed as dist
    sizes = [p.element_size() for p in parts]
    nbytes = [sum(int(c) * sz for c, sz in zip(cr, sizes)) for cr in counts_all]
    m = (max(max(nbytes), 1) + 15) & ~15
    dev = parts[0].device
    pad = torch.empty(world * m, dtype=torch.uint8, device=dev)
    o = rank * m
    for p_, sz in zip(parts, sizes):
        nb = p_.numel() * sz
        if nb:
            pad[o:o + nb].copy_(p_.view(torch.uint8))
        o += nb
    _all_gather_inplace(pad, rank, world, group)
    out = []
    for i, (p_, sz) in enumerate(zip(parts, sizes)):
        chunks = []
        for r in range(world):
            o = r * m + sum(int(counts_all[r][j]) * sizes[j] for j in range(i))
            nb = int(counts_all[r][i]) * sz
            if nb:
                chunks.append(pad[o:o + nb])
        cat = torch.cat(chunks) if chunks else pad[:0]
        # the byte offsets of a list inside a block are multiples of its element size only if the lists before it are:
        # callers order the lists by decreasing element size
        out.append(cat.view(p_.dtype) if cat.numel() else torch.empty(0, dtype=p_.dtype, device=dev))
    return out


def _counts_all(vals, world, dev, group):
    """all-gather of a few host integers per rank -> list (per rank) of lists"""
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(v) for v in vals], dtype=torch.int64, device=dev)
    full = torch.empty(world * t.numel(), dtype=torch.int64, device=dev)
    try:
        dist.all_gather_into_tensor(full, t, group=group)
    except Exception:
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t, group=group)
        full = torch.cat(parts)
    return full.view(world, t.numel()).tolist()


def _all_gather_inplace(full, rank, world, group):
    """all-gather where every rank already holds its own chunk of `full` at the right offset."""
    import torch.distributed as dist
    n = full.numel() // world
    mine = full[rank * n:(rank + 1) * n]
    try:
        dist.all_gather_into_tensor(full, mine, group=group)
    except Exception:
        parts = [full[i * n:(i + 1) * n] for i in range(world)]
        tmp = [p.clone() for p in parts]
        dist.all_gather(tmp, mine.clone(), group=group)
        for p, t in zip(parts, tmp):
            p.copy_(t)


class SharedHostBuffer:
    """uint8 host buffer in POSIX shared memory, mapped by every rank of the group and page-locked in each process
    (cudaHostRegister): with results sliced over the ranks, every rank copies ITS part over ITS PCIe link straight into the
    memory that the one process hosting the reference's single-process tail (rank 0) reads.  Collective constructor."""

    def __init__(self, nbytes, group=None, tag="buf"):
        import torch
        import torch.distributed as dist
        rank = dist.get_rank(group)
        self.nbytes = max(int(nbytes), 1)
        self.path = f"/dev/shm/spb_{os.environ.get('MASTER_PORT', '0')}_{os.getuid()}_{tag}"
        # rank 0 checks that the shared-memory file system has room (a tmpfs lets the file grow past its limit and faults on
        # the first write instead); without it every rank falls back to a private page-locked buffer (`shared` = False:
        # the parts are then on the host, but not in one process)
        ok = [True]
        if rank == 0:
            try:
                st = os.statvfs("/dev/shm")
                ok[0] = st.f_bavail * st.f_frsize > self.nbytes + (64 << 20)
            except OSError:
                ok[0] = False
        dist.broadcast_object_list(ok, src=0, group=group)
        self.shared = bool(ok[0])
        self.registered = False
        if not self.shared:
            self.mm = None
            self.tensor = torch.empty(self.nbytes, dtype=torch.uint8)
            if torch.cuda.is_available():
                self.tensor = self.tensor.pin_memory()
            return
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(self.nbytes)
        dist.barrier(group)
        self.mm = np.memmap(self.path, dtype=np.uint8, mode="r+", shape=(self.nbytes,))
        self.tensor = torch.from_numpy(self.mm)
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), self.nbytes, 0)
            self.registered = int(rc) == 0
        dist.barrier(group)
        if rank == 0:
            os.unlink(self.path)              # the mappings keep the memory alive

    def close(self):
        import torch
        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
            self.registered = False


class Engine:
    """One context = one GPU + one stream.  Mirrors the call order of the C-ABI."""

    def __init__(self, device=0, stream=None, lib=None):
        self.lib = lib if lib is not None else load_library()
        self.device = int(device)
        self.kind = self.lib.spb_build_kind().decode()
        h = C.c_void_p()
        rc = self.lib.spb_create(C.byref(h), int(device), C.c_void_p(stream or 0))
        if rc != 0:
            raise SpbError(rc, "spb_create failed (no CUDA device? there is no CPU fallback)")
        self.h = h
        self.counts = SpbCounts()
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.spb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            msg = self.lib.spb_last_error(self.h).decode(errors="replace")
            raise (DegenerateInputError if rc == -3 else SpbError)(rc, msg)

    # ---- pipeline -----------------------------------------------------------------------
    def load(self, ascii_buf, seq_off, k):
        """ascii_buf: uint8 numpy array / torch tensor (host or device) of concatenated A/C/G/T;
        seq_off: uint64[n_seq + 1] (host)."""
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        self._keep = (ascii_buf, seq_off)
        self._k = int(k)
        self.mg_sliced = False
        self._ck(self.lib.spb_load(self.h, _ptr(ascii_buf), seq_off.ctypes.data, len(seq_off) - 1, int(k)))

    def set_host_output(self, nbp_seqs_host):
        """Page-locked uint8 host tensor / array that is to receive the NBP sequences: compact() then copies them on a second
        stream while origins, components and edge rows are computed (spb_set_host_output); None clears."""
        self._out_keep = nbp_seqs_host
        n = 0 if nbp_seqs_host is None else (nbp_seqs_host.numel() if hasattr(nbp_seqs_host, "numel") else nbp_seqs_host.size)
        self._ck(self.lib.spb_set_host_output(self.h, _ptr(nbp_seqs_host), int(n)))

    def build_dbg(self):
        self._ck(self.lib.spb_build_dbg(self.h, C.byref(self.counts)))
        return self.counts.as_dict()

    def compact(self):
        if getattr(self, "mg_sliced", False):            # the sliced multi-GPU tail has built the graph AND compacted it
            return self.counts.as_dict()
        self._ck(self.lib.spb_compact(self.h, C.byref(self.counts)))
        return self.counts.as_dict()

    def build_dbg_distributed(self, group=None):
        """DBG build over all ranks of `group` (one process per GPU; every rank has loaded the whole input).

        K-mers are hash-partitioned over the ranks: positions are cut into equal slices, each rank turns its
        slice into (hash, position) records, an all-to-all routes every record to the rank owning its hash
        range, the owner deduplicates exactly and an all-to-all returns the first position of every record.
        Mostly-unique inputs then build the graph and compact it SLICED over the ranks (`_mg_sliced_tail`); otherwise the
        first-appearance / orientation bitmaps and the per-position vertex ids are all-gathered, after which every rank
        holds the complete DBG and `compact()` proceeds unchanged (replicated).  The number of ranks must be a power of two
        (one rank: plain `build_dbg()`).  NCCL on GPUs; the CPU test tier drives the same code with the host emulator over gloo."""
        import torch
        import torch.distributed as dist
        G, r = dist.get_world_size(group), dist.get_rank(group)
        if G == 1:
            return self.build_dbg()
        if G & (G - 1):
            # the owner of a k-mer is the top log2(G) bits of its hash: checked here, before any collective is issued
            raise ValueError(f"build_dbg_distributed needs a power-of-two number of ranks (got {G})")
        cuda = self.kind.startswith("cuda")
        dev = torch.device("cuda", self.device) if cuda else torch.device("cpu")
        marks = []

        def mark(name):                      # stage boundaries (CUDA events on the current stream)
            if cuda:
                e = torch.cuda.Event(enable_timing=True)
                e.record(torch.cuda.current_stream(self.device))
                marks.append((name, e))

        def sync():
            if cuda:
                torch.cuda.current_stream(self.device).synchronize()

        T = int(self._keep[1][-1])
        S = C.c_uint64()
        self._ck(self.lib.spb_mg_slice(self.h, G, C.byref(S)))
        S = int(S.value)
        lo, hi = r * S, min((r + 1) * S, T)
        cnt = (C.c_uint64 * G)()
        kp, pp = C.c_void_p(), C.c_void_p()
        mark("start")
        # return protocol: rank 0 samples the duplicate rate of the input (two chunks) and broadcasts the choice
        protocol = os.environ.get("SPB_MG_PROTOCOL")
        if protocol not in ("links", "reps", "owned", "part"):
            flag = torch.zeros(1, dtype=torch.int64, device=dev)
            if r == 0:
                er, dr = C.c_double(), C.c_double()
                self._ck(self.lib.spb_mg_probe(self.h, C.byref(er), C.byref(dr)))
                flag[0] = 1 if dr.value < 0.02 else 0
            dist.broadcast(flag, src=0, group=group)
            protocol = "part" if int(flag.item()) else "reps"
        if protocol == "owned" and int(self._k) < 17:
            protocol = "links"                          # the rolling-hash kernels of that protocol need k >= 17
        self.mg_protocol = protocol
        mark("probe")
        if protocol == "part":
            done = self._mg_part(group, G, r, S, lo, hi, cuda, dev, marks, mark, sync)
            if done is not None:
                return done
            protocol = self.mg_protocol = "links"       # a bucket beyond its region somewhere (masses of identical k-mers): exact path instead
        if protocol == "owned":
            # "owner computes": every rank hashes all positions itself and keeps the k-mers it owns -- no record travels;
            # only the links (position -> first position) of repeated instances are exchanged
            lcnt = (C.c_uint64 * G)()
            lp = C.c_void_p()
            self._ck(self.lib.spb_mg_owned_links(self.h, r, G, lcnt, C.byref(lp)))
            mark("owner_dedup")
            lsend = [int(x) for x in lcnt]
            links = _wrap(lp.value, sum(lsend), torch.int64, cuda, self.device)
            lsc = torch.tensor(lsend, dtype=torch.int64, device=dev)
            lrc = torch.empty_like(lsc)
            dist.all_to_all_single(lrc, lsc, group=group)
            lrecv = [int(x) for x in lrc.tolist()]
            lback = torch.empty(sum(lrecv), dtype=torch.int64, device=dev)
            _exchange(lback, links, lrecv, lsend, r, group)
            mark("exchange_reps")
            sync()
            self._ck(self.lib.spb_mg_apply_links(self.h, lback.data_ptr() if lback.numel() else None, lback.numel(), lo, max(lo, hi)))
            return self._mg_finish(group, G, r, S, lo, hi, cuda, marks, mark, sync, gather_orientation=False)
        self._ck(self.lib.spb_mg_records(self.h, lo, max(lo, hi), G, cnt, C.byref(kp), C.byref(pp)))
        mark("records_partition")
        send = [int(x) for x in cnt]
        keys = _wrap(kp.value, sum(send), torch.int64, cuda, self.device)
        pos = _wrap(pp.value, sum(send), torch.int32, cuda, self.device)
        sc = torch.tensor(send, dtype=torch.int64, device=dev)
        rc_t = torch.empty_like(sc)
        dist.all_to_all_single(rc_t, sc, group=group)
        recv = [int(x) for x in rc_t.tolist()]
        rkeys = torch.empty(sum(recv), dtype=torch.int64, device=dev)
        rpos = torch.empty(sum(recv), dtype=torch.int32, device=dev)
        _exchange(rkeys, keys, recv, send, r, group)
        _exchange(rpos, pos, recv, send, r, group)
        mark("exchange_records")
        sync()
        if protocol == "reps":
            # one representative per record travels back (kept for A/B measurements)
            rep = torch.empty(sum(recv), dtype=torch.int32, device=dev)
            chunk = (C.c_uint64 * G)(*recv)
            self._ck(self.lib.spb_mg_dedup(self.h, rkeys.data_ptr(), rpos.data_ptr(), chunk, G, G, r, rep.data_ptr()))
            mark("owner_dedup")
            back = torch.empty(sum(send), dtype=torch.int32, device=dev)
            _exchange(back, rep, send, recv, r, group)
            mark("exchange_reps")
            sync()
            self._ck(self.lib.spb_mg_scatter(self.h, back.data_ptr()))
        else:
            # "links": only the instances that are not first appearances travel back, as (position -> first position)
            lcnt = (C.c_uint64 * G)()
            lp = C.c_void_p()
            self._ck(self.lib.spb_mg_dedup_links(self.h, rkeys.data_ptr(), rpos.data_ptr(), sum(recv), G, lcnt, C.byref(lp)))
            mark("owner_dedup")
            lsend = [int(x) for x in lcnt]
            links = _wrap(lp.value, sum(lsend), torch.int64, cuda, self.device)
            lsc = torch.tensor(lsend, dtype=torch.int64, device=dev)
            lrc = torch.empty_like(lsc)
            dist.all_to_all_single(lrc, lsc, group=group)
            lrecv = [int(x) for x in lrc.tolist()]
            lback = torch.empty(sum(lrecv), dtype=torch.int64, device=dev)
            _exchange(lback, links, lrecv, lsend, r, group)
            mark("exchange_reps")
            sync()
            self._ck(self.lib.spb_mg_apply_links(self.h, lback.data_ptr() if lback.numel() else None, lback.numel(), lo, max(lo, hi)))
        return self._mg_finish(group, G, r, S, lo, hi, cuda, marks, mark, sync, gather_orientation=True)

    def _mg_part(self, group, G, r, S, lo, hi, cuda, dev, marks, mark, sync):
        """"Part" protocol: slice -> 8-byte records partitioned by (owner, level-1 bucket) [k_roll_partition] -> equal-split
        all-to-all of the regions and their counts -> owner: level-2 partition + shared-memory tables [k_partition2,
        k_bucket_dedup] -> links.  When the links are few (mostly-unique input) every rank receives ALL of them and numbers
        all positions itself (no all-gather of the 4-byte-per-position id array); otherwise the links go to the slice
        owners and the ids are all-gathered.  Returns None when some rank reported an overflow (caller falls back)."""
        import torch
        import torch.distributed as dist
        T = int(self._keep[1][-1])
        rr, nreg, ov = C.c_uint64(), C.c_uint32(), C.c_int()
        rp, cp = C.c_void_p(), C.c_void_p()
        self._ck(self.lib.spb_mg_part_records(self.h, lo, max(lo, hi), G, C.byref(rr), C.byref(nreg), C.byref(rp), C.byref(cp), C.byref(ov)))
        mark("records_partition")
        flag = torch.tensor([int(ov.value)], dtype=torch.int64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        if int(flag.item()):
            return None
        R, NR = int(rr.value), int(nreg.value)
        recs = _wrap(rp.value, G * R, torch.int64, cuda, self.device)
        cnts = _wrap(cp.value, G * NR, torch.int32, cuda, self.device)
        rrecs, rcnts = torch.empty_like(recs), torch.empty_like(cnts)
        dist.all_to_all_single(rcnts, cnts, group=group)
        dist.all_to_all_single(rrecs, recs, group=group)
        # bytes this rank sends over NVLink in the record exchange (fixed-capacity regions: padding included) + the orientation
        # bitmap parts it receives: what `mg_exchange_records_ms` moves
        self.mg_exchange_bytes = (recs.numel() * 8 + cnts.numel() * 4) * (G - 1) // G + (G - 1) * (S // 8)
        self.mg_owned_records = rcnts.sum()             # k-mer instances this rank owns (device scalar: read after the timed region)
        fb, ob, sp = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._ck(self.lib.spb_mg_buffers(self.h, C.byref(fb), C.byref(ob), C.byref(sp)))
        _all_gather_inplace(_wrap(ob.value, G * S // 32, torch.int32, cuda, self.device), r, G, group)   # the owners verify repeats exactly
        mark("exchange_records")
        sync()
        lcnt = (C.c_uint64 * G)()
        lp = C.c_void_p()
        self._ck(self.lib.spb_mg_part_dedup(self.h, rrecs.data_ptr(), rcnts.data_ptr(), G, G, lcnt, C.byref(lp), C.byref(ov)))
        del rrecs, rcnts
        mark("owner_dedup")
        lsend = [int(x) for x in lcnt]
        nl = sum(lsend)
        tot = torch.tensor([int(ov.value), nl], dtype=torch.int64, device=dev)
        allc = [torch.empty_like(tot) for _ in range(G)]
        dist.all_gather(allc, tot, group=group)
        allc = [[int(v) for v in t.tolist()] for t in allc]
        if any(a[0] for a in allc):
            return None
        counts = [a[1] for a in allc]
        links = _wrap(lp.value, nl, torch.int64, cuda, self.device)
        all_links = os.environ.get("SPB_MG_ALL_LINKS", "1")       # "0": never, "2": always (tests), default: when the links are few
        if all_links == "2" or (16 * sum(counts) <= T and all_links != "0"):
            # few links: everybody gets all of them (padded all-gather), applies them and numbers ALL positions locally
            m = max(max(counts), 1)
            pad = torch.zeros(G * m, dtype=torch.int64, device=dev)
            pad[r * m:r * m + nl].copy_(links)
            _all_gather_inplace(pad, r, G, group)
            alll = torch.cat([pad[i * m:i * m + counts[i]] for i in range(G)]) if sum(counts) else pad[:0]
            mark("exchange_reps")
            sync()
            self._ck(self.lib.spb_mg_apply_links(self.h, alll.data_ptr() if alll.numel() else None, alll.numel(), 0, T))
            mark("scatter")
            if os.environ.get("SPB_MG_TAIL", "sliced") == "sliced":
                return self._mg_sliced_tail(group, G, r, cuda, dev, marks, mark, sync)
            self._ck(self.lib.spb_mg_ids(self.h, 0, T))
            mark("bitmaps_ids")
            sync()
            self._ck(self.lib.spb_mg_graph(self.h, C.byref(self.counts)))
            mark("graph")
            if cuda:
                sync()
                self.mg_times = {f"mg_{n}_ms": marks[i - 1][1].elapsed_time(e) for i, (n, e) in enumerate(marks) if i}
            return self.counts.as_dict()
        lsc = torch.tensor(lsend, dtype=torch.int64, device=dev)
        lrc = torch.empty_like(lsc)
        dist.all_to_all_single(lrc, lsc, group=group)
        lrecv = [int(x) for x in lrc.tolist()]
        lback = torch.empty(sum(lrecv), dtype=torch.int64, device=dev)
        _exchange(lback, links, lrecv, lsend, r, group)
        mark("exchange_reps")
        sync()
        self._ck(self.lib.spb_mg_apply_links(self.h, lback.data_ptr() if lback.numel() else None, lback.numel(), lo, max(lo, hi)))
        return self._mg_finish(group, G, r, S, lo, hi, cuda, marks, mark, sync, gather_orientation=False)

    def _raise_code(self, code, local_rc):
        """Raise what status `code` (the worst over all ranks) stands for; the local message if this rank saw it itself."""
        if code == 0:
            return
        msg = self.lib.spb_last_error(self.h).decode(errors="replace") if local_rc == code else "reported by another rank of the group"
        raise (DegenerateInputError if code == -3 else SpbError)(code, msg)

    def _mg_sliced_tail(self, group, G, r, cuda, dev, marks, mark, sync):
        """Graph build + compaction sliced over the ranks (include/superpang_b200.h, "sliced multi-GPU tail"): per-position
        passes on the own slice, per-vertex passes on the own vertices, emission of the own NBP range; only the short lists
        (junction entries, inits, piece bounds, origin pairs) are all-gathered, the contracted graph is computed by every
        rank.  Leaves the engine compacted, with sliced results (`shard`, `complete`)."""
        import torch
        import torch.distributed as dist
        L, h = self.lib, self.h
        self._ck(L.spb_mg_slice_ids(h, r, G))
        mark("slice_ids")
        n, jk, js = C.c_uint64(), C.c_void_p(), C.c_void_p()
        self._ck(L.spb_mg_slice_junctions(h, C.byref(n), C.byref(jk), C.byref(js)))
        ca = _counts_all([n.value], G, dev, group)
        jk_t = _wrap(jk.value, n.value, torch.int64, cuda, self.device)
        js_t = _wrap(js.value, n.value, torch.uint8, cuda, self.device)
        jk_all, js_all = _gather_var([jk_t, js_t], [[c_[0], c_[0]] for c_ in ca], r, G, group)
        # The error flags are rank-local here (a rank only sees the degenerate repeats of ITS positions and vertices), so a status
        # is never raised before every rank knows it: it travels with the next gather of counts / one final reduction, and all
        # ranks raise together instead of leaving the others waiting in a collective.
        ne = C.c_uint64()
        rc_graph = L.spb_mg_slice_graph(h, jk_all.data_ptr() if jk_all.numel() else None, js_all.data_ptr() if js_all.numel() else None,
                                        jk_all.numel(), C.byref(ne))
        del jk_all, js_all
        mark("slice_graph")
        for p_ in range(2):
            ni, ns, nend = C.c_uint32(), C.c_uint32(), C.c_uint32()
            il, pl, ph = C.c_void_p(), C.c_void_p(), C.c_void_p()
            rc_cls = 0 if rc_graph else L.spb_mg_slice_classify(h, C.byref(ni), C.byref(ns), C.byref(nend), C.byref(il), C.byref(pl), C.byref(ph))
            local_rc = rc_graph or rc_cls
            ca = _counts_all([ni.value, ns.value, nend.value, ne.value, -local_rc], G, dev, group)
            self._raise_code(-max(c_[4] for c_ in ca), local_rc)
            lists = [_wrap(il.value, ni.value, torch.int32, cuda, self.device), _wrap(pl.value, ns.value, torch.int32, cuda, self.device),
                     _wrap(ph.value, nend.value, torch.int32, cuda, self.device)]
            init_all, plo_all, phi_all = _gather_var(lists, [c_[:3] for c_ in ca], r, G, group)
            if plo_all.numel() != phi_all.numel():
                raise SpbError(-4, f"internal: {plo_all.numel()} piece starts, {phi_all.numel()} piece ends over the ranks")
            again = C.c_int()
            self._ck(L.spb_mg_slice_contract(h, init_all.data_ptr() if init_all.numel() else None, init_all.numel(),
                                             plo_all.data_ptr() if plo_all.numel() else None, phi_all.data_ptr() if phi_all.numel() else None,
                                             plo_all.numel(), sum(c_[3] for c_ in ca), p_, C.byref(again)))
            del init_all, plo_all, phi_all
            if not again.value:
                break
        mark("slice_classify_rank")
        no, nm = C.c_uint64(), C.c_uint64()
        op, mp, h01 = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._ck(L.spb_mg_slice_emit(h, C.byref(no), C.byref(op), C.byref(nm), C.byref(mp), C.byref(h01)))
        mark("slice_emit_pairs")
        ca = _counts_all([no.value, nm.value], G, dev, group)
        op_all, mp_all = _gather_var([_wrap(op.value, no.value, torch.int64, cuda, self.device), _wrap(mp.value, nm.value, torch.int64, cuda, self.device)],
                                     ca, r, G, group)
        n_seq = len(self._keep[1]) - 1
        has01 = _wrap(h01.value, 2 * n_seq, torch.uint8, cuda, self.device).clone()
        dist.all_reduce(has01, op=dist.ReduceOp.MAX, group=group)
        rc_fin = L.spb_mg_slice_finish(h, op_all.data_ptr() if op_all.numel() else None, op_all.numel(), mp_all.data_ptr() if mp_all.numel() else None,
                                       mp_all.numel(), has01.data_ptr(), C.byref(self.counts))
        worst = torch.tensor([-rc_fin], dtype=torch.int64, device=dev)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX, group=group)
        self._raise_code(-int(worst.item()), rc_fin)
        mark("slice_origins_components")
        self.mg_sliced = True
        if cuda:
            sync()
            self.mg_times = {f"mg_{n_}_ms": marks[i - 1][1].elapsed_time(e) for i, (n_, e) in enumerate(marks) if i}
        return self.counts.as_dict()

    def shard(self, which):
        """Local part of a sliced result as a zero-copy torch view (device memory on the GPU build) and the index of its
        first element (row) in the complete array.  which: 'instance_vertices', 'vertex2coords', 'edges', 'nbp_verts',
        'nbp_seqs' (include/superpang_b200.h, spb_mg_shard)."""
        import torch
        code, dt, width = {"instance_vertices": (0, torch.int32, 1), "vertex2coords": (1, torch.int32, 2), "edges": (2, torch.int32, 2),
                           "nbp_verts": (3, torch.int32, 1), "nbp_seqs": (4, torch.uint8, 1)}[which]
        p_, first, count = C.c_void_p(), C.c_uint64(), C.c_uint64()
        self._ck(self.lib.spb_mg_shard(self.h, code, C.byref(p_), C.byref(first), C.byref(count)))
        return _wrap(p_.value, count.value * width, dt, self.kind.startswith("cuda"), self.device), int(first.value)

    def shard_to_host(self, which, host_tensor):
        """Copy the local part of a sliced result to its place in `host_tensor` (the complete array on the host, e.g. a
        SharedHostBuffer tensor); asynchronous on the engine's stream when the host memory is page-locked
        (spb_mg_shard_to_host).  Returns the number of bytes copied."""
        code = {"instance_vertices": 0, "vertex2coords": 1, "edges": 2, "nbp_verts": 3, "nbp_seqs": 4}[which]
        n = C.c_uint64()
        self._ck(self.lib.spb_mg_shard_to_host(self.h, code, _ptr(host_tensor), C.byref(n)))
        return int(n.value)

    def complete(self):
        """Recompute the sliced passes over the whole input on this rank (no communication): afterwards every getter
        works as on one GPU (spb_mg_complete)."""
        self._ck(self.lib.spb_mg_complete(self.h, C.byref(self.counts)))
        self.mg_sliced = False
        return self.counts.as_dict()

    def _mg_finish(self, group, G, r, S, lo, hi, cuda, marks, mark, sync, gather_orientation):
        """Common tail of the distributed build: bitmap all-gather, ids, id all-gather, graph (replicated)."""
        import torch
        fb, ob, sp = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._ck(self.lib.spb_mg_buffers(self.h, C.byref(fb), C.byref(ob), C.byref(sp)))
        mark("scatter")
        for ptr in ((fb, ob) if gather_orientation else (fb,)):
            full = _wrap(ptr.value, G * S // 32, torch.int32, cuda, self.device)
            _all_gather_inplace(full, r, G, group)
        sync()
        self._ck(self.lib.spb_mg_ids(self.h, lo, max(lo, hi)))
        mark("bitmaps_ids")
        full = _wrap(sp.value, G * S, torch.int32, cuda, self.device)
        _all_gather_inplace(full, r, G, group)
        mark("allgather_ids")
        sync()
        self._ck(self.lib.spb_mg_graph(self.h, C.byref(self.counts)))
        mark("graph")
        if cuda:
            sync()
            self.mg_times = {f"mg_{n}_ms": marks[i - 1][1].elapsed_time(e) for i, (n, e) in enumerate(marks) if i}
        return self.counts.as_dict()

    def refresh_counts(self):
        self._ck(self.lib.spb_get_counts(self.h, C.byref(self.counts)))
        return self.counts.as_dict()

    def stage_stats(self):
        buf = C.create_string_buffer(4096)
        self._ck(self.lib.spb_get_stage_stats(self.h, buf, 4096))
        try:
            return json.loads(buf.value.decode() or "{}")
        except json.JSONDecodeError:
            return {"raw": buf.value.decode()}

    # ---- getters (host numpy copies) -------------------------------------------------------
    def vertex2coords(self):
        out = np.empty((self.counts.n_vertices, 2), dtype=np.uint32)
        self._ck(self.lib.spb_get_vertex2coords(self.h, out.ctypes.data))
        return out

    def edges(self):
        out = np.empty((self.counts.n_edges, 2), dtype=np.uint32)
        self._ck(self.lib.spb_get_edges(self.h, out.ctypes.data))
        return out

    def instance_vertices(self):
        out = np.empty(self.counts.n_bases, dtype=np.uint32)
        self._ck(self.lib.spb_get_instance_vertices(self.h, out.ctypes.data))
        return out

    def seqlimits(self):
        out = np.empty((self.counts.n_seq, 2), dtype=np.uint32)
        self._ck(self.lib.spb_get_seqlimits(self.h, out.ctypes.data))
        return out

    def seq_intervals(self):
        self._ck(self.lib.spb_get_seq_intervals(self.h, None, None))
        self.refresh_counts()
        off = np.empty(self.counts.n_seq + 1, dtype=np.uint64)
        iv = np.empty((self.counts.n_seq_intervals, 2), dtype=np.uint32)
        self._ck(self.lib.spb_get_seq_intervals(self.h, off.ctypes.data, iv.ctypes.data))
        return off, iv

    def components(self):
        out = np.empty(self.counts.n_vertices, dtype=np.int32)
        self._ck(self.lib.spb_get_components(self.h, out.ctypes.data))
        return out

    def nbps(self):
        n = self.counts.n_nbp
        off = np.empty(n + 1, dtype=np.uint64)
        verts = np.empty(self.counts.nbp_vertices, dtype=np.uint32)
        comp = np.empty(n, dtype=np.int32)
        cyc = np.empty(n, dtype=np.uint8)
        self._ck(self.lib.spb_get_nbps(self.h, off.ctypes.data, verts.ctypes.data, comp.ctypes.data, cyc.ctypes.data))
        return off, verts, comp, cyc

    def nbp_seqs(self):
        off = np.empty(self.counts.n_nbp + 1, dtype=np.uint64)
        seq = np.empty(self.counts.nbp_bases, dtype=np.uint8)
        self._ck(self.lib.spb_get_nbp_seqs(self.h, off.ctypes.data, seq.ctypes.data))
        return off, seq

    def write_nbp_fasta(self, path, names=None):
        """The raw NBP sequences as a FASTA file straight from the device buffers (no host copy of the sequence array in
        Python): record i is named names[i] (default "NBP_{i}").  Returns the number of records."""
        import torch
        n = int(self.counts.n_nbp)
        names = [f"NBP_{i}" for i in range(n)] if names is None else list(names)
        assert len(names) == n
        cuda = self.kind.startswith("cuda")
        dev = torch.device("cuda", self.device) if cuda else torch.device("cpu")
        soff = torch.empty(n + 1, dtype=torch.int64, device=dev)
        seq = torch.empty(max(int(self.counts.nbp_bases), 1), dtype=torch.uint8, device=dev)
        self._ck(self.lib.spb_get_nbp_seqs(self.h, soff.data_ptr(), seq.data_ptr()))      # device-to-device
        write_fasta_native(path, names, seq, soff, lib=self.lib)
        return n

    def nbp_graph(self):
        """(off, succ, rev): successors CSR of the NBP graph and the reverse-orientation partner of every raw NBP."""
        ne = C.c_uint64()
        self._ck(self.lib.spb_get_nbp_graph(self.h, C.byref(ne), None, None, None))
        n = self.counts.n_nbp
        off = np.empty(n + 1, dtype=np.uint64)
        succ = np.empty(ne.value, dtype=np.uint32)
        rev = np.empty(n, dtype=np.uint32)
        self._ck(self.lib.spb_get_nbp_graph(self.h, C.byref(ne), off.ctypes.data, succ.ctypes.data, rev.ctypes.data))
        return off, succ, rev

    def nbp_orientation_counts(self):
        """counts[i] = number of input sequences that contain raw NBP i in its own orientation (reference
        get_seqPaths_NBPs / nv2rightOrientation, lib/Assembler.py:1121-1139, :448-453)."""
        out = np.zeros(self.counts.n_nbp, dtype=np.uint32)
        if out.size:
            self._ck(self.lib.spb_get_nbp_orientation_counts(self.h, out.ctypes.data))
        return out

    def seq_nbps(self):
        """(off, ids): CSR over the loaded sequences of the raw NBPs each contains in its own orientation (the sets of the
        reference's get_seqPaths_NBPs, lib/Assembler.py:1121-1139)."""
        n = C.c_uint64()
        self._ck(self.lib.spb_get_seq_nbps(self.h, C.byref(n), None, None))
        off = np.zeros(self.counts.n_seq + 1, dtype=np.uint64)
        ids = np.empty(n.value, dtype=np.uint32)
        self._ck(self.lib.spb_get_seq_nbps(self.h, C.byref(n), off.ctypes.data, ids.ctypes.data))
        return off, ids

    def path_sequences(self, paths, k, nbp_ids=False):
        """Sequences (str) of paths: the reference's reconstruct_sequence (lib/Assembler.py:1038-1118) for arbitrary, e.g.
        joined, paths.  paths: iterables of vertex ids (each >= 2 long, consecutive vertices adjacent), or -- nbp_ids=True --
        of RAW NBP IDS whose NBPs are joined the reference's way (path + succ[1:], :810); those are expanded to vertex
        lists on the device (spb_path_sequences_ex)."""
        paths = [np.asarray(p, dtype=np.uint32) for p in paths]
        if not paths:
            return []
        off = np.zeros(len(paths) + 1, dtype=np.uint64)
        np.cumsum([len(p) for p in paths], out=off[1:])
        items = np.ascontiguousarray(np.concatenate(paths))
        voff = np.zeros(len(paths) + 1, dtype=np.uint64)
        if nbp_ids:
            self._ck(self.lib.spb_path_sequences_ex(self.h, 1, len(paths), off.ctypes.data, items.ctypes.data, voff.ctypes.data, None))
        else:
            voff = off
        out = np.empty(int(voff[-1]) + len(paths) * (k - 1), dtype=np.uint8)
        self._ck(self.lib.spb_path_sequences_ex(self.h, 1 if nbp_ids else 0, len(paths), off.ctypes.data, items.ctypes.data, None, out.ctypes.data))
        b = out.tobytes()
        return [b[int(voff[i]) + i * (k - 1):int(voff[i + 1]) + (i + 1) * (k - 1)].decode() for i in range(len(paths))]

    def path_origins(self, paths, nbp_ids=False, bubble_vertex2origins=None):
        """Per path (concatenation of whole NBPs, as vertex ids or -- nbp_ids=True -- raw NBP ids) the ascending indices of
        the sequences it originates from (reference get_ori2cov_onevertex, lib/Assembler.py:1212-1248, cov >= 1).
        bubble_vertex2origins: {vertex: iterable of sequence indices} left behind by popped bubbles (:922, :1236), united
        with the origins of every vertex the rule looks at."""
        paths = [np.asarray(p, dtype=np.uint32) for p in paths]
        if not paths:
            return []
        off = np.zeros(len(paths) + 1, dtype=np.uint64)
        np.cumsum([len(p) for p in paths], out=off[1:])
        items = np.ascontiguousarray(np.concatenate(paths))
        bub = sorted((int(v), sorted(set(int(x) for x in l))) for v, l in (bubble_vertex2origins or {}).items() if len(l))
        bv = np.asarray([v for v, _ in bub], dtype=np.uint32)
        bo = np.zeros(len(bub) + 1, dtype=np.uint64)
        np.cumsum([len(l) for _, l in bub], out=bo[1:])
        bs = np.asarray([x for _, l in bub for x in l], dtype=np.uint32)
        ooff = np.zeros(len(paths) + 1, dtype=np.uint64)
        cap, n = 64 * len(paths), C.c_uint64()
        for _ in range(2):
            ids = np.empty(max(cap, 1), dtype=np.uint32)
            self._ck(self.lib.spb_path_origins_ex(self.h, 1 if nbp_ids else 0, len(paths), off.ctypes.data, items.ctypes.data, len(bub),
                                                  bv.ctypes.data if len(bub) else None, bo.ctypes.data if len(bub) else None,
                                                  bs.ctypes.data if len(bub) else None, cap, C.byref(n), ooff.ctypes.data, ids.ctypes.data))
            if n.value <= cap:
                break
            cap = n.value
        return [ids[int(ooff[i]):int(ooff[i + 1])].tolist() for i in range(len(paths))]

    def bubble_groups(self):
        """group[i] = smallest raw NBP id among the NBPs sharing (first, last) vertex with NBP i, 0xFFFFFFFF if none
        (the reference's bubble candidates, lib/Assembler.py:728-734)."""
        out = np.full(self.counts.n_nbp, 0xFFFFFFFF, dtype=np.uint32)
        if out.size:
            self._ck(self.lib.spb_get_bubble_groups(self.h, out.ctypes.data))
        return out

    def nbp_origins(self):
        off = np.empty(self.counts.n_nbp + 1, dtype=np.uint64)
        idx = np.empty(self.counts.n_origin_pairs, dtype=np.uint32)
        self._ck(self.lib.spb_get_nbp_origins(self.h, off.ctypes.data, idx.ctypes.data))
        return off, idx
