"""
Order-sensitive 64-bit checksums of the result buffers, computed ON THE DEVICE (torch is only the buffer / reduction
plumbing).  Used by `bench.py` (`result_digest` on every JSON line, so that runs at different GPU counts can be compared)
and by the benchmark-scale parity tests, where copying 10+ GB of results to the host to hash them would dominate.

checksum(x) = sum_i (x_i + 1) * w_i  (mod 2^64),  w_i = odd 64-bit mix of the index i  -- position dependent, so a
permutation or an off-by-one shift of the buffer changes it; x is read in its own integer width.
"""
import numpy as np

_M1 = -7046029254386353131          # 0x9E3779B97F4A7C15 as int64
_M2 = -4658895280553007687          # 0xBF58476D1CE4E5B9 as int64
_CH = 1 << 26


def checksum(t, base=0):
    """64-bit order-sensitive checksum of a 1-D integer torch tensor (device or host); returns a Python int.  `base` = index
    of t[0] in the complete array: the checksum of an array is the sum (mod 2^64) of the checksums of its parts."""
    import torch
    t = t.reshape(-1)
    acc = torch.zeros((), dtype=torch.int64, device=t.device)
    for a in range(0, t.numel(), _CH):
        x = t[a:a + _CH].to(torch.int64)
        i = torch.arange(base + a, base + a + x.numel(), dtype=torch.int64, device=t.device)
        w = (i * _M1) ^ ((i >> 9) * _M2)
        w = (w ^ (w >> 29)) | 1
        acc += ((x + 1) * w).sum()
    return int(acc.item()) & 0xFFFFFFFFFFFFFFFF


def _filled(eng, n, dtype, fill):
    """Caller-owned device tensor of n elements filled by a C-ABI getter (`fill(ptr)`)."""
    import torch
    dev = torch.device("cuda", eng.device) if eng.kind.startswith("cuda") else torch.device("cpu")
    t = torch.empty(max(int(n), 1), dtype=dtype, device=dev)
    eng._ck(fill(t.data_ptr()))
    return t[:int(n)]


def result_digest(eng, graph=True):
    """{name: checksum} over the result buffers of a compacted engine: NBP sequences / origins / vertex lists (+ offsets),
    and with graph=True also vertex2coords, the edge list and the per-instance vertex ids.  Every buffer goes through
    the public getters into a caller-owned device tensor (no host copy)."""
    import torch
    L, h, c = eng.lib, eng.h, eng.counts
    n = int(c.n_nbp)
    out = {}
    soff = _filled(eng, n + 1, torch.int64, lambda p: L.spb_get_nbp_seqs(h, p, None))
    out["nbp_seq_off"] = checksum(soff)
    out["nbp_seqs"] = checksum(_filled(eng, c.nbp_bases, torch.uint8, lambda p: L.spb_get_nbp_seqs(h, None, p)))
    ooff = _filled(eng, n + 1, torch.int64, lambda p: L.spb_get_nbp_origins(h, p, None))
    out["nbp_origin_off"] = checksum(ooff)
    out["nbp_origins"] = checksum(_filled(eng, c.n_origin_pairs, torch.int32, lambda p: L.spb_get_nbp_origins(h, None, p)))
    out["nbp_vert_off"] = checksum(_filled(eng, n + 1, torch.int64, lambda p: L.spb_get_nbps(h, p, None, None, None)))
    out["nbp_verts"] = checksum(_filled(eng, c.nbp_vertices, torch.int32, lambda p: L.spb_get_nbps(h, None, p, None, None)))
    out["nbp_comp"] = checksum(_filled(eng, n, torch.int32, lambda p: L.spb_get_nbps(h, None, None, p, None)))
    out["nbp_cycle"] = checksum(_filled(eng, n, torch.uint8, lambda p: L.spb_get_nbps(h, None, None, None, p)))
    if graph:
        out["vertex2coords"] = checksum(_filled(eng, 2 * c.n_vertices, torch.int32, lambda p: L.spb_get_vertex2coords(h, p)))
        out["edges"] = checksum(_filled(eng, 2 * c.n_edges, torch.int32, lambda p: L.spb_get_edges(h, p)))
        out["instance_vertices"] = checksum(_filled(eng, c.n_bases, torch.int32, lambda p: L.spb_get_instance_vertices(h, p)))
    out["counts"] = {k_: int(getattr(c, k_)) for k_ in ("n_inst", "n_vertices", "n_edges", "n_nbp", "nbp_vertices", "nbp_bases",
                                                         "n_origin_pairs", "n_comp", "n_cycle_comp")}
    return out


def result_digest_sliced(eng, group=None):
    """result_digest(graph=True) of a result that is sliced over the ranks (Engine._mg_sliced_tail): every rank checksums
    its local parts with their global indices, the sums are all-reduced; offsets, origins, components, cycle flags and
    counts are complete on every rank.  Equal to the digest of the same input on one GPU iff the concatenation of the
    local parts is the single-GPU result.  Collective: every rank must call it; all ranks get the digest."""
    import torch
    import torch.distributed as dist
    L, h, c = eng.lib, eng.h, eng.counts
    n = int(c.n_nbp)
    out = {}
    out["nbp_seq_off"] = checksum(_filled(eng, n + 1, torch.int64, lambda p: L.spb_get_nbp_seqs(h, p, None)))
    ooff = _filled(eng, n + 1, torch.int64, lambda p: L.spb_get_nbp_origins(h, p, None))
    out["nbp_origin_off"] = checksum(ooff)
    out["nbp_origins"] = checksum(_filled(eng, c.n_origin_pairs, torch.int32, lambda p: L.spb_get_nbp_origins(h, None, p)))
    out["nbp_vert_off"] = checksum(_filled(eng, n + 1, torch.int64, lambda p: L.spb_get_nbps(h, p, None, None, None)))
    out["nbp_comp"] = checksum(_filled(eng, n, torch.int32, lambda p: L.spb_get_nbps(h, None, None, p, None)))
    out["nbp_cycle"] = checksum(_filled(eng, n, torch.uint8, lambda p: L.spb_get_nbps(h, None, None, None, p)))
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts, rows = {}, {}
    for name, which, width in (("nbp_seqs", "nbp_seqs", 1), ("nbp_verts", "nbp_verts", 1), ("vertex2coords", "vertex2coords", 2),
                               ("edges", "edges", 2), ("instance_vertices", "instance_vertices", 1)):
        t, first = eng.shard(which)
        parts[name] = (t, first, width)
        rows[name] = t.numel() // width
    dev = parts["edges"][0].device
    er = torch.zeros(world, dtype=torch.int64, device=dev)
    er[rank] = rows["edges"]
    dist.all_reduce(er, group=group)
    edge_first = int(er[:rank].sum().item())
    sums = []
    for name in ("nbp_seqs", "nbp_verts", "vertex2coords", "edges", "instance_vertices"):
        t, first, width = parts[name]
        if name == "edges":
            first = edge_first
        if name == "instance_vertices":                        # the single-GPU digest covers the n_bases positions of the input
            t = t[:max(0, min(t.numel(), int(c.n_bases) - first))]
        v = checksum(t, base=first * width)
        sums.append(v - (1 << 64) if v >> 63 else v)
    tot = torch.tensor(sums, dtype=torch.int64, device=dev)
    dist.all_reduce(tot, group=group)                          # int64 sums wrap: mod 2^64
    for name, v in zip(("nbp_seqs", "nbp_verts", "vertex2coords", "edges", "instance_vertices"), tot.tolist()):
        out[name] = int(v) & 0xFFFFFFFFFFFFFFFF
    out["counts"] = {k_: int(getattr(c, k_)) for k_ in ("n_inst", "n_vertices", "n_edges", "n_nbp", "nbp_vertices", "nbp_bases",
                                                         "n_origin_pairs", "n_comp", "n_cycle_comp")}
    return out


def fold(d):
    """One hex string for a digest dict (sha256 of its canonical JSON)."""
    import hashlib
    import json
    return hashlib.sha256(json.dumps(d, sort_keys=True).encode()).hexdigest()[:32]


def independent_vertex_count(ascii_dev, seq_off, k):
    """N_v from a computation that shares nothing with the library: an orientation-symmetric 64-bit polynomial key of
    every k-mer instance from prefix sums (forward string hash + reverse-complement string hash are swapped by taking the
    reverse complement, so their symmetric combination identifies the canonical k-mer), then sort + count distinct.
    ascii_dev: uint8 device tensor of the concatenated input; seq_off: uint64[n+1] host offsets."""
    import torch
    dev = ascii_dev.device
    T = ascii_dev.numel()
    lut = torch.zeros(256, dtype=torch.int64, device=dev)
    for ch, v in zip(b"ACGT", (0, 1, 2, 3)):
        lut[ch] = v
    x = lut[ascii_dev.long()]
    B, Binv = 0x100000001B3 | 1, pow(0x100000001B3 | 1, -1, 1 << 64)

    def s64(v):
        v &= (1 << 64) - 1
        return v - (1 << 64) if v >> 63 else v

    def powers(b):                     # b^i for i in [0, T]: cumprod wraps mod 2^64
        p = torch.full((T + 1,), s64(b), dtype=torch.int64, device=dev)
        p[0] = 1
        return torch.cumprod(p, 0)
    pB, pI = powers(B), powers(Binv)
    zero = torch.zeros(1, dtype=torch.int64, device=dev)
    Sf = torch.cat([zero, torch.cumsum((x + 1) * pI[:T], 0)])          # sum (x_i + 1) Binv^i
    Sr = torch.cat([zero, torch.cumsum((4 - x) * pB[:T], 0)])          # sum (comp_i + 1) B^i
    del x
    keys = []
    so = [int(v) for v in seq_off]
    Bk1 = s64(pow(Binv, k - 1, 1 << 64))
    for a, b in zip(so[:-1], so[1:]):
        if b - a <= k:
            continue
        p = torch.arange(a, b - k + 1, dtype=torch.int64, device=dev)
        wf = (Sf[p + k] - Sf[p]) * pB[p]                                 # sum_i (x[p+i]+1) Binv^i
        wr = (Sr[p + k] - Sr[p]) * pI[p] * Bk1                           # sum_i (comp[p+k-1-i]+1) Binv^i
        lo, hi = torch.minimum(wf, wr), torch.maximum(wf, wr)
        keys.append(lo * _M1 + (hi ^ (hi >> 31)) * _M2)
    del Sf, Sr, pB, pI
    if not keys:
        return 0
    allk = torch.cat(keys)
    del keys
    allk = torch.sort(allk).values
    return int((allk[1:] != allk[:-1]).sum().item()) + 1
