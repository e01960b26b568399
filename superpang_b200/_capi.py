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


class Engine:
    """One context = one GPU + one stream.  Mirrors the call order of the C-ABI."""

    def __init__(self, device=0, stream=None, lib=None):
        self.lib = lib if lib is not None else load_library()
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
        self._ck(self.lib.spb_load(self.h, _ptr(ascii_buf), seq_off.ctypes.data, len(seq_off) - 1, int(k)))

    def build_dbg(self):
        self._ck(self.lib.spb_build_dbg(self.h, C.byref(self.counts)))
        return self.counts.as_dict()

    def compact(self):
        self._ck(self.lib.spb_compact(self.h, C.byref(self.counts)))
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

    def nbp_origins(self):
        off = np.empty(self.counts.n_nbp + 1, dtype=np.uint64)
        idx = np.empty(self.counts.n_origin_pairs, dtype=np.uint32)
        self._ck(self.lib.spb_get_nbp_origins(self.h, off.ctypes.data, idx.ctypes.data))
        return off, idx
