"""ctypes binding of ``include/basal_gpu.h`` (libbasal_gpu.so).

This is test / bench plumbing around the C-ABI, not a compute path: every call
goes straight into the CUDA library and raises when it is missing or when no
sm_100 device can be opened.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libbasal_gpu.so")

BSL_ST_UNMAPPED, BSL_ST_UNIQUE, BSL_ST_MULTI, BSL_ST_FILTERED, BSL_ST_PAIRED = 0, 1, 2, 3, 4


class Params(C.Structure):
    _fields_ = [("from_base", C.c_char), ("to_bases", C.c_char * 7),
                ("seed_size", C.c_uint32), ("index_interval", C.c_uint32), ("max_snp_num", C.c_uint32),
                ("gap", C.c_uint32), ("max_num_hits", C.c_uint32), ("min_insert", C.c_uint32),
                ("max_insert", C.c_uint32), ("chains", C.c_uint32), ("report_repeat_hits", C.c_uint32),
                ("randseed", C.c_uint32), ("max_ns", C.c_uint32), ("min_read_size", C.c_uint32),
                ("max_kmer_ratio", C.c_float), ("reserved", C.c_uint32 * 4)]


class Batch(C.Structure):
    _fields_ = [("n", C.c_uint32), ("readset", C.c_uint32), ("bases", C.c_void_p), ("offsets", C.c_void_p),
                ("index", C.c_void_p), ("first_index", C.c_uint32), ("n_context", C.c_uint32), ("raw_len", C.c_void_p)]


HIT_DTYPE = np.dtype([("loc", "<u4"), ("chr", "<u4"), ("n_hits", "<u4"), ("n_chain0", "<u4"), ("gap_size", "<i4"),
                      ("gap_pos", "<u2"), ("nm", "u1"), ("status", "u1"), ("read_chain", "u1"), ("max_snp", "u1"),
                      ("read_len", "<u2"), ("all_first", "<u4")], align=True)
PAIR_DTYPE = np.dtype([("n_pairs", "<u4"), ("insert", "<u4"), ("chain", "u1"), ("na", "u1"), ("nb", "u1"),
                       ("reserved", "u1"), ("all_first", "<u4")], align=True)
assert HIT_DTYPE.itemsize == 32 and PAIR_DTYPE.itemsize == 16


class IndexInfo(C.Structure):
    _fields_ = [("n_seq", C.c_uint32), ("n_kmers", C.c_uint32), ("sum_length", C.c_uint64), ("n_words", C.c_uint64),
                ("n_entries", C.c_uint64), ("max_kmer_num", C.c_uint32), ("reserved", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("reads", C.c_uint64), ("seed_lookups", C.c_uint64), ("candidates", C.c_uint64), ("hits_added", C.c_uint64),
                ("heavy_reads", C.c_uint64), ("ms_pack", C.c_double), ("ms_search", C.c_double), ("ms_pair", C.c_double),
                ("ms_total", C.c_double), ("kernel_launches", C.c_uint64), ("verify_bytes", C.c_uint64),
                ("ms_device", C.c_double), ("search_launches", C.c_uint64),
                ("ms_lookup", C.c_double), ("ms_verify", C.c_double), ("ms_reduce", C.c_double)]


def parse_v(text: str) -> int:
    """-v as the reference stores it (main.cpp:324-338)."""
    t = float(text)
    if t < 1.0:
        v = int(t * 100 + 0.5) + 100
        return 0 if v == 100 else v
    return min(int(t + 0.5), 15)


def make_params(rule: str = "C:T", s: int = 16, I: int = 4, v: str = "0.1", g: int = 0, w: int = 100, m: int = 28,
                x: int = 1000, n: int = 0, r: int = 1, S: int = 7, f: int = 5, k: float = 5e-7,
                min_read_size: Optional[int] = None, s_given: bool = False) -> Params:
    """Mirror of Param defaults + mGetOptions for the hot-path flags (param.cpp:7-68, main.cpp:272-364)."""
    p = Params()
    p.from_base = rule[0:1].encode()
    p.to_bases = rule[2:].encode()
    p.seed_size, p.index_interval, p.max_snp_num, p.gap, p.max_num_hits = s, I, parse_v(v), min(g, 3), w
    p.min_insert, p.max_insert, p.chains, p.report_repeat_hits, p.randseed, p.max_ns = m, x, n, r, S, f
    if min_read_size is None:
        min_read_size = (s + I - 1) if s_given else 16      # param.cpp:34 vs :112 (only -s recomputes it)
    p.min_read_size = min_read_size
    p.max_kmer_ratio = k
    return p


class _Api:
    """Function table of one shared library exporting <prefix>* symbols of basal_gpu.h."""

    def __init__(self, lib: C.CDLL, prefix: str, has_device: bool):
        self.lib, self.prefix = lib, prefix
        f = lambda name: getattr(lib, prefix + name)
        vp, u64 = C.c_void_p, C.c_uint64
        self.ctx_create = f("ctx_create")
        self.ctx_create.restype = C.c_int
        self.ctx_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(Params)] if has_device else [C.POINTER(vp), C.POINTER(Params)]
        self.ctx_destroy = f("ctx_destroy"); self.ctx_destroy.argtypes = [vp]; self.ctx_destroy.restype = None
        self.index_build = f("index_build"); self.index_build.restype = C.c_int
        self.index_build.argtypes = [vp, vp, vp, vp, C.c_uint32]
        self.index_info_get = f("index_info_get"); self.index_info_get.argtypes = [vp, C.POINTER(IndexInfo)]
        self.index_download = f("index_download"); self.index_download.argtypes = [vp, vp, vp, vp, vp, vp]
        self.align_se = f("align_se"); self.align_se.restype = C.c_int
        self.align_se.argtypes = [vp, C.POINTER(Batch), vp, vp, u64, C.POINTER(u64)]
        self.align_pe = f("align_pe"); self.align_pe.restype = C.c_int
        self.align_pe.argtypes = [vp, C.POINTER(Batch), C.POINTER(Batch), vp, vp, vp, vp, vp, u64, C.POINTER(u64)]
        self.stats_get = f("stats_get"); self.stats_get.argtypes = [vp, C.POINTER(Stats)]
        self.read_budget = f("read_budget"); self.read_budget.restype = C.c_uint32
        self.read_budget.argtypes = [C.POINTER(Params), C.c_uint32, C.c_uint32]
        self.myrand = f("myrand"); self.myrand.restype = C.c_uint32; self.myrand.argtypes = [C.c_uint32, C.c_uint32]
        self.has_device = has_device
        if has_device:
            self.last_error = f("last_error"); self.last_error.restype = C.c_char_p; self.last_error.argtypes = [vp]
            self.host_alloc = f("host_alloc"); self.host_alloc.restype = vp; self.host_alloc.argtypes = [C.c_size_t]
            self.host_free = f("host_free"); self.host_free.argtypes = [vp]; self.host_free.restype = None
            self.abi_version = f("abi_version"); self.abi_version.restype = C.c_int
            self.align_rerun = f("align_rerun"); self.align_rerun.restype = C.c_int
            self.align_rerun.argtypes = [vp, C.POINTER(Batch), C.POINTER(Batch)]


_gpu_api: Optional[_Api] = None


def load_library(path: str = LIB_PATH) -> _Api:
    """dlopen libbasal_gpu.so; raises if it has not been built (no fallback)."""
    global _gpu_api
    if _gpu_api is None:
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `make -C basal_b200/csrc` "
                               f"(or __graft_entry__.build()); there is no CPU fallback")
        _gpu_api = _Api(C.CDLL(path), "bsl_", True)
    return _gpu_api


class BasalError(RuntimeError):
    pass


def _as_u8(a) -> np.ndarray:
    return np.ascontiguousarray(np.frombuffer(a, dtype=np.uint8) if isinstance(a, (bytes, bytearray)) else a, dtype=np.uint8)


class ReadBatch:
    """Host-side batch: concatenated ASCII bases + offsets (bsl_batch)."""

    def __init__(self, bases: np.ndarray, offsets: np.ndarray, readset: int = 0, first_index: int = 0,
                 index: Optional[np.ndarray] = None, raw_len: Optional[np.ndarray] = None, n_context: int = 0):
        self.n_context = n_context          # the first n_context reads only re-establish the carried aligner state (bsl_batch::n_context)
        self.bases = _as_u8(bases)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self.n = len(self.offsets) - 1
        self.readset, self.first_index = readset, first_index
        self.index = None if index is None else np.ascontiguousarray(index, dtype=np.uint32)
        self.raw_len = None if raw_len is None else np.ascontiguousarray(raw_len, dtype=np.uint16)

    @classmethod
    def from_matrix(cls, reads: np.ndarray, readset: int = 0, first_index: int = 0) -> "ReadBatch":
        n, L = reads.shape
        return cls(reads.reshape(-1), np.arange(n + 1, dtype=np.uint64) * L, readset, first_index)

    @classmethod
    def from_strings(cls, seqs, readset: int = 0, first_index: int = 0, n_context: int = 0) -> "ReadBatch":
        lens = np.array([len(s) for s in seqs], dtype=np.uint64)
        off = np.zeros(len(seqs) + 1, dtype=np.uint64); off[1:] = np.cumsum(lens)
        cat = np.frombuffer("".join(seqs).encode(), dtype=np.uint8) if seqs else np.zeros(0, np.uint8)
        return cls(cat, off, readset, first_index, n_context=n_context)

    def struct(self) -> Batch:
        b = Batch()
        b.n, b.readset, b.first_index, b.n_context = self.n, self.readset, self.first_index, self.n_context
        b.bases = self.bases.ctypes.data if self.bases.size else None
        b.offsets = self.offsets.ctypes.data
        b.index = None if self.index is None else self.index.ctypes.data
        b.raw_len = None if self.raw_len is None else self.raw_len.ctypes.data
        return b


class Context:
    """One aligner context (one GPU): mirrors RefSeq + SingleAlign/PairAlign of the reference."""

    def __init__(self, params: Params, device: int = 0, api: Optional[_Api] = None):
        self.api = api or load_library()
        self.params = params
        self._h = C.c_void_p()
        rc = (self.api.ctx_create(C.byref(self._h), device, C.byref(params)) if self.api.has_device
              else self.api.ctx_create(C.byref(self._h), C.byref(params)))
        if rc != 0:
            msg = self.api.last_error(None).decode() if self.api.has_device else "oracle: invalid parameters"
            raise BasalError(f"ctx_create failed ({rc}): {msg}")

    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self.api.last_error(self._h).decode() if self.api.has_device else ""
            raise BasalError(f"{what} failed ({rc}): {msg}")

    def close(self):
        if self._h:
            self.api.ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- index seam
    def index_build(self, cat: np.ndarray, offs: np.ndarray, lens: np.ndarray):
        cat = _as_u8(cat); offs = np.ascontiguousarray(offs, np.uint64); lens = np.ascontiguousarray(lens, np.uint32)
        self._check(self.api.index_build(self._h, cat.ctypes.data, offs.ctypes.data, lens.ctypes.data, len(lens)), "index_build")

    def index_info(self) -> IndexInfo:
        info = IndexInfo(); self.api.index_info_get(self._h, C.byref(info)); return info

    def index_download(self, planes: bool = True):
        info = self.index_info()
        bs = np.empty(info.n_kmers + 1, np.uint32); nf = np.empty(info.n_kmers, np.uint32); loc = np.empty(info.n_entries, np.uint32)
        fw = np.empty(info.n_words, np.uint64) if planes else None; rc = np.empty(info.n_words, np.uint64) if planes else None
        r = self.api.index_download(self._h, bs.ctypes.data, nf.ctypes.data, loc.ctypes.data if loc.size else None,
                                    fw.ctypes.data if planes else None, rc.ctypes.data if planes else None)
        self._check(r, "index_download")
        return bs, nf, loc, fw, rc

    # ---- batch seam
    def align_se(self, batch: ReadBatch, all_cap: int = 0, out: Optional[np.ndarray] = None):
        if out is None:
            out = np.zeros(batch.n, dtype=HIT_DTYPE)
        allh = np.zeros(all_cap, dtype=HIT_DTYPE) if all_cap else None
        n_all = C.c_uint64(0); b = batch.struct()
        rc = self.api.align_se(self._h, C.byref(b), out.ctypes.data, allh.ctypes.data if all_cap else None, all_cap, C.byref(n_all))
        self._check(rc, "align_se")
        return (out, allh[:min(n_all.value, all_cap)]) if all_cap else out

    def align_pe(self, a: ReadBatch, b: ReadBatch, all_cap: int = 0, out=None):
        if out is None:
            oa = np.zeros(a.n, dtype=HIT_DTYPE); ob = np.zeros(a.n, dtype=HIT_DTYPE); op = np.zeros(a.n, dtype=PAIR_DTYPE)
        else:
            oa, ob, op = out
        alla = np.zeros(all_cap, dtype=HIT_DTYPE) if all_cap else None
        allb = np.zeros(all_cap, dtype=HIT_DTYPE) if all_cap else None
        n_all = C.c_uint64(0); sa, sb = a.struct(), b.struct()
        rc = self.api.align_pe(self._h, C.byref(sa), C.byref(sb), oa.ctypes.data, ob.ctypes.data, op.ctypes.data,
                               alla.ctypes.data if all_cap else None, allb.ctypes.data if all_cap else None, all_cap, C.byref(n_all))
        self._check(rc, "align_pe")
        if all_cap:
            k = min(n_all.value, all_cap)
            return oa, ob, op, alla[:k], allb[:k]
        return oa, ob, op

    def align_rerun(self, a: ReadBatch, b: Optional[ReadBatch] = None):
        """Kernels only, on the batch the previous align call left resident on the device (bench hook)."""
        sa = a.struct(); sb = b.struct() if b is not None else None
        self._check(self.api.align_rerun(self._h, C.byref(sa), C.byref(sb) if b is not None else None), "align_rerun")

    def stats(self) -> Stats:
        s = Stats(); self.api.stats_get(self._h, C.byref(s)); return s


def pinned_array(nbytes: int, dtype=np.uint8) -> np.ndarray:
    """numpy view on cudaMallocHost memory (bsl_host_alloc); keep the array alive while in use."""
    api = load_library()
    ptr = api.host_alloc(nbytes)
    if not ptr:
        raise BasalError("bsl_host_alloc failed")
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    arr = np.frombuffer(buf, dtype=np.uint8).view(dtype)
    return arr
