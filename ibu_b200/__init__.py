"""ibu_b200 — B200-native bulk record path of the `ibu` binary format.

Host-side mirror of the reference's interface for this path (same names, argument meaning
and error behaviour as the Rust crate, citations are into noamteyssier/ibu):

    Header, Record (RECORD_DTYPE)            src/constructs/{header,record}.rs
    MmapReader.{len, header, slice}          src/io/mmap.rs:143-270
    MmapReader.process_gpu                   GPU counterpart of process_parallel (mmap.rs:286-332)
    load_to_vec / load_to_device             src/io/reader.rs:510-535 and its device path
    Writer                                   src/io/writer.rs
    IbuError subclasses                      src/error.rs:56-128

Everything calls the C ABI of libibu_b200.so (include/ibu_b200.h) through ctypes; there is
no Python or CPU implementation of the record-processing work in this package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import lib

MAGIC = 0x21554249
VERSION = 2
HEADER_SIZE = 32
RECORD_SIZE = 24
BATCH_SIZE = 1024 * 1024

RECORD_DTYPE = np.dtype([("barcode", "<u8"), ("umi", "<u8"), ("index", "<u8")])
ROW_DTYPE = np.dtype([("barcode", "<u8"), ("n_records", "<u8"), ("n_distinct_umi", "<u8")])

GEN_CLEAN, GEN_DIRTY, GEN_PATTERN, GEN_WHITELIST, GEN_SORTED, GEN_ZIPF = 0, 1, 2, 3, 4, 5
COUNT_WEIGHTED = 8
COUNT_PATH_PARTITION, COUNT_PATH_SORT, COUNT_PATH_LEGACY = 0x10, 0x20, 0x30
PAIRS_WEIGHTED, PAIRS_UNORDERED = 1, 2
OP_REDUCE, OP_TABLE, OP_KEEP, OP_UNPACK = 1, 2, 4, 8
EXCHANGE_AUTO, EXCHANGE_P2P, EXCHANGE_HOST, EXCHANGE_NCCL = 0, 1, 2, 3


def count_lens(bc_len: int, umi_len: int) -> int:
    """IBU_COUNT_LENS: the header's lengths, OR-ed into `mode` of barcode_count / `flags` of pair_table."""
    return ((bc_len & 0x3F) << 8) | ((umi_len & 0x3F) << 16)


# ---- errors (src/error.rs:56-128) ---------------------------------------------------------
class IbuError(Exception):
    code = -1

    def __init__(self, err: _lib.Error | None = None, msg: str | None = None):
        self.sys = int(err.sys) if err is not None else 0
        self.a = int(err.a) if err is not None else 0
        self.b = int(err.b) if err is not None else 0
        text = msg or (err.msg.decode(errors="replace") if err is not None else "")
        super().__init__(text or lib.ibu_strerror(self.code).decode())


class Io(IbuError, OSError):
    code = 1


class InvalidMagicNumber(IbuError):
    code = 3

    @property
    def expected(self):
        return self.a

    @property
    def actual(self):
        return self.b


class TruncatedRecord(IbuError):
    code = 4

    @property
    def pos(self):
        return self.a


class InvalidVersion(InvalidMagicNumber):
    code = 5


class InvalidBarcodeLength(IbuError):
    code = 6


class InvalidUmiLength(IbuError):
    code = 7


class InvalidMapSize(IbuError):
    code = 8


class InvalidIndex(IbuError):
    code = 9

    @property
    def idx(self):
        return self.a

    @property
    def max(self):
        return self.b


class Process(IbuError):
    code = 10


class CudaError(IbuError):
    code = 11


class NcclError(IbuError):
    code = 12


class ArgError(IbuError, ValueError):
    code = 13


class NoMemory(IbuError, MemoryError):
    code = 14


_BY_CODE = {c.code: c for c in (Io, InvalidMagicNumber, TruncatedRecord, InvalidVersion, InvalidBarcodeLength,
                                InvalidUmiLength, InvalidMapSize, InvalidIndex, Process, CudaError, NcclError,
                                ArgError, NoMemory)}


def _check(rc: int, err: _lib.Error):
    if rc != 0:
        raise _BY_CODE.get(rc, IbuError)(err)


def _ptr(x) -> C.c_void_p:
    """Raw address of a device tensor (`.data_ptr()`), numpy array, ctypes pointer or int."""
    if x is None:
        return C.c_void_p(None)
    if isinstance(x, int):
        return C.c_void_p(x)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    if isinstance(x, C.c_void_p):
        return x
    raise TypeError(f"cannot take the address of {type(x)!r}")


def _stream(s) -> C.c_void_p:
    if s is None:
        return C.c_void_p(None)
    return C.c_void_p(getattr(s, "cuda_stream", s))


# ---- Header (src/constructs/header.rs) ------------------------------------------------------
class Header:
    __slots__ = ("_h",)

    def __init__(self, bc_len: int, umi_len: int):
        self._h = _lib.Header()
        lib.ibu_header_init(C.byref(self._h), bc_len, umi_len)

    @classmethod
    def _wrap(cls, raw: _lib.Header) -> "Header":
        h = cls.__new__(cls)
        h._h = raw
        return h

    magic = property(lambda s: s._h.magic)
    version = property(lambda s: s._h.version)
    bc_len = property(lambda s: s._h.bc_len)
    umi_len = property(lambda s: s._h.umi_len)
    flags = property(lambda s: s._h.flags)

    def set_sorted(self):
        lib.ibu_header_set_sorted(C.byref(self._h))

    def sorted(self) -> bool:
        return bool(lib.ibu_header_sorted(C.byref(self._h)))

    def validate(self):
        err = _lib.Error()
        _check(lib.ibu_header_validate(C.byref(self._h), C.byref(err)), err)

    def as_bytes(self) -> bytes:
        return bytes(self._h)

    @classmethod
    def from_bytes(cls, raw: bytes) -> "Header":
        if len(raw) != HEADER_SIZE:  # bytemuck::from_bytes panics on a wrong length
            raise ValueError(f"Header.from_bytes needs exactly {HEADER_SIZE} bytes")
        return cls._wrap(_lib.Header.from_buffer_copy(raw))

    def __eq__(self, other):
        return isinstance(other, Header) and self.as_bytes() == other.as_bytes()

    def __repr__(self):
        return (f"Header(magic={self.magic:#x}, version={self.version}, bc_len={self.bc_len}, "
                f"umi_len={self.umi_len}, flags={self.flags})")


def records(n: int) -> np.ndarray:
    """`vec![Record::default(); n]`."""
    return np.zeros(n, RECORD_DTYPE)


def shard_range(length: int, rank: int, world: int) -> tuple[int, int]:
    """Range of shard `rank`: the thread partition of process_parallel (mmap.rs:297-307)."""
    s, e = C.c_uint64(), C.c_uint64()
    lib.ibu_shard_range(length, rank, world, C.byref(s), C.byref(e))
    return int(s.value), int(e.value)


# ---- results ------------------------------------------------------------------------------
class ReduceResult(dict):
    """n_records, sum_barcode, sum_umi, sum_index, xor_all, n_bad_barcode, n_bad_umi, n_bad_records."""

    @property
    def count_sum(self) -> int:
        """local_sum of the reference's count+sum processor (mmap.rs:359-363), wrapping u64."""
        return (self["sum_barcode"] + self["sum_umi"] + self["sum_index"]) & (2**64 - 1)

    def merge(self, other: "ReduceResult") -> "ReduceResult":
        """on_batch_complete merge: wrapping add, xor for the checksum."""
        out = ReduceResult(self)
        for k, v in other.items():
            out[k] = (out[k] ^ v) if k == "xor_all" else (out[k] + v) & (2**64 - 1)
        return out


# ---- GPU context ----------------------------------------------------------------------------
class GpuContext:
    """One per GPU (one process per GPU under torchrun)."""

    def __init__(self, device: int = 0, chunk_records: int = 0, n_slots: int = 0, copy_threads: int = 0):
        self._h = C.c_void_p()
        cfg = _lib.GpuConfig(chunk_records, n_slots, copy_threads, 0)
        err = _lib.Error()
        _check(lib.ibu_gpu_ctx_create(device, C.byref(cfg), C.byref(self._h), C.byref(err)), err)
        self.device = device
        self.sm_count = int(lib.ibu_gpu_ctx_sm_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            lib.ibu_gpu_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- memory helpers
    def malloc(self, nbytes: int) -> int:
        p, err = C.c_void_p(), _lib.Error()
        _check(lib.ibu_gpu_malloc(self._h, nbytes, C.byref(p), C.byref(err)), err)
        return int(p.value)

    def free(self, d_ptr):
        lib.ibu_gpu_free(self._h, _ptr(d_ptr))

    def h2d(self, d_dst, host: np.ndarray):
        host = np.ascontiguousarray(host)
        err = _lib.Error()
        _check(lib.ibu_gpu_memcpy_h2d(self._h, _ptr(d_dst), _ptr(host), host.nbytes, C.byref(err)), err)

    def d2h(self, host: np.ndarray, d_src):
        assert host.flags.c_contiguous
        err = _lib.Error()
        _check(lib.ibu_gpu_memcpy_d2h(self._h, _ptr(host), _ptr(d_src), host.nbytes, C.byref(err)), err)

    def memset(self, d_dst, value: int, nbytes: int):
        err = _lib.Error()
        _check(lib.ibu_gpu_memset(self._h, _ptr(d_dst), value, nbytes, C.byref(err)), err)

    def synchronize(self, stream=None):
        err = _lib.Error()
        _check(lib.ibu_gpu_synchronize(self._h, _stream(stream), C.byref(err)), err)

    def read_result(self, d_result) -> ReduceResult:
        out = np.zeros(8, np.uint64)
        self.d2h(out, d_result)
        return ReduceResult(zip((k for k, _ in _lib.ReduceResult._fields_), map(int, out)))

    # -- device kernels (stream ordered)
    def validate_reduce_async(self, d_records, n, bc_len, umi_len, d_result, stream=None):
        err = _lib.Error()
        _check(lib.ibu_gpu_validate_reduce_async(self._h, _ptr(d_records), n, bc_len, umi_len, _ptr(d_result),
                                                 _stream(stream), C.byref(err)), err)

    def unpack_async(self, d_records, n, bc_len, umi_len, d_bc_ascii, d_umi_ascii, d_flags=None, d_result=None,
                     stream=None):
        err = _lib.Error()
        _check(lib.ibu_gpu_unpack_async(self._h, _ptr(d_records), n, bc_len, umi_len, _ptr(d_bc_ascii),
                                        _ptr(d_umi_ascii), _ptr(d_flags), _ptr(d_result), _stream(stream),
                                        C.byref(err)), err)

    def pack_async(self, d_bc_ascii, d_umi_ascii, n, bc_len, umi_len, d_records, d_index=None, index_base=0,
                   d_flags=None, d_result=None, stream=None):
        err = _lib.Error()
        _check(lib.ibu_gpu_pack_async(self._h, _ptr(d_bc_ascii), _ptr(d_umi_ascii), _ptr(d_index), index_base, n,
                                      bc_len, umi_len, _ptr(d_records), _ptr(d_flags), _ptr(d_result),
                                      _stream(stream), C.byref(err)), err)

    def generate_records_async(self, d_records, first, n, bc_len, umi_len, mode, param, seed, stream=None):
        err = _lib.Error()
        _check(lib.ibu_gpu_generate_records_async(self._h, _ptr(d_records), first, n, bc_len, umi_len, mode, param,
                                                  seed, _stream(stream), C.byref(err)), err)

    def generate_ascii_async(self, d_ascii, first_row, n_rows, length, dirty_ppm, lower_ppm, seed, stream=None):
        err = _lib.Error()
        _check(lib.ibu_gpu_generate_ascii_async(self._h, _ptr(d_ascii), first_row, n_rows, length, dirty_ppm,
                                                lower_ppm, seed, _stream(stream), C.byref(err)), err)

    def barcode_count_device(self, d_records, n, mode: int = 0, stream=None):
        """Per-barcode table left on the device: (table struct, info).  Release the rows with
        `lib.ibu_gpu_table_free` (or use `barcode_count`, which copies them to the host)."""
        table, err = _lib.BarcodeTable(), _lib.Error()
        _check(lib.ibu_gpu_barcode_count(self._h, _ptr(d_records), n, mode, C.byref(table), _stream(stream),
                                         C.byref(err)), err)
        info = dict(n_rows=int(table.n_rows), n_records=int(table.n_records),
                    n_distinct_pairs=int(table.n_distinct_pairs), input_was_sorted=bool(table.input_was_sorted))
        return table, info

    def table_free(self, table):
        lib.ibu_gpu_table_free(self._h, C.byref(table))

    def barcode_count(self, d_records, n, mode: int = 0, stream=None):
        """Per-barcode table (parallel.rs:79-98 + distinct UMIs): (rows[ROW_DTYPE], info dict)."""
        table, err = _lib.BarcodeTable(), _lib.Error()
        _check(lib.ibu_gpu_barcode_count(self._h, _ptr(d_records), n, mode, C.byref(table), _stream(stream),
                                         C.byref(err)), err)
        rows = np.zeros(int(table.n_rows), ROW_DTYPE)
        info = dict(n_rows=int(table.n_rows), n_records=int(table.n_records),
                    n_distinct_pairs=int(table.n_distinct_pairs), input_was_sorted=bool(table.input_was_sorted))
        try:
            self.table_rows(table, rows)
        finally:
            lib.ibu_gpu_table_free(self._h, C.byref(table))
        return rows, info

    def table_rows(self, table, rows: np.ndarray):
        """The rows of a device table into `rows` (ROW_DTYPE, table.n_rows long)."""
        err = _lib.Error()
        _check(lib.ibu_gpu_table_to_host(self._h, C.byref(table), _ptr(rows), C.byref(err)), err)

    def pair_table(self, d_records, n, weighted: bool = False, stream=None, flags: int = 0):
        """Distinct (barcode, umi) pairs with multiplicities as device records: (ptr, n_pairs).
        `flags`: PAIRS_UNORDERED, count_lens(...), COUNT_PATH_*.  Release with .free(ptr)."""
        p, cnt, err = C.c_void_p(), C.c_uint64(), _lib.Error()
        _check(lib.ibu_gpu_pair_table(self._h, _ptr(d_records), n, int(bool(weighted)) | flags, C.byref(p),
                                      C.byref(cnt), _stream(stream), C.byref(err)), err)
        return int(p.value or 0), int(cnt.value)

    def partition_by_owner(self, d_pairs, n, world: int, d_out, stream=None) -> list[int]:
        """Group pair-table rows by owner(barcode) = splitmix64(barcode) % world; returns bucket sizes."""
        counts, err = (C.c_uint64 * world)(), _lib.Error()
        _check(lib.ibu_gpu_partition_by_owner(self._h, _ptr(d_pairs), n, world, _ptr(d_out), counts, _stream(stream),
                                              C.byref(err)), err)
        return [int(c) for c in counts]

    def memcpy(self, dst, src, nbytes: int):
        err = _lib.Error()
        _check(lib.ibu_gpu_memcpy(self._h, _ptr(dst), _ptr(src), nbytes, C.byref(err)), err)

    def sort_records(self, d_records, n, d_sorted, stream=None):
        """Device sort by Record's Ord (barcode, umi, index) into d_sorted."""
        err = _lib.Error()
        _check(lib.ibu_gpu_sort_records(self._h, _ptr(d_records), n, _ptr(d_sorted), _stream(stream),
                                        C.byref(err)), err)

    # -- host-buffer (end-to-end) paths
    def process_host(self, h_records: np.ndarray, bc_len: int, umi_len: int, on_chunk=None) -> ReduceResult:
        recs = _as_records(h_records)
        res, err = _lib.ReduceResult(), _lib.Error()
        cb, keep = _wrap_cb(on_chunk)
        _check(lib.ibu_gpu_process_host(self._h, _ptr(recs), len(recs), bc_len, umi_len, C.byref(res), cb, None,
                                        C.byref(err)), err)
        return ReduceResult(res.as_dict())

    def process_host_ops(self, h_records: np.ndarray, bc_len: int, umi_len: int, table=False, keep=False, unpack=False,
                         flags=False, table_mode: int = 0, on_chunk=None, **outs):
        """One pass over host records with an operation mask (ibu_gpu_process_host_ops):
        (ReduceResult, ProcessOutput)."""
        recs = _as_records(h_records)

        def call(req, res, cb, err):
            return lib.ibu_gpu_process_host_ops(self._h, _ptr(recs), len(recs), bc_len, umi_len, req, res, cb, None, err)

        return _run_ops(self, call, len(recs), bc_len, umi_len, table, keep, unpack, flags, table_mode, on_chunk, **outs)

    def unpack_host(self, h_records, bc_len, umi_len, bc_out=None, umi_out=None, flags_out=None):
        recs = _as_records(h_records)
        n = len(recs)
        bc = bc_out if bc_out is not None else np.empty((n, bc_len), np.uint8)
        umi = umi_out if umi_out is not None else np.empty((n, umi_len), np.uint8)
        res, err = _lib.ReduceResult(), _lib.Error()
        _check(lib.ibu_gpu_unpack_host(self._h, _ptr(recs), n, bc_len, umi_len, _ptr(bc), _ptr(umi), _ptr(flags_out),
                                       C.byref(res), C.byref(err)), err)
        return bc, umi, ReduceResult(res.as_dict())

    def pack_host(self, bc_ascii, umi_ascii, index=None, index_base=0, out=None, flags_out=None):
        n, bc_len = bc_ascii.shape
        umi_len = umi_ascii.shape[1]
        recs = out if out is not None else np.empty(n, RECORD_DTYPE)
        res, err = _lib.ReduceResult(), _lib.Error()
        _check(lib.ibu_gpu_pack_host(self._h, _ptr(bc_ascii), _ptr(umi_ascii), _ptr(index), index_base, n, bc_len,
                                     umi_len, _ptr(recs), _ptr(flags_out), C.byref(res), C.byref(err)), err)
        return recs, ReduceResult(res.as_dict())


class ProcessOutput:
    """What an ops pass returns besides the reduction: `rows` / `table_info` (OP_TABLE; rows are a
    host copy, ROW_DTYPE), `records` (OP_KEEP: DeviceRecords), `bc_ascii` / `umi_ascii` / `flags` (OP_UNPACK)."""

    rows = table = table_info = records = bc_ascii = umi_ascii = flags = timing = shard_records = None


def _run_ops(ctx: "GpuContext", call, n: int, bc_len: int, umi_len: int, table: bool, keep: bool, unpack: bool,
             flags: bool, table_mode: int, on_chunk, bc_out=None, umi_out=None, flags_out=None, rows_on_device=False):
    out = ProcessOutput()
    req = _lib.ProcessRequest()
    req.ops = OP_REDUCE | (OP_TABLE if table else 0) | (OP_KEEP if keep else 0) | (OP_UNPACK if unpack else 0)
    req.table_mode = table_mode
    tab, dptr = _lib.BarcodeTable(), C.c_void_p()
    req.table = C.pointer(tab)
    req.d_records = C.pointer(dptr)
    if unpack:
        out.bc_ascii = bc_out if bc_out is not None else np.empty((n, bc_len), np.uint8)
        out.umi_ascii = umi_out if umi_out is not None else np.empty((n, umi_len), np.uint8)
        req.h_bc_ascii, req.h_umi_ascii = _ptr(out.bc_ascii), _ptr(out.umi_ascii)
        if flags or flags_out is not None:
            out.flags = flags_out if flags_out is not None else np.empty(n, np.uint8)
            req.h_flags = _ptr(out.flags)
    res, err = _lib.ReduceResult(), _lib.Error()
    cb, keep_cb = _wrap_cb(on_chunk)
    _check(call(C.byref(req), C.byref(res), cb, C.byref(err)), err)
    if table:
        out.table_info = dict(n_rows=int(tab.n_rows), n_records=int(tab.n_records),
                              n_distinct_pairs=int(tab.n_distinct_pairs), input_was_sorted=bool(tab.input_was_sorted))
        if rows_on_device:  # (benchmarks: the caller releases out.table with ctx.table_free)
            out.table = tab
        else:
            out.rows = np.zeros(int(tab.n_rows), ROW_DTYPE)
            try:
                ctx.table_rows(tab, out.rows)
            finally:
                lib.ibu_gpu_table_free(ctx._h, C.byref(tab))
    if keep:
        out.records = DeviceRecords(ctx, int(dptr.value or 0), n)
    return ReduceResult(res.as_dict()), out


def _prefer_bundled_nccl():
    """IBU_EXCHANGE_NCCL loads libnccl.so.2 at run time.  In a Python process the NCCL bundled with
    torch (site-packages/nvidia/nccl/lib) must be the one: a system copy loaded first would shadow
    it for a later `import torch` (same soname)."""
    if "IBU_B200_NCCL_LIB" in os.environ:
        return
    try:
        import importlib.util

        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["IBU_B200_NCCL_LIB"] = cand
                return
    except Exception:
        pass


class GpuGroup:
    """Several GPUs of one box behind one handle (ibu_gpu_group_*): records shard by contiguous range
    (mmap.rs:297-307), one host thread per GPU, results and tables merged inside the library."""

    def __init__(self, devices, chunk_records: int = 0, n_slots: int = 0, copy_threads: int = 0):
        _prefer_bundled_nccl()
        self.devices = list(devices)
        self._h = C.c_void_p()
        cfg = _lib.GpuConfig(chunk_records, n_slots, copy_threads, 0)
        arr = (C.c_int * len(self.devices))(*self.devices)
        err = _lib.Error()
        _check(lib.ibu_gpu_group_create(arr, len(self.devices), C.byref(cfg), C.byref(self._h), C.byref(err)), err)

    def __len__(self):
        return len(self.devices)

    def close(self):
        if getattr(self, "_h", None):
            lib.ibu_gpu_group_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ctx(self, rank: int) -> "GpuContext":
        """Borrowed view of rank's context (owned by the group: do not close)."""
        c = GpuContext.__new__(GpuContext)
        c._h = C.c_void_p(lib.ibu_gpu_group_ctx(self._h, rank))
        c.device = int(lib.ibu_gpu_ctx_device(c._h))
        c.sm_count = int(lib.ibu_gpu_ctx_sm_count(c._h))
        c.close = lambda: None
        return c

    @staticmethod
    def _take_table(tab: "_lib.HostTable"):
        rows = np.zeros(int(tab.n_rows), ROW_DTYPE)
        if tab.n_rows:
            C.memmove(rows.ctypes.data, tab.h_rows, rows.nbytes)
        lib.ibu_free(tab.h_rows)
        info = dict(n_rows=int(tab.n_rows), n_records=int(tab.n_records), n_distinct_pairs=int(tab.n_distinct_pairs))
        return rows, info

    def _process(self, call, table, keep, table_mode, exchange):
        n = len(self.devices)
        req, tab, tim = _lib.GroupRequest(), _lib.HostTable(), _lib.GroupTiming()
        shards, lens = (C.c_void_p * n)(), (C.c_uint64 * n)()
        req.ops = OP_REDUCE | (OP_TABLE if table else 0) | (OP_KEEP if keep else 0)
        req.table_mode, req.exchange = table_mode, exchange
        req.table, req.timing = C.pointer(tab), C.pointer(tim)
        req.d_records = C.cast(shards, C.POINTER(C.c_void_p))
        req.shard_records = C.cast(lens, C.POINTER(C.c_uint64))
        res, err = _lib.ReduceResult(), _lib.Error()
        _check(call(C.byref(req), C.byref(res), C.byref(err)), err)
        out = ProcessOutput()
        out.timing = tim.as_dict()
        out.shard_records = [int(x) for x in lens]
        if table:
            out.rows, out.table_info = self._take_table(tab)
        if keep:
            out.records = [DeviceRecords(self.ctx(r), int(shards[r] or 0), int(lens[r])) for r in range(n)]
        return ReduceResult(res.as_dict()), out

    def process_mmap(self, reader: "MmapReader", start: int = 0, end: int | None = None, table=False, keep=False,
                     table_mode: int = 0, exchange: int = 0):
        """process_parallel across the group: (merged ReduceResult, ProcessOutput with .rows/.table_info,
        .records (per-rank DeviceRecords), .timing, .shard_records)."""
        end_v = 2**64 - 1 if end is None else end
        return self._process(lambda req, res, err: lib.ibu_gpu_group_process_mmap(self._h, reader._h, start, end_v, req, res, err),
                             table, keep, table_mode, exchange)

    def process_host(self, h_records, bc_len: int, umi_len: int, table=False, keep=False, table_mode: int = 0,
                     exchange: int = 0):
        recs = _as_records(h_records)
        return self._process(lambda req, res, err: lib.ibu_gpu_group_process_host(self._h, _ptr(recs), len(recs), bc_len,
                                                                                   umi_len, req, res, err),
                             table, keep, table_mode, exchange)

    def barcode_count(self, d_shards, shard_records, mode: int = 0, exchange: int = 0):
        """Exact table of device-resident shards (one per rank): (rows, info, timing dict)."""
        n = len(self.devices)
        ptrs = (C.c_void_p * n)(*[_ptr(p).value for p in d_shards])
        lens = (C.c_uint64 * n)(*shard_records)
        tab, tim, err = _lib.HostTable(), _lib.GroupTiming(), _lib.Error()
        _check(lib.ibu_gpu_group_barcode_count(self._h, ptrs, lens, mode, exchange, C.byref(tab), C.byref(tim), C.byref(err)), err)
        rows, info = self._take_table(tab)
        return rows, info, tim.as_dict()


class GpuStream:
    """Streaming ingest: the reference's `Reader<R>` (src/io/reader.rs:152-306) feeding the GPU.
    Push the bytes of an .ibu stream in pieces of any size; `finish()` returns the reduction."""

    def __init__(self, ctx: GpuContext):
        self._h = C.c_void_p()
        err = _lib.Error()
        _check(lib.ibu_gpu_stream_open(ctx._h, C.byref(self._h), C.byref(err)), err)

    def push(self, data):
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.view(np.uint8).reshape(-1)
        err = _lib.Error()
        _check(lib.ibu_gpu_stream_push(self._h, _ptr(np.ascontiguousarray(buf)), buf.size, C.byref(err)), err)

    def header(self) -> Header:
        h, err = _lib.Header(), _lib.Error()
        _check(lib.ibu_gpu_stream_header(self._h, C.byref(h), C.byref(err)), err)
        return Header._wrap(h)

    def finish(self) -> ReduceResult:
        res, err = _lib.ReduceResult(), _lib.Error()
        _check(lib.ibu_gpu_stream_finish(self._h, C.byref(res), C.byref(err)), err)
        return ReduceResult(res.as_dict())

    def close(self):
        if getattr(self, "_h", None):
            lib.ibu_gpu_stream_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _as_records(a) -> np.ndarray:
    if isinstance(a, np.ndarray) and a.dtype == RECORD_DTYPE and a.flags.c_contiguous:
        return a
    return np.ascontiguousarray(a, RECORD_DTYPE)


def _wrap_cb(fn):
    if fn is None:
        return _lib.CHUNK_CB(0), None

    def tramp(_user, start, cnt, res):
        try:
            rv = fn(int(start), int(cnt), ReduceResult(res.contents.as_dict()))
            return int(rv or 0)
        except Exception:  # a failing processor maps to IbuError::Process (parallel.rs:338-352)
            return 1

    cb = _lib.CHUNK_CB(tramp)
    return cb, tramp


def launch_count() -> int:
    """Kernels launched by this library in this process."""
    return int(lib.ibu_gpu_launch_count())


def device_count() -> int:
    return int(lib.ibu_gpu_device_count())


# ---- pinned host memory -----------------------------------------------------------------------
class PinnedBuffer:
    """cudaHostAlloc'ed bytes exposed as numpy (the staging source/sink of the end-to-end path)."""

    def __init__(self, nbytes: int):
        p, err = C.c_void_p(), _lib.Error()
        _check(lib.ibu_host_alloc(nbytes, C.byref(p), C.byref(err)), err)
        self.ptr, self.nbytes = int(p.value), nbytes
        self._buf = (C.c_uint8 * max(nbytes, 1)).from_address(self.ptr)

    def array(self, dtype, shape=None, offset: int = 0) -> np.ndarray:
        a = np.frombuffer(self._buf, dtype=np.uint8, count=self.nbytes - offset, offset=offset).view(dtype)
        return a if shape is None else a[: int(np.prod(shape))].reshape(shape)

    def free(self):
        if self.ptr:
            self._buf = None
            lib.ibu_host_free(C.c_void_p(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _RecordView(np.ndarray):
    """ndarray that keeps the MmapReader it borrows from alive."""

    _owner = None

    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)


# ---- MmapReader (src/io/mmap.rs) ----------------------------------------------------------------
class MmapReader:
    def __init__(self, path):
        self._h = C.c_void_p()
        err = _lib.Error()
        _check(lib.ibu_mmap_open(os.fsencode(path), C.byref(self._h), C.byref(err)), err)

    @classmethod
    def new(cls, path) -> "MmapReader":
        return cls(path)

    def clone(self) -> "MmapReader":
        r = MmapReader.__new__(MmapReader)
        r._h = C.c_void_p(lib.ibu_mmap_clone(self._h))
        return r

    def close(self):
        if getattr(self, "_h", None):
            lib.ibu_mmap_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def len(self) -> int:
        return int(lib.ibu_mmap_len(self._h))

    __len__ = len

    def header(self) -> Header:
        return Header._wrap(lib.ibu_mmap_header(self._h))

    def slice(self, start: int, end: int) -> np.ndarray:
        """Zero-copy read-only view of records [start, end); errors as mmap.rs:253-266."""
        out, n, err = C.c_void_p(), C.c_size_t(), _lib.Error()
        _check(lib.ibu_mmap_slice(self._h, start, end, C.byref(out), C.byref(n), C.byref(err)), err)
        buf = (C.c_uint8 * (n.value * RECORD_SIZE)).from_address(out.value)
        a = np.frombuffer(buf, dtype=RECORD_DTYPE).view(_RecordView)
        a.flags.writeable = False
        a._owner = self  # the slice borrows from the map (mmap.rs:253): keep the reader alive
        return a

    def pin(self):
        """Page-lock the mapping so process_gpu DMAs straight from the page cache (optional)."""
        err = _lib.Error()
        _check(lib.ibu_mmap_pin(self._h, C.byref(err)), err)

    def unpin(self):
        lib.ibu_mmap_unpin(self._h)

    def pin_range(self, start: int, end: int):
        """Page-lock only the pages of records [start, end) (one rank's shard)."""
        err = _lib.Error()
        _check(lib.ibu_mmap_pin_range(self._h, start, end, C.byref(err)), err)

    def unpin_range(self, start: int, end: int):
        lib.ibu_mmap_unpin_range(self._h, start, end)

    def process_gpu(self, ctx: GpuContext, start: int = 0, end: int | None = None, on_chunk=None) -> ReduceResult:
        """GPU counterpart of process_parallel for the built-in reductions (mmap.rs:286-332):
        records [start, end) staged through pinned double-buffered streams, validated and
        reduced on device; `on_chunk(start, n, result)` is the on_batch_complete analogue."""
        res, err = _lib.ReduceResult(), _lib.Error()
        cb, keep = _wrap_cb(on_chunk)
        end_v = 2**64 - 1 if end is None else end
        _check(lib.ibu_gpu_process_mmap(ctx._h, self._h, start, end_v, C.byref(res), cb, None, C.byref(err)), err)
        return ReduceResult(res.as_dict())


def _process_gpu_ops(self, ctx: GpuContext, start: int = 0, end: int | None = None, table=False, keep=False,
                     unpack=False, flags=False, table_mode: int = 0, on_chunk=None, **outs):
    """process_gpu with an operation mask (ibu_gpu_process_mmap_ops): one pass over records
    [start, end) that validates / reduces, and optionally unpacks to ASCII, builds the per-barcode
    table (parallel.rs:79-98 + distinct UMIs) and keeps the records on the device.
    Returns (ReduceResult, ProcessOutput)."""
    end_v = 2**64 - 1 if end is None else end
    n = (self.len() if end is None else end) - start
    h = self.header()

    def call(req, res, cb, err):
        return lib.ibu_gpu_process_mmap_ops(ctx._h, self._h, start, end_v, req, res, cb, None, err)

    return _run_ops(ctx, call, max(n, 0), h.bc_len, h.umi_len, table, keep, unpack, flags, table_mode, on_chunk, **outs)


MmapReader.process_gpu_ops = _process_gpu_ops


class _LibOwned:
    """A block allocated by the library, exposed to numpy without a copy and released with
    ibu_free when the last array that views it is collected."""

    def __init__(self, ptr: int, nbytes: int):
        self._ptr = ptr
        self.__array_interface__ = {"data": (ptr, False), "shape": (nbytes,), "typestr": "|u1", "version": 3}

    def __del__(self):
        try:
            lib.ibu_free(C.c_void_p(self._ptr))
        except Exception:
            pass


def load_to_vec(path):
    """(Header, records) — src/io/reader.rs:510-535.  The records array views the library's
    allocation (no copy); it is freed with the array."""
    h, recs, n, err = _lib.Header(), C.c_void_p(), C.c_size_t(), _lib.Error()
    _check(lib.ibu_load_to_vec(os.fsencode(path), C.byref(h), C.byref(recs), C.byref(n), C.byref(err)), err)
    if n.value:
        out = np.asarray(_LibOwned(recs.value, n.value * RECORD_SIZE)).view(RECORD_DTYPE)
    else:
        lib.ibu_free(recs)
        out = np.zeros(0, RECORD_DTYPE)
    return Header._wrap(h), out


class DeviceRecords:
    """Device-resident `[Record]` returned by load_to_device; freed with .free() or the context."""

    def __init__(self, ctx: GpuContext, ptr: int, n: int):
        self.ctx, self.ptr, self.n = ctx, ptr, n

    def data_ptr(self) -> int:
        return self.ptr

    def __len__(self):
        return self.n

    def to_host(self) -> np.ndarray:
        out = np.zeros(self.n, RECORD_DTYPE)
        if self.n:
            self.ctx.d2h(out, self.ptr)
        return out

    def free(self):
        if self.ptr:
            self.ctx.free(self.ptr)
            self.ptr = 0


def load_to_device(ctx: GpuContext, path, start: int = 0, end: int | None = None):
    """Device path of load_to_vec: (Header, DeviceRecords) for records [start, end) of the file."""
    h, recs, n, err = _lib.Header(), C.c_void_p(), C.c_uint64(), _lib.Error()
    end_v = 2**64 - 1 if end is None else end
    _check(lib.ibu_gpu_load_to_device(ctx._h, os.fsencode(path), start, end_v, C.byref(h), C.byref(recs),
                                      C.byref(n), C.byref(err)), err)
    return Header._wrap(h), DeviceRecords(ctx, int(recs.value or 0), int(n.value))


# ---- Writer (src/io/writer.rs) --------------------------------------------------------------------
class Writer:
    def __init__(self, path, header: Header | None, append: bool = False):
        self._h = C.c_void_p()
        err = _lib.Error()
        if header is None:  # Writer::new_headless
            rc = lib.ibu_writer_open_headless(os.fsencode(path), int(append), C.byref(self._h), C.byref(err))
        else:
            rc = lib.ibu_writer_open(os.fsencode(path), C.byref(header._h), C.byref(self._h), C.byref(err))
        _check(rc, err)

    @classmethod
    def from_path(cls, path, header: Header) -> "Writer":
        return cls(path, header)

    def write_record(self, barcode: int, umi: int, index: int):
        rec, err = _lib.Record(barcode, umi, index), _lib.Error()
        _check(lib.ibu_writer_write_record(self._h, C.byref(rec), C.byref(err)), err)

    def write_batch(self, recs: np.ndarray):
        recs = _as_records(recs)
        err = _lib.Error()
        _check(lib.ibu_writer_write_batch(self._h, _ptr(recs), len(recs), C.byref(err)), err)

    def write_device(self, ctx: "GpuContext", d_records, n: int):
        """write_batch fed from HBM: pipelined D2H through pinned buffers, then the file."""
        err = _lib.Error()
        _check(lib.ibu_gpu_write_records(ctx._h, self._h, _ptr(d_records), n, C.byref(err)), err)

    def write_iter(self, it):
        for b, u, i in it:
            self.write_record(int(b), int(u), int(i))

    def records_written(self) -> int:
        return int(lib.ibu_writer_records_written(self._h))

    def finish(self):
        err = _lib.Error()
        _check(lib.ibu_writer_finish(self._h, C.byref(err)), err)

    def close(self):
        if getattr(self, "_h", None):
            lib.ibu_writer_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
