"""oracle_c — TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/libibu_oracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this.  See ibu_oracle.h for the parity status of each piece.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .oracle_np import RECORD_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libibu_oracle.so")


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("ibu_oracle.cpp", "ibu_oracle.h")]
    stale = not os.path.exists(_SO) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


class Header(C.Structure):
    _fields_ = [("magic", C.c_uint32), ("version", C.c_uint32), ("bc_len", C.c_uint32),
                ("umi_len", C.c_uint32), ("flags", C.c_uint64), ("reserved", C.c_uint8 * 8)]


class Record(C.Structure):
    _fields_ = [("barcode", C.c_uint64), ("umi", C.c_uint64), ("index", C.c_uint64)]


class Error(C.Structure):
    _fields_ = [("code", C.c_int32), ("sys", C.c_int32), ("a", C.c_uint64), ("b", C.c_uint64)]


class Reduce(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in (
        "n_records", "sum_barcode", "sum_umi", "sum_index", "xor_all",
        "n_bad_barcode", "n_bad_umi", "n_bad_records")]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class ThreadTrace(C.Structure):
    _fields_ = [("start", C.c_uint64), ("end", C.c_uint64), ("records", C.c_uint64),
                ("batches", C.c_uint64)]


ROW_DTYPE = np.dtype([("barcode", "<u8"), ("n_records", "<u8"), ("n_distinct_umi", "<u8")])

CLONE_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p)
RECORD_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(Record))
BATCH_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)
DROP_FN = C.CFUNCTYPE(None, C.c_void_p)


class ProcessorVtable(C.Structure):
    _fields_ = [("clone", CLONE_FN), ("process_record", RECORD_FN),
                ("on_batch_complete", BATCH_FN), ("drop", DROP_FN)]


ERR_NAMES = {0: "ok", 1: "Io", 2: "Niffler", 3: "InvalidMagicNumber", 4: "TruncatedRecord",
             5: "InvalidVersion", 6: "InvalidBarcodeLength", 7: "InvalidUmiLength",
             8: "InvalidMapSize", 9: "InvalidIndex", 10: "Process"}


class OracleError(Exception):
    def __init__(self, err: Error):
        self.variant = ERR_NAMES.get(err.code, str(err.code))
        self.code, self.sys, self.a, self.b = err.code, err.sys, int(err.a), int(err.b)
        super().__init__(f"{self.variant}(a={self.a}, b={self.b}, sys={self.sys})")


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.orc_mmap_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(Error)]
        L.orc_mmap_close.argtypes = [C.c_void_p]
        L.orc_mmap_len.argtypes = [C.c_void_p]
        L.orc_mmap_len.restype = C.c_size_t
        L.orc_mmap_header.argtypes = [C.c_void_p]
        L.orc_mmap_header.restype = Header
        L.orc_mmap_slice.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_size_t), C.POINTER(Error)]
        L.orc_load_to_vec.argtypes = [C.c_char_p, C.POINTER(Header), C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_size_t), C.POINTER(Error)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_stream_first.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(Record), C.POINTER(Error)]
        L.orc_write_file.argtypes = [C.c_char_p, C.POINTER(Header), C.c_void_p, C.c_size_t, C.c_int,
                                     C.POINTER(Error)]
        L.orc_header_new.argtypes = [C.POINTER(Header), C.c_uint32, C.c_uint32]
        L.orc_header_validate.argtypes = [C.POINTER(Header), C.POINTER(Error)]
        L.orc_process_parallel_reduce.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(Reduce),
                                                  C.POINTER(ThreadTrace), C.c_size_t,
                                                  C.POINTER(C.c_size_t), C.POINTER(Error)]
        L.orc_process_parallel_fail.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.POINTER(Error)]
        L.orc_process_parallel.argtypes = [C.c_void_p, C.POINTER(ProcessorVtable), C.c_void_p,
                                           C.c_size_t, C.POINTER(Error)]
        L.orc_process_parallel_barcodes.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p),
                                                    C.POINTER(C.c_size_t), C.POINTER(Error)]
        L.orc_reduce_records.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_size_t,
                                         C.POINTER(Reduce)]
        L.orc_barcode_table.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
        L.orc_valid_word.argtypes = [C.c_uint64, C.c_uint32]
        L.orc_unpack_word.argtypes = [C.c_uint64, C.c_uint32, C.c_char_p]
        L.orc_pack_word.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_uint64)]
        L.orc_unpack_records.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(Reduce)]
        L.orc_pack_records.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_size_t,
                                       C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t,
                                       C.POINTER(Reduce)]
        L.orc_splitmix64.argtypes = [C.c_uint64]
        L.orc_splitmix64.restype = C.c_uint64
        L.orc_generate_records.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32,
                                           C.c_int, C.c_uint64, C.c_uint64, C.c_size_t]
        L.orc_generate_ascii.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64,
                                         C.c_uint64, C.c_uint64, C.c_size_t]
    return _lib


def _check(rc: int, err: Error):
    if rc != 0:
        err.code = rc
        raise OracleError(err)


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def header_new(bc_len: int, umi_len: int, sorted_: bool = False) -> Header:
    h = Header()
    lib().orc_header_new(C.byref(h), bc_len, umi_len)
    if sorted_:
        h.flags |= 1
    return h


def header_validate(h: Header):
    err = Error()
    _check(lib().orc_header_validate(C.byref(h), C.byref(err)), err)


class MmapReader:
    """MmapReader restatement (src/io/mmap.rs:99-270)."""

    def __init__(self, path: str):
        self._h = C.c_void_p()
        err = Error()
        _check(lib().orc_mmap_open(os.fsencode(path), C.byref(self._h), C.byref(err)), err)

    def close(self):
        if self._h:
            lib().orc_mmap_close(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __len__(self):
        return int(lib().orc_mmap_len(self._h))

    def header(self) -> Header:
        return lib().orc_mmap_header(self._h)

    def slice(self, start: int, end: int) -> np.ndarray:
        out, n, err = C.c_void_p(), C.c_size_t(), Error()
        _check(lib().orc_mmap_slice(self._h, start, end, C.byref(out), C.byref(n), C.byref(err)), err)
        buf = (C.c_uint8 * (n.value * 24)).from_address(out.value)
        return np.frombuffer(buf, dtype=RECORD_DTYPE).copy()

    def process_parallel_reduce(self, num_threads: int, max_trace: int = 1024):
        red, err, nt = Reduce(), Error(), C.c_size_t()
        trace = (ThreadTrace * max_trace)()
        _check(lib().orc_process_parallel_reduce(self._h, num_threads, C.byref(red), trace, max_trace,
                                                 C.byref(nt), C.byref(err)), err)
        tr = [dict(start=t.start, end=t.end, records=t.records, batches=t.batches)
              for t in trace[: min(nt.value, max_trace)]]
        return red.as_dict(), tr

    def process_parallel_fail(self, num_threads: int, fail_index: int):
        err = Error()
        _check(lib().orc_process_parallel_fail(self._h, num_threads, fail_index, C.byref(err)), err)

    def process_parallel(self, vt: ProcessorVtable, proc, num_threads: int):
        err = Error()
        _check(lib().orc_process_parallel(self._h, C.byref(vt), proc, num_threads, C.byref(err)), err)

    def process_parallel_barcodes(self, num_threads: int) -> np.ndarray:
        rows, n, err = C.c_void_p(), C.c_size_t(), Error()
        _check(lib().orc_process_parallel_barcodes(self._h, num_threads, C.byref(rows), C.byref(n),
                                                   C.byref(err)), err)
        return _take_rows(rows, n.value)


def _take_rows(rows: C.c_void_p, n: int) -> np.ndarray:
    if n == 0:
        lib().orc_free(rows)
        return np.zeros(0, ROW_DTYPE)
    buf = (C.c_uint8 * (n * 24)).from_address(rows.value)
    out = np.frombuffer(buf, dtype=ROW_DTYPE).copy()
    lib().orc_free(rows)
    return out


def load_to_vec(path: str):
    h, recs, n, err = Header(), C.c_void_p(), C.c_size_t(), Error()
    _check(lib().orc_load_to_vec(os.fsencode(path), C.byref(h), C.byref(recs), C.byref(n), C.byref(err)), err)
    if n.value:
        buf = (C.c_uint8 * (n.value * 24)).from_address(recs.value)
        out = np.frombuffer(buf, dtype=RECORD_DTYPE).copy()
    else:
        out = np.zeros(0, RECORD_DTYPE)
    lib().orc_free(recs)
    return h, out


def stream_first(data: bytes):
    rec, err = Record(), Error()
    rc = lib().orc_stream_first(data, len(data), C.byref(rec), C.byref(err))
    if rc == -1:
        return None
    _check(rc, err)
    return (rec.barcode, rec.umi, rec.index)


def write_file(path: str, header: Header, records: np.ndarray, mode: int = 1):
    records = np.ascontiguousarray(records, RECORD_DTYPE)
    err = Error()
    _check(lib().orc_write_file(os.fsencode(path), C.byref(header), _ptr(records), len(records), mode,
                                C.byref(err)), err)


def num_cpus() -> int:
    return int(lib().orc_num_cpus())


def reduce_records(records: np.ndarray, bc_len: int, umi_len: int, num_threads: int = 0) -> dict:
    records = np.ascontiguousarray(records, RECORD_DTYPE)
    red = Reduce()
    lib().orc_reduce_records(_ptr(records), len(records), bc_len, umi_len, num_threads, C.byref(red))
    return red.as_dict()


def barcode_table(records: np.ndarray):
    records = np.ascontiguousarray(records, RECORD_DTYPE)
    rows, n, pairs = C.c_void_p(), C.c_size_t(), C.c_uint64()
    lib().orc_barcode_table(_ptr(records), len(records), C.byref(rows), C.byref(n), C.byref(pairs))
    return _take_rows(rows, n.value), int(pairs.value)


def unpack_word(w: int, length: int) -> bytes:
    buf = C.create_string_buffer(length)
    lib().orc_unpack_word(w, length, buf)
    return buf.raw


def pack_word(s: bytes):
    w = C.c_uint64()
    bad = lib().orc_pack_word(s, len(s), C.byref(w))
    return int(w.value), bool(bad)


def valid_word(w: int, length: int) -> bool:
    return bool(lib().orc_valid_word(w, length))


def unpack_records(records: np.ndarray, bc_len: int, umi_len: int, num_threads: int = 0,
                   bc_out: np.ndarray | None = None, umi_out: np.ndarray | None = None,
                   flags: np.ndarray | None = None):
    records = np.ascontiguousarray(records, RECORD_DTYPE)
    n = len(records)
    bc = bc_out if bc_out is not None else np.empty((n, bc_len), np.uint8)
    umi = umi_out if umi_out is not None else np.empty((n, umi_len), np.uint8)
    fl = flags if flags is not None else np.empty(n, np.uint8)
    red = Reduce()
    lib().orc_unpack_records(_ptr(records), n, bc_len, umi_len, _ptr(bc), _ptr(umi), _ptr(fl),
                             num_threads, C.byref(red))
    return bc, umi, fl, red.as_dict()


def pack_records(bc: np.ndarray, umi: np.ndarray, index: np.ndarray | None = None, index_base: int = 0,
                 num_threads: int = 0, out: np.ndarray | None = None):
    bc = np.ascontiguousarray(bc, np.uint8)
    umi = np.ascontiguousarray(umi, np.uint8)
    n = bc.shape[0]
    recs = out if out is not None else np.empty(n, RECORD_DTYPE)
    flags = np.empty(n, np.uint8)
    if index is not None:
        index = np.ascontiguousarray(index, np.uint64)
    red = Reduce()
    lib().orc_pack_records(_ptr(bc), _ptr(umi), _ptr(index), index_base, n, bc.shape[1], umi.shape[1],
                           _ptr(recs), _ptr(flags), num_threads, C.byref(red))
    return recs, flags, red.as_dict()


def splitmix64(x: int) -> int:
    return int(lib().orc_splitmix64(x))


def generate_records(first: int, n: int, bc_len: int, umi_len: int, mode: int, param: int, seed: int,
                     num_threads: int = 0, out: np.ndarray | None = None) -> np.ndarray:
    recs = out if out is not None else np.empty(n, RECORD_DTYPE)
    lib().orc_generate_records(_ptr(recs), first, n, bc_len, umi_len, mode, param, seed, num_threads)
    return recs


def generate_ascii(first_row: int, n_rows: int, length: int, dirty_ppm: int, lower_ppm: int, seed: int,
                   num_threads: int = 0) -> np.ndarray:
    rows = np.empty((n_rows, length), np.uint8)
    lib().orc_generate_ascii(_ptr(rows), first_row, n_rows, length, dirty_ppm, lower_ppm, seed, num_threads)
    return rows
