"""ctypes binding of libibu_b200.so — the C ABI declared in include/ibu_b200.h.

There is no fallback: if the shared library (built by `make -C ibu_b200/csrc` or
`__graft_entry__.build()`) is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libibu_b200.so")


class Header(C.Structure):
    """`#[repr(C)] struct Header` (src/constructs/header.rs:44-61)."""

    _fields_ = [("magic", C.c_uint32), ("version", C.c_uint32), ("bc_len", C.c_uint32),
                ("umi_len", C.c_uint32), ("flags", C.c_uint64), ("reserved", C.c_uint8 * 8)]


class Record(C.Structure):
    """`#[repr(C)] struct Record` (src/constructs/record.rs:58-66)."""

    _fields_ = [("barcode", C.c_uint64), ("umi", C.c_uint64), ("index", C.c_uint64)]


class Error(C.Structure):
    _fields_ = [("code", C.c_int32), ("sys", C.c_int32), ("a", C.c_uint64), ("b", C.c_uint64),
                ("msg", C.c_char * 232)]


class ReduceResult(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in (
        "n_records", "sum_barcode", "sum_umi", "sum_index", "xor_all",
        "n_bad_barcode", "n_bad_umi", "n_bad_records")]

    def as_dict(self) -> dict:
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class GpuConfig(C.Structure):
    _fields_ = [("chunk_records", C.c_uint32), ("n_slots", C.c_uint32), ("copy_threads", C.c_uint32),
                ("reserved", C.c_uint32)]


class BarcodeTable(C.Structure):
    _fields_ = [("d_rows", C.c_void_p), ("n_rows", C.c_uint64), ("n_records", C.c_uint64),
                ("n_distinct_pairs", C.c_uint64), ("input_was_sorted", C.c_uint32),
                ("reserved", C.c_uint32)]


class ProcessRequest(C.Structure):
    _fields_ = [("ops", C.c_uint32), ("table_mode", C.c_int32), ("table", C.POINTER(BarcodeTable)),
                ("d_records", C.POINTER(C.c_void_p)), ("h_bc_ascii", C.c_void_p), ("h_umi_ascii", C.c_void_p),
                ("h_flags", C.c_void_p)]


class HostTable(C.Structure):
    _fields_ = [("h_rows", C.c_void_p), ("n_rows", C.c_uint64), ("n_records", C.c_uint64),
                ("n_distinct_pairs", C.c_uint64)]


class GroupTiming(C.Structure):
    _fields_ = [("ingest_ms", C.c_double), ("local_ms", C.c_double), ("exchange_ms", C.c_double),
                ("owner_ms", C.c_double), ("gather_ms", C.c_double), ("table_ms", C.c_double),
                ("total_ms", C.c_double), ("pairs_local", C.c_uint64), ("bytes_sent", C.c_uint64),
                ("exchange", C.c_uint32), ("reserved", C.c_uint32)]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class GroupRequest(C.Structure):
    _fields_ = [("ops", C.c_uint32), ("table_mode", C.c_int32), ("exchange", C.c_uint32), ("reserved", C.c_uint32),
                ("table", C.POINTER(HostTable)), ("d_records", C.POINTER(C.c_void_p)),
                ("shard_records", C.POINTER(C.c_uint64)), ("timing", C.POINTER(GroupTiming))]


CHUNK_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(ReduceResult))

_P = C.POINTER
_vp, _u64, _u32, _sz, _int = C.c_void_p, C.c_uint64, C.c_uint32, C.c_size_t, C.c_int
_err = _P(Error)

# name -> (restype, argtypes); every symbol include/ibu_b200.h declares
SIGNATURES = {
    "ibu_strerror": (C.c_char_p, [_int]),
    "ibu_version": (C.c_char_p, []),
    "ibu_header_init": (None, [_P(Header), _u32, _u32]),
    "ibu_header_set_sorted": (None, [_P(Header)]),
    "ibu_header_sorted": (_int, [_P(Header)]),
    "ibu_header_validate": (_int, [_P(Header), _err]),
    "ibu_mmap_open": (_int, [C.c_char_p, _P(_vp), _err]),
    "ibu_mmap_clone": (_vp, [_vp]),
    "ibu_mmap_close": (None, [_vp]),
    "ibu_mmap_len": (_sz, [_vp]),
    "ibu_mmap_header": (Header, [_vp]),
    "ibu_mmap_slice": (_int, [_vp, _sz, _sz, _P(_vp), _P(_sz), _err]),
    "ibu_load_to_vec": (_int, [C.c_char_p, _P(Header), _P(_vp), _P(_sz), _err]),
    "ibu_free": (None, [_vp]),
    "ibu_shard_range": (None, [_u64, _u32, _u32, _P(_u64), _P(_u64)]),
    "ibu_writer_open": (_int, [C.c_char_p, _P(Header), _P(_vp), _err]),
    "ibu_writer_open_headless": (_int, [C.c_char_p, _int, _P(_vp), _err]),
    "ibu_writer_write_record": (_int, [_vp, _P(Record), _err]),
    "ibu_writer_write_batch": (_int, [_vp, _vp, _sz, _err]),
    "ibu_writer_records_written": (_u64, [_vp]),
    "ibu_writer_finish": (_int, [_vp, _err]),
    "ibu_writer_close": (None, [_vp]),
    "ibu_gpu_device_count": (_int, []),
    "ibu_gpu_ctx_create": (_int, [_int, _P(GpuConfig), _P(_vp), _err]),
    "ibu_gpu_ctx_destroy": (None, [_vp]),
    "ibu_gpu_ctx_device": (_int, [_vp]),
    "ibu_gpu_ctx_sm_count": (_int, [_vp]),
    "ibu_gpu_launch_count": (_u64, []),
    "ibu_gpu_synchronize": (_int, [_vp, _vp, _err]),
    "ibu_gpu_malloc": (_int, [_vp, _sz, _P(_vp), _err]),
    "ibu_gpu_free": (None, [_vp, _vp]),
    "ibu_gpu_memcpy_h2d": (_int, [_vp, _vp, _vp, _sz, _err]),
    "ibu_gpu_memcpy_d2h": (_int, [_vp, _vp, _vp, _sz, _err]),
    "ibu_gpu_memset": (_int, [_vp, _vp, _int, _sz, _err]),
    "ibu_host_alloc": (_int, [_sz, _P(_vp), _err]),
    "ibu_host_free": (None, [_vp]),
    "ibu_host_register": (_int, [_vp, _sz, _int, _err]),
    "ibu_host_unregister": (None, [_vp]),
    "ibu_host_stream_copy": (None, [_vp, _vp, _sz, _u32]),
    "ibu_gpu_validate_reduce_async": (_int, [_vp, _vp, _u64, _u32, _u32, _vp, _vp, _err]),
    "ibu_gpu_unpack_async": (_int, [_vp, _vp, _u64, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _err]),
    "ibu_gpu_pack_async": (_int, [_vp, _vp, _vp, _vp, _u64, _u64, _u32, _u32, _vp, _vp, _vp, _vp, _err]),
    "ibu_gpu_barcode_count": (_int, [_vp, _vp, _u64, _int, _P(BarcodeTable), _vp, _err]),
    "ibu_gpu_table_free": (None, [_vp, _P(BarcodeTable)]),
    "ibu_gpu_table_to_host": (_int, [_vp, _P(BarcodeTable), _vp, _err]),
    "ibu_gpu_pair_table": (_int, [_vp, _vp, _u64, _int, _P(_vp), _P(_u64), _vp, _err]),
    "ibu_gpu_partition_by_owner": (_int, [_vp, _vp, _u64, _u32, _vp, _P(_u64), _vp, _err]),
    "ibu_gpu_memcpy": (_int, [_vp, _vp, _vp, _sz, _err]),
    "ibu_gpu_sort_records": (_int, [_vp, _vp, _u64, _vp, _vp, _err]),
    "ibu_gpu_generate_records_async": (_int, [_vp, _vp, _u64, _u64, _u32, _u32, _int, _u64, _u64, _vp, _err]),
    "ibu_gpu_generate_ascii_async": (_int, [_vp, _vp, _u64, _u64, _u32, _u64, _u64, _u64, _vp, _err]),
    "ibu_mmap_pin": (_int, [_vp, _err]),
    "ibu_mmap_unpin": (None, [_vp]),
    "ibu_gpu_process_mmap": (_int, [_vp, _vp, _u64, _u64, _P(ReduceResult), CHUNK_CB, _vp, _err]),
    "ibu_gpu_process_host": (_int, [_vp, _vp, _u64, _u32, _u32, _P(ReduceResult), CHUNK_CB, _vp, _err]),
    "ibu_gpu_process_mmap_ops": (_int, [_vp, _vp, _u64, _u64, _P(ProcessRequest), _P(ReduceResult), CHUNK_CB, _vp, _err]),
    "ibu_gpu_process_host_ops": (_int, [_vp, _vp, _u64, _u32, _u32, _P(ProcessRequest), _P(ReduceResult), CHUNK_CB, _vp,
                                  _err]),
    "ibu_mmap_pin_range": (_int, [_vp, _u64, _u64, _err]),
    "ibu_mmap_unpin_range": (None, [_vp, _u64, _u64]),
    "ibu_gpu_group_create": (_int, [_P(_int), _u32, _P(GpuConfig), _P(_vp), _err]),
    "ibu_gpu_group_destroy": (None, [_vp]),
    "ibu_gpu_group_size": (_u32, [_vp]),
    "ibu_gpu_group_ctx": (_vp, [_vp, _u32]),
    "ibu_gpu_group_process_mmap": (_int, [_vp, _vp, _u64, _u64, _P(GroupRequest), _P(ReduceResult), _err]),
    "ibu_gpu_group_process_host": (_int, [_vp, _vp, _u64, _u32, _u32, _P(GroupRequest), _P(ReduceResult), _err]),
    "ibu_gpu_group_barcode_count": (_int, [_vp, _P(_vp), _P(_u64), _int, _u32, _P(HostTable), _P(GroupTiming), _err]),
    "ibu_gpu_stream_open": (_int, [_vp, _P(_vp), _err]),
    "ibu_gpu_stream_push": (_int, [_vp, _vp, _sz, _err]),
    "ibu_gpu_stream_header": (_int, [_vp, _P(Header), _err]),
    "ibu_gpu_stream_finish": (_int, [_vp, _P(ReduceResult), _err]),
    "ibu_gpu_stream_close": (None, [_vp]),
    "ibu_gpu_load_to_device": (_int, [_vp, C.c_char_p, _u64, _u64, _P(Header), _P(_vp), _P(_u64), _err]),
    "ibu_gpu_write_records": (_int, [_vp, _vp, _vp, _u64, _err]),
    "ibu_gpu_unpack_host": (_int, [_vp, _vp, _u64, _u32, _u32, _vp, _vp, _vp, _P(ReduceResult), _err]),
    "ibu_gpu_pack_host": (_int, [_vp, _vp, _vp, _vp, _u64, _u64, _u32, _u32, _vp, _vp, _P(ReduceResult), _err]),
}


def _prefer_bundled_nccl() -> None:
    """IBU_EXCHANGE_NCCL loads libnccl.so.2 at run time.  In a Python process that may import torch
    LATER, the library must pick the NCCL torch ships (nvidia/nccl/lib): the loader keeps one object
    per soname, so a system libnccl loaded first would later be handed to torch, which then fails to
    import on the newer symbols it needs.  Only sets IBU_B200_NCCL_LIB when the caller has not."""
    if os.environ.get("IBU_B200_NCCL_LIB"):
        return
    try:
        import importlib.util

        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec else ()):
            path = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(path):
                os.environ["IBU_B200_NCCL_LIB"] = path
                return
    except (ImportError, ValueError, AttributeError):
        pass


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C ibu_b200/csrc` (or "
            "`python -c 'import __graft_entry__ as g; g.build()'`). ibu_b200 has no CPU fallback.")
    _prefer_bundled_nccl()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype, fn.argtypes = res, args
    return lib


lib = _load()
