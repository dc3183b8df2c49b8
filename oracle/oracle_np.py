"""oracle_np — TEST INFRASTRUCTURE ONLY.

Second, independent (numpy) restatement of the reference's bulk record path, used to
cross-check the C++ oracle (ibu_oracle.cpp) and, through it, the CUDA path.  Nothing
under ibu_b200/ may import this.  Citations are into the reference crate
(noamteyssier/ibu); parity status is in ibu_oracle.h (pack/unpack/validation/barcode
table: PARITY UNPINNED — bitnuc is not vendored; conventions from record.rs:19-27).
"""
from __future__ import annotations

import numpy as np

MAGIC = 0x21554249  # src/constructs/header.rs:5
VERSION = 2  # header.rs:6
HEADER_SIZE = 32  # header.rs:7
RECORD_SIZE = 24  # src/constructs/record.rs:3
BATCH_SIZE = 1024 * 1024  # src/io/mmap.rs:284

HEADER_DTYPE = np.dtype(
    [("magic", "<u4"), ("version", "<u4"), ("bc_len", "<u4"), ("umi_len", "<u4"),
     ("flags", "<u8"), ("reserved", "u1", (8,))]
)  # header.rs:44-61
RECORD_DTYPE = np.dtype([("barcode", "<u8"), ("umi", "<u8"), ("index", "<u8")])  # record.rs:58-66
assert HEADER_DTYPE.itemsize == HEADER_SIZE and RECORD_DTYPE.itemsize == RECORD_SIZE

U64 = np.uint64


def header_bytes(bc_len: int, umi_len: int, sorted_: bool = False) -> bytes:
    """Header::new (+ set_sorted) as bytes (header.rs:84-93, 111-113, 203-205)."""
    h = np.zeros((), dtype=HEADER_DTYPE)
    h["magic"], h["version"], h["bc_len"], h["umi_len"] = MAGIC, VERSION, bc_len, umi_len
    h["flags"] = 1 if sorted_ else 0
    return h.tobytes()


def validate_header(raw: bytes) -> tuple[str, tuple]:
    """Header::validate (header.rs:167-187): ('ok', ()) or (variant, payload)."""
    h = np.frombuffer(raw[:HEADER_SIZE], dtype=HEADER_DTYPE)[0]
    if int(h["magic"]) != MAGIC:
        return "InvalidMagicNumber", (MAGIC, int(h["magic"]))
    if int(h["version"]) != VERSION:
        return "InvalidVersion", (VERSION, int(h["version"]))
    if int(h["bc_len"]) == 0 or int(h["bc_len"]) > 32:
        return "InvalidBarcodeLength", (int(h["bc_len"]),)
    if int(h["umi_len"]) == 0 or int(h["umi_len"]) > 32:
        return "InvalidUmiLength", (int(h["umi_len"]),)
    return "ok", ()


def file_bytes(bc_len: int, umi_len: int, records: np.ndarray, sorted_: bool = False) -> bytes:
    """Writer output: header ‖ records (writer.rs:129-143, 315-351)."""
    return header_bytes(bc_len, umi_len, sorted_) + np.ascontiguousarray(records, RECORD_DTYPE).tobytes()


def read_file(path: str):
    """MmapReader::new semantics via np.memmap (mmap.rs:143-161)."""
    raw = np.memmap(path, dtype=np.uint8, mode="r")
    status, payload = validate_header(bytes(raw[:HEADER_SIZE]))
    if status != "ok":
        raise ValueError((status, payload))
    if (raw.size - HEADER_SIZE) % RECORD_SIZE:
        raise ValueError(("InvalidMapSize", ()))
    n = (raw.size - HEADER_SIZE) // RECORD_SIZE
    recs = np.frombuffer(raw, dtype=RECORD_DTYPE, count=n, offset=HEADER_SIZE)
    return np.frombuffer(bytes(raw[:HEADER_SIZE]), dtype=HEADER_DTYPE)[0], recs


def partition(length: int, num_threads: int) -> list[tuple[int, int]]:
    """Thread ranges of process_parallel (mmap.rs:297-307): last takes the remainder."""
    per, rem = divmod(length, num_threads)
    return [(i * per, (i + 1) * per + (rem if i == num_threads - 1 else 0)) for i in range(num_threads)]


def batches(start: int, end: int) -> list[tuple[int, int]]:
    """Batch loop of one thread (mmap.rs:311-320)."""
    out, b = [], start
    while b < end:
        out.append((b, min(b + BATCH_SIZE, end)))
        b += BATCH_SIZE
    return out


def low_mask(length: int) -> int:
    return (1 << (2 * length)) - 1 if length < 32 else (1 << 64) - 1


def reduce_records(records: np.ndarray, bc_len: int, umi_len: int) -> dict:
    """Built-in reductions (mmap.rs:359-363; examples/parallel.rs:22-27; roundtrip.rs:84-87)."""
    b, u, x = records["barcode"], records["umi"], records["index"]
    with np.errstate(over="ignore"):
        sb, su, sx = (int(np.sum(v, dtype=U64)) for v in (b, u, x))
    xor = int(np.bitwise_xor.reduce(b ^ u ^ x)) if len(records) else 0
    bb = (b & U64(~low_mask(bc_len) & (2**64 - 1))) != 0
    bu = (u & U64(~low_mask(umi_len) & (2**64 - 1))) != 0
    return dict(n_records=len(records), sum_barcode=sb, sum_umi=su, sum_index=sx, xor_all=xor,
                n_bad_barcode=int(bb.sum()), n_bad_umi=int(bu.sum()), n_bad_records=int((bb | bu).sum()))


_LUT = np.frombuffer(b"ACGT", dtype=np.uint8)


def unpack_words(words: np.ndarray, length: int) -> np.ndarray:
    """[n] u64 -> [n][length] ASCII; base i at bits [2i, 2i+1] (record.rs:19-27, bitnuc order)."""
    shifts = (2 * np.arange(length, dtype=U64))[None, :]
    codes = (words.astype(U64)[:, None] >> shifts) & U64(3)
    return _LUT[codes.astype(np.intp)]


def pack_rows(rows: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """[n][length] ASCII -> ([n] u64, [n] bad flag); ACGT any case, anything else flagged."""
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    up = rows & 0xDF
    bad = ~((up == 0x41) | (up == 0x43) | (up == 0x47) | (up == 0x54))
    c1 = (rows >> 1) & 3
    code = (c1 ^ (c1 >> 1)).astype(U64)
    shifts = (2 * np.arange(rows.shape[1], dtype=U64))[None, :]
    words = np.bitwise_or.reduce(code << shifts, axis=1) if rows.shape[1] else np.zeros(len(rows), U64)
    return words.astype(U64), bad.any(axis=1)


def barcode_table(records: np.ndarray) -> np.ndarray:
    """Per-barcode (n_records, n_distinct_umi), sorted by barcode (parallel.rs:79-98 + distinct UMIs)."""
    out_dt = np.dtype([("barcode", "<u8"), ("n_records", "<u8"), ("n_distinct_umi", "<u8")])
    if len(records) == 0:
        return np.zeros(0, out_dt)
    order = np.lexsort((records["umi"], records["barcode"]))
    b, u = records["barcode"][order], records["umi"][order]
    new_b = np.ones(len(b), bool)
    new_b[1:] = b[1:] != b[:-1]
    new_p = new_b.copy()
    new_p[1:] |= u[1:] != u[:-1]
    seg = np.cumsum(new_b) - 1
    out = np.zeros(int(seg[-1]) + 1, out_dt)
    out["barcode"] = b[new_b]
    out["n_records"] = np.bincount(seg).astype(U64)
    out["n_distinct_umi"] = np.bincount(seg, weights=new_p).astype(U64)
    return out


# ---- synthetic generators (DESIGN.md §Synthetic data) -------------------------------------
def splitmix64(x):
    x = np.asarray(x, dtype=U64)
    with np.errstate(over="ignore"):
        z = x + U64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> U64(30))) * U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> U64(27))) * U64(0x94D049BB133111EB)
        return z ^ (z >> U64(31))


def generate_records(first: int, n: int, bc_len: int, umi_len: int, mode: int, param: int, seed: int) -> np.ndarray:
    i = np.arange(first, first + n, dtype=U64)
    key = splitmix64(U64(seed) ^ splitmix64(i))
    rb, ru = splitmix64(key ^ U64(1)), splitmix64(key ^ U64(2))
    mb, mu = U64(low_mask(bc_len)), U64(low_mask(umi_len))
    out = np.zeros(n, RECORD_DTYPE)
    out["index"] = i
    if mode in (0, 1):
        out["barcode"], out["umi"] = rb & mb, ru & mu
        if mode == 1:
            rd = splitmix64(key ^ U64(3))
            hit = (rd % U64(1000000)) < U64(param)
            top = (rd >> U64(63)).astype(bool)
            out["barcode"] = np.where(hit & top, rb, out["barcode"])
            out["umi"] = np.where(hit & ~top, ru, out["umi"])
    elif mode == 2:
        with np.errstate(over="ignore"):
            out["barcode"], out["umi"] = i % U64(1000000), (i * U64(31)) % U64(1000000)
    elif mode == 3:
        nb, us = param & 0xFFFFFFFF, param >> 32
        nb = nb or 1000
        out["barcode"] = splitmix64((rb % U64(nb)) ^ U64(seed) ^ U64(0xB)) & mb
        out["umi"] = ((ru % U64(us)) if us else ru) & mu
    elif mode == 5:
        nb, us = (param & 0xFFFFFFFF) or 1000, param >> 32
        e = rb % U64(nb.bit_length())
        one = U64(1)
        r = (((one << e) - one) + (splitmix64(key ^ U64(7)) & ((one << e) - one))) % U64(nb)
        out["barcode"] = splitmix64(r ^ U64(seed) ^ U64(0xB)) & mb
        out["umi"] = ((ru % U64(us)) if us else ru) & mu
    elif mode == 4:
        rpb, dup = (param & 0xFFFFFFFF) or 1000, (param >> 32) or 1
        out["barcode"], out["umi"] = (i // U64(rpb)) & mb, ((i % U64(rpb)) // U64(dup)) & mu
    else:
        raise ValueError(mode)
    return out


def generate_ascii(first_row: int, n_rows: int, length: int, dirty_ppm: int, lower_ppm: int, seed: int) -> np.ndarray:
    r = np.arange(first_row, first_row + n_rows, dtype=U64)
    key = splitmix64(U64(seed) ^ splitmix64(r))
    w, rd, rl = splitmix64(key ^ U64(4)), splitmix64(key ^ U64(5)), splitmix64(key ^ U64(6))
    rows = unpack_words(w, length).copy()
    lower = (rl % U64(1000000)) < U64(lower_ppm)
    rows[lower] |= 0x20
    dirty = np.nonzero((rd % U64(1000000)) < U64(dirty_ppm))[0]
    rows[dirty, ((rd[dirty] >> U64(40)) % U64(length)).astype(np.intp)] = ord("N")
    return rows
