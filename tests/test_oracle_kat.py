"""Pins the CPU oracle against every known-answer test the reference holds for the bulk
record path.  Citations: reference crate paths."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle_c as oc
from oracle import oracle_np as on


def recs(triples):
    a = np.zeros(len(triples), on.RECORD_DTYPE)
    for i, (b, u, x) in enumerate(triples):
        a[i] = (b, u, x)
    return a


def pattern(n, fb=lambda i: i, fu=lambda i: 2 * i, fx=lambda i: 3 * i):
    i = np.arange(n, dtype=np.uint64)
    a = np.zeros(n, on.RECORD_DTYPE)
    a["barcode"], a["umi"], a["index"] = fb(i), fu(i), fx(i)
    return a


def write(path, records, bc=16, umi=12, mode=1):
    oc.write_file(path, oc.header_new(bc, umi), records, mode)


# ---- layout ---------------------------------------------------------------------------
def test_sizes():  # header.rs:247-251, record.rs:148-152
    assert C.sizeof(oc.Header) == 32 and C.sizeof(oc.Record) == 24
    assert on.HEADER_DTYPE.itemsize == 32 and on.RECORD_DTYPE.itemsize == 24


def test_magic_bytes():  # header.rs:373-378
    assert on.MAGIC.to_bytes(4, "little") == b"IBU!"


def test_header_bytes_kat():  # header.rs:84-93 + layout 48-61
    want = bytes.fromhex("49425521" "02000000" "10000000" "0c000000" + "00" * 16)
    assert bytes(oc.header_new(16, 12)) == want
    assert on.header_bytes(16, 12) == want
    s = bytearray(want)
    s[16] = 1
    assert bytes(oc.header_new(16, 12, sorted_=True)) == bytes(s)  # set_sorted: flags bit 0
    assert on.header_bytes(16, 12, True) == bytes(s)


@pytest.mark.parametrize("bc,umi", [(1, 1), (16, 12), (32, 32)])
def test_validate_ok(bc, umi):  # header.rs:272-280
    oc.header_validate(oc.header_new(bc, umi))
    assert on.validate_header(on.header_bytes(bc, umi))[0] == "ok"


def test_validate_errors_and_order():  # header.rs:282-348, order 167-187
    h = oc.header_new(16, 12)
    h.magic = 0x12345678
    with pytest.raises(oc.OracleError) as e:
        oc.header_validate(h)
    assert (e.value.variant, e.value.a, e.value.b) == ("InvalidMagicNumber", 0x21554249, 0x12345678)
    h = oc.header_new(16, 12)
    h.version = 1
    with pytest.raises(oc.OracleError) as e:
        oc.header_validate(h)
    assert (e.value.variant, e.value.a, e.value.b) == ("InvalidVersion", 2, 1)
    for bad in (0, 33):
        with pytest.raises(oc.OracleError) as e:
            oc.header_validate(oc.header_new(bad, 12))
        assert (e.value.variant, e.value.a) == ("InvalidBarcodeLength", bad)
        with pytest.raises(oc.OracleError) as e:
            oc.header_validate(oc.header_new(16, bad))
        assert (e.value.variant, e.value.a) == ("InvalidUmiLength", bad)
    # first failure wins: bad magic AND bad version AND bad lengths -> magic
    h = oc.header_new(0, 0)
    h.magic, h.version = 1, 7
    with pytest.raises(oc.OracleError) as e:
        oc.header_validate(h)
    assert e.value.variant == "InvalidMagicNumber"
    h.magic = on.MAGIC
    with pytest.raises(oc.OracleError) as e:
        oc.header_validate(h)
    assert e.value.variant == "InvalidVersion"
    h.version = 2
    with pytest.raises(oc.OracleError) as e:
        oc.header_validate(h)
    assert e.value.variant == "InvalidBarcodeLength"
    raw = bytearray(on.header_bytes(0, 0))
    assert on.validate_header(bytes(raw))[0] == "InvalidBarcodeLength"


# ---- writer / file size -----------------------------------------------------------------
@pytest.mark.parametrize("mode", [0, 1])
def test_file_size_and_bytes(tmp_ibu, mode):  # writer.rs:645,673,693; README.md:84-85
    r = recs([(1, 2, 3), (4, 5, 6)])
    write(tmp_ibu, r, mode=mode)
    assert os.path.getsize(tmp_ibu) == 80
    assert open(tmp_ibu, "rb").read() == on.file_bytes(16, 12, r)


@pytest.mark.parametrize("n", [0, 1, 49151, 49152, 49153, 100_000])
def test_writer_buffer_boundaries(tmp_ibu, n):  # writer.rs:10 (48 Ki records), 321-351
    r = pattern(n)
    for mode in (0, 1):
        write(tmp_ibu, r, mode=mode)
        assert os.path.getsize(tmp_ibu) == 32 + 24 * n
        assert open(tmp_ibu, "rb").read() == on.file_bytes(16, 12, r)


# ---- mmap reader ------------------------------------------------------------------------
def test_mmap_creation(tmp_ibu):  # mmap.rs:375-394
    write(tmp_ibu, recs([(1, 2, 3), (4, 5, 6), (7, 8, 9)]))
    m = oc.MmapReader(tmp_ibu)
    assert len(m) == 3 and m.header().bc_len == 16 and m.header().umi_len == 12


def test_mmap_slice(tmp_ibu):  # mmap.rs:396-423
    write(tmp_ibu, pattern(100))
    m = oc.MmapReader(tmp_ibu)
    full = m.slice(0, 100)
    assert len(full) == 100 and tuple(full[0]) == (0, 0, 0) and tuple(full[99]) == (99, 198, 297)
    part = m.slice(10, 20)
    assert len(part) == 10 and tuple(part[0]) == (10, 20, 30) and tuple(part[9]) == (19, 38, 57)
    one = m.slice(50, 51)
    assert len(one) == 1 and tuple(one[0]) == (50, 100, 150)


def test_mmap_slice_errors(tmp_ibu):  # mmap.rs:425-452
    write(tmp_ibu, recs([(1, 2, 3)]))
    m = oc.MmapReader(tmp_ibu)
    for (s, e), want in [((0, 2), (2, 1)), ((1, 1), (1, 1)), ((1, 0), (0, 1))]:
        with pytest.raises(oc.OracleError) as ex:
            m.slice(s, e)
        assert (ex.value.variant, ex.value.a, ex.value.b) == ("InvalidIndex", *want)


def test_parallel_reduce_kat(tmp_ibu):  # mmap.rs:454-481
    write(tmp_ibu, pattern(10_000))
    red, trace = oc.MmapReader(tmp_ibu).process_parallel_reduce(4)
    assert red["n_records"] == 10_000
    assert red["sum_barcode"] + red["sum_umi"] + red["sum_index"] == 299_970_000
    nt = min(4, oc.num_cpus())
    assert [(t["start"], t["end"]) for t in trace] == on.partition(10_000, nt)


def test_parallel_auto_threads(tmp_ibu):  # mmap.rs:483-500
    write(tmp_ibu, pattern(1000, fu=lambda i: 0 * i, fx=lambda i: 0 * i))
    red, trace = oc.MmapReader(tmp_ibu).process_parallel_reduce(0)
    assert red["n_records"] == 1000 and len(trace) == oc.num_cpus()


def test_empty_file(tmp_ibu):  # mmap.rs:502-519, reader.rs:699-720
    write(tmp_ibu, recs([]))
    m = oc.MmapReader(tmp_ibu)
    assert len(m) == 0
    red, trace = m.process_parallel_reduce(2)
    assert red["n_records"] == 0
    assert all(t["batches"] == 0 for t in trace)  # empty ranges never call on_batch_complete
    h, r = oc.load_to_vec(tmp_ibu)
    assert len(r) == 0 and h.bc_len == 16


def test_large_file(tmp_ibu):  # mmap.rs:546-565
    n = 100_000
    write(tmp_ibu, pattern(n, lambda i: i % 1000, lambda i: i % 500, lambda i: i))
    m = oc.MmapReader(tmp_ibu)
    assert len(m) == n
    mid = m.slice(50_000, 50_010)
    assert len(mid) == 10 and int(mid[0]["index"]) == 50_000


def test_batch_size_constant():  # mmap.rs:567-573
    assert on.BATCH_SIZE == 1 << 20 and on.BATCH_SIZE * 24 < 100 * 1024 * 1024


def test_batching_and_remainder(tmp_ibu):  # mmap.rs:297-320: last thread takes the remainder
    n = 2 * (1 << 20) + 12345
    r = np.zeros(n, on.RECORD_DTYPE)
    r["index"] = np.arange(n)
    write(tmp_ibu, r)
    m = oc.MmapReader(tmp_ibu)
    for nt in (1, 2, 3):
        red, trace = m.process_parallel_reduce(nt)
        eff = min(nt, oc.num_cpus())
        assert red["n_records"] == n and red["sum_index"] == n * (n - 1) // 2
        assert [(t["start"], t["end"]) for t in trace] == on.partition(n, eff)
        for t in trace:
            assert t["records"] == t["end"] - t["start"]
            assert t["batches"] == len(on.batches(t["start"], t["end"]))


def test_len_smaller_than_threads(tmp_ibu):  # rpt == 0: only the last thread works
    write(tmp_ibu, pattern(1))
    red, trace = oc.MmapReader(tmp_ibu).process_parallel_reduce(2)
    if len(trace) == 2:
        assert [t["records"] for t in trace] == [0, 1]
    assert red["n_records"] == 1


def test_error_processor(tmp_ibu):  # parallel.rs:338-352, mmap.rs:326-328
    write(tmp_ibu, pattern(1000, fx=lambda i: i))
    m = oc.MmapReader(tmp_ibu)
    with pytest.raises(oc.OracleError) as e:
        m.process_parallel_fail(2, 5)
    assert e.value.variant == "Process"
    m.process_parallel_fail(2, 10_000)  # never matches: Ok(())


def test_vtable_processor_defaults(tmp_ibu):  # parallel.rs:100-190 (defaults), 438-458
    write(tmp_ibu, pattern(10))
    seen = []
    vt = oc.ProcessorVtable(oc.CLONE_FN(0), oc.RECORD_FN(lambda s, r: seen.append(r.contents.index) or 0),
                            oc.BATCH_FN(0), oc.DROP_FN(0))
    oc.MmapReader(tmp_ibu).process_parallel(vt, None, 1)
    assert sorted(seen) == [3 * i for i in range(10)]


# ---- load_to_vec / stream ---------------------------------------------------------------
def test_load_to_vec(tmp_ibu):  # reader.rs:668-697
    r = recs([(1, 2, 3), (4, 5, 6), (7, 8, 9)])
    write(tmp_ibu, r)
    h, got = oc.load_to_vec(tmp_ibu)
    assert (h.bc_len, h.umi_len) == (16, 12) and np.array_equal(got, r)
    hn, gn = on.read_file(tmp_ibu)
    assert np.array_equal(gn, r) and int(hn["bc_len"]) == 16


def test_bad_size(tmp_ibu):  # reader.rs:722-741, mmap.rs:155-157
    write(tmp_ibu, recs([(1, 2, 3), (4, 5, 6)]))
    with open(tmp_ibu, "r+b") as f:
        f.truncate(os.path.getsize(tmp_ibu) - 5)
    with pytest.raises(oc.OracleError) as e:
        oc.load_to_vec(tmp_ibu)
    assert e.value.variant == "InvalidMapSize"
    with pytest.raises(oc.OracleError) as e:
        oc.MmapReader(tmp_ibu)
    assert e.value.variant == "InvalidMapSize"


def test_truncated_stream():  # reader.rs:618-636
    data = on.file_bytes(16, 12, recs([(1, 2, 3)]))
    assert oc.stream_first(data) == (1, 2, 3)
    with pytest.raises(oc.OracleError) as e:
        oc.stream_first(data[:-5])
    assert e.value.variant == "TruncatedRecord" and e.value.a == 32
    assert oc.stream_first(data[:32]) is None


def test_bad_header_file(tmp_ibu):
    with open(tmp_ibu, "wb") as f:
        f.write(b"\0" * 56)
    with pytest.raises(oc.OracleError) as e:
        oc.MmapReader(tmp_ibu)
    assert e.value.variant == "InvalidMagicNumber"


# ---- record order / occupancy -------------------------------------------------------------
def test_record_order():  # record.rs:163-232: lexicographic (barcode, umi, index)
    r = recs([(2, 0, 0), (1, 5, 9), (1, 5, 3), (1, 2, 7)])
    order = np.lexsort((r["index"], r["umi"], r["barcode"]))
    assert [tuple(x) for x in r[order]] == [(1, 2, 7), (1, 5, 3), (1, 5, 9), (2, 0, 0)]
    assert np.array_equal(np.sort(r, order=["barcode", "umi", "index"]), r[order])


def test_low_bits_occupancy():  # record.rs:312-321: 16 bases = 32 bits
    assert oc.valid_word((1 << 32) - 1, 16) and not oc.valid_word(1 << 32, 16)
    assert oc.valid_word(2**64 - 1, 32)
    assert oc.pack_word(b"T" * 16) == ((1 << 32) - 1, False)


# ---- example pattern closed forms (examples/parallel.rs:65-69, roundtrip.rs:34-38) --------
def test_example_pattern_closed_form():
    n = 1_000_000
    r = oc.generate_records(0, n, 16, 12, 2, 0, 0)
    assert np.array_equal(r, on.generate_records(0, n, 16, 12, 2, 0, 0))
    red = oc.reduce_records(r, 16, 12)
    assert (red["sum_barcode"], red["sum_umi"], red["sum_index"]) == (499_999_500_000,) * 3
    assert red["xor_all"] == 0 and red["n_records"] == n
    table, pairs = oc.barcode_table(r)
    assert len(table) == 1_000_000 and pairs == 1_000_000
    assert np.all(table["n_records"] == 1) and np.all(table["n_distinct_umi"] == 1)
