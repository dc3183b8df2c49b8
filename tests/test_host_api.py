"""Host side of the product library through the C ABI (no GPU): the same known answers the
oracle is pinned to, now asked of libibu_b200.so, plus ABI completeness."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import ibu_b200 as ibu
from ibu_b200 import _lib
from oracle import oracle_c as oc
from oracle import oracle_np as on

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pattern(n, fb=lambda i: i, fu=lambda i: 2 * i, fx=lambda i: 3 * i):
    i = np.arange(n, dtype=np.uint64)
    a = ibu.records(n)
    a["barcode"], a["umi"], a["index"] = fb(i), fu(i), fx(i)
    return a


def write(path, recs, bc=16, umi=12):
    with ibu.Writer(path, ibu.Header(bc, umi)) as w:
        w.write_batch(recs)
        w.finish()


def test_abi_exports_every_declared_symbol():
    text = open(os.path.join(ROOT, "include", "ibu_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(ibu_[a-z0-9_]+)\s*\(", text))
    declared -= {"ibu_chunk_cb"}
    assert len(declared) >= 45
    so = C.CDLL(_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(so, s)]
    assert not missing, missing
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_product_path_does_not_touch_the_oracle():
    """oracle/ is test infrastructure: nothing under ibu_b200/ may import, link or execute it, and the
    library carries device code for sm_100a only (no fallback path to fall back to)."""
    import glob
    import shutil
    import subprocess

    pkg = os.path.join(ROOT, "ibu_b200")
    for path in glob.glob(os.path.join(pkg, "**", "*"), recursive=True):
        if not os.path.isfile(path) or not path.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
            continue
        text = open(path, errors="replace").read()
        code = re.sub(r"#.*|//.*|/\*.*?\*/|\"\"\".*?\"\"\"", "", text, flags=re.S if path.endswith(".py") else 0)
        assert not re.search(r"^\s*(from|import)\s+oracle\b", code, flags=re.M), path
        assert "libibu_oracle" not in code and "oracle_c" not in code and "oracle_np" not in code, path
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd, ldd
    if shutil.which("cuobjdump"):
        elf = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
        archs = set(re.findall(r"sm_\d+a?", elf))
        assert archs == {"sm_100a"}, archs


def test_host_stream_copy_any_alignment_and_size():
    """The non-temporal staging copy (head to 16-byte alignment, 64-byte body, tail) is exact."""
    rng = np.random.default_rng(7)
    src = rng.integers(0, 256, 40 << 20, dtype=np.uint8)
    dst = np.zeros(src.size + 256, np.uint8)
    for so, do, n, thr in [(0, 0, 0, 1), (1, 3, 1, 1), (5, 9, 63, 1), (3, 1, 64, 1), (7, 16, 4097, 2), (0, 5, 1 << 20, 3),
                           (11, 13, (33 << 20) + 77, 4), (0, 0, 40 << 20, 0)]:
        dst[:] = 0xEE
        _lib.lib.ibu_host_stream_copy(dst.ctypes.data + do, src.ctypes.data + so, n, thr)
        assert np.array_equal(dst[do:do + n], src[so:so + n]), (so, do, n, thr)
        assert (dst[:do] == 0xEE).all() and (dst[do + n:do + n + 64] == 0xEE).all(), (so, do, n, thr)


def test_cpp_example_builds_against_the_header():
    """examples/roundtrip.cpp is a plain C++ caller of include/ibu_b200.h (no Python): it must
    compile and link against the shared library.  (It needs a B200 to run.)"""
    import subprocess

    subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True, capture_output=True)
    assert os.access(os.path.join(ROOT, "examples", "roundtrip"), os.X_OK)


def test_struct_layouts():
    assert C.sizeof(_lib.Header) == 32 and C.sizeof(_lib.Record) == 24
    assert C.sizeof(_lib.ReduceResult) == 64 and C.sizeof(_lib.Error) == 256
    assert ibu.RECORD_DTYPE.itemsize == 24 and ibu.BATCH_SIZE == 1 << 20


def test_header_kats():  # header.rs:84-93, 111-113, 167-187, 373-378
    h = ibu.Header(16, 12)
    assert h.as_bytes() == on.header_bytes(16, 12) == bytes(oc.header_new(16, 12))
    assert not h.sorted()
    h.set_sorted()
    assert h.sorted() and h.as_bytes()[16] == 1
    assert ibu.Header.from_bytes(h.as_bytes()) == h
    for bc, umi in [(1, 1), (16, 12), (32, 32)]:
        ibu.Header(bc, umi).validate()
    with pytest.raises(ibu.InvalidBarcodeLength) as e:
        ibu.Header(0, 12).validate()
    assert e.value.a == 0 and "1-32" in str(e.value)
    with pytest.raises(ibu.InvalidBarcodeLength):
        ibu.Header(33, 12).validate()
    with pytest.raises(ibu.InvalidUmiLength):
        ibu.Header(16, 0).validate()
    with pytest.raises(ibu.InvalidUmiLength):
        ibu.Header(16, 33).validate()
    raw = bytearray(ibu.Header(0, 0).as_bytes())
    raw[0:4] = (0x12345678).to_bytes(4, "little")
    raw[4:8] = (1).to_bytes(4, "little")
    with pytest.raises(ibu.InvalidMagicNumber) as e:
        ibu.Header.from_bytes(bytes(raw)).validate()
    assert (e.value.expected, e.value.actual) == (0x21554249, 0x12345678)
    assert "0x21554249" in str(e.value) and "0x12345678" in str(e.value)
    raw[0:4] = ibu.MAGIC.to_bytes(4, "little")
    with pytest.raises(ibu.InvalidVersion) as e:
        ibu.Header.from_bytes(bytes(raw)).validate()
    assert (e.value.expected, e.value.actual) == (2, 1)


@pytest.mark.parametrize("n", [0, 2, 49152, 49153, 100_000])
def test_writer_bytes_match_oracle(tmp_path, n):  # writer.rs:129-143, 260-351
    r = pattern(n)
    a, b, c = (str(tmp_path / x) for x in "abc")
    write(a, r)
    oc.write_file(b, oc.header_new(16, 12), r, 1)
    with ibu.Writer(c, ibu.Header(16, 12)) as w:  # record-at-a-time path
        if n <= 49153:
            w.write_iter((int(x["barcode"]), int(x["umi"]), int(x["index"])) for x in r)
        else:
            w.write_batch(r[: n // 2])
            w.write_batch(r[n // 2:])
        assert w.records_written() == n
    want = on.file_bytes(16, 12, r)
    assert open(a, "rb").read() == open(b, "rb").read() == open(c, "rb").read() == want
    assert os.path.getsize(a) == 32 + 24 * n


def test_headless_ingest(tmp_path):  # writer.rs:169-179, 477-482: shards appended after a header
    main, shard = str(tmp_path / "m.ibu"), str(tmp_path / "s.bin")
    r = pattern(1000)
    with ibu.Writer(main, ibu.Header(16, 12)) as w:
        w.write_batch(r[:400])
    with ibu.Writer(shard, None) as w:
        w.write_batch(r[400:])
    assert os.path.getsize(shard) == 600 * 24
    with ibu.Writer(main, None, append=True) as w:
        w.write_batch(np.fromfile(shard, ibu.RECORD_DTYPE))
    assert open(main, "rb").read() == on.file_bytes(16, 12, r)


def test_mmap_reader_kats(tmp_ibu):  # mmap.rs:375-452, 521-565
    write(tmp_ibu, pattern(100))
    m = ibu.MmapReader.new(tmp_ibu)
    assert m.len() == 100 and m.header().bc_len == 16 and m.header().umi_len == 12
    full = m.slice(0, 100)
    assert len(full) == 100 and tuple(full[0]) == (0, 0, 0) and tuple(full[99]) == (99, 198, 297)
    part = m.slice(10, 20)
    assert len(part) == 10 and tuple(part[0]) == (10, 20, 30) and tuple(part[9]) == (19, 38, 57)
    assert tuple(m.slice(50, 51)[0]) == (50, 100, 150)
    c = m.clone()
    assert c.len() == m.len() and c.header() == m.header()
    assert np.array_equal(c.slice(0, 2), m.slice(0, 2))
    assert c.slice(0, 1).ctypes.data == m.slice(0, 1).ctypes.data  # same Arc<Mmap>
    m.close()
    assert tuple(c.slice(99, 100)[0]) == (99, 198, 297)  # the clone keeps the map alive


def test_mmap_slice_errors(tmp_ibu):  # mmap.rs:425-452
    write(tmp_ibu, pattern(1))
    m = ibu.MmapReader(tmp_ibu)
    for (s, e), want in [((0, 2), (2, 1)), ((1, 1), (1, 1)), ((1, 0), (0, 1))]:
        with pytest.raises(ibu.InvalidIndex) as ex:
            m.slice(s, e)
        assert (ex.value.idx, ex.value.max) == want


def test_empty_and_bad_files(tmp_ibu, tmp_path):  # mmap.rs:502-519, reader.rs:699-741
    write(tmp_ibu, pattern(0))
    assert ibu.MmapReader(tmp_ibu).len() == 0
    h, r = ibu.load_to_vec(tmp_ibu)
    assert len(r) == 0 and (h.bc_len, h.umi_len) == (16, 12)
    write(tmp_ibu, pattern(2))
    with open(tmp_ibu, "r+b") as f:
        f.truncate(os.path.getsize(tmp_ibu) - 5)
    with pytest.raises(ibu.InvalidMapSize):
        ibu.MmapReader(tmp_ibu)
    with pytest.raises(ibu.InvalidMapSize):
        ibu.load_to_vec(tmp_ibu)
    with pytest.raises(ibu.Io):
        ibu.MmapReader(str(tmp_path / "missing.ibu"))
    with pytest.raises(ibu.Io):
        ibu.load_to_vec(str(tmp_path / "missing.ibu"))
    bad = str(tmp_path / "bad.ibu")
    open(bad, "wb").write(b"\0" * 56)
    with pytest.raises(ibu.InvalidMagicNumber):
        ibu.MmapReader(bad)
    open(bad, "wb").write(b"IBU!")
    with pytest.raises(ibu.Io):  # shorter than a header: a panic upstream (mmap.rs:149)
        ibu.MmapReader(bad)


def test_load_to_vec_matches_oracle(tmp_ibu):  # reader.rs:668-697
    r = oc.generate_records(0, 100_000, 16, 12, 1, 50_000, 1)
    write(tmp_ibu, r)
    h, got = ibu.load_to_vec(tmp_ibu)
    ho, want = oc.load_to_vec(tmp_ibu)
    assert np.array_equal(got, want) and np.array_equal(got, r)
    assert h.as_bytes() == bytes(ho)
    m = ibu.MmapReader(tmp_ibu)
    assert np.array_equal(m.slice(50_000, 50_010), r[50_000:50_010])


@pytest.mark.parametrize("length,world", [(10_000, 4), (1000, 8), (1, 2), (0, 3), (2**40 + 7, 8)])
def test_shard_range_is_process_parallel_partition(length, world):  # mmap.rs:297-307
    got = [ibu.shard_range(length, r, world) for r in range(world)]
    assert got == on.partition(length, world)
    assert got[0][0] == 0 and got[-1][1] == length


def test_gpu_entry_points_fail_loudly_without_a_device():
    if ibu.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(ibu.CudaError):
        ibu.GpuContext(0)


# ---- the reference's own writer / error tests, asked of the product library -----------------------
def test_writer_reference_tests(tmp_path):  # writer.rs:629-760
    p = str(tmp_path / "w.ibu")
    w = ibu.Writer(p, ibu.Header(16, 12))  # test_writer_creation: header written immediately
    assert w.records_written() == 0 and os.path.getsize(p) == 32
    w.write_record(0x1234, 0x5678, 42)  # test_single_record_write
    assert w.records_written() == 1 and os.path.getsize(p) == 32  # still buffered
    w.finish()
    assert os.path.getsize(p) == 32 + 24
    w.write_batch(pattern(3))  # test_batch_write
    assert w.records_written() == 4
    w.close()
    assert os.path.getsize(p) == 32 + 4 * 24
    q = str(tmp_path / "h.bin")
    w = ibu.Writer(q, None)  # test_writer_headless
    assert w.records_written() == 0 and os.path.getsize(q) == 0
    w.close()
    w = ibu.Writer(p, ibu.Header(16, 12))  # test_buffer_flushing: 48 Ki records fill the buffer exactly
    full = 48 * 1024
    w.write_iter((i, 0, 0) for i in range(full))
    assert os.path.getsize(p) == 32  # not flushed yet
    w.write_record(999, 0, 0)  # one more record triggers the flush
    assert os.path.getsize(p) == 32 + full * 24
    w.write_batch(pattern(100_000))  # test_large_batch_direct_write: bypasses the buffer, after flushing it
    assert w.records_written() == full + 1 + 100_000
    assert os.path.getsize(p) == 32 + (full + 1 + 100_000) * 24
    w.close()
    w = ibu.Writer(p, ibu.Header(20, 10))  # test_writer_roundtrip
    r = ibu.records(2)
    r[0], r[1] = (0x12345, 0x67890, 100), (0xABCDE, 0xF0123, 200)
    w.write_batch(r)
    w.close()  # Drop flushes (writer.rs:519-523)
    h, got = ibu.load_to_vec(p)
    assert (h.bc_len, h.umi_len) == (20, 10) and np.array_equal(got, r)
    bad = ibu.Header(0, 99)  # Writer::new does not validate the header (writer.rs:129-133)
    with ibu.Writer(p, bad):
        pass
    with pytest.raises(ibu.InvalidBarcodeLength):
        ibu.MmapReader(p)


def test_error_display_messages(tmp_path):  # error.rs:179-260: Display strings carry the payload
    raw = bytearray(ibu.Header(16, 12).as_bytes())
    raw[0:4] = (0x12345678).to_bytes(4, "little")
    with pytest.raises(ibu.IbuError) as e:
        ibu.Header.from_bytes(bytes(raw)).validate()
    assert "0x21554249" in str(e.value) and "0x12345678" in str(e.value)
    raw = bytearray(ibu.Header(16, 12).as_bytes())
    raw[4:8] = (1).to_bytes(4, "little")
    with pytest.raises(ibu.IbuError) as e:
        ibu.Header.from_bytes(bytes(raw)).validate()
    assert "expected (2)" in str(e.value) and "found (1)" in str(e.value)
    with pytest.raises(ibu.IbuError) as e:
        ibu.Header(33, 12).validate()
    assert "33" in str(e.value) and "1-32" in str(e.value)
    with pytest.raises(ibu.IbuError) as e:
        ibu.Header(16, 0).validate()
    assert "0" in str(e.value) and "1-32" in str(e.value)
    p = str(tmp_path / "x.ibu")
    with ibu.Writer(p, ibu.Header(16, 12)) as w:
        w.write_batch(pattern(50))
    with open(p, "r+b") as f:
        f.truncate(32 + 24 * 50 - 1)
    with pytest.raises(ibu.IbuError) as e:
        ibu.MmapReader(p)
    assert "not a multiple" in str(e.value)
    with open(p, "r+b") as f:
        f.truncate(32 + 24 * 50 - 24)
    with pytest.raises(ibu.IbuError) as e:
        ibu.MmapReader(p).slice(0, 100)
    assert "100" in str(e.value) and "49" in str(e.value)
    for code, text in [(1, "I/O error"), (8, "not a multiple"), (10, "Processing error")]:
        assert text in _lib.lib.ibu_strerror(code).decode()
