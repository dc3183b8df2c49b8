"""GPU end-to-end paths through the C ABI with HOST buffers: the device counterparts of
MmapReader::process_parallel (mmap.rs:286-332) and load_to_vec (reader.rs:510-535), and the
host->host unpack/pack pipelines, against the oracle's process_parallel on the same files."""
import numpy as np
import pytest

import ibu_b200 as ibu
from oracle import oracle_c as oc
from oracle import oracle_np as on

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = ibu.GpuContext(0, chunk_records=1 << 18, n_slots=3, copy_threads=4)
    yield c
    c.close()


def write(path, recs, bc=16, umi=12):
    with ibu.Writer(path, ibu.Header(bc, umi)) as w:
        w.write_batch(recs)


@pytest.mark.parametrize("n", [0, 1, 10_000, (1 << 18), (1 << 18) + 1, 1_000_000])
def test_process_gpu_matches_process_parallel(ctx, tmp_ibu, n):
    """configs[0] shape: Writer -> MmapReader -> process (count+sum, field sums, xor, validation)."""
    recs = oc.generate_records(0, n, 16, 12, 1, 20_000, 42)
    write(tmp_ibu, recs)
    want, _ = oc.MmapReader(tmp_ibu).process_parallel_reduce(0)
    reader = ibu.MmapReader(tmp_ibu)
    chunks = []
    got = reader.process_gpu(ctx, on_chunk=lambda s, c, r: chunks.append((s, c, r)))
    assert got == want
    # on_batch_complete analogue: chunk order, full coverage, never called for an empty file
    csz = 1 << 18
    assert [(s, c) for s, c, _ in chunks] == [(s, min(csz, n - s)) for s in range(0, n, csz)]
    merged = ibu.ReduceResult({k: 0 for k in want})
    for _, _, r in chunks:
        merged = merged.merge(r)
    if n:
        assert merged == want


def test_process_gpu_reference_kat(ctx, tmp_ibu):  # mmap.rs:454-481
    i = np.arange(10_000, dtype=np.uint64)
    recs = ibu.records(10_000)
    recs["barcode"], recs["umi"], recs["index"] = i, 2 * i, 3 * i
    write(tmp_ibu, recs)
    got = ibu.MmapReader(tmp_ibu).process_gpu(ctx)
    assert got["n_records"] == 10_000 and got.count_sum == 299_970_000


def test_process_gpu_from_a_pinned_mapping(ctx, tmp_ibu):
    """ibu_mmap_pin: the mapping itself is page-locked (read-only) and DMA'd without staging; the
    results, slices and clones of the reader are unchanged, and unpin returns to the staged path."""
    n = 600_001
    recs = oc.generate_records(0, n, 16, 12, 1, 30_000, 11)
    write(tmp_ibu, recs)
    want, _ = oc.MmapReader(tmp_ibu).process_parallel_reduce(0)
    reader = ibu.MmapReader(tmp_ibu)
    staged = reader.process_gpu(ctx)
    reader.pin()
    try:
        assert reader.process_gpu(ctx) == want == staged
        assert reader.process_gpu(ctx, 12_345, 500_000) == oc.reduce_records(recs[12_345:500_000], 16, 12)
        assert np.array_equal(reader.slice(7, 1000), recs[7:1000])
        assert reader.clone().process_gpu(ctx) == want
    finally:
        reader.unpin()
    assert reader.process_gpu(ctx) == want


def test_process_gpu_shards_merge_to_whole(ctx, tmp_ibu):
    """Range sharding (mmap.rs:297-307 with n = #GPUs): shard results merge to the file's."""
    n = 777_777
    recs = oc.generate_records(0, n, 16, 12, 1, 50_000, 3)
    write(tmp_ibu, recs)
    reader = ibu.MmapReader(tmp_ibu)
    whole = reader.process_gpu(ctx)
    for world in (2, 3, 8):
        merged = ibu.ReduceResult({k: 0 for k in whole})
        for rank in range(world):
            s, e = ibu.shard_range(n, rank, world)
            merged = merged.merge(reader.process_gpu(ctx, s, e))
        assert merged == whole == oc.reduce_records(recs, 16, 12)
    with pytest.raises(ibu.InvalidIndex):
        reader.process_gpu(ctx, 0, n + 1)


def test_process_gpu_callback_error_aborts(ctx, tmp_ibu):  # parallel.rs:338-352 -> IbuError::Process
    write(tmp_ibu, oc.generate_records(0, 1_000_000, 16, 12, 0, 0, 1))
    calls = []

    def boom(start, cnt, res):
        calls.append(start)
        return 1 if start >= (1 << 18) else 0

    with pytest.raises(ibu.Process):
        ibu.MmapReader(tmp_ibu).process_gpu(ctx, on_chunk=boom)
    assert calls == [0, 1 << 18]


def test_chunk_callback_cannot_reenter_its_context(ctx, tmp_ibu):
    """on_chunk runs while the context's chunk slots are in use: a host-buffer call on the same
    context from inside it is refused (IBU_ERR_ARG), not deadlocked; device-pointer calls are fine."""
    recs = oc.generate_records(0, 600_000, 16, 12, 0, 0, 2)
    write(tmp_ibu, recs)
    seen = []

    def cb(start, cnt, res):
        try:
            ctx.process_host(recs[:1000], 16, 12)
            seen.append("ran")
        except ibu.ArgError:
            seen.append("refused")
        return 0

    red = ibu.MmapReader(tmp_ibu).process_gpu(ctx, on_chunk=cb)
    assert red == oc.reduce_records(recs, 16, 12) and seen and set(seen) == {"refused"}
    assert ctx.process_host(recs[:1000], 16, 12)["n_records"] == 1000  # and the context is usable afterwards


def test_process_host_pinned_and_pageable(ctx):
    n = 600_001
    recs = oc.generate_records(9, n, 20, 10, 1, 10_000, 8)
    want = oc.reduce_records(recs, 20, 10)
    assert ctx.process_host(recs, 20, 10) == want
    pin = ibu.PinnedBuffer(recs.nbytes)
    p = pin.array(ibu.RECORD_DTYPE)
    p[:] = recs
    assert ctx.process_host(p, 20, 10) == want
    del p
    pin.free()


@pytest.mark.parametrize("n", [0, 3, 300_000, 1_000_003])
def test_load_to_device_matches_load_to_vec(ctx, tmp_ibu, n):
    recs = oc.generate_records(0, n, 16, 12, 0, 0, 4)
    write(tmp_ibu, recs)
    h, d = ibu.load_to_device(ctx, tmp_ibu)
    ho, want = oc.load_to_vec(tmp_ibu)
    assert h.as_bytes() == bytes(ho) and len(d) == n
    assert np.array_equal(d.to_host(), want)
    d.free()
    if n > 10:
        h, d = ibu.load_to_device(ctx, tmp_ibu, 5, n - 2)  # a shard of the file
        assert np.array_equal(d.to_host(), want[5:n - 2])
        d.free()


def test_load_to_device_errors(ctx, tmp_ibu, tmp_path):
    write(tmp_ibu, oc.generate_records(0, 10, 16, 12, 0, 0, 4))
    with open(tmp_ibu, "r+b") as f:
        f.truncate(32 + 24 * 10 - 5)
    with pytest.raises(ibu.InvalidMapSize):
        ibu.load_to_device(ctx, tmp_ibu)
    with pytest.raises(ibu.Io):
        ibu.load_to_device(ctx, str(tmp_path / "nope.ibu"))


@pytest.mark.parametrize("bc,umi", [(16, 12), (32, 32), (9, 5)])
@pytest.mark.parametrize("pinned", [False, True])
def test_unpack_host_and_pack_host(ctx, bc, umi, pinned):
    n = 700_003
    recs = oc.generate_records(0, n, bc, umi, 1, 30_000, 6)
    ob, ou, of, ores = oc.unpack_records(recs, bc, umi, 0)
    bufs = []
    if pinned:
        def alloc(dtype, shape):
            b = ibu.PinnedBuffer(int(np.prod(shape)) * np.dtype(dtype).itemsize)
            bufs.append(b)
            return b.array(dtype, shape)
        src = alloc(ibu.RECORD_DTYPE, (n,))
        src[:] = recs
        gb, gu, gf = alloc(np.uint8, (n, bc)), alloc(np.uint8, (n, umi)), alloc(np.uint8, (n,))
    else:
        src, gb, gu, gf = recs, np.zeros((n, bc), np.uint8), np.zeros((n, umi), np.uint8), np.zeros(n, np.uint8)
    _, _, res = ctx.unpack_host(src, bc, umi, gb, gu, gf)
    assert np.array_equal(gb, ob) and np.array_equal(gu, ou) and np.array_equal(gf, of)
    for k in ("n_records", "n_bad_barcode", "n_bad_umi", "n_bad_records"):
        assert res[k] == ores[k]
    assert res == oc.reduce_records(recs, bc, umi)  # sums / xor ride along, merged across chunks
    # pack the decoded rows back: masked originals, explicit index array
    idx = np.ascontiguousarray(recs["index"])
    back, pres = ctx.pack_host(gb, gu, index=idx)
    want, _, wres = oc.pack_records(ob, ou, idx)
    assert np.array_equal(back, want) and pres["n_bad_records"] == wres["n_bad_records"] == 0
    back2, _ = ctx.pack_host(np.ascontiguousarray(gb), np.ascontiguousarray(gu), index_base=0)
    assert np.array_equal(back2["index"], np.arange(n, dtype=np.uint64))
    assert np.array_equal(back2["barcode"], recs["barcode"] & np.uint64(on.low_mask(bc)))
    del src, gb, gu, gf
    for b in bufs:
        b.free()


@pytest.mark.parametrize("n", [0, 5, (1 << 18) * 3 + 17])
def test_sort_on_device_and_write_sorted_file(ctx, tmp_path, n):
    """'Next' row 1+2 of SURVEY §8f: load_to_device -> sort by Record's Ord -> Writer device path
    with a truthful set_sorted header; the file equals the oracle's sorted file byte for byte and
    its table streams through the sorted fast path."""
    src, dst = str(tmp_path / "in.ibu"), str(tmp_path / "sorted.ibu")
    recs = oc.generate_records(0, n, 16, 12, 3, (16 << 32) | 700, 41)
    write(src, recs)
    h, d = ibu.load_to_device(ctx, src)
    out = ctx.malloc(max(24 * n, 1))
    ctx.sort_records(d, n, out)
    hs = ibu.Header(h.bc_len, h.umi_len)
    hs.set_sorted()
    with ibu.Writer(dst, hs) as w:
        w.write_device(ctx, out, n)
        assert w.records_written() == n
    want = recs[np.lexsort((recs["index"], recs["umi"], recs["barcode"]))]
    assert open(dst, "rb").read() == on.file_bytes(16, 12, want, sorted_=True)
    if n:
        rows, info = ctx.barcode_count(out, n, 1)
        assert info["input_was_sorted"] and np.array_equal(rows, on.barcode_table(recs))
    ctx.free(out)
    d.free()
    assert ibu.MmapReader(dst).header().sorted()


# ---- streaming ingest: the Reader<R> semantics (reader.rs:152-306) on the GPU path --------------------
@pytest.mark.parametrize("n", [0, 1, 1000, (1 << 18), (1 << 18) * 2 + 12345])
@pytest.mark.parametrize("piece", [1 << 20, 4097, 24 * 1000 + 7])
def test_stream_ingest_matches_oracle(ctx, n, piece):
    recs = oc.generate_records(0, n, 16, 12, 1, 30_000, 51)
    data = on.file_bytes(16, 12, recs)
    with ibu.GpuStream(ctx) as st:
        for off in range(0, len(data), piece):  # pieces cut anywhere: mid-header, mid-record
            st.push(data[off:off + piece])
        assert st.header().as_bytes() == data[:32]
        got = st.finish()
    assert got == oc.reduce_records(recs, 16, 12)


def test_stream_ingest_errors(ctx):
    good = on.file_bytes(16, 12, oc.generate_records(0, 10, 16, 12, 0, 0, 52))
    with ibu.GpuStream(ctx) as st:  # reader.rs:618-636: truncated record
        st.push(good[:-5])
        with pytest.raises(ibu.TruncatedRecord) as e:
            st.finish()
        assert e.value.pos == 32 + 24 * 9
    with pytest.raises(oc.OracleError) as oe:  # the oracle reports the same position for 1 record - 5 bytes
        oc.stream_first(good[:32 + 24 - 5])
    assert oe.value.variant == "TruncatedRecord" and oe.value.a == 32
    with ibu.GpuStream(ctx) as st:
        st.push(good[:32 + 24 - 5])
        with pytest.raises(ibu.TruncatedRecord) as e:
            st.finish()
        assert e.value.pos == 32
    bad = bytearray(good)
    bad[0] = 0
    with ibu.GpuStream(ctx) as st:  # Reader::new validates the header first
        with pytest.raises(ibu.InvalidMagicNumber):
            st.push(bytes(bad))
    with ibu.GpuStream(ctx) as st:  # shorter than a header: read_exact fails
        st.push(good[:20])
        with pytest.raises(ibu.Io):
            st.finish()
    with ibu.GpuStream(ctx) as st:  # one stream per context at a time
        with pytest.raises(ibu.ArgError):
            ibu.GpuStream(ctx)
        st.push(good)
        assert st.finish()["n_records"] == 10
    assert ctx.process_host(np.frombuffer(good[32:], ibu.RECORD_DTYPE), 16, 12)["n_records"] == 10  # lock released


def test_stream_outlives_its_context_and_blocks_host_calls():
    """A stream owns its context's chunk slots between open and close without holding a lock across
    calls: host-buffer calls on the context are refused meanwhile, another thread may close it, and a
    stream that outlives its context fails cleanly instead of touching freed memory."""
    import threading

    good = on.file_bytes(16, 12, oc.generate_records(0, 1000, 16, 12, 0, 0, 53))
    recs = np.frombuffer(good[32:], ibu.RECORD_DTYPE)
    c = ibu.GpuContext(0, chunk_records=1 << 16)
    st = ibu.GpuStream(c)
    st.push(good[:5000])
    with pytest.raises(ibu.ArgError):
        c.process_host(recs, 16, 12)
    c.close()  # the context goes first
    with pytest.raises(ibu.ArgError):
        st.push(good[5000:])
    with pytest.raises(ibu.ArgError):
        st.finish()
    st.close()
    c = ibu.GpuContext(0, chunk_records=1 << 16)
    st = ibu.GpuStream(c)
    st.push(good)
    assert st.finish()["n_records"] == 1000
    t = threading.Thread(target=st.close)  # closed by a thread that did not open it
    t.start()
    t.join()
    assert c.process_host(recs, 16, 12)["n_records"] == 1000
    c.close()


# ---- one pass with an operation mask: ingest -> validate/reduce (+ unpack) (+ per-barcode table) ----
def zipf_file_records(n, seed, dirty=True):
    recs = oc.generate_records(0, n, 16, 12, 5, (64 << 32) | 20_000, seed)  # Zipf-ish barcodes, heavy duplication
    if dirty:  # a few words wider than the header allows (examples/random.rs:46)
        rng = np.random.default_rng(seed)
        hit = rng.random(n) < 0.003
        recs["umi"][hit] |= np.uint64(1 << 45)
    return recs


@pytest.mark.parametrize("n", [0, 1, 70_000, (1 << 18) * 3 + 17, 1_500_003])
@pytest.mark.parametrize("pre_sorted", [False, True])
def test_process_ops_table_matches_oracle(ctx, tmp_ibu, n, pre_sorted):
    """mmap.rs:312-320 driving the HashMap<barcode, count> processor of parallel.rs:79-98 (+ distinct
    UMIs): the table of the whole file out of the same pass that validates and reduces it."""
    recs = zipf_file_records(n, 61)
    if pre_sorted:
        recs = recs[np.lexsort((recs["index"], recs["umi"], recs["barcode"]))]
    write(tmp_ibu, recs)
    want_red, _ = oc.MmapReader(tmp_ibu).process_parallel_reduce(0)
    want = on.barcode_table(recs)
    reader = ibu.MmapReader(tmp_ibu)
    chunks = []
    red, out = reader.process_gpu_ops(ctx, table=True, on_chunk=lambda s, c, r: chunks.append((s, c)))
    assert red == want_red
    assert np.array_equal(out.rows, want)
    assert out.table_info["n_records"] == n and out.table_info["n_distinct_pairs"] == int(want["n_distinct_umi"].sum())
    if n > 1000:
        assert out.table_info["input_was_sorted"] == pre_sorted
    csz = 1 << 18
    assert chunks == [(s, min(csz, n - s)) for s in range(0, n, csz)]
    # a sub-range of the file, forced through the partition path (small inputs default to the legacy one)
    if n > 1000:
        a, b = n // 7, n - n // 5
        red, out = reader.process_gpu_ops(ctx, a, b, table=True, table_mode=2 | ibu.COUNT_PATH_PARTITION)
        assert red == oc.reduce_records(recs[a:b], 16, 12)
        assert np.array_equal(out.rows, on.barcode_table(recs[a:b]))


def test_process_ops_unpack_table_keep_in_one_pass(ctx, tmp_ibu):
    """decode + validate + count: ASCII, counters, the table and the resident records from ONE pass."""
    n = 900_001
    recs = zipf_file_records(n, 62)
    write(tmp_ibu, recs)
    reader = ibu.MmapReader(tmp_ibu)
    red, out = reader.process_gpu_ops(ctx, table=True, keep=True, unpack=True, flags=True)
    ob, ou, ofl, ores = oc.unpack_records(recs, 16, 12, 0)
    assert np.array_equal(out.bc_ascii, ob) and np.array_equal(out.umi_ascii, ou) and np.array_equal(out.flags, ofl)
    assert red == oc.reduce_records(recs, 16, 12)
    assert np.array_equal(out.rows, on.barcode_table(recs))
    assert np.array_equal(out.records.to_host(), recs)
    out.records.free()
    # host-array form, pinned source
    pin = ibu.PinnedBuffer(recs.nbytes)
    h = pin.array(ibu.RECORD_DTYPE, (n,))
    h[:] = recs
    red2, out2 = ctx.process_host_ops(h, 16, 12, table=True)
    assert red2 == red and np.array_equal(out2.rows, out.rows)
    del h
    pin.free()


def test_process_ops_table_exact_layout_when_duplicates_dominate(ctx, tmp_ibu):
    """The reference's example pattern (examples/parallel.rs:65-69): 10^6 distinct pairs however long the
    file is — the bucket loads are too uneven for a uniform layout and the exact one takes over."""
    n = 3_000_000
    recs = oc.generate_records(0, n, 16, 12, 2, 0, 0)
    write(tmp_ibu, recs)
    red, out = ibu.MmapReader(tmp_ibu).process_gpu_ops(ctx, table=True)
    assert red["n_records"] == n and len(out.rows) == 1_000_000
    assert np.array_equal(out.rows["barcode"], np.arange(1_000_000, dtype=np.uint64))
    assert np.all(out.rows["n_records"] == 3) and np.all(out.rows["n_distinct_umi"] == 1)


def test_process_ops_argument_errors(ctx, tmp_ibu):
    write(tmp_ibu, oc.generate_records(0, 10, 16, 12, 0, 0, 1))
    reader = ibu.MmapReader(tmp_ibu)
    from ibu_b200 import _lib
    import ctypes as C
    req, res, err = _lib.ProcessRequest(), _lib.ReduceResult(), _lib.Error()
    req.ops = ibu.OP_TABLE  # no table pointer
    rc = _lib.lib.ibu_gpu_process_mmap_ops(ctx._h, reader._h, 0, 2**64 - 1, C.byref(req), C.byref(res), _lib.CHUNK_CB(0), None,
                                           C.byref(err))
    assert rc == 13 and err.code == 13
    with pytest.raises(ibu.InvalidIndex):
        reader.process_gpu_ops(ctx, 5, 11, table=True)
