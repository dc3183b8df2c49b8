"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs — bit-exact (all arithmetic is u64/u8 integer).  Run with `-m gpu` on a B200."""
import numpy as np
import pytest

import ibu_b200 as ibu
from oracle import oracle_c as oc
from oracle import oracle_np as on

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 2, 63, 64, 65, 127, 128, 129, 1000, 128 * 1000 + 77, (1 << 20) + 1, 3_000_005]
U64 = np.uint64


@pytest.fixture(scope="module")
def ctx():
    c = ibu.GpuContext(0, chunk_records=1 << 20, n_slots=3)
    yield c
    c.close()


class Dev:
    """Device buffer from the library's own allocator (256-byte aligned)."""

    def __init__(self, ctx, nbytes, host=None):
        self.ctx, self.nbytes = ctx, nbytes
        self.ptr = ctx.malloc(max(nbytes, 1))
        if host is not None and nbytes:
            ctx.h2d(self.ptr, host)

    def data_ptr(self):
        return self.ptr

    def get(self, dtype, shape):
        out = np.zeros(shape, dtype)
        if out.nbytes:
            self.ctx.d2h(out, self.ptr)
        return out

    def free(self):
        self.ctx.free(self.ptr)


def gpu_reduce(ctx, recs, bc, umi):
    d = Dev(ctx, recs.nbytes, recs)
    r = Dev(ctx, 64)
    ctx.memset(r.ptr, 0xAB, 64)  # the kernel must overwrite, not accumulate
    ctx.validate_reduce_async(d, len(recs), bc, umi, r)
    ctx.synchronize()
    out = ctx.read_result(r.ptr)
    d.free(), r.free()
    return out


# ---- K1 -----------------------------------------------------------------------------------
@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("bc,umi,mode,param", [(16, 12, 0, 0), (16, 12, 1, 10_000), (16, 12, 1, 1_000_000),
                                               (32, 32, 1, 1_000_000), (1, 1, 1, 500_000), (7, 31, 1, 300_000),
                                               (16, 12, 2, 0)])
def test_validate_reduce(ctx, n, bc, umi, mode, param):
    recs = oc.generate_records(5, n, bc, umi, mode, param, 1234)
    assert gpu_reduce(ctx, recs, bc, umi) == oc.reduce_records(recs, bc, umi)


def test_validate_reduce_both_words_bad(ctx):
    """Adjacent invalid barcode+umi at every lane/tile position (the shuffle pairing path)."""
    rng = np.random.default_rng(0)
    n = 128 * 40 + 17
    recs = ibu.records(n)
    recs["barcode"] = rng.integers(0, 2**32, n, dtype=U64)
    recs["umi"] = rng.integers(0, 2**24, n, dtype=U64)
    recs["index"] = rng.integers(0, 2**64, n, dtype=U64)
    kind = rng.integers(0, 4, n)
    recs["barcode"][kind & 1 == 1] |= U64(1 << 40)
    recs["umi"][kind & 2 == 2] |= U64(1 << 50)
    got = gpu_reduce(ctx, recs, 16, 12)
    assert got == oc.reduce_records(recs, 16, 12) == on.reduce_records(recs, 16, 12)
    assert got["n_bad_records"] == int((kind != 0).sum())
    recs["barcode"] |= U64(1 << 63)
    recs["umi"] |= U64(1 << 63)
    got = gpu_reduce(ctx, recs, 16, 12)
    assert got["n_bad_records"] == n and got == oc.reduce_records(recs, 16, 12)


def test_reference_kat_on_gpu(ctx):
    """mmap.rs:454-481: 10 000 records (i, 2i, 3i): count 10 000, sum 299 970 000."""
    i = np.arange(10_000, dtype=U64)
    recs = ibu.records(10_000)
    recs["barcode"], recs["umi"], recs["index"] = i, 2 * i, 3 * i
    got = gpu_reduce(ctx, recs, 16, 12)
    assert got["n_records"] == 10_000 and got.count_sum == 299_970_000


# ---- K2 -----------------------------------------------------------------------------------
def gpu_unpack(ctx, recs, bc, umi, flags=True):
    n = len(recs)
    d = Dev(ctx, recs.nbytes, recs)
    db, du, df, dr = Dev(ctx, n * bc), Dev(ctx, n * umi), Dev(ctx, n), Dev(ctx, 64)
    ctx.memset(db.ptr, 0, max(n * bc, 1)), ctx.memset(du.ptr, 0, max(n * umi, 1))
    ctx.unpack_async(d, n, bc, umi, db, du, df if flags else None, dr)
    ctx.synchronize()
    out = db.get(np.uint8, (n, bc)), du.get(np.uint8, (n, umi)), df.get(np.uint8, n), ctx.read_result(dr.ptr)
    for x in (d, db, du, df, dr):
        x.free()
    return out


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("bc,umi", [(16, 12), (16, 16), (32, 32), (12, 8), (5, 3), (32, 12), (1, 1), (20, 10), (31, 32),
                                    (16, 32), (28, 16), (16, 10), (30, 2), (18, 26)])
def test_unpack(ctx, n, bc, umi):
    recs = oc.generate_records(0, n, bc, umi, 1, 100_000, 99)
    gb, gu, gf, gr = gpu_unpack(ctx, recs, bc, umi)
    ob, ou, of, orr = oc.unpack_records(recs, bc, umi, 2)
    assert np.array_equal(gb, ob) and np.array_equal(gu, ou) and np.array_equal(gf, of)
    for k in ("n_records", "n_bad_barcode", "n_bad_umi", "n_bad_records"):
        assert gr[k] == orr[k], k
    assert gr == oc.reduce_records(recs, bc, umi)  # the unpack pass also carries K1's reductions


def test_unpack_without_flags_and_result(ctx):
    recs = oc.generate_records(0, 70_001, 16, 12, 0, 0, 5)
    gb, gu, _, _ = gpu_unpack(ctx, recs, 16, 12, flags=False)
    assert np.array_equal(gb, on.unpack_words(recs["barcode"], 16))
    assert np.array_equal(gu, on.unpack_words(recs["umi"], 12))


def test_unpack_codec_kats(ctx):  # record.rs:19-27 + bitnuc order
    recs = ibu.records(3)
    recs["barcode"] = [0xE4, 2**64 - 1, 0]
    recs["umi"] = [0xFFFFFF, 0x1B, 2**64 - 1]
    gb, gu, gf, _ = gpu_unpack(ctx, recs, 4, 12)
    assert gb.tobytes() == b"ACGT" + b"TTTT" + b"AAAA"
    assert gu.tobytes() == b"T" * 12 + b"TGCA" + b"A" * 8 + b"T" * 12
    assert gf.tolist() == [0, 1, 2]


@pytest.mark.parametrize("n", [1, 127, 128, 129, 1023, 1024, 1025, 128 * 8 * 5 + 3])
@pytest.mark.parametrize("bc,umi", [(16, 12), (32, 32), (15, 9), (20, 10), (1, 32)])
def test_kernels_stay_inside_their_buffers(ctx, n, bc, umi):
    """One tile per warp, the grid rounded up to whole CTAs, one warp for the ragged tail: guard
    zones around every output of K2 and K3 (and K2's input tail) must come back untouched."""
    G = 4096  # guard bytes on both sides (keeps 16/32-byte alignment)
    recs = oc.generate_records(0, n, bc, umi, 1, 200_000, 17)

    def guarded(nbytes, payload=None):
        host = np.full(G + nbytes + G, 0xA5, np.uint8)
        if payload is not None:
            host[G:G + nbytes] = np.frombuffer(payload.tobytes(), np.uint8)
        return Dev(ctx, host.nbytes, host), host

    def check(dev, nbytes):
        back = dev.get(np.uint8, G + nbytes + G)
        assert (back[:G] == 0xA5).all() and (back[G + nbytes:] == 0xA5).all()
        return back[G:G + nbytes]

    d_recs, _ = guarded(recs.nbytes, recs)
    d_bc, _ = guarded(n * bc)
    d_umi, _ = guarded(n * umi)
    d_fl, _ = guarded(n)
    res = Dev(ctx, 64)
    ctx.unpack_async(d_recs.ptr + G, n, bc, umi, d_bc.ptr + G, d_umi.ptr + G, d_fl.ptr + G, res)
    ctx.synchronize()
    ob, ou, of, _ = oc.unpack_records(recs, bc, umi, 1)
    gb, gu, gf = check(d_bc, n * bc), check(d_umi, n * umi), check(d_fl, n)
    assert np.array_equal(gb.reshape(n, bc), ob) and np.array_equal(gu.reshape(n, umi), ou) and np.array_equal(gf, of)
    # pack the rows back into a guarded record buffer
    d_out, _ = guarded(n * 24)
    ctx.pack_async(d_bc.ptr + G, d_umi.ptr + G, n, bc, umi, d_out.ptr + G, d_flags=d_fl.ptr + G, d_result=res)
    ctx.synchronize()
    back = check(d_out, n * 24).view(ibu.RECORD_DTYPE)
    check(d_fl, n)
    mask = lambda length: U64(2**64 - 1) if length == 32 else U64((1 << (2 * length)) - 1)  # noqa: E731
    assert np.array_equal(back["barcode"], recs["barcode"] & mask(bc))
    assert np.array_equal(back["umi"], recs["umi"] & mask(umi))
    assert np.array_equal(back["index"], np.arange(n, dtype=U64))
    for d in (d_recs, d_bc, d_umi, d_fl, d_out, res):
        d.free()


def test_result_blocks_under_concurrent_streams(ctx):
    """K1/K2/K3 accumulate into a 16-entry ring of spread result blocks folded by a second kernel;
    48 launches in flight on 8 streams (3 x the ring) must each get exactly their own result."""
    import torch

    streams = [torch.cuda.Stream() for _ in range(8)]
    jobs = []
    for j in range(48):
        n = 20_000 + 1_777 * j
        recs = oc.generate_records(1000 * j, n, 16, 12, 1, 50_000 + 1000 * j, 100 + j)
        d = Dev(ctx, recs.nbytes, recs)
        jobs.append((recs, d, Dev(ctx, n * 16), Dev(ctx, n * 12), Dev(ctx, 64), Dev(ctx, 64), Dev(ctx, 64), Dev(ctx, n * 24)))
    ctx.synchronize()
    for j, (recs, d, b, u, r1, r2, r3, back) in enumerate(jobs):
        st = streams[j % 8]
        ctx.validate_reduce_async(d, len(recs), 16, 12, r1, st)
        ctx.unpack_async(d, len(recs), 16, 12, b, u, None, r2, st)
        ctx.pack_async(b, u, len(recs), 16, 12, back, d_result=r3, stream=st)
    for st in streams:
        st.synchronize()
    for recs, d, b, u, r1, r2, r3, back in jobs:
        want = oc.reduce_records(recs, 16, 12)
        assert ctx.read_result(r1.ptr) == want
        assert ctx.read_result(r2.ptr) == want
        got3 = ctx.read_result(r3.ptr)
        assert got3["n_records"] == len(recs) and got3["n_bad_records"] == 0 and got3["sum_index"] == 0
        for x in (d, b, u, r1, r2, r3, back):
            x.free()


# ---- K3 -----------------------------------------------------------------------------------
def gpu_pack(ctx, bc_rows, umi_rows, index=None, index_base=0):
    n, bc = bc_rows.shape
    umi = umi_rows.shape[1]
    db, du = Dev(ctx, bc_rows.nbytes, bc_rows), Dev(ctx, umi_rows.nbytes, umi_rows)
    di = Dev(ctx, 8 * n, index) if index is not None else None
    dr, df, res = Dev(ctx, 24 * n), Dev(ctx, n), Dev(ctx, 64)
    ctx.pack_async(db, du, n, bc, umi, dr, d_index=di, index_base=index_base, d_flags=df, d_result=res)
    ctx.synchronize()
    out = dr.get(ibu.RECORD_DTYPE, n), df.get(np.uint8, n), ctx.read_result(res.ptr)
    for x in (db, du, dr, df, res) + ((di,) if di else ()):
        x.free()
    return out


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("bc,umi", [(32, 32), (16, 16), (16, 12), (32, 16), (16, 32), (5, 3), (1, 1), (31, 7), (20, 28),
                                    (16, 10), (30, 2), (18, 26), (12, 12)])
def test_pack(ctx, n, bc, umi):
    b = oc.generate_ascii(3, n, bc, 30_000, 200_000, 7)
    u = oc.generate_ascii(3, n, umi, 30_000, 200_000, 8)
    idx = (np.arange(n, dtype=U64) * U64(0x9E3779B97F4A7C15)) if n % 2 else None
    gr, gf, gres = gpu_pack(ctx, b, u, idx, 1 << 40)
    orr, of, ores = oc.pack_records(b, u, idx, 1 << 40)
    assert np.array_equal(gr, orr) and np.array_equal(gf, of)
    for k in ("n_records", "n_bad_barcode", "n_bad_umi", "n_bad_records"):
        assert gres[k] == ores[k], k


def test_pack_every_byte_value(ctx):
    """Validity and the deterministic fill for all 256 byte values in every row position."""
    rows = np.full((256 * 32, 32), ord("A"), np.uint8)
    for pos in range(32):
        rows[pos * 256:(pos + 1) * 256, pos] = np.arange(256)
    gr, gf, _ = gpu_pack(ctx, rows, rows[:, :16].copy())
    orr, of, _ = oc.pack_records(rows, rows[:, :16].copy())
    assert np.array_equal(gr, orr) and np.array_equal(gf, of)
    valid = np.isin(np.arange(256), np.frombuffer(b"ACGTacgt", np.uint8))
    assert np.array_equal((gf[:256] & 1) == 0, valid)


# ---- generators ---------------------------------------------------------------------------
@pytest.mark.parametrize("mode,param", [(0, 0), (1, 100_000), (2, 0), (3, (4096 << 32) | 10_000), (4, (5 << 32) | 1000),
                                        (5, (20 << 32) | 100_000)])
@pytest.mark.parametrize("bc,umi", [(16, 12), (32, 32), (3, 9)])
def test_device_generators_match_oracle(ctx, mode, param, bc, umi):
    n = 100_003
    d = Dev(ctx, 24 * n)
    ctx.generate_records_async(d, 77, n, bc, umi, mode, param, 42)
    ctx.synchronize()
    assert np.array_equal(d.get(ibu.RECORD_DTYPE, n), oc.generate_records(77, n, bc, umi, mode, param, 42))
    d.free()
    a = Dev(ctx, n * bc)
    ctx.generate_ascii_async(a, 5, n, bc, 20_000, 100_000, 9)
    ctx.synchronize()
    assert np.array_equal(a.get(np.uint8, (n, bc)), oc.generate_ascii(5, n, bc, 20_000, 100_000, 9))
    a.free()


# ---- size-independent properties at BASELINE sizes ----------------------------------------------
@pytest.mark.parametrize("bc,umi,n", [(16, 12, 100_000_000), (32, 32, 100_000_000)])
def test_full_size_roundtrip(ctx, bc, umi, n):
    """configs[1]/[2] sizes: generate -> K1, unpack (K2) -> pack (K3) -> K1 again.  The repacked
    records equal the masked originals, checked by a checksum of checksums (sums + xor) and by
    closed-form counts of the injected invalid words."""
    recs, b, u, back, res = (Dev(ctx, s) for s in (24 * n, bc * n, umi * n, 24 * n, 64))
    ctx.generate_records_async(recs, 0, n, bc, umi, ibu.GEN_DIRTY, 10_000, 2024)
    ctx.validate_reduce_async(recs, n, bc, umi, res)
    ctx.synchronize()
    r0 = ctx.read_result(res.ptr)
    ctx.unpack_async(recs, n, bc, umi, b, u, None, res)
    ctx.synchronize()
    r1 = ctx.read_result(res.ptr)
    for k in ("n_records", "n_bad_barcode", "n_bad_umi", "n_bad_records"):
        assert r0[k] == r1[k], k
    assert r0["n_records"] == n
    if bc < 32:
        assert 0.8 * n / 100 < r0["n_bad_records"] < 1.2 * n / 100  # 10 000 ppm carry a raw word
    else:
        assert r0["n_bad_records"] == 0  # a 32-base word has no invalid bits
    ctx.pack_async(b, u, n, bc, umi, back, d_result=res)
    ctx.synchronize()
    r2 = ctx.read_result(res.ptr)
    assert r2["n_records"] == n and r2["n_bad_records"] == 0  # unpack only emits ACGT
    # clean twin: same generator without the dirt == masked originals
    ctx.generate_records_async(recs, 0, n, bc, umi, ibu.GEN_CLEAN, 0, 2024)
    ctx.validate_reduce_async(recs, n, bc, umi, res)
    ctx.synchronize()
    want = ctx.read_result(res.ptr)
    ctx.validate_reduce_async(back, n, bc, umi, res)
    ctx.synchronize()
    got = ctx.read_result(res.ptr)
    assert got == want and got["n_bad_records"] == 0
    assert got["sum_index"] == n * (n - 1) // 2
    # spot-check a window against the oracle
    w0, wn = 77_777_777, 4096
    host = np.zeros(wn, ibu.RECORD_DTYPE)
    ctx.d2h(host, back.ptr + 24 * w0)
    assert np.array_equal(host, oc.generate_records(w0, wn, bc, umi, 0, 0, 2024))
    for x in (recs, b, u, back, res):
        x.free()


def _d2h_slab(ctx, dptr, first, count, width, dtype=np.uint8):
    out = np.empty((count, width) if width else count, dtype)
    ctx.d2h(out, dptr + first * out.dtype.itemsize * (width or 1))
    return out


def test_full_size_unpack_every_byte_against_the_oracle(ctx):
    """configs[1], whole arrays: the 10^8 x 16 barcode bytes, 10^8 x 12 umi bytes, 10^8 flag bytes and the
    8-word result of ONE K2 launch over 10^8 dirty records equal the oracle's, slab by slab (the
    oracle regenerates the same records from the shared counter-based generator)."""
    n, bc, umi, slab = 100_000_000, 16, 12, 20_000_000
    recs, b, u, f, res = (Dev(ctx, s) for s in (24 * n, bc * n, umi * n, n, 64))
    ctx.generate_records_async(recs, 0, n, bc, umi, ibu.GEN_DIRTY, 10_000, 2024)
    ctx.unpack_async(recs, n, bc, umi, b, u, f, res)
    ctx.synchronize()
    got = ctx.read_result(res.ptr)
    total = dict.fromkeys(got, 0)
    for first in range(0, n, slab):
        want = oc.generate_records(first, slab, bc, umi, 1, 10_000, 2024)
        assert np.array_equal(_d2h_slab(ctx, recs.ptr, first, slab, 0, ibu.RECORD_DTYPE), want), first
        ob, ou, of, ucnt = oc.unpack_records(want, bc, umi, 0)
        ored = oc.reduce_records(want, bc, umi)  # (the oracle's unpack carries the counters only)
        assert all(ucnt[k] == ored[k] for k in ("n_records", "n_bad_barcode", "n_bad_umi", "n_bad_records"))
        assert np.array_equal(_d2h_slab(ctx, b.ptr, first, slab, bc), ob), first
        assert np.array_equal(_d2h_slab(ctx, u.ptr, first, slab, umi), ou), first
        assert np.array_equal(_d2h_slab(ctx, f.ptr, first, slab, 0), of), first
        for k in total:
            total[k] = (total[k] ^ ored[k]) if k == "xor_all" else (total[k] + ored[k]) % (1 << 64)
    assert got == total
    for x in (recs, b, u, f, res):
        x.free()


def test_full_size_pack_every_record_against_the_oracle(ctx):
    """configs[2], whole array: 10^8 ASCII pairs bc32/umi32 (1 % of the rows of either input carry an 'N', 5 % are lower
    case) -> Records; every record, every flag byte and the result equal the oracle's."""
    n, bc, umi, slab = 100_000_000, 32, 32, 20_000_000
    b, u, out, f, res = (Dev(ctx, s) for s in (bc * n, umi * n, 24 * n, n, 64))
    ctx.generate_ascii_async(b, 0, n, bc, 10_000, 50_000, 5)
    ctx.generate_ascii_async(u, 0, n, umi, 10_000, 50_000, 6)
    ctx.pack_async(b, u, n, bc, umi, out, d_flags=f, d_result=res, index_base=7)
    ctx.synchronize()
    got = ctx.read_result(res.ptr)
    total = dict.fromkeys(got, 0)
    for first in range(0, n, slab):
        hb, hu = oc.generate_ascii(first, slab, bc, 10_000, 50_000, 5), oc.generate_ascii(first, slab, umi, 10_000, 50_000, 6)
        want, wflags, wred = oc.pack_records(hb, hu, None, 7 + first)
        assert np.array_equal(_d2h_slab(ctx, out.ptr, first, slab, 0, ibu.RECORD_DTYPE), want), first
        assert np.array_equal(_d2h_slab(ctx, f.ptr, first, slab, 0), wflags), first
        for k in ("n_records", "n_bad_barcode", "n_bad_umi", "n_bad_records"):
            total[k] += wred[k]
    assert all(got[k] == total[k] for k in ("n_records", "n_bad_barcode", "n_bad_umi", "n_bad_records"))
    assert 0.005 * n < got["n_bad_records"] < 0.03 * n
    for x in (b, u, out, f, res):
        x.free()


# ---- K4: per-barcode table --------------------------------------------------------------------
def gpu_table(ctx, recs, mode=0):
    d = Dev(ctx, recs.nbytes, recs)
    rows, info = ctx.barcode_count(d, len(recs), mode)
    d.free()
    return rows, info


def sort_records(recs):
    return recs[np.lexsort((recs["index"], recs["umi"], recs["barcode"]))]  # Record's Ord (record.rs:58)


@pytest.mark.parametrize("n", [0, 1, 2, 255, 2047, 2048, 2049, 100_003, 3_000_005])
@pytest.mark.parametrize("mode,param", [(3, (64 << 32) | 1000), (3, (4 << 32) | 50), (2, 0), (0, 0), (1, 300_000)])
def test_barcode_table_sorted_stream(ctx, n, mode, param):
    recs = sort_records(oc.generate_records(0, n, 16, 12, mode, param, 21))
    rows, info = gpu_table(ctx, recs)
    want = on.barcode_table(recs)
    assert info["input_was_sorted"] and info["n_records"] == n
    assert np.array_equal(rows, want)
    assert info["n_distinct_pairs"] == int(want["n_distinct_umi"].sum())
    if 0 < n <= 100_003:
        t, pairs = oc.barcode_table(recs)
        assert np.array_equal(rows, t) and pairs == info["n_distinct_pairs"]


@pytest.mark.parametrize("n", [1, 2, 2049, 100_003, 3_000_005])
@pytest.mark.parametrize("bc,umi,mode,param", [(16, 12, 3, (64 << 32) | 1000), (16, 12, 2, 0), (16, 12, 0, 0),
                                               (16, 12, 1, 300_000), (32, 32, 3, (1000 << 32) | 77), (5, 3, 0, 0)])
def test_barcode_table_unsorted(ctx, n, bc, umi, mode, param):
    recs = oc.generate_records(0, n, bc, umi, mode, param, 22)
    rows, info = gpu_table(ctx, recs)
    want = on.barcode_table(recs)
    assert np.array_equal(rows, want)
    assert info["n_distinct_pairs"] == int(want["n_distinct_umi"].sum())
    rows2, info2 = gpu_table(ctx, recs, mode=2)  # without the streaming attempt: the same table
    assert np.array_equal(rows2, want) and not info2["input_was_sorted"]
    # every implementation of the unsorted path, with and without the header's lengths as a hint
    for path in (ibu.COUNT_PATH_PARTITION, ibu.COUNT_PATH_LEGACY):
        for lens in (0, ibu.count_lens(bc, umi)):
            rows3, info3 = gpu_table(ctx, recs, mode=2 | path | lens)
            assert np.array_equal(rows3, want), (path, lens)
            assert info3["n_distinct_pairs"] == int(want["n_distinct_umi"].sum()) and info3["n_records"] == n


def test_barcode_table_require_sorted_mode(ctx):
    recs = oc.generate_records(0, 50_000, 16, 12, 0, 0, 23)
    rows, info = gpu_table(ctx, recs, mode=1)
    assert not info["input_was_sorted"] and len(rows) == 0
    rows, info = gpu_table(ctx, sort_records(recs), mode=1)
    assert info["input_was_sorted"] and np.array_equal(rows, on.barcode_table(recs))


def test_barcode_table_reference_pattern_closed_form(ctx):
    """examples/parallel.rs:65-69 pattern at N = 10^7: 10^6 barcodes x 10 records x 1 UMI."""
    n = 10_000_000
    d = Dev(ctx, 24 * n)
    ctx.generate_records_async(d, 0, n, 16, 12, ibu.GEN_PATTERN, 0, 0)
    ctx.synchronize()
    rows, info = ctx.barcode_count(d, n)
    d.free()
    assert not info["input_was_sorted"] and len(rows) == 1_000_000 and info["n_distinct_pairs"] == 1_000_000
    assert np.array_equal(rows["barcode"], np.arange(1_000_000, dtype=U64))
    assert np.all(rows["n_records"] == 10) and np.all(rows["n_distinct_umi"] == 1)


@pytest.mark.parametrize("gen,param,shape", [(2, 0, "pattern"), (3, (20 << 32) | 1_000_000, "whitelist")])
def test_barcode_table_three_partition_levels_closed_form(ctx, gen, param, shape):
    """1.5 x 10^8 unsorted records: 2^18 buckets, i.e. three levels of the staged partition (two levels
    cover up to 2^17).  The table of the whole input equals the sum of the tables of its thirds
    (n_records adds by barcode; both are checked against closed forms where one exists)."""
    n = 150_000_000
    d = Dev(ctx, 24 * n)
    ctx.generate_records_async(d, 0, n, 16, 12, gen, param, 9)
    ctx.synchronize()
    rows, info = ctx.barcode_count(d, n, mode=ibu.count_lens(16, 12))
    assert not info["input_was_sorted"] and int(rows["n_records"].sum()) == n
    assert np.all(np.diff(rows["barcode"].astype(np.int64)) > 0)
    if shape == "pattern":  # examples/parallel.rs:65-69: barcode i % 10^6 always meets umi 31 i % 10^6
        assert len(rows) == 1_000_000 and info["n_distinct_pairs"] == 1_000_000
        assert np.all(rows["n_records"] == 150) and np.all(rows["n_distinct_umi"] == 1)
    else:
        third = n // 3
        acc = {}
        for k in range(3):  # each third takes the two-level path
            part, _ = ctx.barcode_count(d.ptr + 24 * third * k, third, mode=ibu.count_lens(16, 12))
            assert int(part["n_records"].sum()) == third
            acc[k] = part
        cat = np.concatenate(list(acc.values()))
        order = np.argsort(cat["barcode"], kind="stable")
        bc, cnt = cat["barcode"][order], cat["n_records"][order]
        head = np.ones(len(bc), bool)
        head[1:] = bc[1:] != bc[:-1]
        assert np.array_equal(bc[head], rows["barcode"])
        assert np.array_equal(np.add.reduceat(cnt, np.flatnonzero(head)), rows["n_records"])
        assert np.all(rows["n_distinct_umi"] <= 20) and info["n_distinct_pairs"] == int(rows["n_distinct_umi"].sum())
        # umi space 20, 150 records per barcode on average: nearly every barcode has seen all 20 umis
        assert (rows["n_distinct_umi"] == 20).mean() > 0.95
    d.free()


def test_barcode_table_capacity_retry(ctx):
    """More distinct barcodes than the optimistic 8 Mi-row table: the exact-size second pass."""
    n = 9_000_000
    recs = oc.generate_records(0, n, 16, 12, 0, 0, 24)
    recs = recs[np.lexsort((recs["umi"], recs["barcode"]))]
    rows, info = gpu_table(ctx, recs)
    want = on.barcode_table(recs)
    assert info["input_was_sorted"] and len(rows) > (8 << 20)
    assert np.array_equal(rows, want)


def zipf_records(n, n_barcodes, umi_space, seed, bc=16, umi=12):
    """10x-like shape: a whitelist of barcodes drawn log-uniformly (Zipf-ish),
    UMIs uniform in a small space, unsorted."""
    rng = np.random.default_rng(seed)
    white = rng.integers(0, 1 << (2 * bc), n_barcodes, dtype=U64)
    rank = np.minimum((n_barcodes ** rng.random(n)).astype(np.int64), n_barcodes - 1)
    recs = np.zeros(n, ibu.RECORD_DTYPE)
    recs["barcode"] = white[rank]
    recs["umi"] = rng.integers(0, umi_space, n, dtype=U64)
    recs["index"] = np.arange(n, dtype=U64)
    return recs


@pytest.mark.parametrize("n,n_barcodes,umi_space", [(70_000, 300, 50), (1_000_003, 5000, 40), (3_000_005, 100_000, 1 << 24),
                                                    (2_000_000, 1, 1 << 24), (2_000_000, 1_000_000, 1)])
def test_barcode_table_partition_path_skewed(ctx, n, n_barcodes, umi_space):
    """Heavy barcodes (one holds ~1/ln(B) of the records) and heavy duplication: the partition path's
    buckets are balanced by the hash of the pair, whatever the barcode distribution."""
    recs = zipf_records(n, n_barcodes, umi_space, 41)
    want = on.barcode_table(recs)
    for mode in (2 | ibu.COUNT_PATH_PARTITION | ibu.count_lens(16, 12), 0):
        rows, info = gpu_table(ctx, recs, mode=mode)
        assert np.array_equal(rows, want)
        assert info["n_distinct_pairs"] == int(want["n_distinct_umi"].sum())


def test_barcode_table_partition_path_wide_words(ctx):
    """Words wider than the header allows (examples/random.rs:46 writes a full random u64 into a 12-base
    UMI) go through the side list and land in the same table — including the barcode 2^64 - 1."""
    n = 1_500_000
    recs = zipf_records(n, 20_000, 64, 42)
    rng = np.random.default_rng(43)
    wide = rng.random(n) < 0.01
    recs["umi"][wide] = rng.integers(0, 1 << 63, int(wide.sum()), dtype=U64) | U64(1 << 40)
    recs["barcode"][rng.random(n) < 0.002] |= U64(1 << 50)
    recs["barcode"][:7] = U64(2**64 - 1)
    recs["umi"][:3] = U64(2**64 - 1)
    recs["umi"][3:7] = np.arange(1, 5, dtype=U64)
    want = on.barcode_table(recs)
    for lens in (0, ibu.count_lens(16, 12)):
        rows, info = gpu_table(ctx, recs, mode=2 | ibu.COUNT_PATH_PARTITION | lens)
        assert np.array_equal(rows, want)
        assert info["n_distinct_pairs"] == int(want["n_distinct_umi"].sum())
    assert rows["barcode"][-1] == U64(2**64 - 1) and rows["n_records"][-1] == 7 and rows["n_distinct_umi"][-1] == 5


def test_barcode_table_random_rs_full_width_umis(ctx):
    """examples/random.rs:35-47: barcode in [0, 1000), umi a full random u64, index in [0, 10000) —
    almost every umi word is invalid for umi12 (n (1 - 2^-40) of them) and every pair is distinct."""
    n = 1_000_000
    rng = np.random.default_rng(44)
    recs = np.zeros(n, ibu.RECORD_DTYPE)
    recs["barcode"] = rng.integers(0, 1000, n, dtype=U64)
    recs["umi"] = rng.integers(0, 2**64, n, dtype=U64)
    recs["index"] = rng.integers(0, 10000, n, dtype=U64)
    red = gpu_reduce(ctx, recs, 16, 12)
    assert red == oc.reduce_records(recs, 16, 12)
    assert red["n_bad_umi"] == int((recs["umi"] >> U64(24) != 0).sum()) and red["n_bad_umi"] > n - 10 and red["n_bad_barcode"] == 0
    want = on.barcode_table(recs)
    for mode in (0, 2 | ibu.COUNT_PATH_PARTITION | ibu.count_lens(16, 12)):
        rows, info = gpu_table(ctx, recs, mode=mode)
        assert np.array_equal(rows, want) and len(rows) == 1000


def test_barcode_table_about_as_many_barcodes_as_records(ctx):
    """3 x 10^7 random records (bc16 / umi12, 1 % with an unmasked word): nearly every record has a barcode
    of its own — the ordered form of the partition path (k4_ordered.cuh), chosen by the sample.  Whole
    table against the oracle; the forced sort fallback must give the same rows."""
    import os

    n = 30_000_000
    recs = oc.generate_records(0, n, 16, 12, 1, 10_000, 77)
    recs[n // 2:n // 2 + 1_000_000] = recs[:1_000_000]  # some repeated records ...
    recs["umi"][n // 2:n // 2 + 500_000] ^= U64(5)      # ... half of them with another UMI
    want = on.barcode_table(recs)
    old = os.environ.get("IBU_B200_K4_ORDERED")
    os.environ["IBU_B200_K4_ORDERED"] = "2"  # strict: falling back to the sort is an error here
    try:
        rows, info = gpu_table(ctx, recs, mode=2 | ibu.count_lens(16, 12))
    finally:
        os.environ.pop("IBU_B200_K4_ORDERED") if old is None else os.environ.__setitem__("IBU_B200_K4_ORDERED", old)
    assert np.array_equal(rows, want) and info["n_distinct_pairs"] == int(want["n_distinct_umi"].sum())
    rows2, info2 = gpu_table(ctx, recs, mode=2 | ibu.COUNT_PATH_SORT)
    assert np.array_equal(rows2, want) and info2["n_distinct_pairs"] == info["n_distinct_pairs"]


def test_barcode_table_ordered_path_exact_layout_after_overflow(ctx):
    """Barcodes whose 7th bit from the top is 0 for 70 % of the records: the first partition level (top 6
    bits) is even, the final buckets are not, the uniform layout of the last level overflows and the
    ordered path lays it out again exactly from a histogram — still without the sort fallback."""
    import os

    n = 3_000_000
    recs = oc.generate_records(0, n, 16, 12, 0, 0, 78)
    rng = np.random.default_rng(79)
    recs["barcode"][rng.random(n) < 0.7] &= ~U64(1 << 25)
    want = on.barcode_table(recs)
    old = os.environ.get("IBU_B200_K4_ORDERED")
    os.environ["IBU_B200_K4_ORDERED"] = "2"
    try:
        rows, info = gpu_table(ctx, recs, mode=2 | ibu.COUNT_PATH_PARTITION | ibu.count_lens(16, 12))
    finally:
        os.environ.pop("IBU_B200_K4_ORDERED") if old is None else os.environ.__setitem__("IBU_B200_K4_ORDERED", old)
    assert np.array_equal(rows, want) and info["n_distinct_pairs"] == int(want["n_distinct_umi"].sum())


def test_barcode_table_sorted_full_size_closed_form(ctx):
    """10^8 sorted records (1000 per barcode, 5 per umi): 10^5 rows x 1000 records x 200 UMIs,
    streamed in one pass (24 B/record)."""
    n = 100_000_000
    d = Dev(ctx, 24 * n)
    ctx.generate_records_async(d, 0, n, 16, 12, ibu.GEN_SORTED, (5 << 32) | 1000, 0)
    ctx.synchronize()
    rows, info = ctx.barcode_count(d, n, mode=1)
    d.free()
    assert info["input_was_sorted"] and len(rows) == 100_000 and info["n_distinct_pairs"] == 20_000_000
    assert np.array_equal(rows["barcode"], np.arange(100_000, dtype=U64))
    assert np.all(rows["n_records"] == 1000) and np.all(rows["n_distinct_umi"] == 200)


# ---- device sort by Record's Ord, pair tables, weighted merge ----------------------------------
def gpu_sort(ctx, recs):
    n = len(recs)
    d, o = Dev(ctx, recs.nbytes, recs), Dev(ctx, 24 * n)
    ctx.sort_records(d, n, o)
    out = o.get(ibu.RECORD_DTYPE, n)
    d.free(), o.free()
    return out


@pytest.mark.parametrize("n", [0, 1, 2, 2047, 2048, 2049, 100_003, 3_000_005])
@pytest.mark.parametrize("bc,umi,mode,param", [(16, 12, 3, (64 << 32) | 1000), (16, 12, 1, 500_000), (32, 32, 0, 0),
                                               (16, 12, 2, 0), (16, 12, 4, (5 << 32) | 1000), (3, 2, 0, 0)])
@pytest.mark.parametrize("index_order", ["descending", "ascending", "one swap"])
def test_sort_records_matches_record_ord(ctx, n, bc, umi, mode, param, index_order):
    """record.rs:29-32,58: lexicographic (barcode, umi, index).  Input that already comes in index order
    skips the index passes (the sort is stable); a single out-of-order pair must bring them back."""
    recs = oc.generate_records(0, n, bc, umi, mode, param, 31)
    if n > 10 and index_order == "descending":
        recs["index"] = recs["index"][::-1].copy()  # index must be sorted as the third key, not kept
    if n > 10 and index_order == "one swap":
        recs["barcode"][n - 2], recs["umi"][n - 2] = recs["barcode"][n - 1], recs["umi"][n - 1]
        recs["index"][n - 2], recs["index"][n - 1] = recs["index"][n - 1], recs["index"][n - 2]
    got = gpu_sort(ctx, recs)
    assert np.array_equal(got, sort_records(recs))
    if n:
        assert np.array_equal(gpu_sort(ctx, got), got)  # idempotent
        rows, info = gpu_table(ctx, got, mode=1)        # and the sorted fast path accepts it
        assert info["input_was_sorted"] and np.array_equal(rows, on.barcode_table(recs))


@pytest.mark.parametrize("word,bit", [("barcode", 60), ("umi", 41), ("index", 35)])
def test_sort_records_bits_the_sample_misses(ctx, word, bit):
    """ibu_gpu_sort_records guesses the digits to sort from 2^16 sampled records and confirms the guess
    with exact masks: ONE record with a bit far above every other's must still end up in its place
    (the digit passes it needs are not in the guess), and so must an index word that breaks the order."""
    n = 2_500_003
    recs = oc.generate_records(0, n, 16, 12, 0, 0, 33)
    recs[word][n // 2 + 7] |= np.uint64(1) << np.uint64(bit)
    got = gpu_sort(ctx, recs)
    assert np.array_equal(got, sort_records(recs))


@pytest.mark.parametrize("shape", ["random", "repeats", "pattern", "uneven top bits", "sorted already", "few umis"])
@pytest.mark.parametrize("index_order", ["ascending", "descending", "shuffled"])
def test_sort_records_by_partition(ctx, shape, index_order):
    """From 2^20 records on ibu_gpu_sort_records partitions (key, index) pairs by the key's top bits and sorts
    every final bucket in shared memory (k4_sort_records_msd).  IBU_B200_SORT_MSD=2: handing the input to the
    LSD sort is an error here, so these shapes prove that path; the result is Record's Ord (record.rs:58)."""
    import os

    n = 2_300_007
    rng = np.random.default_rng(91)
    if shape == "pattern":
        recs = oc.generate_records(0, n, 16, 12, 2, 0, 34)
    elif shape == "sorted already":
        recs = oc.generate_records(0, n, 16, 12, 4, (5 << 32) | 1000, 34)
    else:
        recs = oc.generate_records(0, n, 16, 12, 0, 0, 34)
    if shape == "repeats":  # whole records again (equal keys AND equal indices), and keys again with other indices
        recs[n // 2:n // 2 + 400_000] = recs[:400_000]
        recs["barcode"][n - 300_000:] = recs["barcode"][:300_000]
        recs["umi"][n - 300_000:] = recs["umi"][:300_000]
    if shape == "uneven top bits":  # the uniform layout of the last level overflows: exact layout from a histogram
        recs["barcode"][rng.random(n) < 0.7] &= ~U64(1 << 25)
    if shape == "few umis":  # 16 UMIs: long runs of one barcode's records differ only far down the key
        recs["barcode"] >>= U64(12)
        recs["umi"] &= U64(15)
    if index_order == "descending":
        recs["index"] = recs["index"][::-1].copy()
    elif index_order == "shuffled":
        recs["index"] = rng.permutation(n).astype(U64) * U64(1 << 20)  # (wide index words)
    want = sort_records(recs)
    old = os.environ.get("IBU_B200_SORT_MSD")
    # ("sorted already": barcodes 0..2299 fill 56 % of their 12-bit range, the first level's uniform layout
    # overflows and the LSD sort takes over — allowed, the result must be right either way)
    os.environ["IBU_B200_SORT_MSD"] = "1" if shape == "sorted already" else "2"
    try:
        got = gpu_sort(ctx, recs)
    finally:
        os.environ.pop("IBU_B200_SORT_MSD") if old is None else os.environ.__setitem__("IBU_B200_SORT_MSD", old)
    assert np.array_equal(got, want)
    os.environ["IBU_B200_SORT_MSD"] = "0"  # and the LSD sort agrees
    try:
        got = gpu_sort(ctx, recs)
    finally:
        os.environ.pop("IBU_B200_SORT_MSD") if old is None else os.environ.__setitem__("IBU_B200_SORT_MSD", old)
    assert np.array_equal(got, want)


def np_pair_table(recs, weighted=False):
    order = np.lexsort((recs["umi"], recs["barcode"]))
    b, u, w = recs["barcode"][order], recs["umi"][order], recs["index"][order]
    head = np.ones(len(b), bool)
    head[1:] = (b[1:] != b[:-1]) | (u[1:] != u[:-1])
    seg = np.cumsum(head) - 1
    out = np.zeros(int(head.sum()), ibu.RECORD_DTYPE)
    out["barcode"], out["umi"] = b[head], u[head]
    out["index"] = np.bincount(seg, weights=w.astype(np.float64) if weighted else None).astype(U64) if len(b) else 0
    return out


@pytest.mark.parametrize("n", [1, 2049, 100_003, 2_000_003])
@pytest.mark.parametrize("mode,param,pre_sorted", [(3, (64 << 32) | 1000, False), (3, (8 << 32) | 50, True), (2, 0, False),
                                                   (4, (5 << 32) | 1000, True)])
def test_pair_table(ctx, n, mode, param, pre_sorted):
    recs = oc.generate_records(0, n, 16, 12, mode, param, 32)
    if pre_sorted:
        recs = sort_records(recs)
    d = Dev(ctx, recs.nbytes, recs)
    ptr, n_pairs = ctx.pair_table(d, n)
    got = np.zeros(n_pairs, ibu.RECORD_DTYPE)
    ctx.d2h(got, ptr)
    ctx.free(ptr)
    want = np_pair_table(recs)
    assert np.array_equal(got, want) and int(got["index"].sum()) == n
    for flags in (ibu.COUNT_PATH_PARTITION, ibu.COUNT_PATH_PARTITION | ibu.PAIRS_UNORDERED | ibu.count_lens(16, 12)):
        ptr, n_pairs = ctx.pair_table(d, n, flags=flags)
        got = np.zeros(n_pairs, ibu.RECORD_DTYPE)
        ctx.d2h(got, ptr)
        ctx.free(ptr)
        if flags & ibu.PAIRS_UNORDERED:
            got = got[np.lexsort((got["umi"], got["barcode"]))]
        assert np.array_equal(got, want), flags
    d.free()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_weighted_merge_of_shard_pair_tables_is_exact(ctx, world):
    """distinct-UMI counts are not additive across shards; exchanging the shards'
    de-duplicated (barcode, umi, count) tables and counting them weighted is exact."""
    n = 1_000_003
    recs = oc.generate_records(0, n, 16, 12, 3, (32 << 32) | 5000, 33)  # unsorted, heavy duplication
    shards = []
    for r in range(world):
        s, e = ibu.shard_range(n, r, world)
        d = Dev(ctx, 24 * (e - s), recs[s:e])
        ptr, k = ctx.pair_table(d, e - s)
        part = np.zeros(k, ibu.RECORD_DTYPE)
        ctx.d2h(part, ptr)
        ctx.free(ptr), d.free()
        shards.append(part)
    merged = np.concatenate(shards)  # what an all-gather / all-to-all of the pair tables delivers
    assert len(merged) < n
    d = Dev(ctx, merged.nbytes, merged)
    rows, info = ctx.barcode_count(d, len(merged), ibu.COUNT_WEIGHTED)
    assert np.array_equal(rows, on.barcode_table(recs))
    assert int(rows["n_records"].sum()) == n
    rows, info = ctx.barcode_count(d, len(merged), ibu.COUNT_WEIGHTED | 2 | ibu.COUNT_PATH_PARTITION)
    d.free()
    assert np.array_equal(rows, on.barcode_table(recs)) and info["n_records"] == len(merged)


def test_weighted_count_of_near_distinct_pairs(ctx):
    """The owner count of a multi-GPU merge over near-distinct data: (barcode, umi, multiplicity) rows with about
    as many barcodes as rows.  Neither table form of the partition path takes weighted near-distinct input; the
    sort behind the segment pass does — by partition (k4_sort_records_msd) and, forced, by LSD digits."""
    import os

    n = 2_200_003
    rng = np.random.default_rng(5)
    recs = oc.generate_records(0, n, 16, 12, 0, 0, 35)
    recs[n - 200_000:] = recs[:200_000]                       # the same pair from two shards
    recs["umi"][n - 100_000:] ^= U64(3)                       # ... and the same barcode with another UMI
    recs["index"] = rng.integers(1, 6, n).astype(U64)         # multiplicities
    order = np.lexsort((recs["umi"], recs["barcode"]))
    b, u, w = recs["barcode"][order], recs["umi"][order], recs["index"][order]
    new_bc = np.ones(n, bool)
    new_bc[1:] = b[1:] != b[:-1]
    new_pair = new_bc.copy()
    new_pair[1:] |= u[1:] != u[:-1]
    seg = np.cumsum(new_bc) - 1
    want_bc = b[new_bc]
    want_rec = np.bincount(seg, weights=w.astype(np.float64)).astype(U64)
    want_dist = np.bincount(seg, weights=new_pair.astype(np.float64)).astype(U64)
    d = Dev(ctx, recs.nbytes, recs)
    old = os.environ.get("IBU_B200_SORT_MSD")
    try:
        for env in (None, "0"):
            if env is not None:
                os.environ["IBU_B200_SORT_MSD"] = env
            rows, info = ctx.barcode_count(d, n, ibu.COUNT_WEIGHTED | 2)
            assert np.array_equal(rows["barcode"], want_bc) and np.array_equal(rows["n_records"], want_rec)
            assert np.array_equal(rows["n_distinct_umi"], want_dist)
    finally:
        os.environ.pop("IBU_B200_SORT_MSD", None) if old is None else os.environ.__setitem__("IBU_B200_SORT_MSD", old)
        d.free()


def test_partition_by_owner_and_emulated_all_to_all(ctx):
    """The exchange of ibu_b200.distributed.exact_barcode_table emulated on one GPU: per-shard pair
    tables -> owner buckets -> every owner counts what it would receive -> concatenated rows."""
    from ibu_b200 import distributed as ibd

    n, world = 600_011, 4
    recs = oc.generate_records(0, n, 16, 12, 3, (32 << 32) | 3000, 34)
    inbox = [[] for _ in range(world)]
    for r in range(world):
        s, e = ibu.shard_range(n, r, world)
        d = Dev(ctx, 24 * (e - s), recs[s:e])
        ptr, k = ctx.pair_table(d, e - s)
        out = Dev(ctx, 24 * k)
        counts = ctx.partition_by_owner(ptr, k, world, out)
        buckets = out.get(ibu.RECORD_DTYPE, k)
        ctx.free(ptr), d.free(), out.free()
        assert sum(counts) == k
        off = 0
        for o, c in enumerate(counts):
            part = buckets[off:off + c]
            assert np.all(ibd.owner_of(part["barcode"], world) == o)
            inbox[o].append(part)
            off += c
    tables = []
    for o in range(world):
        got = np.concatenate(inbox[o])
        d = Dev(ctx, got.nbytes, got)
        rows, _ = ctx.barcode_count(d, len(got), ibu.COUNT_WEIGHTED)
        d.free()
        tables.append(rows)
    cat = np.concatenate(tables)
    assert np.array_equal(cat[np.argsort(cat["barcode"])], on.barcode_table(recs))
