"""Cross-checks the two independent oracles (C++ scalar vs numpy) on the semantics the
reference does not pin (pack/unpack/validation/barcode table), plus the
codec known answers derived from record.rs:19-27 and bitnuc's published convention."""
import numpy as np
import pytest

from oracle import oracle_c as oc
from oracle import oracle_np as on


def test_codec_kats():  # record.rs:19-27 + bitnuc: first base in the least-significant bits
    assert oc.pack_word(b"ACGT") == (0xE4, False)
    assert oc.unpack_word(0xE4, 4) == b"ACGT"
    assert oc.pack_word(b"acgt") == (0xE4, False)
    assert oc.pack_word(b"T" * 12) == (0xFFFFFF, False)
    assert oc.pack_word(b"T" * 32) == (2**64 - 1, False)
    assert oc.pack_word(b"A" * 7) == (0, False)
    assert oc.pack_word(b"ACGN")[1] is True
    assert oc.unpack_word(2**64 - 1, 32) == b"T" * 32
    assert oc.unpack_word(0xFFFF_FFFF_FFFF_FFFF, 3) == b"TTT"  # bits >= 2L ignored
    for c in range(256):  # validity: exactly ACGTacgt
        assert oc.pack_word(bytes([c]))[1] == (chr(c) not in "ACGTacgt")
        c1 = (c >> 1) & 3
        assert oc.pack_word(bytes([c]))[0] == c1 ^ (c1 >> 1)


def test_kernel_validity_identity_exhaustive():
    """K3's table-free validity test (kernels.cuh pack4_top) against the oracle on every byte value
    and on random 32-bit groups: ((c & 0xD9) ^ (T ? 0x11 : 0)) == 0x41, T = bit 2 & ~bit 1."""
    for c in range(256):
        t = (c >> 2) & ~(c >> 1) & 1
        assert (((c & 0xD9) ^ (0x11 * t)) != 0x41) == oc.pack_word(bytes([c]))[1], c
    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(b"ACGTacgtNn@BDEHPQSUVWXd`\x00\xff", np.uint8)
    for _ in range(2000):
        group = alphabet[rng.integers(0, len(alphabet), 4)]
        w = int.from_bytes(group.tobytes(), "little")
        s1, s2 = w >> 1, w >> 2
        t = s2 & ~s1 & 0x01010101
        bad = (((w & 0xD9D9D9D9) ^ (t * 0x11)) ^ 0x41414141) != 0
        code = (s1 & 0x03030303) ^ (s2 & 0x01010101)
        top = ((code * 0x01041040) & 0xFFFFFFFF) >> 24
        want_w, want_bad = oc.pack_word(group.tobytes())
        assert bad == want_bad and (want_bad or top == want_w), group


@pytest.mark.parametrize("length", [1, 2, 5, 12, 16, 31, 32])
def test_codec_roundtrip_cross(length):
    rng = np.random.default_rng(length)
    w = rng.integers(0, 2**64, 1000, dtype=np.uint64) & np.uint64(on.low_mask(length))
    rows = on.unpack_words(w, length)
    for i in range(0, 1000, 97):
        assert oc.unpack_word(int(w[i]), length) == rows[i].tobytes()
        assert oc.pack_word(rows[i].tobytes()) == (int(w[i]), False)
    back, bad = on.pack_rows(rows)
    assert np.array_equal(back, w) and not bad.any()


@pytest.mark.parametrize("mode,param", [(0, 0), (1, 10_000), (1, 1_000_000), (2, 0), (3, 1000),
                                        (3, (4096 << 32) | 50_000), (4, (5 << 32) | 1000), (4, 0), (5, (20 << 32) | 100_000),
                                        (5, 1), (5, 0)])
@pytest.mark.parametrize("bc,umi", [(16, 12), (32, 32), (1, 1), (7, 31)])
def test_generators_agree(mode, param, bc, umi):
    a = oc.generate_records(123, 20_000, bc, umi, mode, param, 42)
    b = on.generate_records(123, 20_000, bc, umi, mode, param, 42)
    assert np.array_equal(a, b)
    # chunked generation is position-independent (counter-based)
    c = oc.generate_records(123 + 5000, 100, bc, umi, mode, param, 42)
    assert np.array_equal(c, a[5000:5100])
    ra, rb = oc.reduce_records(a, bc, umi), on.reduce_records(b, bc, umi)
    assert ra == rb
    if mode == 0:
        assert ra["n_bad_records"] == 0
    if mode == 1 and param == 1_000_000 and bc < 32 and umi < 31:
        assert ra["n_bad_records"] > 15_000


@pytest.mark.parametrize("length", [1, 12, 16, 32])
def test_ascii_generator_and_pack(length):
    a = oc.generate_ascii(7, 5000, length, 20_000, 100_000, 9)
    b = on.generate_ascii(7, 5000, length, 20_000, 100_000, 9)
    assert np.array_equal(a, b)
    u = oc.generate_ascii(8, 5000, length, 0, 0, 10)
    recs, flags, red = oc.pack_records(a, u, None, 1000)
    wa, bad_a = on.pack_rows(a)
    wu, bad_u = on.pack_rows(u)
    assert np.array_equal(recs["barcode"], wa) and np.array_equal(recs["umi"], wu)
    assert np.array_equal(recs["index"], 1000 + np.arange(5000, dtype=np.uint64))
    assert np.array_equal(flags, bad_a.astype(np.uint8) | (bad_u.astype(np.uint8) << 1))
    assert red["n_bad_barcode"] == int(bad_a.sum()) and red["n_bad_umi"] == 0
    assert bad_a.sum() > 0 and not bad_u.any()


@pytest.mark.parametrize("bc,umi", [(16, 12), (32, 32), (3, 5)])
def test_unpack_records_cross(bc, umi):
    r = oc.generate_records(0, 30_000, bc, umi, 1, 50_000, 3)
    b, u, fl, red = oc.unpack_records(r, bc, umi, 3)
    assert np.array_equal(b, on.unpack_words(r["barcode"], bc))
    assert np.array_equal(u, on.unpack_words(r["umi"], umi))
    want = on.reduce_records(r, bc, umi)
    for k in ("n_records", "n_bad_barcode", "n_bad_umi", "n_bad_records"):
        assert red[k] == want[k]
    bb = (r["barcode"] & np.uint64(~on.low_mask(bc) & (2**64 - 1))) != 0
    bu = (r["umi"] & np.uint64(~on.low_mask(umi) & (2**64 - 1))) != 0
    assert np.array_equal(fl, bb.astype(np.uint8) | (bu.astype(np.uint8) << 1))
    # unpack -> pack round trip recovers the masked words
    recs, flags, _ = oc.pack_records(b, u, r["index"])
    assert np.array_equal(recs["barcode"], r["barcode"] & np.uint64(on.low_mask(bc)))
    assert np.array_equal(recs["umi"], r["umi"] & np.uint64(on.low_mask(umi)))
    assert not flags.any()


@pytest.mark.parametrize("mode,param", [(3, (64 << 32) | 500), (2, 0), (0, 0)])
def test_barcode_table_cross(mode, param):
    r = oc.generate_records(0, 50_000, 16, 12, mode, param, 11)
    t, pairs = oc.barcode_table(r)
    tn = on.barcode_table(r)
    assert np.array_equal(t, tn)
    assert pairs == int(tn["n_distinct_umi"].sum()) and int(t["n_records"].sum()) == 50_000
    assert np.all(np.diff(t["barcode"].astype(object)) > 0) if len(t) > 1 else True


def test_barcode_table_via_process_parallel(tmp_ibu):  # parallel.rs:79-98 processor through the driver
    r = oc.generate_records(0, 40_000, 16, 12, 3, (32 << 32) | 300, 5)
    oc.write_file(tmp_ibu, oc.header_new(16, 12), r)
    t = oc.MmapReader(tmp_ibu).process_parallel_barcodes(3)
    assert np.array_equal(t, on.barcode_table(r))
