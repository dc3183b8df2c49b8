"""The committed golden fixtures (tests/golden/, written by tests/golden/make_golden.py) against
the two oracles, the host side of the library, and — on a GPU — the CUDA path through the C ABI.
The fixtures anchor on the reference's own known-answer tests (mmap.rs:454-481, header.rs:373-378,
writer.rs:645) where those exist; the codec / validation / table vectors are PARITY UNPINNED
 and frozen here so that no later change can move them silently."""
import json
import os

import numpy as np
import pytest

import ibu_b200 as ibu
from oracle import oracle_c as oc
from oracle import oracle_np as on

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
G = json.load(open(os.path.join(HERE, "golden.json")))
KAT = os.path.join(HERE, G["ref_kat_10000"]["file"])
U64 = np.uint64


def recs_from_hex(h: str) -> np.ndarray:
    return np.frombuffer(bytes.fromhex(h), on.RECORD_DTYPE).copy()


# ------------------------------------------------------------------ CPU: oracles and host API
def test_generator_script_reproduces_the_committed_fixtures(tmp_path):
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main(str(tmp_path))
    for name in ("golden.json", "ref_kat_10000.ibu"):
        assert open(tmp_path / name, "rb").read() == open(os.path.join(HERE, name), "rb").read(), name



def test_kat_file_bytes_and_oracles():
    blob = open(KAT, "rb").read()
    assert len(blob) == G["ref_kat_10000"]["bytes"] == 32 + 24 * 10_000  # writer.rs:645,673
    assert blob[:32].hex() == G["headers"]["16_12_0"] and blob[:4] == b"IBU!"  # header.rs:373-378
    hdr, recs = on.read_file(KAT)
    assert (int(hdr["bc_len"]), int(hdr["umi_len"])) == (16, 12) and len(recs) == 10_000
    assert on.reduce_records(recs, 16, 12) == G["ref_kat_10000"]["reduce"]
    want, trace = oc.MmapReader(KAT).process_parallel_reduce(4)  # mmap.rs:454-481: 4 threads
    assert want == G["ref_kat_10000"]["reduce"] and sum(t["records"] for t in trace) == 10_000
    assert (want["sum_barcode"] + want["sum_umi"] + want["sum_index"]) % 2**64 == G["ref_kat_10000"]["count_sum"]


def test_kat_file_through_the_host_api():
    reader = ibu.MmapReader(KAT)
    assert reader.len() == 10_000 and reader.header().as_bytes().hex() == G["headers"]["16_12_0"]
    sl = reader.slice(100, 200)  # mmap.rs:396-423
    assert sl["barcode"].tolist() == list(range(100, 200)) and sl["index"][0] == 300
    hdr, recs = ibu.load_to_vec(KAT)  # reader.rs:668-697
    assert len(recs) == 10_000 and np.array_equal(recs["umi"], 2 * np.arange(10_000, dtype=U64))


def test_headers_golden():
    for key, want in G["headers"].items():
        bc, umi, srt = map(int, key.split("_"))
        h = ibu.Header(bc, umi)
        if srt:
            h.set_sorted()
        assert h.as_bytes().hex() == want == on.header_bytes(bc, umi, bool(srt)).hex()


def test_codec_golden_on_both_oracles():
    for v in G["codec"]:
        s, w = v["ascii"].encode(), int(v["word"], 16)
        assert oc.pack_word(s) == (w, False)
        assert oc.unpack_word(w, len(s)) == s.upper()
        assert int(on.pack_rows(np.frombuffer(s, np.uint8).reshape(1, -1))[0][0]) == w
    for v in G["codec_invalid"]:
        assert oc.pack_word(v["ascii"].encode())[1] is True
        assert bool(on.pack_rows(np.frombuffer(v["ascii"].encode(), np.uint8).reshape(1, -1))[1][0])


def test_generators_and_splitmix_golden():
    for x, v in G["splitmix64"].items():
        assert oc.splitmix64(int(x)) == int(v, 16) == int(on.splitmix64(U64(int(x))))
    for name, g in G["generators"].items():
        if name == "ascii":
            rows = oc.generate_ascii(g["first_row"], g["n_rows"], g["len"], g["dirty_ppm"], g["lower_ppm"], g["seed"])
            assert rows.tobytes().decode() == g["rows"]
            continue
        want = recs_from_hex(g["records"])
        for gen in (oc.generate_records, on.generate_records):
            assert np.array_equal(gen(g["first"], g["n"], g["bc"], g["umi"], g["mode"], g["param"], g["seed"]), want), name


def test_dirty_batch_and_table_golden():
    d = G["dirty_64"]
    recs = recs_from_hex(d["records"])
    ob, ou, of, _ = oc.unpack_records(recs, 16, 12, 1)
    assert ob.tobytes().decode() == d["bc_ascii"] and ou.tobytes().decode() == d["umi_ascii"] and of.tolist() == d["flags"]
    assert oc.reduce_records(recs, 16, 12) == d["reduce"] == on.reduce_records(recs, 16, 12)
    w = G["whitelist_3000"]
    g = w["gen"]
    recs = oc.generate_records(0, g["n"], g["bc"], g["umi"], g["mode"], g["param"], g["seed"])
    rows, pairs = oc.barcode_table(recs)
    assert [[int(r["barcode"]), int(r["n_records"]), int(r["n_distinct_umi"])] for r in rows] == w["rows"]
    assert pairs == w["n_distinct_pairs"]


# ------------------------------------------------------------------ GPU: the CUDA path
@pytest.fixture(scope="module")
def ctx():
    c = ibu.GpuContext(0, chunk_records=4096, n_slots=3)  # 3 chunks for the 10 000-record file
    yield c
    c.close()


@pytest.mark.gpu
def test_gpu_kat_file(ctx):
    reader = ibu.MmapReader(KAT)
    chunks = []
    got = reader.process_gpu(ctx, on_chunk=lambda s, n, r: chunks.append((s, n)))
    assert got == G["ref_kat_10000"]["reduce"] and got.count_sum == G["ref_kat_10000"]["count_sum"]
    assert chunks == [(0, 4096), (4096, 4096), (8192, 1808)]
    hdr, dev = ibu.load_to_device(ctx, KAT)
    assert np.array_equal(dev.to_host(), ibu.load_to_vec(KAT)[1])
    dev.free()


@pytest.mark.gpu
def test_gpu_codec_golden(ctx):
    by_len = {}
    for v in G["codec"]:
        by_len.setdefault(len(v["ascii"]), []).append(v)
    for length, vs in by_len.items():
        rows = np.frombuffer("".join(v["ascii"] for v in vs).encode(), np.uint8).reshape(len(vs), length)
        words = np.array([int(v["word"], 16) for v in vs], U64)
        # pack the golden rows (as barcode and as umi), then unpack the golden words
        back, res = ctx.pack_host(rows, rows)
        assert np.array_equal(back["barcode"], words) and np.array_equal(back["umi"], words) and res["n_bad_records"] == 0
        recs = ibu.records(len(vs))
        recs["barcode"], recs["umi"] = words, words
        gb, gu, _ = ctx.unpack_host(recs, length, length)
        upper = np.frombuffer("".join(v["ascii"].upper() for v in vs).encode(), np.uint8).reshape(len(vs), length)
        assert np.array_equal(gb, upper) and np.array_equal(gu, upper)
    for v in G["codec_invalid"]:
        rows = np.frombuffer(v["ascii"].encode(), np.uint8).reshape(1, -1)
        _, res = ctx.pack_host(rows, rows)
        assert res["n_bad_records"] == 1 and res["n_bad_barcode"] == 1 and res["n_bad_umi"] == 1


@pytest.mark.gpu
def test_gpu_dirty_batch_table_and_generators(ctx):
    d = G["dirty_64"]
    recs = recs_from_hex(d["records"])
    flags = np.zeros(len(recs), np.uint8)
    gb, gu, res = ctx.unpack_host(recs, 16, 12, flags_out=flags)
    assert gb.tobytes().decode() == d["bc_ascii"] and gu.tobytes().decode() == d["umi_ascii"]
    assert flags.tolist() == d["flags"] and res == d["reduce"]
    # device generators reproduce the committed records, and the table of the whitelist set
    for name, g in G["generators"].items():
        if name == "ascii":
            dptr = ctx.malloc(g["n_rows"] * g["len"])
            ctx.generate_ascii_async(dptr, g["first_row"], g["n_rows"], g["len"], g["dirty_ppm"], g["lower_ppm"], g["seed"])
            ctx.synchronize()
            out = np.zeros(g["n_rows"] * g["len"], np.uint8)
            ctx.d2h(out, dptr)
            ctx.free(dptr)
            assert out.tobytes().decode() == g["rows"]
            continue
        dptr = ctx.malloc(24 * g["n"])
        ctx.generate_records_async(dptr, g["first"], g["n"], g["bc"], g["umi"], g["mode"], g["param"], g["seed"])
        ctx.synchronize()
        out = ibu.records(g["n"])
        ctx.d2h(out, dptr)
        ctx.free(dptr)
        assert np.array_equal(out, recs_from_hex(g["records"])), name
    w = G["whitelist_3000"]
    g = w["gen"]
    dptr = ctx.malloc(24 * g["n"])
    ctx.generate_records_async(dptr, 0, g["n"], g["bc"], g["umi"], g["mode"], g["param"], g["seed"])
    rows, info = ctx.barcode_count(dptr, g["n"])
    ctx.free(dptr)
    assert [[int(r["barcode"]), int(r["n_records"]), int(r["n_distinct_umi"])] for r in rows] == w["rows"]
    assert info["n_distinct_pairs"] == w["n_distinct_pairs"]
