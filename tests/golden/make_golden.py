#!/usr/bin/env python
"""make_golden.py — writes the committed golden fixtures of tests/golden/.

The reference (noamteyssier/ibu) is a Rust crate and there is no Rust toolchain in the build
image, so no fixture here is an output of the reference binary.  They are produced by the numpy
oracle (oracle/oracle_np.py), cross-checked against the C++ oracle before anything is written,
and anchored on the reference's own known-answer tests wherever one exists:

  ref_kat_10000.ibu   the file of the reference's process_parallel test (src/io/mmap.rs:454-481):
                      Header::new(16, 12) + records (i, 2i, 3i), i < 10 000 — 240 032 bytes
                      (32 + 24 N, src/io/writer.rs:645,673).  Expected count 10 000, sum 299 970 000.
  golden.json         header bytes (src/constructs/header.rs:44-61, 84-93, 373-378), the KAT's
                      reductions, 2-bit codec vectors (record.rs:19-27 + bitnuc's LSB-first order:
                      as_2bit("ACGT") = 0xE4 — PARITY UNPINNED), validation
                      flags, a per-barcode table, and the first records / rows of every synthetic
                      generator mode (pins the counter-based generators shared by oracle and device).

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import oracle_c as oc  # noqa: E402
from oracle import oracle_np as on  # noqa: E402

U64 = np.uint64


def hx(a) -> str:
    return np.ascontiguousarray(a).tobytes().hex()


def main(out_dir: str = HERE):
    out = {}
    # ---- the reference's KAT file ----------------------------------------------------------
    i = np.arange(10_000, dtype=U64)
    recs = np.zeros(10_000, on.RECORD_DTYPE)
    recs["barcode"], recs["umi"], recs["index"] = i, 2 * i, 3 * i
    blob = on.file_bytes(16, 12, recs)
    assert len(blob) == 32 + 24 * 10_000
    with open(os.path.join(out_dir, "ref_kat_10000.ibu"), "wb") as f:
        f.write(blob)
    red = on.reduce_records(recs, 16, 12)
    assert red == oc.reduce_records(recs, 16, 12)
    assert red["n_records"] == 10_000
    assert (red["sum_barcode"] + red["sum_umi"] + red["sum_index"]) % 2**64 == 299_970_000  # mmap.rs:478-480
    out["ref_kat_10000"] = {"file": "ref_kat_10000.ibu", "bytes": len(blob), "reduce": red, "count_sum": 299_970_000}

    # ---- header bytes ------------------------------------------------------------------------
    out["headers"] = {f"{bc}_{umi}_{int(srt)}": on.header_bytes(bc, umi, srt).hex()
                      for bc, umi, srt in [(16, 12, False), (16, 12, True), (1, 1, False), (32, 32, False)]}
    for k, v in out["headers"].items():
        bc, umi, srt = map(int, k.split("_"))
        assert bytes(oc.header_new(bc, umi, bool(srt))).hex() == v
        assert v.startswith("49425521" + "02000000")  # "IBU!" little-endian magic, version 2 (header.rs:5-6, 373-378)

    # ---- codec vectors -----------------------------------------------------------------------
    codec = [("ACGT", 0xE4), ("acgt", 0xE4), ("T" * 16, 0xFFFF_FFFF), ("T" * 12, 0xFF_FFFF), ("T" * 32, 2**64 - 1),
             ("A" * 7, 0), ("C", 1), ("G", 2), ("GATTACA", None), ("AcGtAcGtAcGtAcGt", None)]
    rng = np.random.default_rng(20240601)
    for length in (1, 5, 10, 12, 15, 16, 20, 31, 32):
        for _ in range(4):
            codec.append(("".join("ACGT"[k] for k in rng.integers(0, 4, length)), None))
    vectors = []
    for s, want in codec:
        w, bad = oc.pack_word(s.encode())
        rows = np.frombuffer(s.encode(), np.uint8).reshape(1, -1)
        w_np, bad_np = on.pack_rows(rows)
        assert not bad and not bad_np[0] and int(w_np[0]) == w and (want is None or w == want), s
        assert oc.unpack_word(w, len(s)) == s.upper().encode() == on.unpack_words(np.array([w], U64), len(s))[0].tobytes()
        vectors.append({"ascii": s, "word": f"{w:#018x}"})
    out["codec"] = vectors
    out["codec_invalid"] = [{"ascii": s, "bad": True} for s in ("ACGN", "ACG-", "NNNN", "ACGU", "acgn", "AC GT")]
    for v in out["codec_invalid"]:
        assert oc.pack_word(v["ascii"].encode())[1] is True

    # ---- validation of dirty records + unpack of a small batch ---------------------------------
    dirty = on.generate_records(0, 64, 16, 12, 1, 250_000, 11)  # GEN_DIRTY, 25 % of the records
    assert np.array_equal(dirty, oc.generate_records(0, 64, 16, 12, 1, 250_000, 11))
    ob, ou, of, ores = oc.unpack_records(dirty, 16, 12, 1)
    assert np.array_equal(ob, on.unpack_words(dirty["barcode"], 16)) and np.array_equal(ou, on.unpack_words(dirty["umi"], 12))
    out["dirty_64"] = {"records": hx(dirty), "bc_ascii": ob.tobytes().decode(), "umi_ascii": ou.tobytes().decode(),
                       "flags": of.tolist(), "reduce": on.reduce_records(dirty, 16, 12)}
    assert out["dirty_64"]["reduce"] == oc.reduce_records(dirty, 16, 12) and sum(f != 0 for f in of) == ores["n_bad_records"] > 0

    # ---- per-barcode table ---------------------------------------------------------------------
    wl = on.generate_records(0, 3000, 16, 12, 3, (8 << 32) | 40, 5)  # GEN_WHITELIST: 40 barcodes, 8 UMIs each
    assert np.array_equal(wl, oc.generate_records(0, 3000, 16, 12, 3, (8 << 32) | 40, 5))
    table = on.barcode_table(wl)
    t_c, pairs = oc.barcode_table(wl)
    assert np.array_equal(table, t_c) and len(table) == 40 and int(table["n_records"].sum()) == 3000
    out["whitelist_3000"] = {"gen": {"n": 3000, "bc": 16, "umi": 12, "mode": 3, "param": (8 << 32) | 40, "seed": 5},
                             "rows": [[int(r["barcode"]), int(r["n_records"]), int(r["n_distinct_umi"])] for r in table],
                             "n_distinct_pairs": int(pairs)}

    # ---- generators ------------------------------------------------------------------------------
    gens = {}
    for name, mode, param in [("clean", 0, 0), ("dirty", 1, 500_000), ("pattern", 2, 0), ("whitelist", 3, (16 << 32) | 100),
                              ("sorted", 4, (5 << 32) | 1000)]:
        a = on.generate_records(123_456_789, 8, 16, 12, mode, param, 42)
        assert np.array_equal(a, oc.generate_records(123_456_789, 8, 16, 12, mode, param, 42)), name
        gens[name] = {"first": 123_456_789, "n": 8, "bc": 16, "umi": 12, "mode": mode, "param": param, "seed": 42, "records": hx(a)}
    rows = on.generate_ascii(77, 6, 20, 300_000, 300_000, 9)
    assert np.array_equal(rows, oc.generate_ascii(77, 6, 20, 300_000, 300_000, 9))
    gens["ascii"] = {"first_row": 77, "n_rows": 6, "len": 20, "dirty_ppm": 300_000, "lower_ppm": 300_000, "seed": 9,
                     "rows": rows.tobytes().decode()}
    out["generators"] = gens
    out["splitmix64"] = {str(x): f"{int(on.splitmix64(U64(x))):#018x}" for x in (0, 1, 42, 2**63)}
    for x, v in out["splitmix64"].items():
        assert int(v, 16) == oc.splitmix64(int(x))
    assert out["splitmix64"]["0"] == "0xe220a8397b1dcdaf"  # the published first output of splitmix64 seeded with 0

    with open(os.path.join(out_dir, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote", os.path.join(out_dir, "ref_kat_10000.ibu"), "and golden.json")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else HERE)
