"""Randomised (hypothesis) parity of the CUDA path against the oracle: arbitrary lengths, sizes,
seeds and dirt levels, through the C ABI.  Everything is integer/byte work: bit-exact."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import ibu_b200 as ibu
from oracle import oracle_c as oc
from oracle import oracle_np as on

pytestmark = pytest.mark.gpu

_ctx = None


def ctx():
    global _ctx
    if _ctx is None:
        _ctx = ibu.GpuContext(0, chunk_records=1 << 16, n_slots=2)
    return _ctx


SETTINGS = dict(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
lens = st.integers(1, 32)
sizes = st.one_of(st.integers(0, 300), st.integers(0, 70_000))


@settings(**SETTINGS)
@given(bc=lens, umi=lens, n=sizes, seed=st.integers(0, 2**32), ppm=st.sampled_from([0, 1_000, 300_000, 1_000_000]))
def test_unpack_pack_reduce_through_host_buffers(bc, umi, n, seed, ppm):
    c = ctx()
    recs = oc.generate_records(seed % 1000, n, bc, umi, 1, ppm, seed)
    gb, gu, res = c.unpack_host(recs, bc, umi)
    ob, ou, _, ores = oc.unpack_records(recs, bc, umi, 1)
    assert np.array_equal(gb, ob) and np.array_equal(gu, ou)
    assert all(res[k] == ores[k] for k in ("n_records", "n_bad_barcode", "n_bad_umi", "n_bad_records"))
    assert c.process_host(recs, bc, umi) == oc.reduce_records(recs, bc, umi, 1)
    back, pres = c.pack_host(gb, gu, index=np.ascontiguousarray(recs["index"]))
    assert pres["n_bad_records"] == 0
    assert np.array_equal(back["barcode"], recs["barcode"] & np.uint64(on.low_mask(bc)))
    assert np.array_equal(back["umi"], recs["umi"] & np.uint64(on.low_mask(umi)))
    assert np.array_equal(back["index"], recs["index"])


@settings(**SETTINGS)
@given(bc=lens, umi=lens, n=sizes, seed=st.integers(0, 2**32), dirty=st.sampled_from([0, 50_000, 1_000_000]),
       lower=st.sampled_from([0, 500_000]))
def test_pack_dirty_ascii(bc, umi, n, seed, dirty, lower):
    c = ctx()
    b = oc.generate_ascii(seed % 77, n, bc, dirty, lower, seed)
    u = oc.generate_ascii(seed % 77, n, umi, dirty, lower, seed + 1)
    flags = np.zeros(n, np.uint8)
    got, res = c.pack_host(b, u, index_base=seed, flags_out=flags)
    want, wflags, wres = oc.pack_records(b, u, None, seed)
    assert np.array_equal(got, want) and np.array_equal(flags, wflags)
    assert all(res[k] == wres[k] for k in ("n_records", "n_bad_barcode", "n_bad_umi", "n_bad_records"))


@settings(**SETTINGS)
@given(n=sizes, seed=st.integers(0, 2**32), nb=st.integers(1, 5000), us=st.integers(1, 300),
       presort=st.booleans(), weird=st.booleans())
def test_barcode_table_sort_and_pairs(n, seed, nb, us, presort, weird):
    c = ctx()
    recs = oc.generate_records(0, n, 16, 12, 3, (us << 32) | nb, seed)
    if weird and n:  # full 64-bit words, including the all-ones key the hash table reserves
        recs["barcode"][:: max(1, n // 7)] = np.uint64(2**64 - 1)
        recs["umi"][:: max(1, n // 5)] = np.uint64(2**64 - 1)
    order = np.lexsort((recs["index"], recs["umi"], recs["barcode"]))
    if presort:
        recs = recs[order]
    d = c.malloc(max(24 * n, 32))
    o = c.malloc(max(24 * n, 32))
    if n:
        c.h2d(d, recs)
    rows, info = c.barcode_count(d, n)
    assert np.array_equal(rows, on.barcode_table(recs))
    if n:
        if presort:
            assert info["input_was_sorted"]
        c.sort_records(d, n, o)
        got = np.zeros(n, ibu.RECORD_DTYPE)
        c.d2h(got, o)
        assert np.array_equal(got, recs[np.lexsort((recs["index"], recs["umi"], recs["barcode"]))])
    c.free(d), c.free(o)


@settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))
@given(n=st.integers(1, 400_000), seed=st.integers(0, 2**32), nb=st.integers(1, 200_000), us=st.integers(1, 1 << 24),
       gen=st.sampled_from([3, 5, 2, 1]), weird=st.sampled_from([0, 0, 1, 2]), hint=st.booleans(),
       chunk=st.sampled_from([1 << 12, 1 << 14, 1 << 16]), ranks=st.sampled_from([1, 2, 3]))
def test_partition_path_everywhere(n, seed, nb, us, gen, weird, hint, chunk, ranks):
    """The partition path (forced, whatever the sample would choose) over random shapes — whitelist,
    Zipf, the example pattern, dirty random records; a few records with words wider than the key
    layout or the all-ones key — through its three routes: a resident array, the chunked ingest
    pipeline (records arrive piece by piece on several streams) and the multi-rank group (pair mode +
    weighted owner count).  All three must equal the oracle's table."""
    param = {3: (us << 32) | nb, 5: (us << 32) | nb, 2: 0, 1: 20_000}[gen]
    recs = oc.generate_records(seed % 997, n, 16, 12, gen, param, seed)
    if weird and n:
        recs["umi"][:: max(1, n // 11)] |= np.uint64(1 << 55)  # wider than umi12: the side list
        if weird == 2:
            recs["barcode"][:: max(1, n // 3)] = np.uint64(2**64 - 1)
    want = on.barcode_table(recs)
    mode = 2 | ibu.COUNT_PATH_PARTITION | (ibu.count_lens(16, 12) if hint else 0)
    c = ctx()
    d = c.malloc(max(24 * n, 32))
    c.h2d(d, recs)
    rows, info = c.barcode_count(d, n, mode)
    c.free(d)
    assert np.array_equal(rows, want) and info["n_distinct_pairs"] == int(want["n_distinct_umi"].sum())
    with ibu.GpuContext(0, chunk_records=chunk, n_slots=3) as c2:
        red, out = c2.process_host_ops(recs, 16, 12, table=True, table_mode=mode)
        assert red == oc.reduce_records(recs, 16, 12, 1) and np.array_equal(out.rows, want)
    with ibu.GpuGroup([0] * ranks, chunk_records=chunk) as g:
        red, out = g.process_host(recs, 16, 12, table=True, table_mode=mode)
        assert red == oc.reduce_records(recs, 16, 12, 1) and np.array_equal(out.rows, want)


class _Env:
    """An environment variable for the duration of a block (the library reads these per call)."""

    def __init__(self, key, value):
        self.key, self.value = key, value

    def __enter__(self):
        import os

        self.old = os.environ.get(self.key)
        os.environ[self.key] = self.value

    def __exit__(self, *exc):
        import os

        if self.old is None:
            os.environ.pop(self.key, None)
        else:
            os.environ[self.key] = self.old


@settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))
@given(n=st.integers(1, 400_000), seed=st.integers(0, 2**32), lens=st.sampled_from([(16, 12), (16, 16), (12, 8), (20, 10), (9, 5)]),
       dup=st.sampled_from([0, 0, 3, 40]), ppm=st.sampled_from([0, 10_000, 200_000]), weird=st.sampled_from([0, 1, 2]),
       hint=st.booleans(), chunk=st.sampled_from([1 << 12, 1 << 16]))
def test_ordered_path_everywhere(n, seed, lens, dup, ppm, weird, hint, chunk):
    """The ordered form of the partition path (about as many barcodes as records: buckets cut by the
    barcode's top bits, sorted in shared memory, rows written in barcode order), forced and strict
    (IBU_B200_K4_ORDERED=2: handing the input to the sort fallback is an error) over random
    barcodes with repeats of a record, of a barcode with other UMIs, words wider than the layout
    (merged from the side list, inside a bucket's range and beyond the last bucket) and the all-ones
    barcode — as a resident array and through the chunked ingest pipeline."""
    bc, umi = lens
    recs = oc.generate_records(seed % 997, n, bc, umi, 1 if ppm else 0, ppm, seed)
    rng = np.random.default_rng(seed)
    if dup:  # repeats: whole records, and barcodes that come back with another UMI
        src = rng.integers(0, n, n // 2)
        recs[n - len(src):] = recs[src]
        again = rng.integers(0, n, n // dup + 1)
        recs["umi"][again] = rng.integers(0, 1 << (2 * umi), len(again), dtype=np.uint64)
    if weird and n:
        recs["umi"][:: max(1, n // 11)] |= np.uint64(1 << (2 * umi + 1)) if 2 * umi + 1 < 64 else np.uint64(0)
        if weird == 2:
            recs["barcode"][:: max(1, n // 3)] = np.uint64(2**64 - 1)
            recs["barcode"][1:: max(1, n // 5)] |= np.uint64(1 << 63)
    want = on.barcode_table(recs)
    mode = 2 | ibu.COUNT_PATH_PARTITION | (ibu.count_lens(bc, umi) if hint else 0)
    c = ctx()
    # strict unless a fifth of the records are wider than the layout (the side list takes an eighth), or the
    # layout has to come from a sample of words like that
    strict = ppm <= 10_000 and (hint or not (ppm or weird))
    with _Env("IBU_B200_K4_ORDERED", "2" if strict else "1"):
        d = c.malloc(max(24 * n, 32))
        c.h2d(d, recs)
        rows, info = c.barcode_count(d, n, mode)
        c.free(d)
        assert np.array_equal(rows, want) and info["n_distinct_pairs"] == int(want["n_distinct_umi"].sum())
        with ibu.GpuContext(0, chunk_records=chunk, n_slots=3) as c2:
            red, out = c2.process_host_ops(recs, bc, umi, table=True, table_mode=mode)
            assert red == oc.reduce_records(recs, bc, umi, 1) and np.array_equal(out.rows, want)
