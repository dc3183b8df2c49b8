"""The multi-GPU group behind the C ABI (ibu_gpu_group_*): range sharding with the reference's
partition rule (mmap.rs:297-307), host merge of the reduction results (mmap.rs:365-372) and the
exact merge of the per-barcode tables through the exchange of de-duplicated pairs — no
torch.distributed anywhere in the data path.  On a box with one GPU the ranks share it (the whole
protocol still runs: grouping by owner, pulls, weighted count, gather); with two or more the
copies cross NVLink and NCCL is exercised as well."""
import numpy as np
import pytest

import ibu_b200 as ibu
from oracle import oracle_c as oc
from oracle import oracle_np as on

pytestmark = pytest.mark.gpu


def devices(world):
    n = ibu.device_count()
    return [r % n for r in range(world)]


def write(path, recs, bc=16, umi=12):
    with ibu.Writer(path, ibu.Header(bc, umi)) as w:
        w.write_batch(recs)


def skewed(n, seed):
    recs = oc.generate_records(0, n, 16, 12, 5, (64 << 32) | 30_000, seed)  # Zipf-ish barcodes, heavy duplication
    rng = np.random.default_rng(seed)
    recs["umi"][rng.random(n) < 0.002] |= np.uint64(1 << 50)  # words wider than the header allows
    return recs


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("exchange", [ibu.EXCHANGE_P2P, ibu.EXCHANGE_HOST])
def test_group_process_mmap_table_is_exact(tmp_ibu, world, exchange):
    n = 700_003
    recs = skewed(n, 71)
    write(tmp_ibu, recs)
    want_red, _ = oc.MmapReader(tmp_ibu).process_parallel_reduce(0)
    want = on.barcode_table(recs)
    reader = ibu.MmapReader(tmp_ibu)
    with ibu.GpuGroup(devices(world), chunk_records=1 << 16) as g:
        red, out = g.process_mmap(reader, table=True, exchange=exchange, table_mode=ibu.COUNT_PATH_PARTITION)
        assert red == want_red
        assert np.array_equal(out.rows, want)
        assert out.table_info["n_records"] == n and out.table_info["n_distinct_pairs"] == int(want["n_distinct_umi"].sum())
        assert out.shard_records == [ibu.shard_range(n, r, world)[1] - ibu.shard_range(n, r, world)[0] for r in range(world)]
        assert out.timing["exchange"] == exchange and out.timing["pairs_local"] >= out.table_info["n_distinct_pairs"]
        # default paths (small shards take the legacy one), a sub-range, reduction only
        red2, out2 = g.process_mmap(reader, 1000, n - 77, table=True)
        assert red2 == oc.reduce_records(recs[1000:n - 77], 16, 12)
        assert np.array_equal(out2.rows, on.barcode_table(recs[1000:n - 77]))
        red3, out3 = g.process_mmap(reader)
        assert red3 == want_red and out3.rows is None


def test_group_keep_and_resident_barcode_count(tmp_ibu):
    n, world = 400_001, 4
    recs = skewed(n, 72)
    write(tmp_ibu, recs)
    reader = ibu.MmapReader(tmp_ibu)
    with ibu.GpuGroup(devices(world), chunk_records=1 << 16) as g:
        red, out = g.process_mmap(reader, keep=True)
        assert red == oc.reduce_records(recs, 16, 12)
        for r, d in enumerate(out.records):
            s, e = ibu.shard_range(n, r, world)
            assert np.array_equal(d.to_host(), recs[s:e])
        rows, info, timing = g.barcode_count([d.ptr for d in out.records], [len(d) for d in out.records],
                                             mode=ibu.count_lens(16, 12))
        assert np.array_equal(rows, on.barcode_table(recs)) and info["n_records"] == n
        assert timing["total_ms"] > 0
        for d in out.records:
            d.free()


def test_group_sorted_file_and_host_records():
    """A file sorted by Record's Ord: the shards' streaming passes emit their pairs, runs cut by a shard
    boundary meet on their owner."""
    n, world = 300_000, 3
    recs = oc.generate_records(0, n, 16, 12, 4, (7 << 32) | 1000, 0)  # sorted: 1000 records / barcode, 7 / umi
    with ibu.GpuGroup(devices(world), chunk_records=1 << 16) as g:
        red, out = g.process_host(recs, 16, 12, table=True)
        assert red == oc.reduce_records(recs, 16, 12)
        assert np.array_equal(out.rows, on.barcode_table(recs))


def test_group_empty_and_errors(tmp_ibu):
    write(tmp_ibu, oc.generate_records(0, 0, 16, 12, 0, 0, 1))
    reader = ibu.MmapReader(tmp_ibu)
    with ibu.GpuGroup(devices(2)) as g:
        red, out = g.process_mmap(reader, table=True)
        assert red["n_records"] == 0 and len(out.rows) == 0
        with pytest.raises(ibu.InvalidIndex):
            g.process_mmap(reader, 0, 5)
    with pytest.raises(ibu.ArgError):
        ibu.GpuGroup([])
    with pytest.raises(ibu.ArgError):
        ibu.GpuGroup([99])


@pytest.mark.skipif(ibu.device_count() < 2, reason="NCCL needs one distinct GPU per rank")
def test_group_nccl_exchange(tmp_ibu):
    n = 1_000_003
    recs = skewed(n, 73)
    write(tmp_ibu, recs)
    world = min(ibu.device_count(), 8)
    with ibu.GpuGroup(list(range(world)), chunk_records=1 << 17) as g:
        red, out = g.process_mmap(ibu.MmapReader(tmp_ibu), table=True, exchange=ibu.EXCHANGE_NCCL,
                                  table_mode=ibu.COUNT_PATH_PARTITION)
        assert np.array_equal(out.rows, on.barcode_table(recs)) and out.timing["exchange"] == ibu.EXCHANGE_NCCL


def test_group_nccl_refused_on_shared_devices():
    recs = oc.generate_records(0, 10_000, 16, 12, 0, 0, 1)
    with ibu.GpuGroup([0, 0]) as g:
        with pytest.raises(ibu.NcclError):
            g.process_host(recs, 16, 12, table=True, exchange=ibu.EXCHANGE_NCCL)
