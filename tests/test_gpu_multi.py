"""Real multi-GPU path (NCCL over NVLink): needs >= 2 visible B200s, skipped otherwise.  One
process per GPU; records sharded by contiguous range; results merged by ibu_b200.distributed."""
import os
import socket

import numpy as np
import pytest

import ibu_b200 as ibu
from oracle import oracle_c as oc
from oracle import oracle_np as on

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_dir):
    import torch
    import torch.distributed as dist

    from ibu_b200 import distributed as ibd

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        ctx = ibu.GpuContext(rank)
        s, e = ibd.my_shard(n)
        recs = torch.empty((e - s) * 24, dtype=torch.uint8, device=dev)
        # unsorted whitelist data with heavy duplication, generated per shard on its own GPU
        ctx.generate_records_async(recs, s, e - s, 16, 12, ibu.GEN_WHITELIST, (32 << 32) | 3000, 35)
        res = torch.zeros(8, dtype=torch.int64, device=dev)
        ctx.validate_reduce_async(recs, e - s, 16, 12, res)
        ctx.synchronize()
        merged = ibd.merge_results(ctx.read_result(res), device=dev)
        table = ibd.exact_barcode_table(ctx, recs, e - s, dev)
        np.save(os.path.join(out_dir, f"table{rank}.npy"), table)
        np.save(os.path.join(out_dir, f"res{rank}.npy"), np.array([merged[k] for k in ibd._FIELDS], np.uint64))
        ctx.close()
    finally:
        dist.destroy_process_group()


def test_sharded_reduce_and_exact_table_over_nccl(tmp_path):
    world = ibu.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(world, 4)
    import torch.multiprocessing as mp

    from ibu_b200 import distributed as ibd

    n = 3_000_017
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    recs = oc.generate_records(0, n, 16, 12, 3, (32 << 32) | 3000, 35)
    want_red, want_table = oc.reduce_records(recs, 16, 12), on.barcode_table(recs)
    for r in range(world):
        got = dict(zip(ibd._FIELDS, map(int, np.load(tmp_path / f"res{r}.npy"))))
        assert got == want_red
        assert np.array_equal(np.load(tmp_path / f"table{r}.npy"), want_table)
