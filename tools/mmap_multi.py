#!/usr/bin/env python
"""mmap_multi.py — end-to-end ingest FROM AN MMAP'ED FILE across GPUs (configs[3] shape; run under
torchrun).  Rank 0 writes an .ibu file of the reference's example pattern into /dev/shm through
the Writer; every rank opens it with MmapReader and processes its contiguous shard
(mmap.rs:297-307 rule) with ibu_gpu_process_mmap: pread -> non-temporal copy -> pinned chunk ->
H2D -> K1, counters all-reduced.  The host cores are shared by the ranks (copy threads =
cores / world), so this measures the box's staging ceiling, next to the per-rank link rate.
Rank 0 prints JSON lines; results are checked against the closed form."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import ibu_b200 as ibu  # noqa: E402
from ibu_b200 import distributed as ibd  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=400_000_000)
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = args.records
    path = os.path.join(args.dir, f"ibu_mmap_multi_{n}.ibu")
    cores = os.cpu_count() or 1
    ctx = ibu.GpuContext(local, copy_threads=max(2, cores // world))

    if rank == 0:  # the file: generated on the device, written through the Writer's device path
        t0 = time.perf_counter()
        step = 64 << 20
        buf = torch.empty(min(n, step) * 24, dtype=torch.uint8, device=dev)
        with ibu.Writer(path, ibu.Header(16, 12)) as w:
            for s in range(0, n, step):
                cnt = min(step, n - s)
                ctx.generate_records_async(buf, s, cnt, 16, 12, ibu.GEN_PATTERN, 0, 0)
                ctx.synchronize()
                w.write_device(ctx, buf, cnt)
        del buf
        print(json.dumps(dict(stage="write file (device generator -> Writer device path)", records=n,
                              gb=24 * n / 1e9, sec=time.perf_counter() - t0)), flush=True)
    dist.barrier()

    reader = ibu.MmapReader(path)
    assert reader.len() == n
    s, e = ibd.my_shard(n)
    best, merged = 1e30, None
    for _ in range(args.reps):
        dist.barrier()
        t0 = time.perf_counter()
        merged = ibd.merge_results(reader.process_gpu(ctx, s, e), device=dev)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        best = min(best, float(dt[0]))
    ok = merged["n_records"] == n and merged["sum_index"] == (n * (n - 1) // 2) % 2**64 and merged["n_bad_records"] == 0
    if rank == 0:
        print(json.dumps(dict(stage="mmap ingest + validate + count (staged), all ranks", world=world, records=n,
                              copy_threads_per_rank=max(2, cores // world), host_cores=cores, sec=best,
                              grec_s=n / best / 1e9, gb_s_total=24 * n / best / 1e9, gb_s_per_gpu=24 * n / best / 1e9 / world,
                              closed_form_ok=bool(ok))), flush=True)
        from oracle import oracle_c as oc
        m = oc.MmapReader(path)
        t0 = time.perf_counter()
        want, _ = m.process_parallel_reduce(0)
        t_cpu = time.perf_counter() - t0
        print(json.dumps(dict(stage="cpu oracle process_parallel on the same file, all host threads", cores=oc.num_cpus(),
                              sec=t_cpu, grec_s=n / t_cpu / 1e9, gb_s=24 * n / t_cpu / 1e9,
                              matches_gpu=bool(want == dict(merged)))), flush=True)
    dist.barrier()
    reader.close()
    ctx.close()
    if rank == 0:
        os.unlink(path)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
