"""getclip of ONE indexed BAM on several GPUs of one node: one process per GPU (torchrun), every rank loads only its shard.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        -m seeksv_b200.mgpu getclip -o tumor tumor.sort.bam            # coordinate-range shards (default)
    python -m torch.distributed.run ... -m seeksv_b200.mgpu getclip --by chromosome -o tumor tumor.sort.bam

Writes the reference's four files (prefix.clip.gz, .clip.fq.gz, .unmapped_1.fq.gz, .unmapped_2.fq.gz), byte-identical after
decompression to `seeksv getclip` on the whole file (tests/test_gpu_parity.py). The options -t -q -s -o mean what they mean
for `seeksv getclip` (seeksv.cpp:128-155). Needs tumor.sort.bam.bai (the index the reference's getsv requires anyway).
The sharding rules are in seeksv_b200/sharding.py; without torchrun this runs as a single rank.
"""
from __future__ import annotations

import argparse
import os
import sys


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="seeksv_b200.mgpu")
    ap.add_argument("command", choices=["getclip"])
    ap.add_argument("-t", type=float, default=0.9, dest="match_rate")
    ap.add_argument("-q", type=int, default=1, dest="min_mapq")
    ap.add_argument("-s", action="store_true", dest="save_low_quality")
    ap.add_argument("-o", default="output", dest="prefix")
    ap.add_argument("--by", choices=["range", "chromosome"], default="range")
    ap.add_argument("--bai", default=None)
    ap.add_argument("bam")
    a = ap.parse_args(argv)

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        keep = os.dup(1)           # NCCL announces itself on stdout
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        finally:
            os.dup2(keep, 1)
            os.close(keep)
        os.environ.setdefault("SEEKSV_B200_THREADS", str(max(2, (os.cpu_count() or 2) // world)))
    from . import lib, sharding
    ctx = lib.Context(local)
    kw = dict(match_rate=a.match_rate, min_mapq=a.min_mapq, save_low_quality=a.save_low_quality)
    try:
        if a.by == "chromosome":
            worker = sharding.open_ref_shard(ctx, a.bam, rank, world, a.bai, **kw)
            texts = sharding.sharded_getclip(worker, dist)
        else:
            probe = lib.Bam.open_refs(ctx, a.bam, 0, 0, a.bai)
            n_ref = len(probe.ref_names)
            probe.close()
            steps = 1
            while True:   # every rank computes the same plan from the index; a context without a mapped record is widened
                plan = sharding.plan_range_shards(a.bam, a.bai, n_ref, world, context_steps=steps)[rank]
                worker = sharding.RangeShardWorker(ctx, a.bam, plan, **kw)
                ok = worker.context_has_mapped_record()
                oks = sharding.all_gather_objects(ok, dist)
                if all(oks) or steps > 64:
                    break
                worker.close()
                steps *= 2
            texts = sharding.sharded_getclip_ranges(worker, dist)
        worker.close()
        if rank == 0:
            for ext, t in zip((".clip.gz", ".clip.fq.gz", ".unmapped_1.fq.gz", ".unmapped_2.fq.gz"), texts):
                lib.write_gz(a.prefix + ext, t.encode("latin-1"))
            print("[GetSClipReads] finished!", file=sys.stderr)
    finally:
        ctx.close()
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
