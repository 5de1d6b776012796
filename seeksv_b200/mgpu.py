"""getclip / getsv / somatic of ONE indexed BAM on several GPUs of one node: one process per GPU (torchrun), every rank loads only
its shard of the genome.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        -m seeksv_b200.mgpu getclip -o tumor tumor.sort.bam             # coordinate-range shards (default) or --by chromosome
    python -m torch.distributed.run ... -m seeksv_b200.mgpu getsv tumor.clip.bam tumor.sort.bam tumor.clip.gz out.sv.txt out.unmapped.fq.gz
    python -m torch.distributed.run ... -m seeksv_b200.mgpu somatic normal.sort.bam normal.clip.gz tumor.sv.txt tumor.somatic.sv.txt

The files are the reference's, byte-identical after decompression to the single-process commands (tests/test_gpu_parity.py); the
options mean what they mean for `seeksv` (seeksv.cpp:128-410; getsv / somatic take the single-process command's option letters
after `--`). Needs the .bai (the index the reference's getsv requires anyway). The sharding rules are in seeksv_b200/sharding.py:
getclip clusters every breakpoint key on exactly one rank (halo + key ownership), getsv / somatic run the per-record passes on every
rank's own records and add the statistics, pair counts and depths up with NCCL collectives on device tensors; rank 0 does the
order-dependent junction bookkeeping and writes the files. Without torchrun this runs as a single rank.
"""
from __future__ import annotations

import argparse
import os
import struct
import sys
import tempfile


def _init_dist():
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            torch.cuda.set_device(local)
            keep = os.dup(1)           # NCCL announces itself on stdout
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            finally:
                os.dup2(keep, 1)
                os.close(keep)
        os.environ.setdefault("SEEKSV_B200_THREADS", str(max(2, (os.cpu_count() or 2) // world)))
    return rank, world, local, dist


def open_range_worker(ctx, dist, bam_path, bai, rank, world, **kw):
    """this rank's coordinate-range shard, loaded; every rank computes the same plan from the index, a context without a
    mapped-branch record is widened"""
    from . import lib, sharding
    probe = lib.Bam.open_refs(ctx, bam_path, 0, 0, bai)
    n_ref = len(probe.ref_names)
    probe.close()
    steps = 1
    while True:
        plan = sharding.plan_range_shards(bam_path, bai, n_ref, world, context_steps=steps)[rank]
        worker = sharding.RangeShardWorker(ctx, bam_path, plan, **kw)
        oks = sharding.all_gather_objects(worker.context_has_mapped_record(), dist)
        if all(oks) or steps > 64:
            return worker
        worker.close()
        steps *= 2


def _gather_bytes(parts, dist, device):
    """every rank's list of byte strings on rank 0 (device tensors over NCCL point-to-point; sizes first)"""
    from . import sharding
    if dist is None:
        return [parts]
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = sharding._all_gather_i64([len(p) for p in parts], dist, device)
    if rank != 0:
        for p in parts:
            if len(p):
                dist.send(torch.frombuffer(bytearray(p), dtype=torch.uint8).to(device), 0)
        return None
    out = [parts]
    for src in range(1, world):
        got = []
        for n in sizes[src]:
            if n:
                t = torch.empty(n, dtype=torch.uint8, device=device)
                dist.recv(t, src)
                got.append(t.cpu().numpy().tobytes())
            else:
                got.append(b"")
        out.append(got)
    return out


def run_getclip(ctx, dist, device, a):
    from . import lib, sharding
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    kw = dict(match_rate=a.match_rate, min_mapq=a.min_mapq, save_low_quality=a.save_low_quality)
    if a.by == "chromosome":
        worker = sharding.open_ref_shard(ctx, a.bam, rank, world, a.bai, **kw)
        lasts = sharding.all_gather_objects(worker.last_mapped_tid(), dist)
        cl = worker.bam.getclip_handle(prev_tid=sharding.prev_tids(lasts)[rank], export_unmapped=True, **kw)
    else:
        worker = open_range_worker(ctx, dist, a.bam, a.bai, rank, world, **kw)
        p = worker.plan
        cl = None
        if worker.bam is not None:
            cl = worker.bam.getclip_handle(prev_tid=p.prev_tid, export_unmapped=True, key_range=(p.key_lo, p.key_hi), halo_bytes=p.halo_bytes, **kw)
    mine = [cl.text(0), cl.text(1), cl.unmapped_records()] if cl is not None else [b"", b"", b""]
    every = _gather_bytes(mine, dist, device)
    names, lens = (worker.bam.ref_names, worker.bam.ref_lens) if worker.bam is not None else ([], [])
    if rank == 0:
        if a.by == "chromosome":
            clip, fq = b"".join(e[0] for e in every), b"".join(e[1] for e in every)
        else:
            clip, fq = sharding.merge_range_texts_fast([(e[0], e[1]) for e in every])
        records = b"".join(e[2] for e in every)
        u1 = u2 = b""
        if records:   # mates are paired by name across the whole file: the shards' unmapped-branch records, in file order, as one stream
            mini = lib.Bam.from_host(ctx, records, 0, len(names))
            mini.set_refs(names, lens)
            cu = mini.getclip_handle(unmapped_only=True, **kw)
            u1, u2 = cu.text(2), cu.text(3)
            cu.close()
            mini.close()
        for ext, t in zip((".clip.gz", ".clip.fq.gz", ".unmapped_1.fq.gz", ".unmapped_2.fq.gz"), (clip, fq, u1, u2)):
            lib.write_gz(a.prefix + ext, t)
        print("[GetSClipReads] finished!", file=sys.stderr)
    if cl is not None:
        cl.close()
    worker.close()


def shard_for_passes(ctx, dist, bam_path, bai, rank, world):
    """this rank's own records for the additive getsv / somatic passes: a coordinate-range shard, the halo left out"""
    from . import sharding
    worker = open_range_worker(ctx, dist, bam_path, bai, rank, world)
    if worker.bam is not None:
        worker.bam.set_own_offset(worker.plan.halo_bytes)
    return worker


class _NoRecords:
    """a rank without records (more ranks than cut points): contributes zeros"""

    def __init__(self, device):
        self.device = device

    def insert_partial(self, min_mapq, take):
        return (0, 0, 0, 0)

    def insert_sq(self, min_mapq, take, mean):
        return 0

    def pairs_depth(self, min_mapq, mean, dev, times, junctions, windows):
        import torch
        return torch.zeros(max(len(junctions) + sum(w[2] - w[1] + 1 for w in windows), 1), dtype=torch.int32, device=self.device)


def _write_results(path, n, mean, dev, counts, depth):
    with open(path, "wb") as f:
        f.write(b"SVBR" + struct.pack("<3i", int(min(n, 2 ** 31 - 1)), mean, dev) + struct.pack("<i", len(counts)) + struct.pack("<%di" % len(counts), *counts))
        f.write(struct.pack("<i", len(depth)))
        import numpy as np
        f.write(np.asarray(depth, dtype="<i4").tobytes())


def _opt(argv, letter, default, cast=int):
    v = default
    for i, x in enumerate(argv):
        if x == "-" + letter and i + 1 < len(argv):
            v = cast(argv[i + 1])
    return v


def run_getsv(ctx, dist, device, a):
    """a.rest = the arguments of `seeksv getsv` (options and the five files)"""
    from . import lib, sharding
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    files = [x for i, x in enumerate(a.rest) if not x.startswith("-") and (i == 0 or a.rest[i - 1] not in ("-F", "-B", "-t", "-l", "-q", "-Q", "-w", "-n", "-a", "-b", "-d", "-e", "-m", "-i", "-R", "-f", "-T", "-L"))]
    if len(files) != 5 or "-F" in a.rest or "-B" in a.rest:
        raise SystemExit("mgpu getsv: five files, no -F / -B (use the single-process command for those)")
    clip_aln, bam_path, clip_gz = files[0], files[1], files[2]
    min_mapq, pairs_used, flank, flank_len = _opt(a.rest, "q", 20), _opt(a.rest, "n", 5000000), _opt(a.rest, "l", 50), _opt(a.rest, "L", 200)
    with_depth = "-D" not in a.rest
    worker = shard_for_passes(ctx, dist, bam_path, a.bai, rank, world)
    names, lens = sharding.all_gather_objects((worker.bam.ref_names, worker.bam.ref_lens) if worker.bam is not None else None, dist)[0]
    gw = sharding.GpuShardWorker(worker.bam, device) if worker.bam is not None else _NoRecords(device)
    juncs, wins = lib.plan_getsv(clip_aln, clip_gz, names, lens, flank, flank_len)      # host, the same on every rank
    if not with_depth:
        wins = []
    n = mean = dev = 0
    if pairs_used >= 100000:
        n, mean, dev = sharding.sharded_insert_stats(gw, dist, device, min_mapq, pairs_used)
    else:
        juncs = []
    t = sharding.sharded_pairs_depth(gw, dist, min_mapq, mean, dev, 4, juncs, wins).cpu().numpy()
    rc = 0
    if rank == 0:
        with tempfile.NamedTemporaryFile(suffix=".svbr", delete=False) as tf:
            res = tf.name
        _write_results(res, n, mean, dev, t[:len(juncs)].tolist() if pairs_used >= 100000 else [], t[len(juncs):len(juncs) + sum(w[2] - w[1] + 1 for w in wins)])
        os.environ["SEEKSV_B200_SHARD_RESULTS"] = res
        try:
            rc = lib.run_cli(["getsv"] + list(a.rest))
        finally:
            del os.environ["SEEKSV_B200_SHARD_RESULTS"]
            os.unlink(res)
    worker.close()
    return rc


def run_somatic(ctx, dist, device, a):
    """a.rest = the arguments of `seeksv somatic` (options and the four files)"""
    from . import lib, sharding
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    files = [x for i, x in enumerate(a.rest) if not x.startswith("-") and (i == 0 or a.rest[i - 1] not in ("-t", "-q", "-l", "-m", "-n"))]
    if len(files) != 4:
        raise SystemExit("mgpu somatic: four files")
    normal_bam, normal_clip, tumor_sv = files[0], files[1], files[2]
    rate, min_mapq, offset = _opt(a.rest, "t", 0.9, float), _opt(a.rest, "q", 20), _opt(a.rest, "l", 30)
    min_len, pairs_used = _opt(a.rest, "m", 10), _opt(a.rest, "n", 5000000)
    worker = shard_for_passes(ctx, dist, normal_bam, a.bai, rank, world)
    names, _ = sharding.all_gather_objects((worker.bam.ref_names, worker.bam.ref_lens) if worker.bam is not None else None, dist)[0]
    gw = sharding.GpuShardWorker(worker.bam, device) if worker.bam is not None else _NoRecords(device)
    n = mean = dev = 0
    if pairs_used >= 100000:
        n, mean, dev = sharding.sharded_insert_stats(gw, dist, device, min_mapq, pairs_used)
    juncs = lib.plan_somatic(normal_clip, tumor_sv, names, rate, offset, min_len, mean)
    t = sharding.sharded_pairs_depth(gw, dist, min_mapq, mean, dev, 4, juncs, []).cpu().numpy()
    rc = 0
    if rank == 0:
        with tempfile.NamedTemporaryFile(suffix=".svbr", delete=False) as tf:
            res = tf.name
        _write_results(res, n, mean, dev, t[:len(juncs)].tolist(), [])
        os.environ["SEEKSV_B200_SHARD_RESULTS"] = res
        try:
            rc = lib.run_cli(["somatic"] + list(a.rest))
        finally:
            del os.environ["SEEKSV_B200_SHARD_RESULTS"]
            os.unlink(res)
    worker.close()
    return rc


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="seeksv_b200.mgpu")
    ap.add_argument("command", choices=["getclip", "getsv", "somatic"])
    ap.add_argument("-t", type=float, default=0.9, dest="match_rate")
    ap.add_argument("-q", type=int, default=1, dest="min_mapq")
    ap.add_argument("-s", action="store_true", dest="save_low_quality")
    ap.add_argument("-o", default="output", dest="prefix")
    ap.add_argument("--by", choices=["range", "chromosome"], default="range")
    ap.add_argument("--bai", default=None)
    ap.add_argument("bam", nargs="?", help="getclip: the BAM; getsv / somatic: `--` and then the command's own arguments")
    argv = list(sys.argv[1:] if argv is None else argv)
    rest = []
    if "--" in argv:
        k = argv.index("--")
        argv, rest = argv[:k], argv[k + 1:]
    a = ap.parse_intermixed_args(argv)
    a.rest = rest if rest else ([a.bam] if a.bam else [])
    rank, world, local, dist = _init_dist()
    from . import lib
    ctx = lib.Context(local)
    device = "cuda:%d" % local
    rc = 0
    try:
        if a.command == "getclip":
            if len(a.rest) != 1:
                raise SystemExit("mgpu getclip: one BAM")
            a.bam = a.rest[0]
            run_getclip(ctx, dist, device, a)
        elif a.command == "getsv":
            rc = run_getsv(ctx, dist, device, a)
        else:
            rc = run_somatic(ctx, dist, device, a)
    finally:
        ctx.close()
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
