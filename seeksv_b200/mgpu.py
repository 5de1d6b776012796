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
import time


class _Phases:
    """wall-clock phase marks of one command on this rank (SEEKSV_B200_TIMING=1 prints them to stderr)"""

    def __init__(self, name, rank):
        self.on, self.name, self.rank, self.t, self.rows = bool(os.environ.get("SEEKSV_B200_TIMING")), name, rank, time.perf_counter(), []

    def mark(self, what):
        if self.on:
            now = time.perf_counter()
            self.rows.append("%s %.1f" % (what, 1e3 * (now - self.t)))
            self.t = now

    def done(self):
        if self.on:
            os.write(2, ("[time] mgpu %s rank %d: %s ms\n" % (self.name, self.rank, ", ".join(self.rows))).encode())


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


def assemble_block_files(prefix, lists):
    """lists[r] = the (chromosome, side) blocks rank r wrote with lib.write_range_blocks(prefix + ".part<r>", ...): concatenates the
    block files into prefix.clip.gz / prefix.clip.fq.gz - per chromosome (order of first appearance over the ranks) the '5' blocks of
    all ranks, then the '3' blocks - and removes them."""
    from . import lib
    order, groups = [], {}
    for r, bl in enumerate(lists):
        for b, (chrom, side) in enumerate(bl):
            if chrom not in groups:
                order.append(chrom)
                groups[chrom] = {b"5": [], b"3": []}
            groups[chrom][side].append("%s.part%d.%d" % (prefix, r, b))
    files = [f for chrom in order for side in (b"5", b"3") for f in groups[chrom][side]]
    for ext, pext in ((".clip.gz", ".clip.gz"), (".clip.fq.gz", ".fq.gz")):
        if not files:
            lib.write_gz(prefix + ext, b"")
            continue
        with open(prefix + ext, "wb") as out:
            for f in files:
                with open(f + pext, "rb") as src:
                    n = os.fstat(src.fileno()).st_size
                    off = 0
                    while off < n:
                        off += os.sendfile(out.fileno(), src.fileno(), off, n - off)
                os.unlink(f + pext)


def run_getclip(ctx, dist, device, a):
    from . import lib, sharding
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    kw = dict(match_rate=a.match_rate, min_mapq=a.min_mapq, save_low_quality=a.save_low_quality)
    ph = _Phases("getclip", rank)
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
    ph.mark("open+getclip")
    # Every rank compresses the (chromosome, side) blocks of ITS clip / clip.fq text into gzip files of their own; gzip members
    # concatenate, so rank 0 only has to put the files in order (per chromosome the '5' blocks of all ranks, then the '3' blocks -
    # clip_reads.h:300-345,423-438). No text crosses ranks and nothing is compressed twice. Only the unmapped-branch records travel
    # (NCCL): mates are paired by name across the whole file.
    part = "%s.part%d" % (a.prefix, rank)
    threads = int(os.environ.get("SEEKSV_B200_THREADS", "0"))
    blocks = lib.write_range_blocks(part, cl.text(0), cl.text(1), threads) if cl is not None else []
    ph.mark("block files")
    lists = sharding.all_gather_objects(blocks, dist)
    # the shards' unmapped-branch records go to rank 0 from HBM to HBM (NCCL point-to-point on the library's own export buffer) and
    # are paired there as one stream, in file order; they never pass through host memory
    import torch
    dptr, n_exp = cl.export_device() if cl is not None else (0, 0)
    sizes = [s[0] for s in sharding._all_gather_i64([n_exp], dist, device)]
    if rank != 0 and n_exp:
        dist.send(sharding._as_tensor(dptr, n_exp, device), 0)
    names, lens = (worker.bam.ref_names, worker.bam.ref_lens) if worker.bam is not None else ([], [])
    if rank == 0:
        total = sum(sizes)
        records = torch.empty(total + 256, dtype=torch.uint8, device=device)
        records[total:].zero_()
        off = 0
        for src, k in enumerate(sizes):
            if k and src == 0:
                records[off:off + k].copy_(sharding._as_tensor(dptr, k, device))
            elif k:
                dist.recv(records[off:off + k], src)
            off += k
        torch.cuda.synchronize()
        ph.mark("gather")
        import threading
        assembler = threading.Thread(target=assemble_block_files, args=(a.prefix, lists))   # (sendfile: no GIL held) next to the pairing
        assembler.start()
        u1 = u2 = None
        if total:   # mates are paired by name across the whole file; the two FASTQ files are compressed on the device
            mini = lib.Bam.from_device(ctx, records.data_ptr(), total, 0, len(names), keep=records)
            mini.set_refs(names, lens)
            cu = mini.getclip_handle(unmapped_only=True, gz_outputs=True, **kw)
            ph.mark("pair:device")
            u1, u2 = cu.gz(2), cu.gz(3)
            ph.mark("pair:gz")
            cu.close()
            mini.close()
        for ext, image in ((".unmapped_1.fq.gz", u1), (".unmapped_2.fq.gz", u2)):
            if image is None:
                lib.write_gz(a.prefix + ext, b"")
            else:
                with open(a.prefix + ext, "wb") as f:
                    f.write(image)
        ph.mark("write")
        assembler.join()
        ph.mark("assemble (rest)")
        print("[GetSClipReads] finished!", file=sys.stderr)
    else:
        ph.mark("gather")
    if cl is not None:
        cl.close()
    worker.close()
    ph.mark("close")
    ph.done()


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
    """a.rest = the arguments of `seeksv getsv` (options and the five files). Rank 0 runs the command itself (one join of the clip
    files, the junction bookkeeping, the output files) with a shard provider registered: when the command has merged its junctions
    it hands them over, rank 0 broadcasts them, every rank runs the BAM passes on its own records and the results are added up
    with NCCL collectives on device tensors. The ranks load their shards while rank 0 reads and joins the clip files."""
    import threading
    from . import lib, sharding
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    files = [x for i, x in enumerate(a.rest) if not x.startswith("-") and (i == 0 or a.rest[i - 1] not in ("-F", "-B", "-t", "-l", "-q", "-Q", "-w", "-n", "-a", "-b", "-d", "-e", "-m", "-i", "-R", "-f", "-T", "-L"))]
    if len(files) != 5 or "-F" in a.rest or "-B" in a.rest:
        raise SystemExit("mgpu getsv: five files, no -F / -B (use the single-process command for those)")
    bam_path = files[1]
    ph = _Phases("getsv", rank)
    box = {}

    def load_shard():
        try:
            # (dist = None: a context without a mapped-branch record is widened by this rank alone - the width of a shard's context
            # changes nothing in the other ranks' plans - so that no collective runs on this helper thread)
            box["worker"] = shard_for_passes(ctx, None, bam_path, a.bai, rank, world)
        except Exception as e:      # noqa: BLE001
            box["error"] = e

    def passes(juncs, wins, min_mapq, pairs_used, times):
        """on every rank, with the same arguments: (n, mean, dev, counts, depth)"""
        loader.join()
        if "error" in box:
            raise box["error"]
        worker = box["worker"]
        gw = sharding.GpuShardWorker(worker.bam, device) if worker.bam is not None else _NoRecords(device)
        n = mean = dev = 0
        if pairs_used >= 100000:
            n, mean, dev = sharding.sharded_insert_stats(gw, dist, device, min_mapq, pairs_used)
        t = sharding.sharded_pairs_depth(gw, dist, min_mapq, mean, dev, times, juncs, wins).cpu().numpy()
        return n, mean, dev, t[:len(juncs)], t[len(juncs):len(juncs) + sum(w[2] - w[1] + 1 for w in wins)]

    loader = threading.Thread(target=load_shard)
    loader.start()
    rc = 0
    if rank == 0:
        def provider(juncs, wins, min_mapq, pairs_used, times):
            ph.mark("join (command)")
            box["called"] = True
            if dist is not None:
                dist.broadcast_object_list([juncs, wins, min_mapq, pairs_used, times], src=0)
            out = passes(juncs, wins, min_mapq, pairs_used, times)
            ph.mark("passes")
            return out
        keep = lib.set_shard_provider(provider)
        threads = os.environ.get("SEEKSV_B200_THREADS")
        if dist is not None:    # the one join of the run: all host cores (the other ranks only load their shards meanwhile)
            os.environ["SEEKSV_B200_THREADS"] = str(os.cpu_count() or 2)
        try:
            rc = lib.run_cli(["getsv"] + list(a.rest))
        finally:
            lib.set_shard_provider(None)
            del keep
            if threads is not None:
                os.environ["SEEKSV_B200_THREADS"] = threads
            elif dist is not None:
                del os.environ["SEEKSV_B200_THREADS"]
        ph.mark("output")
        if dist is not None and not box.get("called"):      # the command stopped before its passes: release the other ranks
            dist.broadcast_object_list([None] * 5, src=0)
    else:
        plan = [None] * 5
        dist.broadcast_object_list(plan, src=0)
        if plan[0] is not None:
            passes(*plan)
        ph.mark("passes")
    loader.join()
    if "worker" in box:
        box["worker"].close()
    ph.mark("close")
    ph.done()
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
