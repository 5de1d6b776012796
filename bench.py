#!/usr/bin/env python
"""bench.py - BAM records/sec through getclip + getsv (BASELINE.json metric) on synthetic BAMs.

    python bench.py --gpus 1 --steps 5 --warmup 3            # our arm (CUDA path through the C ABI)
    python bench.py --impl reference --steps 2 --warmup 1    # the reference's own CPU implementation (oracle/_ref)
    torchrun --nproc-per-node N ... bench.py --gpus N ...     # one chromosome-sized shard per GPU, weak scaling

Workload (config.workload = "C2"): BASELINE.json configs[1] - synthetic 30x chr21-size (46,709,983 bp) paired-end
150 bp BAM with 500 planted deletions / inversions / moved segments (tools/svsim.cpp, seed 20261017). The external
`bwa mem` realign step between the two commands is replaced by tools/minialign.cpp and is NOT timed, exactly as the
reference pipeline keeps it outside seeksv.

One step = one getclip pass + one getsv pass over the BAM:
  value : inputs (uncompressed BAM stream) already resident in HBM; record index + soft-clip scan + sort + clustering
          + text emission, then record index + decode + insert-size statistics + discordant-pair support + window depth.
  e2e   : the same two commands end to end through the C-ABI entry point the CLI is (svb_main: file image in host
          memory -> host-thread BGZF inflate -> pinned staging -> cudaMemcpyAsync -> kernels -> results back to host
          -> output files written), i.e. what `seeksv getclip ...; seeksv getsv ...` costs a user.
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
WORK = os.environ.get("SEEKSV_B200_BENCH_DIR", "/tmp/seeksv_b200_bench")
BIN = os.path.join(ROOT, "seeksv_b200", "bin")
REF_SEEKSV = os.path.join(ROOT, "oracle", "_ref", "seeksv")
C2_LEN = 46709983
SAMPLE_LEN = 5000000      # bounded sample for the CPU arm: ~1.0 M records, ~10 s of single-core reference work
SEED = 20261017


def ensure_tools():
    from seeksv_b200 import build as b
    b.build()      # library, CLI, and the svsim / minialign helpers


def make_bam(prefix, contig, length, seed, nsv):
    if not os.path.exists(prefix + ".bam"):
        subprocess.run([os.path.join(BIN, "svsim"), "--out", prefix, "--genome", "%s:%d" % (contig, length), "--cov", "30",
                        "--nsv", str(nsv), "--seed", str(seed)], check=True, stderr=subprocess.DEVNULL)
    return prefix + ".bam"


def realign(prefix, clip_fq_gz):
    """the external realign step (untimed): clip.fq.gz -> clip.sam"""
    sam = prefix + ".clip.sam"
    with open(sam, "w") as o:
        subprocess.run([os.path.join(BIN, "minialign"), prefix + ".fa", clip_fq_gz], check=True, stdout=o)
    return sam


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md), sampled in-process through NVML - spawning
    nvidia-smi five times a second takes driver locks and measurably slows the host side of the end-to-end path."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nv = None

    def run(self):
        nv = self.nv
        while not self.stop_flag and nv is not None:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, mx, r))
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [r[0] for r in self.rows]
        mx = [r[1] for r in self.rows]
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        reasons = sorted(name for name, bit in bits.items() if any(r[2] & bit for r in self.rows))
        return {"sm_mhz": int(statistics.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvml"}


def time_reference(bam, sam_and_clip, work, tag):
    """wall-clock of the reference CLI: getclip, then getsv (its own inflate, index queries and pileup included)"""
    pre = os.path.join(work, tag)
    t0 = time.perf_counter()
    subprocess.run([REF_SEEKSV, "getclip", "-o", pre, bam], check=True, stderr=subprocess.DEVNULL)
    t1 = time.perf_counter()
    sam = sam_and_clip(pre)
    t2 = time.perf_counter()
    with open(pre + ".stdout", "w") as o:
        subprocess.run([REF_SEEKSV, "getsv", sam, bam, pre + ".clip.gz", pre + ".sv", pre + ".unm"], check=True, stdout=o,
                       stderr=subprocess.DEVNULL)
    t3 = time.perf_counter()
    return (t1 - t0) + (t3 - t2)


def count_records(bam):
    import gzip
    import struct
    side = bam + ".nrec"
    if os.path.exists(side) and os.path.getmtime(side) >= os.path.getmtime(bam):
        return int(open(side).read())
    n = 0
    with gzip.open(bam, "rb") as f:
        data = f.read()
    o = 8 + struct.unpack_from("<i", data, 4)[0]
    nref = struct.unpack_from("<i", data, o)[0]
    o += 4
    for _ in range(nref):
        o += 8 + struct.unpack_from("<i", data, o)[0]
    while o + 4 <= len(data):
        o += 4 + struct.unpack_from("<i", data, o)[0]
        n += 1
    with open(side, "w") as f:
        f.write(str(n))
    return n


REF_WALL_BUDGET = float(os.environ.get("SEEKSV_B200_REF_BUDGET_S", "200"))


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation (oracle/_ref/seeksv, the unmodified sources compiled by
    oracle/build_ref.sh) on the SAME input as the GPU arm - the full C2 BAM - rank 0 only, 1 thread (the reference has none to
    use). One step is ~40-75 s of CPU work, so the run is bounded by a wall budget instead of a sample: warm-up steps are skipped
    (a step is two fresh processes; nothing is cached across steps but the page cache, warmed by generating the input) and timed
    steps stop when the budget is used up; at least one step is timed. `steps_run` says how many were."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.makedirs(WORK, exist_ok=True)
    if not os.path.exists(REF_SEEKSV):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/seeksv not built (needs /root/reference at build time)"}))
        return
    ensure_tools()
    pre = os.path.join(WORK, "c2_chr21_%d" % args.genome_len)
    nsv = max(1, int(500 * args.genome_len / C2_LEN))
    bam = make_bam(pre, "chr21", args.genome_len, SEED, nsv)
    n_rec = count_records(bam)
    sam_cache = {}

    def sam_for(p):
        if "sam" not in sam_cache:
            if os.path.abspath(pre + ".fa") != os.path.abspath(p + ".fa"):
                shutil.copy(pre + ".fa", p + ".fa")
            sam_cache["sam"] = realign(p, p + ".clip.fq.gz")
        return sam_cache["sam"]
    t_start = time.perf_counter()
    ts = []
    for _ in range(max(1, args.steps)):
        ts.append(time_reference(bam, sam_for, WORK, "refarm"))
        if time.perf_counter() - t_start + ts[-1] > REF_WALL_BUDGET:
            break
    t = sum(ts)
    v = n_rec * len(ts) / t
    sample = "the full workload: svsim chr21:%d 30x 150bp PE (%d records), getclip+getsv CLI wall-clock, %d step(s) inside a %d s budget" % (
        args.genome_len, n_rec, len(ts), int(REF_WALL_BUDGET))
    print(json.dumps({
        "impl": "reference", "metric": "BAM records/sec getclip+getsv", "value": v, "unit": "records/s", "n_gpus": args.gpus,
        "steps": args.steps, "steps_run": len(ts), "warmup": args.warmup, "warmup_run": 0, "ms_per_step": 1e3 * t / len(ts), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": "C2" if args.genome_len == C2_LEN else "C2-shape-%dbp" % args.genome_len, "records": n_rec, "threads": 1,
                   "same_input_as_gpu_arm": True},
        "cpu_baseline": {"value": v, "unit": "records/s", "cores": 1, "kind": "reference", "sample": sample,
                         "host_cores": os.cpu_count()},
        "e2e": {"value": v, "unit": "records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def ncu_traffic(kernel_scope):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel behind a timer scope, from the NEWEST committed
    `ncu --set full` raw CSV under profiles/ that holds the kernel (page raw; the workload is seeded, so the byte counts are
    those of this run's input). Returns (bytes or None, source string)."""
    import csv
    import glob
    kernel = {"rec_walk_fused": "rec_walk<(bool)1, (bool)1>", "rec_walk_clip": "rec_walk<(bool)1, (bool)0>",
              "rec_walk_rows": "rec_walk<(bool)0, (bool)1>", "rec_walk_count": "rec_walk<(bool)0, (bool)0>"}.get(kernel_scope, kernel_scope)
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full_*raw.csv")), key=os.path.getmtime, reverse=True)
    files.sort(key=lambda f: os.path.basename(f)[:3], reverse=True)     # newest round first
    for f in files:
        try:
            rows = list(csv.reader(open(f)))
            hdr, units = rows[0], rows[1]
            ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            for r in rows[2:]:
                if kernel.replace(" ", "").replace("(bool)", "") in r[ki].replace(" ", "").replace("(bool)", ""):
                    tot = float(r[ri]) * unit.get(units[ri], 1.0) + float(r[wi]) * unit.get(units[wi], 1.0)
                    return tot, "profiles/%s (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch on C2)" % os.path.basename(f)
        except Exception:
            continue
    return None, "no committed ncu --set full capture holds this kernel"


def partitioned_main(args, torch, dist, rank, world, local):
    """--gpus N > 1: ONE coordinate-sorted multi-chromosome BAM (3 contigs per GPU, N x the C2 genome: weak scaling), partitioned by
    coordinate range across the ranks (seeksv_b200/sharding.py: cuts at .bai linear-index offsets, halo + breakpoint-key ownership),
    every rank's shard resident in HBM. One step per rank = getclip on its shard (keys it owns; its pass over the records also
    leaves getsv's rows) -> all-to-all of the unmapped-branch records by read-name group over NVLink (device buffers) and pairing
    of the received group -> getsv's statistics, pair support and depth on its own records, combined with NCCL collectives on
    device tensors (all_gather of 4 integers, all_reduce of the counts / depths). The untimed prologue checks the sharded results
    against the same commands on the whole file on one GPU."""
    import ctypes as C
    import gzip
    import hashlib
    import seeksv_b200 as S
    from seeksv_b200 import lib as SL, mgpu, sharding
    device = "cuda:%d" % local
    n_contigs = 3 * world
    clen = args.genome_len // 3
    pre = os.path.join(WORK, "part_%dx%d" % (n_contigs, clen))
    bam_path = pre + ".bam"
    if rank == 0 and not os.path.exists(bam_path):
        genome = ",".join("chr%d:%d" % (i + 1, clen) for i in range(n_contigs))
        subprocess.run([os.path.join(BIN, "svsim"), "--out", pre, "--genome", genome, "--cov", "30", "--nsv", str(max(1, int(500 * world * args.genome_len / C2_LEN))),
                        "--seed", str(SEED)], check=True, stderr=subprocess.DEVNULL)
    dist.barrier()
    ctx = S.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream)
    worker = mgpu.open_range_worker(ctx, dist, bam_path, None, rank, world)
    plan = worker.plan
    assert worker.bam is not None, "a rank without records: the BAM is too small for this many ranks"
    names, lens = worker.bam.ref_names, worker.bam.ref_lens
    dptr, nbytes, first = worker.bam.device_stream()

    # ---- whole file on rank 0 (untimed): realign hand-off, getsv plan, and the results the shards have to reproduce -------------
    ref = None
    if rank == 0:
        whole = S.Bam.open(ctx, bam_path)
        clip = whole.getclip()
        with gzip.open(pre + ".clip.gz", "wb", compresslevel=1) as f:
            f.write(clip[0])
        with gzip.open(pre + ".clip.fq.gz", "wb", compresslevel=1) as f:
            f.write(clip[1])
        sam = realign(pre, pre + ".clip.fq.gz")
        juncs, wins = S.plan_getsv(sam, pre + ".clip.gz", names, lens, 50, 200)
        st, cnts, deps = whole.getsv_passes(juncs, wins, 20, 5000000, 4)
        n_total, whole_bytes = whole.n_records, whole.record_bytes
        whole.close()
        ref = dict(clip=clip, st=st, cnts=cnts, deps=[x for d in deps for x in d], n_clusters=clip[0].count(b"\n"))
        plan_obj = [juncs, wins, n_total, whole_bytes, sam]
    else:
        plan_obj = [None] * 5
    dist.broadcast_object_list(plan_obj, src=0)
    juncs, wins, n_total, whole_bytes, sam = plan_obj
    nj, n_pos = len(juncs), sum(w[2] - w[1] + 1 for w in wins)

    class Step:
        keep = None

    # the pairing of the received name group runs on a second context (own stream, own workspace) in a helper thread, next to the
    # getsv passes of the main thread: the two do not depend on each other
    import queue
    import threading
    ctx2 = S.Context(local)
    jobs, done = queue.Queue(), queue.Queue()

    def pair_thread():
        torch.cuda.set_device(local)
        while True:
            job = jobs.get()
            if job is None:
                return
            exchange, recv, n_recv = job
            try:
                exchange.wait()
                torch.cuda.current_stream().synchronize()
                mini = S.Bam.from_device(ctx2, recv.data_ptr(), n_recv, 0, len(names))
                mini.set_refs(names, lens)
                done.put((mini, mini.getclip_handle(unmapped_only=True)))
            except Exception as e:      # noqa: BLE001
                done.put(e)
    pairer = threading.Thread(target=pair_thread, daemon=True)
    pairer.start()
    vec = torch.zeros(world + 4, dtype=torch.int64, device=device)           # [bytes for every name group | n, sum x, sum x^2, large]
    vec_all = torch.zeros(world * (world + 4), dtype=torch.int64, device=device)
    part = torch.zeros(4, dtype=torch.int64, device=device)
    sizes_host = torch.zeros(world, dtype=torch.int64).pin_memory()
    cap_flag = torch.zeros(1, dtype=torch.int32, device=device)      # a pile-up at libbam's cap cut by a shard boundary (sharding.py)
    phases = {}

    def shard_step(keep=False):
        t0 = time.perf_counter()
        b = S.Bam.from_device(ctx, dptr, nbytes, first, len(names))
        b.set_refs(names, lens)
        b.set_own_offset(plan.halo_bytes)
        cl = b.getclip_handle(prev_tid=plan.prev_tid, export_unmapped=True, key_range=(plan.key_lo, plan.key_hi), halo_bytes=plan.halo_bytes,
                              with_rows=True, export_partitions=world)
        t1 = time.perf_counter()
        # one all_gather carries the exchange's byte counts and the shard's insert-size sums (device tensors, no read-back before it)
        parts = cl.export_parts(world)
        for r in range(world):
            sizes_host[r] = parts[r + 1] - parts[r]
        vec[:world].copy_(sizes_host, non_blocking=True)
        b.insert_partial_async(20, -1, vec.data_ptr() + 8 * world)
        dist.all_gather_into_tensor(vec_all, vec)
        every = vec_all.view(world, world + 4).tolist()
        t2 = time.perf_counter()
        # the unmapped-branch records travel all-to-all by read-name group (NVLink, device buffers) while the statistics are settled
        send_sizes = [int(x) for x in every[rank][:world]]
        recv_sizes = [int(every[src][rank]) for src in range(world)]
        dptr_e, n_e = cl.export_device()
        send = sharding._as_tensor(dptr_e, max(n_e, 1), device)[:n_e]
        recv = torch.zeros(sum(recv_sizes) + 256, dtype=torch.uint8, device=device)
        exchange = dist.all_to_all_single(recv[:sum(recv_sizes)], send, recv_sizes, send_sizes, async_op=True)
        jobs.put((exchange, recv, sum(recv_sizes)))
        counts = [e[world] for e in every]
        takes = sharding.prefix_cutoffs(counts, 5000000)
        if takes == counts:
            n, sx, sxx, big = (sum(e[world + k] for e in every) for k in range(4))
        else:       # the -n cut falls inside a shard: that shard sums its prefix again, one small all_reduce
            if takes[rank] == counts[rank]:
                part.copy_(vec[world:])
            elif takes[rank] == 0:
                part.zero_()
            else:
                b.insert_partial_async(20, takes[rank], part.data_ptr())
            dist.all_reduce(part)
            n, sx, sxx, big = part.tolist()
        assert big == 0, "insert sizes above 46340: the wrap-exact path is sharding.sharded_insert_stats"
        mean = sx // n if n else 0
        dev = sharding.mean_dev(n, sx, sxx - 2 * mean * sx + n * mean * mean)[1] if n else 0
        t3 = time.perf_counter()
        gw = sharding.GpuShardWorker(b, device)
        gw._arrays = prepared
        t = sharding.sharded_pairs_depth(gw, dist, 20, mean, dev, 4, juncs, wins, cap_flag=cap_flag)   # (flag checked after the timed loops)
        t4 = time.perf_counter()
        got = done.get()
        if isinstance(got, Exception):
            raise got
        mini, cu = got
        t5 = time.perf_counter()
        if keep:
            Step.keep = (cl.text(0), cl.text(1), cu.text(2), cu.text(3), (n, mean, dev), t.cpu().numpy().copy())
        cu.close()
        mini.close()
        cl.close()
        b.close()
        for k, v in (("getclip", t1 - t0), ("gather", t2 - t1), ("stats", t3 - t2), ("pairs_depth", t4 - t3), ("unmapped", t5 - t4)):
            phases[k] = phases.get(k, 0.0) + v
        return 4 * (nj + n_pos)

    probe = sharding.GpuShardWorker(None, device)
    probe.prepare(juncs, wins)
    prepared = probe._arrays

    # ---- the check (untimed): merged shard results == whole-file results ----------------------------------------------------------
    with torch.cuda.stream(stream):
        shard_step(keep=True)
    torch.cuda.synchronize()
    mine = Step.keep
    every = sharding.all_gather_objects((mine[0], mine[1], mine[2], mine[3]), dist)
    if rank == 0:
        clip_m, fq_m = sharding.merge_range_texts_fast([(e[0], e[1]) for e in every])
        assert clip_m == ref["clip"][0] and fq_m == ref["clip"][1], "sharded getclip differs from the whole-file run"

        def records(text):
            lines = text.split(b"\n")[:-1]
            return [b"\n".join(lines[i:i + 4]) for i in range(0, len(lines), 4)]
        for which in (2, 3):          # every name group's FASTQ == the whole-file FASTQ restricted to the names of the group, in order
            groups = [[] for _ in range(world)]
            for r in records(ref["clip"][which]):
                name = r[1:r.index(b"\n")].rsplit(b"/", 1)[0]
                groups[sharding.fnv1a64(name) % world].append(r)
            for g in range(world):
                assert records(every[g][which]) == groups[g], "sharded unmapped-mate pairing differs from the whole-file run (group %d)" % g
        n, mean, dev = mine[4]
        import math
        assert (mean, dev) == (ref["st"][2], int(math.sqrt(ref["st"][3] / ref["st"][0]))), "sharded insert-size statistics differ"
        assert mine[5][:nj].tolist() == ref["cnts"] and mine[5][nj:nj + n_pos].tolist() == ref["deps"], "sharded pair support / depth differ"
    dist.barrier()
    Step.keep = None

    # ---- timing ---------------------------------------------------------------------------------------------------------------------
    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            ev0.record()
            t0 = time.perf_counter()
            for _ in range(steps):
                last = fn()
            ev1.record()
        barrier()
        wall = time.perf_counter() - t0
        t = torch.tensor([max(ev0.elapsed_time(ev1), 0.0), wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item(), t[1].item(), last

    for _ in range(args.warmup):
        with torch.cuda.stream(stream):
            shard_step()
    sampler = ClockSampler(local)
    sampler.start()
    phases.clear()
    ms_dev, wall_dev, d2h = timed(shard_step, args.steps)
    host_phases = {k: round(1e3 * v / args.steps, 4) for k, v in phases.items()}
    ctx.prof(True)
    ctx.prof_reset()
    ms_prof, _, _ = timed(shard_step, args.steps)
    prof = ctx.prof_read()
    ctx.prof(False)
    assert int(cap_flag.item()) == 0, "a pile-up of >= 8000 reads is cut by a shard boundary: the sharded depth is not the whole-file depth"

    # ---- end to end: the two commands on N GPUs through the multi-GPU entry points, BGZF file -> output files --------------------
    out_dir = os.path.join(WORK, "pout")
    if rank == 0:
        os.makedirs(out_dir, exist_ok=True)
    dist.barrier()

    class A:
        pass

    def e2e_step():
        a = A()
        a.match_rate, a.min_mapq, a.save_low_quality, a.by, a.bai = 0.9, 1, False, "range", None
        a.bam, a.prefix = bam_path, os.path.join(out_dir, "x")
        mgpu.run_getclip(ctx, dist, device, a)
        dist.barrier()
        a.rest = [sam, bam_path, os.path.join(out_dir, "x.clip.gz"), os.path.join(out_dir, "x.sv"), os.path.join(out_dir, "x.unm")]
        rc = mgpu.run_getsv(ctx, dist, device, a)
        assert rc == 0
        return 0

    wall_e2e = float("nan")
    if not args.value_only:
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved, saved_out = os.dup(2), os.dup(1)
        if not os.environ.get("SEEKSV_B200_TIMING"):      # (the commands' progress lines; the phase timings go to stderr too)
            os.dup2(devnull, 2)
        sys.stdout.flush()
        os.dup2(devnull, 1)
        try:
            e2e_step()
            _, wall_e2e, _ = timed(e2e_step, max(1, min(args.steps, 3)))
            wall_e2e /= max(1, min(args.steps, 3))
        finally:
            os.dup2(saved, 2)
            os.dup2(saved_out, 1)
        if rank == 0:     # the files of the multi-GPU commands against the whole-file run
            with gzip.open(os.path.join(out_dir, "x.clip.gz"), "rb") as f:
                assert f.read() == ref["clip"][0], "multi-GPU getclip file differs from the whole-file run"
    sampler.stop_flag = True
    sampler.join(timeout=2)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        kern = {k: v for k, v in prof.items() if v["launches"] > 0 and "wall" not in k}
        dom = "rec_walk_fused"
        value = n_total * args.steps / (ms_dev / 1e3)
        roofline = None
        if dom in kern:
            v = kern[dom]
            avg_ms = v["ms"] / v["launches"]
            per_launch = v["bytes"] / v["launches"]
            ach = per_launch / (avg_ms * 1e-3) / 1e9
            step_s = ms_dev / args.steps * 1e-3
            roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                        "peak_source": "measured" if peaks else "fallback", "algorithmic_bytes_per_launch": per_launch, "avg_launch_ms": avg_ms,
                        "kernels_ms_per_step": {k: round(x["ms"] / args.steps, 4) for k, x in sorted(kern.items())},
                        "step": {"algorithmic_bytes_all_gpus": 2.0 * whole_bytes, "achieved_all_gpus": 2.0 * whole_bytes / step_s / 1e9,
                                 "frac_per_gpu": 2.0 * whole_bytes / step_s / 1e9 / peak / world, "ms_per_step_with_timers": ms_prof / args.steps,
                                 "host_ms_per_phase_rank0": host_phases,
                                 "note": "rank 0's kernels; 2 x record bytes of the whole BAM / step time, per GPU"}}
        line = {
            "metric": "BAM records/sec getclip+getsv", "value": value, "unit": "records/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "C2 x %d: one BAM of %d contigs x %d bp (C3's shape at N x the C2 size)" % (world, n_contigs, clen),
                       "records": n_total, "record_bytes": whole_bytes, "records_per_gpu": n_total // world, "clusters": ref["n_clusters"],
                       "junction_candidates": nj, "depth_windows": len(wins),
                       "sharding": "coordinate ranges of one coordinate-sorted BAM (cuts at .bai linear-index offsets, equal compressed bytes; halo + "
                                   "breakpoint-key ownership), one shard per GPU resident in HBM",
                       "collectives": "NCCL on device tensors: all_gather(4 x int64) + all_reduce for the insert-size statistics, all_to_all_single of the "
                                      "unmapped-branch records by read-name group, all_reduce(int32 x %d) of pair counts and depths" % (nj + n_pos),
                       "checked": "sharded clip / clip.fq / unmapped FASTQ / statistics / pair counts / depths == the whole-file run (untimed prologue)",
                       "l2_note": "each shard (%.2f GB) is far larger than the 126 MB L2; no flush needed" % (nbytes / 1e9)},
            "e2e": {"value": n_total / wall_e2e if wall_e2e == wall_e2e else None, "unit": "records/s", "ms_per_step": 1e3 * wall_e2e,
                    "h2d_bytes_per_step": 2 * os.path.getsize(bam_path),
                    "d2h_bytes_per_step": (len(ref["clip"][0]) + len(ref["clip"][1]) + 4 * (nj + n_pos)
                                           + int(sum(os.path.getsize(os.path.join(out_dir, f)) for f in os.listdir(out_dir) if f.startswith("x.unmapped")))
                                           ) if wall_e2e == wall_e2e else None,
                    "d2h_note": "all ranks together: the clip / clip.fq TEXT of every shard (device -> pinned host, compressed into gzip block files by "
                                "that rank's host threads), the two unmapped FASTQ files compressed on rank 0's device, pair counts and depths",
                    "path": "seeksv_b200.mgpu getclip + getsv on N ranks: BGZF file (page cache) -> each rank loads its shard twice -> device passes -> "
                            "NCCL merges -> rank 0 writes the reference's files"},
            "gpu_launches": int(sum(v["launches"] for v in kern.values())),
            "clocks": sampler.summary(), "roofline": roofline,
        }
        print(json.dumps(line))
    jobs.put(None)
    worker.close()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--genome-len", type=int, default=C2_LEN, help="chromosome length of the synthetic BAM (C2: chr21 size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--value-only", action="store_true", help="only the HBM-resident step (for ncu launch lists): no e2e legs, no CPU baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL announces its version on stdout at the first collective: keep stdout for the one JSON line
        sys.stdout.flush()
        keep = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.all_reduce(torch.zeros(1, device="cuda"))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(keep, 1)
            os.close(keep)
        # the ranks share the host: split the cores for the host side of the commands (staging copies, parsing, gzip)
        os.environ.setdefault("SEEKSV_B200_THREADS", str(max(2, (os.cpu_count() or 2) // world)))
    os.environ["SEEKSV_B200_DEVICE"] = str(local)     # the CLI entry point (svb_main) picks its GPU from the environment
    os.makedirs(WORK, exist_ok=True)
    if rank == 0:
        ensure_tools()
    if world > 1:
        dist.barrier()
    import seeksv_b200 as S
    if world > 1:
        return partitioned_main(args, torch, dist, rank, world, local)

    # ---- inputs (untimed): one chromosome-sized BAM per rank - the genome is partitioned by chromosome ----------
    contig = "chr21" if world == 1 else "chr%d" % (rank + 1)
    pre = os.path.join(WORK, "c2_%s_%d" % (contig, args.genome_len))
    nsv = max(1, int(500 * args.genome_len / C2_LEN))
    bam_path = make_bam(pre, contig, args.genome_len, SEED + rank, nsv)
    ctx = S.Context(local)
    file_img = torch.from_file(bam_path, shared=False, size=os.path.getsize(bam_path), dtype=torch.uint8).pin_memory()
    resident = S.Bam.from_bgzf(ctx, file_img)          # uncompressed stream in HBM (kept for the whole run)
    dptr, nbytes, first = resident.device_stream()
    names, lens = resident.ref_names, resident.ref_lens
    n_rec, rec_bytes = resident.n_records, resident.record_bytes
    with open(bam_path + ".nrec", "w") as f:      # (spares the reference arm a Python walk over the records)
        f.write(str(n_rec))
    # realign hand-off and getsv plan from our own getclip output (parity with the reference is tested elsewhere)
    clip = resident.getclip()
    import gzip
    with gzip.open(pre + ".clip.gz", "wb", compresslevel=1) as f:
        f.write(clip[0])
    with gzip.open(pre + ".clip.fq.gz", "wb", compresslevel=1) as f:
        f.write(clip[1])
    sam = realign(pre, pre + ".clip.fq.gz")
    juncs, wins = S.plan_getsv(sam, pre + ".clip.gz", names, lens, 50, 200)
    n_clusters = clip[0].count(b"\n")
    # the merge buffer has the same length on every rank (shards find different numbers of junction candidates)
    n_merge = torch.tensor([max(1, len(juncs))], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(n_merge, op=dist.ReduceOp.MAX)
    # final candidate merge: two buffer sets, so that the merge of step i overlaps the kernels of step i + 1
    n_slots = int(n_merge.item())
    counts_dev = [torch.zeros(n_slots, dtype=torch.int32, device="cuda") for _ in range(2)]
    merged_dev = [torch.empty(n_slots * max(world, 1), dtype=torch.int32, device="cuda") for _ in range(2)]
    pending_merge = [None, None]
    step_no = [0]
    stage_host = None
    stream = torch.cuda.ExternalStream(ctx.stream)

    # C arrays for the getsv passes are built once; results land in preallocated host arrays (no per-step marshalling)
    import ctypes as C
    from seeksv_b200 import lib as SL
    nj, nw = len(juncs), len(wins)
    j_arr = (SL.Junction * max(nj, 1))(*[SL.Junction(ut, up, dt, dp, us.encode(), ds.encode(), b"") for ut, up, us, dt, dp, ds in juncs])
    w_arr = (SL.Window * max(nw, 1))(*[SL.Window(*w) for w in wins])
    n_pos = sum(w[2] - w[1] + 1 for w in wins)
    cnt_host = torch.zeros(max(nj, 1), dtype=torch.int32).pin_memory()
    dep_host = torch.zeros(max(n_pos, 1), dtype=torch.int32).pin_memory()
    stage_host = [torch.zeros_like(cnt_host).pin_memory() for _ in range(2)]
    cnt_arr = C.cast(cnt_host.data_ptr(), C.POINTER(C.c_int32))
    dep_arr = C.cast(dep_host.data_ptr(), C.POINTER(C.c_int32))

    gp = SL.GetsvParams(20, 4, 5000000)
    st_arr = (C.c_int64 * 4)()
    fused_walk = [True]

    def device_step():
        """getclip + getsv on the HBM-resident stream, as `seeksv run` does them: ONE handle over the stream (chunk guesses made
        once), getclip's pass over the records also leaves getsv's rows (with_rows), the getsv passes are one fused call. The
        four texts stay in HBM (svb_clusters_text would copy them on request); counts and depth come back to pinned memory.
        fused_walk off: the getsv passes stream the records themselves (two passes per step, as two separate commands)."""
        b = S.Bam.from_device(ctx, dptr, nbytes, first, len(names))
        b.set_refs(names, lens)
        sizes = b.getclip_sizes(with_rows=fused_walk[0])
        b.getsv_passes_raw(gp, j_arr, nj, w_arr, nw, st_arr, cnt_arr, dep_arr)
        b.close()
        if world > 1:   # final candidate merge: every rank learns every shard's support counts (small NCCL allgather)
            k = step_no[0] & 1
            step_no[0] += 1
            if pending_merge[k] is not None:
                pending_merge[k].wait()      # the buffers of two steps ago are free again
            stage_host[k].copy_(cnt_host)     # (the next step overwrites cnt_host while this copy may still be queued)
            counts_dev[k][:cnt_host.numel()].copy_(stage_host[k], non_blocking=True)
            pending_merge[k] = dist.all_gather_into_tensor(merged_dev[k], counts_dev[k], async_op=True)
        return 4 * nj + 4 * n_pos + 32

    out_dir = os.path.join(WORK, "out_%d" % rank)
    os.makedirs(out_dir, exist_ok=True)

    def e2e_step():
        """the two commands through svb_main (what the CLI runs), files in the page cache"""
        rc = S.run_cli(["getclip", "-o", os.path.join(out_dir, "x"), bam_path])
        assert rc == 0
        rc = S.run_cli(["getsv", sam, bam_path, os.path.join(out_dir, "x.clip.gz"), os.path.join(out_dir, "x.sv"),
                        os.path.join(out_dir, "x.unm")])
        assert rc == 0

    def fused_step():
        """the same two commands inside one `seeksv run`: the BAM is uploaded and inflated once and stays resident"""
        rc = S.run_cli(["run", "--", "getclip", "-o", os.path.join(out_dir, "y"), bam_path, "--",
                        "getsv", sam, bam_path, os.path.join(out_dir, "y.clip.gz"), os.path.join(out_dir, "y.sv"),
                        os.path.join(out_dir, "y.unm")])
        assert rc == 0

    def barrier():
        for k in range(2):               # every merge that is still in flight belongs to the region that ends here
            if pending_merge[k] is not None:
                pending_merge[k].wait()
                pending_merge[k] = None
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            ev0.record()
            t0 = time.perf_counter()
            for _ in range(steps):
                last = fn()
            ev1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = max(ev0.elapsed_time(ev1), 0.0)
        t = torch.tensor([ms, wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item(), t[1].item(), last

    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(2)
    for _ in range(args.warmup):
        device_step()
    sampler = ClockSampler(local)
    sampler.start()
    ms_dev, wall_dev, d2h = timed(device_step, args.steps)      # the headline: no timers inside
    ctx.prof(True)                                              # the same steps again with the per-kernel CUDA-event timers on
    ctx.prof_reset()
    ms_prof, _, _ = timed(device_step, args.steps)
    prof = ctx.prof_read()
    ctx.prof(False)
    fused_walk[0] = False          # for comparison: the same step with two passes over the records (not the headline)
    for _ in range(2):
        device_step()
    ms_two, _, _ = timed(device_step, args.steps)
    fused_walk[0] = True
    os.dup2(devnull, 2)          # the commands print the reference's progress lines on stderr
    sys.stdout.flush()
    saved_out = os.dup(1)
    os.dup2(devnull, 1)          # ... and getsv lists filtered junctions on stdout
    wall_e2e = wall_fused = float("nan")
    try:
        if not args.value_only:
            for _ in range(max(1, args.warmup)):
                e2e_step()
            ms_e2e, wall_e2e, _ = timed(e2e_step, args.steps)
            for _ in range(max(1, args.warmup)):
                fused_step()
            _, wall_fused, _ = timed(fused_step, args.steps)
    finally:
        os.dup2(saved, 2)
        os.dup2(saved_out, 1)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    total_rec = torch.tensor([n_rec], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_rec)
    total_rec = total_rec.item()
    value = total_rec * args.steps / (ms_dev / 1e3)
    e2e = total_rec * args.steps / wall_e2e    # svb_main runs on its own context/stream: host clock, max over ranks

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        full_pass = {"guess_starts", "rec_walk_fused", "rec_walk_clip", "rec_walk_rows", "rec_walk_count"}
        kern = {k: v for k, v in prof.items() if v["launches"] > 0 and "wall" not in k}
        # the roofline is quoted for the dominant FULL-PASS kernel (the ones that stream the records); the candidate-side scopes
        # (sorts, clustering: ~2 % of the data, many tiny launches) are listed in kernels_ms_per_step but have no HBM roofline
        passes = {k: v for k, v in kern.items() if k in full_pass}
        dom = max(passes, key=lambda k: passes[k]["ms"] / passes[k]["launches"]) if passes else (max(kern, key=lambda k: kern[k]["ms"]) if kern else None)
        roofline = None
        if dom:
            v = kern[dom]
            per_launch = (v["bytes"] / v["launches"]) if v["bytes"] else (rec_bytes if dom in full_pass else 0)
            avg_ms = v["ms"] / v["launches"]
            ach = per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
            traffic, traffic_src = ncu_traffic(dom) if args.genome_len == C2_LEN else (None, "not the C2 workload")
            roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "traffic": traffic, "traffic_source": traffic_src,
                        "dram_frac": (traffic / (avg_ms * 1e-3) / 1e9 / peak) if traffic and avg_ms > 0 else None,
                        "peak_source": "measured" if peaks else "fallback",
                        "algorithmic_bytes_per_launch": per_launch, "avg_launch_ms": avg_ms,
                        "kernels_ms_per_step": {k: round(x["ms"] / args.steps, 4) for k, x in sorted(kern.items())}}
            # the whole step against the same roofline: getclip and getsv each stream every record once
            step_s = ms_dev / args.steps * 1e-3
            step_gbs = 2.0 * rec_bytes / step_s / 1e9 if step_s > 0 else 0.0
            two_s = ms_two / args.steps * 1e-3
            roofline["step"] = {"algorithmic_bytes": 2.0 * rec_bytes, "achieved": step_gbs, "frac": step_gbs / peak,
                                "kernel_ms": round(sum(x["ms"] for x in kern.values()) / args.steps, 4),
                                "ms_per_step_with_timers": ms_prof / args.steps,
                                "passes_over_records": 1,
                                "frac_one_pass_basis": (rec_bytes / step_s / 1e9 / peak) if step_s > 0 else None,
                                "two_pass_step": {"ms_per_step": ms_two / args.steps, "frac": (2.0 * rec_bytes / two_s / 1e9 / peak) if two_s > 0 else None,
                                                  "note": "the same step with with_rows off: the getsv passes stream the records themselves, as two separate commands do"},
                                "note": "per GPU (rank 0). algorithmic_bytes = 2 x record bytes: getclip and getsv each have to stream every record "
                                        "(SURVEY.md section 8d); the resident pipeline timed here (`seeksv run`) serves both from ONE pass of the record "
                                        "walker (passes_over_records), frac_one_pass_basis divides by that. Kernel scopes on the side stream overlap "
                                        "the main stream, so kernel_ms can exceed the step"}
        line = {
            "metric": "BAM records/sec getclip+getsv", "value": value, "unit": "records/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "C2" if args.genome_len == C2_LEN else "C2-shape-%dbp" % args.genome_len,
                       "records_per_gpu": n_rec, "record_bytes_per_gpu": rec_bytes, "clusters": n_clusters,
                       "junction_candidates": len(juncs), "depth_windows": len(wins),
                       "sharding": "one chromosome-sized BAM per GPU" if world > 1 else "single GPU",
                       "l2_note": "inputs (%.2f GB per pass) are far larger than the 126 MB L2; no flush needed" % (rec_bytes / 1e9)},
            "e2e": {"value": e2e, "unit": "records/s", "h2d_bytes_per_step": 2 * os.path.getsize(bam_path), "d2h_bytes_per_step": int(sum(os.path.getsize(os.path.join(out_dir, f)) for f in os.listdir(out_dir) if f.startswith("x.") and f.endswith(".gz")) + 4 * nj + 4 * n_pos),
                    "h2d_note": "each command uploads the BGZF file image (inflated on the device to %d bytes)" % nbytes,
                    "ms_per_step": 1e3 * wall_e2e / args.steps, "path": "svb_main getclip + svb_main getsv, BGZF file image -> outputs"},
            "e2e_fused": {"value": total_rec * args.steps / wall_fused, "unit": "records/s", "ms_per_step": 1e3 * wall_fused / args.steps,
                          "h2d_bytes_per_step": os.path.getsize(bam_path),
                          "path": "svb_main run -- getclip -- getsv (one process, BAM loaded once and kept in HBM; not the headline)"},
            "value_d2h_bytes_per_step": int(d2h),
            "gpu_launches": int(sum(v["launches"] for v in kern.values())),
            "clocks": sampler.summary(), "roofline": roofline,
        }
        if args.value_only:
            pass
        elif world == 1 and not args.no_cpu_baseline and os.path.exists(REF_SEEKSV):
            spre = os.path.join(WORK, "sample")
            sbam = make_bam(spre, "chr21", SAMPLE_LEN, SEED, max(1, int(500 * SAMPLE_LEN / C2_LEN)))
            sn = count_records(sbam)

            def sam_for(p):
                shutil.copy(spre + ".fa", p + ".fa")
                return realign(p, p + ".clip.fq.gz")
            t = time_reference(sbam, sam_for, WORK, "cpu_baseline")
            line["cpu_baseline"] = {"value": sn / t, "unit": "records/s", "cores": 1, "kind": "reference", "host_cores": os.cpu_count(),
                                    "sample": "svsim chr21:%d 30x (%d records), reference seeksv getclip+getsv CLI wall-clock, 1 thread "
                                              "(the reference has no parallelism)" % (SAMPLE_LEN, sn)}
        elif world == 1:
            line["cpu_baseline"] = {"value": None, "unit": "records/s", "cores": 1, "kind": "reference",
                                    "sample": "unavailable: oracle/_ref/seeksv not built"}
        print(json.dumps(line))
    resident.close()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
        # skip interpreter teardown: freeing pinned tensors after NCCL has torn its context down aborts the process
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
