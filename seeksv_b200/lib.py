"""ctypes bindings of include/seeksv_b200.h (host-side mirror used by tests, bench.py and sharded runs)."""
from __future__ import annotations

import ctypes as C
import os
import sys
from typing import List, Optional, Sequence, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib_path() -> str:
    return os.path.join(HERE, "libseeksv_b200.so")


def cli_path() -> str:
    return os.path.join(HERE, "bin", "seeksv")


class SvbError(RuntimeError):
    pass


class GetclipParams(C.Structure):
    _fields_ = [("match_rate", C.c_double), ("min_mapq", C.c_int32), ("save_low_quality", C.c_int32),
                ("prev_tid", C.c_int32), ("export_unmapped_records", C.c_int32), ("key_filter", C.c_int32),
                ("key_lo_tid", C.c_int32), ("key_lo_pos", C.c_int32), ("key_hi_tid", C.c_int32), ("key_hi_pos", C.c_int32),
                ("halo_bytes", C.c_uint64), ("gz_outputs", C.c_int32), ("with_rows", C.c_int32),
                ("unmapped_only", C.c_int32), ("export_partitions", C.c_int32)]


class GetsvParams(C.Structure):
    _fields_ = [("min_mapq", C.c_int32), ("times", C.c_int32), ("max_pairs", C.c_int64)]


class Junction(C.Structure):
    _fields_ = [("up_tid", C.c_int32), ("up_pos", C.c_int32), ("down_tid", C.c_int32), ("down_pos", C.c_int32),
                ("up_strand", C.c_char), ("down_strand", C.c_char), ("pad_", C.c_char * 2)]


class PairParams(C.Structure):
    _fields_ = [("min_mapq", C.c_int32), ("mean_insert", C.c_int32), ("deviation", C.c_int32), ("times", C.c_int32)]


class Window(C.Structure):
    _fields_ = [("tid", C.c_int32), ("begin", C.c_int32), ("end", C.c_int32)]


EXPORTS = [
    "svb_abi_version", "svb_ctx_create", "svb_ctx_destroy", "svb_last_error", "svb_ctx_stream", "svb_prof_enable",
    "svb_prof_reset", "svb_prof_read", "svb_bam_from_device", "svb_bam_from_host", "svb_bam_from_bgzf", "svb_bam_open",
    "svb_bam_free", "svb_bam_device_stream", "svb_bam_copy_stream", "svb_inflate_bgzf", "svb_bam_n_records", "svb_bam_record_bytes", "svb_bam_n_ref", "svb_bam_ref_name", "svb_bam_ref_len",
    "svb_bam_set_refs", "svb_getclip", "svb_clusters_free", "svb_clusters_count", "svb_clusters_candidates",
    "svb_clusters_text", "svb_insert_stats", "svb_discordant_support", "svb_window_depth", "svb_plan_getsv", "svb_free",
    "svb_write_gz", "svb_read_gz", "svb_bam_open_refs", "svb_bai_first_offsets", "svb_bam_last_mapped_tid", "svb_clusters_unmapped_records",
    "svb_clusters_gz", "svb_gzip_text", "svb_bam_open_voffsets", "svb_bai_linear_offsets", "svb_bam_peek_record", "svb_voffset_distance",
    "svb_sam_to_stream", "svb_main", "svb_getsv_passes", "svb_clusters_text_len", "svb_clusters_export_device",
    "svb_clusters_export_parts", "svb_bam_set_own_offset", "svb_insert_partial", "svb_insert_sq", "svb_pairs_depth", "svb_plan_somatic", "svb_insert_partial_async",
    "svb_write_range_blocks", "svb_set_shard_provider", "svb_clip_join", "svb_read_gz_device",
]


def load():
    """dlopen the CUDA library. Raises if it has not been built: there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise SvbError("%s is missing - run `python -m seeksv_b200.build` (no CPU fallback exists)" % p)
    L = C.CDLL(p)
    vp, u64, i32, i64 = C.c_void_p, C.c_uint64, C.c_int32, C.c_int64
    L.svb_abi_version.restype = C.c_int
    L.svb_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.svb_ctx_destroy.argtypes = [vp]
    L.svb_ctx_destroy.restype = None
    L.svb_last_error.argtypes = [vp]
    L.svb_last_error.restype = C.c_char_p
    L.svb_ctx_stream.argtypes = [vp]
    L.svb_ctx_stream.restype = vp
    L.svb_prof_enable.argtypes = [vp, C.c_int]
    L.svb_prof_enable.restype = None
    L.svb_prof_reset.argtypes = [vp]
    L.svb_prof_reset.restype = None
    L.svb_prof_read.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(i64),
                                C.POINTER(C.c_double)]
    L.svb_bam_from_device.argtypes = [vp, vp, u64, u64, i32, C.POINTER(vp)]
    L.svb_bam_from_host.argtypes = [vp, vp, u64, u64, i32, C.POINTER(vp)]
    L.svb_bam_from_bgzf.argtypes = [vp, vp, u64, C.c_int, C.POINTER(vp)]
    L.svb_bam_open.argtypes = [vp, C.c_char_p, C.c_int, C.POINTER(vp)]
    L.svb_bam_free.argtypes = [vp]
    L.svb_bam_free.restype = None
    L.svb_bam_device_stream.argtypes = [vp, C.POINTER(vp), C.POINTER(u64), C.POINTER(u64)]
    L.svb_bam_copy_stream.argtypes = [vp, vp, u64, u64]
    L.svb_inflate_bgzf.argtypes = [vp, vp, u64, vp, u64, C.POINTER(u64)]
    for f in ("svb_bam_n_records", "svb_bam_record_bytes"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = u64
    L.svb_bam_n_ref.argtypes = [vp]
    L.svb_bam_n_ref.restype = i32
    L.svb_bam_ref_name.argtypes = [vp, i32]
    L.svb_bam_ref_name.restype = C.c_char_p
    L.svb_bam_ref_len.argtypes = [vp, i32]
    L.svb_bam_ref_len.restype = C.c_uint32
    L.svb_bam_set_refs.argtypes = [vp, i32, C.POINTER(C.c_char_p), C.POINTER(C.c_uint32)]
    L.svb_getclip.argtypes = [vp, vp, C.POINTER(GetclipParams), C.POINTER(vp)]
    L.svb_clusters_free.argtypes = [vp]
    L.svb_clusters_free.restype = None
    L.svb_clusters_count.argtypes = [vp]
    L.svb_clusters_count.restype = u64
    L.svb_clusters_candidates.argtypes = [vp]
    L.svb_clusters_candidates.restype = u64
    L.svb_clusters_text.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(u64)]
    L.svb_clusters_text_len.argtypes = [vp, C.c_int, C.POINTER(u64)]
    L.svb_getsv_passes.argtypes = [vp, vp, C.POINTER(GetsvParams), C.POINTER(Junction), u64, C.POINTER(Window), u64, C.POINTER(i64),
                                   C.POINTER(i32), C.POINTER(i32)]
    L.svb_clusters_export_device.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
    L.svb_clusters_export_parts.argtypes = [vp, C.POINTER(u64), i32]
    L.svb_bam_set_own_offset.argtypes = [vp, u64]
    L.svb_insert_partial.argtypes = [vp, vp, i32, i64, C.POINTER(i64)]
    L.svb_insert_partial_async.argtypes = [vp, vp, i32, i64, vp]
    L.svb_insert_sq.argtypes = [vp, vp, i32, i64, i32, C.POINTER(i64)]
    L.svb_pairs_depth.argtypes = [vp, vp, C.POINTER(PairParams), C.POINTER(Junction), u64, C.POINTER(Window), u64, vp, vp]
    L.svb_gzip_text.argtypes = [vp, C.c_char_p, C.c_uint64, C.POINTER(vp), C.POINTER(u64)]
    L.svb_clusters_gz.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(u64)]
    L.svb_clusters_unmapped_records.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(u64)]
    L.svb_insert_stats.argtypes = [vp, vp, i32, i64, C.POINTER(i64)]
    L.svb_discordant_support.argtypes = [vp, vp, C.POINTER(Junction), u64, C.POINTER(PairParams), C.POINTER(i32)]
    L.svb_window_depth.argtypes = [vp, vp, C.POINTER(Window), u64, i32, C.POINTER(i32)]
    L.svb_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    L.svb_plan_getsv.argtypes = [C.c_char_p, C.c_char_p, i32, C.POINTER(C.c_char_p), C.POINTER(C.c_uint32), i32, i32,
                                 C.POINTER(C.POINTER(Junction)), C.POINTER(u64), C.POINTER(C.POINTER(Window)),
                                 C.POINTER(u64)]
    L.svb_plan_somatic.argtypes = [C.c_char_p, C.c_char_p, C.c_double, i32, i32, i32, i32, C.POINTER(C.c_char_p), C.POINTER(C.POINTER(Junction)),
                                   C.POINTER(u64)]
    L.svb_bam_open_refs.argtypes = [vp, C.c_char_p, C.c_char_p, i32, i32, C.c_int, C.POINTER(vp)]
    L.svb_bam_open_voffsets.argtypes = [vp, C.c_char_p, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(vp)]
    L.svb_bai_linear_offsets.argtypes = [C.c_char_p, i32, C.POINTER(C.c_uint64), C.c_int64]
    L.svb_bai_linear_offsets.restype = C.c_int64
    L.svb_bam_peek_record.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(i32), C.POINTER(i32)]
    L.svb_voffset_distance.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    L.svb_bam_last_mapped_tid.argtypes = [vp, vp, C.POINTER(i32), C.POINTER(i32)]
    L.svb_bai_first_offsets.argtypes = [C.c_char_p, C.POINTER(C.c_uint64), C.c_int64]
    L.svb_bai_first_offsets.restype = C.c_int64
    L.svb_write_gz.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_int]
    L.svb_read_gz.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.svb_sam_to_stream.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.svb_free.argtypes = [vp]
    L.svb_free.restype = None
    L.svb_write_range_blocks.argtypes = [C.c_char_p, vp, u64, vp, u64, C.c_int, C.POINTER(vp), C.POINTER(u64)]
    L.svb_set_shard_provider.argtypes = [vp, vp]
    L.svb_set_shard_provider.restype = None
    L.svb_read_gz_device.argtypes = [vp, C.c_char_p, C.POINTER(vp), C.POINTER(u64)]
    L.svb_clip_join.argtypes = [vp, vp, u64, vp, u64, vp, u64, vp, u64, vp, u64, C.POINTER(vp), C.POINTER(u64)]
    _lib = L
    return L


# int provider(junctions, n_j, windows, n_w, min_mapq, pairs_used, times, int64 stats[3], int32 counts[n_j], int32 depth[positions], user)
SHARD_PROVIDER = C.CFUNCTYPE(C.c_int, C.POINTER(Junction), C.c_uint64, C.POINTER(Window), C.c_uint64, C.c_int32, C.c_int32, C.c_int32,
                             C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p)


def set_shard_provider(fn):
    """svb_set_shard_provider: `fn(juncs, wins, min_mapq, pairs_used, times) -> (n, mean, dev, counts, depth)` is called once by the
    next `run_cli(["getsv", ...])` after its junction merge (sharded runs, seeksv_b200/mgpu.py); None unregisters. Returns the ctypes
    callback, which the caller has to keep alive while it is registered."""
    L = load()
    if fn is None:
        L.svb_set_shard_provider(None, None)
        return None

    def thunk(pj, nj, pw, nw, min_mapq, pairs_used, times, stats, counts, depth, _user):
        try:
            juncs = [(pj[i].up_tid, pj[i].up_pos, pj[i].up_strand.decode(), pj[i].down_tid, pj[i].down_pos, pj[i].down_strand.decode())
                     for i in range(nj)]
            wins = [(pw[i].tid, pw[i].begin, pw[i].end) for i in range(nw)]
            n, mean, dev, c, d = fn(juncs, wins, min_mapq, pairs_used, times)
            stats[0], stats[1], stats[2] = int(n), int(mean), int(dev)
            if len(c) != nj or len(d) != sum(w[2] - w[1] + 1 for w in wins):
                return 2
            import numpy as np
            c, d = np.ascontiguousarray(c, dtype=np.int32), np.ascontiguousarray(d, dtype=np.int32)
            if nj:
                C.memmove(counts, c.ctypes.data, 4 * nj)
            if len(d):
                C.memmove(depth, d.ctypes.data, 4 * len(d))
            return 0
        except Exception as e:      # noqa: BLE001 - an exception must not unwind through the C frames
            sys.stderr.write("[seeksv_b200] shard provider: %r\n" % (e,))
            return 1
    cb = SHARD_PROVIDER(thunk)
    L.svb_set_shard_provider(C.cast(cb, C.c_void_p), None)
    return cb


def write_range_blocks(part_prefix: str, clip: bytes, fq: bytes, threads: int = 0):
    """svb_write_range_blocks: this rank's clip / clip.fq texts as gzip files per (chromosome, side) block; returns [(chromosome, side)]"""
    L = load()
    out, n = C.c_void_p(), C.c_uint64()
    pc, pf = C.c_char_p(clip), C.c_char_p(fq)      # (pointers into the bytes objects: no copy)
    rc = L.svb_write_range_blocks(part_prefix.encode(), C.cast(pc, C.c_void_p), len(clip), C.cast(pf, C.c_void_p), len(fq), threads,
                                  C.byref(out), C.byref(n))
    if rc != 0:
        raise SvbError("svb_write_range_blocks = %d" % rc)
    try:
        text = C.string_at(out, n.value)
    finally:
        L.svb_free(out)
    return [tuple(line.split(b"\t")) for line in text.split(b"\n")[:-1]]


class Context:
    """One GPU (svb_ctx)."""

    def __init__(self, device: int = 0):
        self.L = load()
        self.h = C.c_void_p()
        rc = self.L.svb_ctx_create(device, C.byref(self.h))
        if rc != 0:
            raise SvbError("svb_ctx_create(%d) = %d: %s" % (device, rc, self.L.svb_last_error(None).decode()))

    def check(self, rc: int, what: str):
        if rc != 0:
            raise SvbError("%s = %d: %s" % (what, rc, self.L.svb_last_error(self.h).decode()))

    @property
    def stream(self) -> int:
        return self.L.svb_ctx_stream(self.h) or 0

    def prof(self, on: bool = True):
        self.L.svb_prof_enable(self.h, 1 if on else 0)

    def prof_reset(self):
        self.L.svb_prof_reset(self.h)

    def prof_read(self):
        cap = 64
        names = (C.c_char_p * cap)()
        ms = (C.c_double * cap)()
        n_l = (C.c_int64 * cap)()
        by = (C.c_double * cap)()
        n = self.L.svb_prof_read(self.h, cap, names, ms, n_l, by)
        return {names[i].decode(): dict(ms=ms[i], launches=n_l[i], bytes=by[i]) for i in range(min(n, cap))}

    def close(self):
        if self.h:
            self.L.svb_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def inflate_bgzf(ctx: "Context", image: bytes) -> bytes:
    """device-side BGZF inflate of an arbitrary image (tests of inflate.cu)"""
    n = C.c_uint64()
    src = (C.c_char * len(image)).from_buffer_copy(image)
    ctx.check(ctx.L.svb_inflate_bgzf(ctx.h, src, len(image), None, 0, C.byref(n)), "svb_inflate_bgzf(size)")
    out = (C.c_char * max(1, n.value))()
    ctx.check(ctx.L.svb_inflate_bgzf(ctx.h, src, len(image), out, n.value, C.byref(n)), "svb_inflate_bgzf")
    return bytes(out[:n.value])


class Clusters:
    """Result of svb_getclip, kept as a handle: the four texts (and, for shards, the exported unmapped-branch records) stay in HBM
    until they are asked for."""

    def __init__(self, ctx: "Context", handle):
        self.ctx, self.h = ctx, handle

    @property
    def n_clusters(self) -> int:
        return self.ctx.L.svb_clusters_count(self.h)

    def text_len(self, which: int) -> int:
        n = C.c_uint64()
        self.ctx.L.svb_clusters_text_len(self.h, which, C.byref(n))
        return n.value

    def text(self, which: int) -> bytes:
        d, n = C.c_char_p(), C.c_uint64()
        self.ctx.check(self.ctx.L.svb_clusters_text(self.h, which, C.byref(d), C.byref(n)), "svb_clusters_text")
        return C.string_at(d, n.value) if n.value else b""

    def gz(self, which: int) -> bytes:
        """gzip file image of output `which` (gz_outputs)"""
        d, n = C.c_char_p(), C.c_uint64()
        self.ctx.check(self.ctx.L.svb_clusters_gz(self.h, which, C.byref(d), C.byref(n)), "svb_clusters_gz")
        return C.string_at(d, n.value) if n.value else b""

    def export_device(self) -> Tuple[int, int]:
        """(device pointer, bytes) of the exported unmapped-branch records"""
        d, n = C.c_void_p(), C.c_uint64()
        self.ctx.check(self.ctx.L.svb_clusters_export_device(self.h, C.byref(d), C.byref(n)), "svb_clusters_export_device")
        return d.value or 0, n.value

    def export_parts(self, n_parts: int) -> List[int]:
        arr = (C.c_uint64 * (n_parts + 1))()
        self.ctx.check(self.ctx.L.svb_clusters_export_parts(self.h, arr, n_parts), "svb_clusters_export_parts")
        return list(arr)

    def unmapped_records(self) -> bytes:
        d, n = C.c_char_p(), C.c_uint64()
        self.ctx.L.svb_clusters_unmapped_records(self.h, C.byref(d), C.byref(n))
        return C.string_at(d, n.value) if n.value else b""

    def close(self):
        if self.h:
            if self.ctx.h:      # (a handle that outlived its context - e.g. after an exception - is dropped, not freed into a dead context)
                self.ctx.L.svb_clusters_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Bam:
    """A BAM (or shard) resident in HBM (svb_bam)."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self.h = ctx, handle
        self._keep = None

    @classmethod
    def open(cls, ctx: Context, path: str, threads: int = 0) -> "Bam":
        h = C.c_void_p()
        ctx.check(ctx.L.svb_bam_open(ctx.h, path.encode(), threads, C.byref(h)), "svb_bam_open(%s)" % path)
        return cls(ctx, h)

    @classmethod
    def open_refs(cls, ctx: Context, path: str, tid_begin: int, tid_end: int, bai: str = None, threads: int = 0) -> "Bam":
        """records of references [tid_begin, tid_end) of an indexed BAM (one rank's chromosome shard)"""
        h = C.c_void_p()
        ctx.check(ctx.L.svb_bam_open_refs(ctx.h, path.encode(), bai.encode() if bai else None, tid_begin, tid_end, threads, C.byref(h)),
                  "svb_bam_open_refs(%s, %d, %d)" % (path, tid_begin, tid_end))
        return cls(ctx, h)

    @classmethod
    def open_voffsets(cls, ctx: Context, path: str, v_begin: int, v_end: Optional[int], threads: int = 0) -> "Bam":
        """records between two BGZF virtual offsets (record boundaries from the .bai); v_begin 0 = first record, v_end None = EOF"""
        h = C.c_void_p()
        ctx.check(ctx.L.svb_bam_open_voffsets(ctx.h, path.encode(), v_begin, 2 ** 64 - 1 if v_end is None else v_end, threads, C.byref(h)),
                  "svb_bam_open_voffsets(%s)" % path)
        return cls(ctx, h)

    def last_mapped_tid(self):
        """tid of the last mapped-branch record (None if the shard has none): the next shard's prev_tid"""
        has, tid = C.c_int32(), C.c_int32()
        self.ctx.check(self.ctx.L.svb_bam_last_mapped_tid(self.ctx.h, self.h, C.byref(has), C.byref(tid)), "svb_bam_last_mapped_tid")
        return tid.value if has.value else None

    @classmethod
    def from_bgzf(cls, ctx: Context, file_bytes, threads: int = 0) -> "Bam":
        """file image of a .bam in host memory (bytes / numpy uint8 / pinned torch tensor pointer+len tuple)"""
        ptr, n, keep = _host_ptr(file_bytes)
        h = C.c_void_p()
        ctx.check(ctx.L.svb_bam_from_bgzf(ctx.h, ptr, n, threads, C.byref(h)), "svb_bam_from_bgzf")
        b = cls(ctx, h)
        return b

    @classmethod
    def from_host(cls, ctx: Context, stream, first_record: int, n_ref: int) -> "Bam":
        ptr, n, keep = _host_ptr(stream)
        h = C.c_void_p()
        ctx.check(ctx.L.svb_bam_from_host(ctx.h, ptr, n, first_record, n_ref, C.byref(h)), "svb_bam_from_host")
        return cls(ctx, h)

    @classmethod
    def from_device(cls, ctx: Context, dptr: int, nbytes: int, first_record: int, n_ref: int, keep=None) -> "Bam":
        h = C.c_void_p()
        ctx.check(ctx.L.svb_bam_from_device(ctx.h, C.c_void_p(dptr), nbytes, first_record, n_ref, C.byref(h)),
                  "svb_bam_from_device")
        b = cls(ctx, h)
        b._keep = keep
        return b

    def set_refs(self, names: Sequence[str], lengths: Sequence[int]):
        n = len(names)
        arr = (C.c_char_p * n)(*[s.encode() for s in names])
        ln = (C.c_uint32 * n)(*lengths)
        self.ctx.check(self.ctx.L.svb_bam_set_refs(self.h, n, arr, ln), "svb_bam_set_refs")

    def device_stream(self) -> Tuple[int, int, int]:
        """(device pointer, nbytes, first_record) of the resident uncompressed stream"""
        d, n, f = C.c_void_p(), C.c_uint64(), C.c_uint64()
        self.ctx.check(self.ctx.L.svb_bam_device_stream(self.h, C.byref(d), C.byref(n), C.byref(f)), "svb_bam_device_stream")
        return d.value or 0, n.value, f.value

    def copy_stream(self) -> bytes:
        """the whole resident uncompressed stream, copied back to the host (tests)"""
        _, n, _ = self.device_stream()
        buf = (C.c_char * n)()
        self.ctx.check(self.ctx.L.svb_bam_copy_stream(self.h, buf, 0, n), "svb_bam_copy_stream")
        return bytes(buf)

    @property
    def n_records(self) -> int:
        return self.ctx.L.svb_bam_n_records(self.h)

    @property
    def record_bytes(self) -> int:
        return self.ctx.L.svb_bam_record_bytes(self.h)

    @property
    def ref_names(self) -> List[str]:
        L = self.ctx.L
        return [L.svb_bam_ref_name(self.h, t).decode() for t in range(L.svb_bam_n_ref(self.h))]

    @property
    def ref_lens(self) -> List[int]:
        L = self.ctx.L
        return [L.svb_bam_ref_len(self.h, t) for t in range(L.svb_bam_n_ref(self.h))]

    def getclip(self, match_rate=0.9, min_mapq=1, save_low_quality=False, prev_tid=0, export_unmapped=False, key_range=None,
                halo_bytes=0, with_rows=False):
        """(clip, clip.fq, unmapped_1, unmapped_2) decompressed file contents. export_unmapped (shards): the unmapped branch is
        not paired here; its packed records are left in self.last_unmapped_records for the merging rank."""
        p = GetclipParams(match_rate, min_mapq, 1 if save_low_quality else 0, prev_tid, 1 if export_unmapped else 0)
        if key_range is not None:   # ((lo_tid, lo_pos), (hi_tid, hi_pos)): breakpoint keys owned by this range shard
            (p.key_lo_tid, p.key_lo_pos), (p.key_hi_tid, p.key_hi_pos) = key_range
            p.key_filter = 1
        p.halo_bytes = halo_bytes
        p.with_rows = 1 if with_rows else 0
        out = C.c_void_p()
        self.ctx.check(self.ctx.L.svb_getclip(self.ctx.h, self.h, C.byref(p), C.byref(out)), "svb_getclip")
        try:
            res = []
            for which in range(4):
                d = C.c_char_p()
                n = C.c_uint64()
                self.ctx.L.svb_clusters_text(out, which, C.byref(d), C.byref(n))
                res.append(C.string_at(d, n.value) if n.value else b"")
            d, n = C.c_char_p(), C.c_uint64()
            self.ctx.L.svb_clusters_unmapped_records(out, C.byref(d), C.byref(n))
            self.last_unmapped_records = C.string_at(d, n.value) if n.value else b""
            self.last_clusters = self.ctx.L.svb_clusters_count(out)
            self.last_candidates = self.ctx.L.svb_clusters_candidates(out)
            return tuple(res)
        finally:
            self.ctx.L.svb_clusters_free(out)

    def getclip_gz(self, match_rate=0.9, min_mapq=1, save_low_quality=False, prev_tid=0):
        """the four outputs as gzip file images, compressed on the device (what the CLI writes)"""
        p = GetclipParams(match_rate, min_mapq, 1 if save_low_quality else 0, prev_tid)
        p.gz_outputs = 1
        out = C.c_void_p()
        self.ctx.check(self.ctx.L.svb_getclip(self.ctx.h, self.h, C.byref(p), C.byref(out)), "svb_getclip")
        try:
            res = []
            for which in range(4):
                d, n = C.c_char_p(), C.c_uint64()
                self.ctx.L.svb_clusters_gz(out, which, C.byref(d), C.byref(n))
                res.append(C.string_at(d, n.value) if n.value else b"")
            return tuple(res)
        finally:
            self.ctx.L.svb_clusters_free(out)

    def getclip_sizes(self, match_rate=0.9, min_mapq=1, save_low_quality=False, prev_tid=0, gz=False, with_rows=False,
                      fetch=False) -> Tuple[int, int, int, int]:
        """svb_getclip without turning the four texts into Python objects: returns their lengths. The texts stay in HBM unless
        fetch is set (then they are copied to pinned host memory, as svb_clusters_text does on request)."""
        p = GetclipParams(match_rate, min_mapq, 1 if save_low_quality else 0, prev_tid)
        p.gz_outputs = 1 if gz else 0
        p.with_rows = 1 if with_rows else 0
        out = C.c_void_p()
        self.ctx.check(self.ctx.L.svb_getclip(self.ctx.h, self.h, C.byref(p), C.byref(out)), "svb_getclip")
        try:
            res = []
            for which in range(4):
                d = C.c_char_p()
                n = C.c_uint64()
                if gz:
                    self.ctx.L.svb_clusters_gz(out, which, C.byref(d), C.byref(n))
                elif fetch:
                    self.ctx.L.svb_clusters_text(out, which, C.byref(d), C.byref(n))
                else:
                    self.ctx.L.svb_clusters_text_len(out, which, C.byref(n))
                res.append(n.value)
            self.last_clusters = self.ctx.L.svb_clusters_count(out)
            return tuple(res)
        finally:
            self.ctx.L.svb_clusters_free(out)

    def getclip_handle(self, match_rate=0.9, min_mapq=1, save_low_quality=False, prev_tid=0, export_unmapped=False, key_range=None,
                       halo_bytes=0, with_rows=False, unmapped_only=False, export_partitions=0, gz_outputs=False) -> Clusters:
        """svb_getclip, results left in HBM behind a Clusters handle (gz_outputs: the four outputs are compressed on the device,
        Clusters.gz gives the file images)"""
        p = GetclipParams(match_rate, min_mapq, 1 if save_low_quality else 0, prev_tid, 1 if export_unmapped else 0)
        if key_range is not None:
            (p.key_lo_tid, p.key_lo_pos), (p.key_hi_tid, p.key_hi_pos) = key_range
            p.key_filter = 1
        p.halo_bytes, p.with_rows = halo_bytes, 1 if with_rows else 0
        p.unmapped_only, p.export_partitions = 1 if unmapped_only else 0, export_partitions
        p.gz_outputs = 1 if gz_outputs else 0
        out = C.c_void_p()
        self.ctx.check(self.ctx.L.svb_getclip(self.ctx.h, self.h, C.byref(p), C.byref(out)), "svb_getclip")
        return Clusters(self.ctx, out)

    def set_own_offset(self, own_offset: int):
        self.ctx.check(self.ctx.L.svb_bam_set_own_offset(self.h, own_offset), "svb_bam_set_own_offset")

    def insert_partial(self, min_mapq=20, take=-1):
        """{taken, sum, sum of squares, records above 46340} over the first `take` qualifying own records (additive over shards)"""
        out = (C.c_int64 * 4)()
        self.ctx.check(self.ctx.L.svb_insert_partial(self.ctx.h, self.h, min_mapq, take, out), "svb_insert_partial")
        return tuple(out)

    def insert_partial_async(self, min_mapq, take, device_ptr: int):
        """svb_insert_partial_async: the four values go to device memory at device_ptr, in stream order (no read-back)"""
        self.ctx.check(self.ctx.L.svb_insert_partial_async(self.ctx.h, self.h, min_mapq, take, C.c_void_p(device_ptr)), "svb_insert_partial_async")

    def insert_sq(self, min_mapq, take, mean) -> int:
        out = C.c_int64()
        self.ctx.check(self.ctx.L.svb_insert_sq(self.ctx.h, self.h, min_mapq, take, mean, C.byref(out)), "svb_insert_sq")
        return out.value

    def pairs_depth_raw(self, pair_params, junction_array, n_j, window_array, n_w, counts_ptr, depth_ptr):
        """svb_pairs_depth; counts_ptr / depth_ptr are integers (host or DEVICE addresses)"""
        self.ctx.check(self.ctx.L.svb_pairs_depth(self.ctx.h, self.h, C.byref(pair_params), junction_array, n_j, window_array, n_w,
                                                  C.c_void_p(counts_ptr), C.c_void_p(depth_ptr)), "svb_pairs_depth")

    def getsv_passes_raw(self, params, junction_array, n_j, window_array, n_w, stats_array, counts_array, depth_array):
        """svb_getsv_passes on prebuilt C arrays (bench.py): insert-size statistics, pair support and window depth, one read-back"""
        self.ctx.check(self.ctx.L.svb_getsv_passes(self.ctx.h, self.h, C.byref(params), junction_array, n_j, window_array, n_w, stats_array,
                                                   counts_array, depth_array), "svb_getsv_passes")

    def getsv_passes(self, junctions, windows, min_mapq=20, max_pairs=5000000, times=4):
        """(stats[4], counts per junction, depth lists per window) from one fused call"""
        nj, nw = len(junctions), len(windows)
        j_arr = (Junction * max(nj, 1))(*[Junction(ut, up, dt, dp, us.encode(), ds.encode(), b"") for ut, up, us, dt, dp, ds in junctions])
        w_arr = (Window * max(nw, 1))(*[Window(*w) for w in windows])
        tot = sum(w[2] - w[1] + 1 for w in windows)
        st, cnt, dep = (C.c_int64 * 4)(), (C.c_int32 * max(nj, 1))(), (C.c_int32 * max(tot, 1))()
        self.getsv_passes_raw(GetsvParams(min_mapq, times, max_pairs), j_arr, nj, w_arr, nw, st, cnt, dep)
        res, o = [], 0
        for w in windows:
            k = w[2] - w[1] + 1
            res.append(list(dep[o:o + k]))
            o += k
        return tuple(st), list(cnt[:nj]), res

    def discordant_support_raw(self, junction_array, n, pair_params, counts_array):
        self.ctx.check(self.ctx.L.svb_discordant_support(self.ctx.h, self.h, junction_array, n, C.byref(pair_params), counts_array),
                       "svb_discordant_support")

    def window_depth_raw(self, window_array, n, min_mapq, depth_array):
        self.ctx.check(self.ctx.L.svb_window_depth(self.ctx.h, self.h, window_array, n, min_mapq, depth_array), "svb_window_depth")

    def insert_stats(self, min_mapq=20, max_pairs=5000000):
        out = (C.c_int64 * 4)()
        self.ctx.check(self.ctx.L.svb_insert_stats(self.ctx.h, self.h, min_mapq, max_pairs, out), "svb_insert_stats")
        return tuple(out)

    def discordant_support(self, junctions, min_mapq, mean, dev, times=4) -> List[int]:
        """junctions: (up_tid, up_pos, up_strand, down_tid, down_pos, down_strand)"""
        n = len(junctions)
        arr = (Junction * max(n, 1))()
        for i, (ut, up, us, dt, dp, ds) in enumerate(junctions):
            arr[i] = Junction(ut, up, dt, dp, us.encode(), ds.encode(), b"")
        cnt = (C.c_int32 * max(n, 1))()
        pp = PairParams(min_mapq, mean, dev, times)
        self.ctx.check(self.ctx.L.svb_discordant_support(self.ctx.h, self.h, arr, n, C.byref(pp), cnt),
                       "svb_discordant_support")
        return list(cnt[:n])

    def window_depth(self, windows, min_mapq) -> List[List[int]]:
        """windows: sorted disjoint (tid, begin1, end1); returns per-window depth lists"""
        n = len(windows)
        arr = (Window * max(n, 1))(*[Window(*w) for w in windows])
        tot = sum(w[2] - w[1] + 1 for w in windows)
        out = (C.c_int32 * max(tot, 1))()
        self.ctx.check(self.ctx.L.svb_window_depth(self.ctx.h, self.h, arr, n, min_mapq, out), "svb_window_depth")
        res, o = [], 0
        for w in windows:
            k = w[2] - w[1] + 1
            res.append(list(out[o:o + k]))
            o += k
        return res

    def close(self):
        if self.h:
            if self.ctx.h:      # (see Clusters.close)
                self.ctx.L.svb_bam_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _host_ptr(obj):
    """(void*, nbytes, keepalive) for bytes / bytearray / numpy arrays / torch CPU tensors"""
    if isinstance(obj, bytes):       # a pointer into the bytes object itself: no copy (the callee only reads)
        p = C.c_char_p(obj)
        return C.cast(p, C.c_void_p), len(obj), (p, obj)
    if isinstance(obj, bytearray):
        buf = (C.c_char * len(obj)).from_buffer(obj)
        return C.cast(buf, C.c_void_p), len(obj), buf
    if hasattr(obj, "data_ptr"):     # torch tensor
        return C.c_void_p(obj.data_ptr()), obj.numel() * obj.element_size(), obj
    if hasattr(obj, "ctypes"):       # numpy
        return C.c_void_p(obj.ctypes.data), obj.nbytes, obj
    raise TypeError("unsupported host buffer")


def plan_getsv(clip_alignments: str, clip_file: str, ref_names: Sequence[str], ref_lens: Sequence[int], reach: int = 50,
               flank_len: int = 200):
    """Host-side planning of getsv (svb_plan_getsv): junction tuples in output order and merged depth windows."""
    L = load()
    n = len(ref_names)
    names = (C.c_char_p * n)(*[s.encode() for s in ref_names])
    lens = (C.c_uint32 * n)(*ref_lens)
    pj, pw = C.POINTER(Junction)(), C.POINTER(Window)()
    nj, nw = C.c_uint64(), C.c_uint64()
    rc = L.svb_plan_getsv(clip_alignments.encode(), clip_file.encode(), n, names, lens, reach, flank_len, C.byref(pj),
                          C.byref(nj), C.byref(pw), C.byref(nw))
    if rc != 0:
        raise SvbError("svb_plan_getsv = %d" % rc)
    try:
        juncs = [(pj[i].up_tid, pj[i].up_pos, pj[i].up_strand.decode(), pj[i].down_tid, pj[i].down_pos,
                  pj[i].down_strand.decode()) for i in range(nj.value)]
        wins = [(pw[i].tid, pw[i].begin, pw[i].end) for i in range(nw.value)]
    finally:
        L.svb_free(pj)
        L.svb_free(pw)
    return juncs, wins


JOIN_LINE_BYTES, JOIN_ALN_BYTES, JOIN_CAND_BYTES = 20, 32, 28     # sizeof svb_join_line / svb_join_aln / svb_join_cand


def clip_join_raw(ctx: "Context", lines: bytes, seqs: bytes, alns: bytes, names: bytes, cigars: bytes) -> bytes:
    """svb_clip_join on packed arrays (the structs of include/seeksv_b200.h as raw little-endian bytes): the junction candidates
    of the device join, sorted, as raw svb_join_cand bytes."""
    L = load()
    out, n = C.c_void_p(), C.c_uint64()
    assert len(lines) % JOIN_LINE_BYTES == 0 and len(alns) % JOIN_ALN_BYTES == 0 and len(cigars) % 4 == 0
    rc = L.svb_clip_join(ctx.h, lines, len(lines) // JOIN_LINE_BYTES, seqs, len(seqs), alns, len(alns) // JOIN_ALN_BYTES, names, len(names),
                         cigars, len(cigars) // 4, C.byref(out), C.byref(n))
    ctx.check(rc, "svb_clip_join")
    try:
        return C.string_at(out, n.value * JOIN_CAND_BYTES)
    finally:
        L.svb_free(out)


def plan_somatic(normal_clip: str, tumor_sv: str, ref_names: Sequence[str], match_rate=0.9, offset=30, min_len=10, mean_insert=0):
    """junction tuples `somatic` asks the normal BAM about (svb_plan_somatic), in the command's order"""
    L = load()
    n = len(ref_names)
    names = (C.c_char_p * n)(*[s.encode() for s in ref_names])
    pj, nj = C.POINTER(Junction)(), C.c_uint64()
    rc = L.svb_plan_somatic(normal_clip.encode(), tumor_sv.encode(), match_rate, offset, min_len, mean_insert, n, names, C.byref(pj), C.byref(nj))
    if rc != 0:
        raise SvbError("svb_plan_somatic = %d" % rc)
    try:
        return [(pj[i].up_tid, pj[i].up_pos, pj[i].up_strand.decode(), pj[i].down_tid, pj[i].down_pos, pj[i].down_strand.decode())
                for i in range(nj.value)]
    finally:
        L.svb_free(pj)


def gzip_text(ctx: "Context", data: bytes) -> bytes:
    """gzip file image of `data`, compressed on the device (gzip.cu)"""
    p, n = C.c_void_p(), C.c_uint64()
    ctx.check(ctx.L.svb_gzip_text(ctx.h, data, len(data), C.byref(p), C.byref(n)), "svb_gzip_text")
    try:
        return C.string_at(p, n.value)
    finally:
        ctx.L.svb_free(p)


def bai_first_offsets(bai_path: str):
    """BGZF virtual offset of the first record of every reference (None where the reference has no records); host only"""
    L = load()
    n = L.svb_bai_first_offsets(bai_path.encode(), None, 0)
    if n < 0:
        raise SvbError("svb_bai_first_offsets(%s) = %d" % (bai_path, n))
    arr = (C.c_uint64 * max(1, n))()
    L.svb_bai_first_offsets(bai_path.encode(), arr, n)
    return [None if arr[i] == 2 ** 64 - 1 else int(arr[i]) for i in range(n)]


def bai_linear_offsets(bai_path: str, tid: int):
    """linear index of one reference: virtual offset of the first record overlapping each 16 kb window (0 = none); host only"""
    L = load()
    n = L.svb_bai_linear_offsets(bai_path.encode(), tid, None, 0)
    if n < 0:
        raise SvbError("svb_bai_linear_offsets(%s, %d) = %d" % (bai_path, tid, n))
    arr = (C.c_uint64 * max(1, n))()
    L.svb_bai_linear_offsets(bai_path.encode(), tid, arr, n)
    return [int(arr[i]) for i in range(n)]


def peek_record(bam_path: str, voffset: int):
    """(tid, 0-based pos) of the record at a virtual offset; host only"""
    tid, pos = C.c_int32(), C.c_int32()
    rc = load().svb_bam_peek_record(bam_path.encode(), voffset, C.byref(tid), C.byref(pos))
    if rc != 0:
        raise SvbError("svb_bam_peek_record(%s, %d) = %d" % (bam_path, voffset, rc))
    return tid.value, pos.value


def voffset_distance(bam_path: str, v_a: int, v_b: int) -> int:
    """uncompressed bytes between two virtual offsets; host only"""
    n = C.c_uint64()
    rc = load().svb_voffset_distance(bam_path.encode(), v_a, v_b, C.byref(n))
    if rc != 0:
        raise SvbError("svb_voffset_distance = %d" % rc)
    return n.value


def write_gz(path: str, data: bytes, threads: int = 0) -> None:
    """multi-member gzip writer of the CLI outputs (svb_write_gz); no GPU involved"""
    rc = load().svb_write_gz(path.encode(), data, len(data), threads)
    if rc != 0:
        raise SvbError("svb_write_gz(%s) = %d" % (path, rc))


def sam_to_stream(path: str):
    """SAM text file -> (uncompressed BAM byte stream, offset of the first record): the host-side conversion behind Bam.open / getsv
    for inputs that are not .bam (svb_sam_to_stream)"""
    L = load()
    p, n, first = C.c_void_p(), C.c_uint64(), C.c_uint64()
    rc = L.svb_sam_to_stream(path.encode(), C.byref(p), C.byref(n), C.byref(first))
    if rc != 0:
        raise SvbError("svb_sam_to_stream(%s) = %d" % (path, rc))
    try:
        return C.string_at(p, n.value), first.value
    finally:
        L.svb_free(p)


def read_gz_device(ctx: "Context", path: str) -> bytes:
    """svb_read_gz_device: a gzip file of the device writer (members of <= 64 KiB of text), inflated on the GPU"""
    L = load()
    p, n = C.c_void_p(), C.c_uint64()
    ctx.check(L.svb_read_gz_device(ctx.h, path.encode(), C.byref(p), C.byref(n)), "svb_read_gz_device(%s)" % path)
    return C.string_at(p, n.value)


def read_gz(path: str) -> bytes:
    """reader of gzip / plain text inputs (svb_read_gz); member-parallel for files written by write_gz"""
    L = load()
    p, n = C.c_void_p(), C.c_uint64()
    rc = L.svb_read_gz(path.encode(), C.byref(p), C.byref(n))
    if rc != 0:
        raise SvbError("svb_read_gz(%s) = %d" % (path, rc))
    try:
        return C.string_at(p, n.value)
    finally:
        L.svb_free(p)


def run_cli(args: Sequence[str]) -> int:
    """svb_main in-process (same as executing bin/seeksv)."""
    L = load()
    argv = [b"seeksv"] + [a.encode() for a in args]
    arr = (C.c_char_p * (len(argv) + 1))(*argv, None)
    return L.svb_main(len(argv), arr)
