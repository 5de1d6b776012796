"""CPU-only checks of the C-ABI boundary: the library builds for sm_100a without a GPU, loads, exports every symbol that
include/seeksv_b200.h declares (and the ctypes mirror binds them all), refuses to run without a device (no CPU fallback),
and the CLI keeps the reference's argument handling (seeksv.cpp:26-58,128-155,366-410)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import GOLDEN, ROOT, read_text


@pytest.fixture(scope="module")
def lib():
    from seeksv_b200 import build, lib_path
    build.build()
    return ctypes.CDLL(lib_path())


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "seeksv_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(svb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    from seeksv_b200 import lib as pylib
    assert sorted(pylib.EXPORTS) == names, "seeksv_b200/lib.py must bind exactly the header's entry points"


def test_abi_version_and_no_cpu_fallback(lib):
    assert lib.svb_abi_version() == 2
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = ctypes.c_void_p()
    rc = lib.svb_ctx_create(0, ctypes.byref(ctx))
    assert rc == -1 and not ctx.value          # SVB_ERR_NO_DEVICE: nothing can run without the CUDA device
    lib.svb_last_error.restype = ctypes.c_char_p
    assert b"no CPU fallback" in lib.svb_last_error(None)


def test_product_never_touches_the_oracle():
    """no import / include / exec of anything under oracle/ in the product tree (comments may mention the tests)"""
    pat = re.compile(r"(^\s*(from|import)\s+oracle)|(#include\s*[\"<][^\n]*oracle)|(oracle/_ref)|(oracle\.)", re.M)
    for d, _, files in os.walk(os.path.join(ROOT, "seeksv_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f), errors="ignore").read()
                assert not pat.search(src), (d, f)


def _cli(args):
    from seeksv_b200 import cli_path
    return subprocess.run([cli_path()] + args, capture_output=True, text=True)


def test_cli_argument_surface(lib):
    r = _cli([])
    assert r.returncode == 1 and "Usage: seeksv <command> [options]" in r.stderr
    r = _cli(["frobnicate"])
    assert r.returncode == 1 and "[seeksv] unrecognized command 'frobnicate'" in r.stderr
    for cmd, frag in (("getclip", "<input.sorted.bam>"), ("getsv", "<output SVs>"), ("somatic", "<input tumor SV file>")):
        r = _cli([cmd])
        assert r.returncode == 1 and frag in r.stderr
    assert _cli(["getsv", "-l", "91", "a", "b", "c", "d", "e"]).returncode == 1      # -l must be 0..90 (seeksv.cpp:191)
    assert _cli(["somatic", "-l", "90", "a", "b", "c", "d"]).returncode == 1         # -l must be 0..89 (seeksv.cpp:386)
    assert _cli(["cluster", "x"]).returncode == 0                                    # recognised, dispatch disabled


@pytest.mark.parametrize("d,s", [("example", "cancer"), ("example", "normal"), ("micro", "tumor"), ("micro", "normal"),
                                 ("fuzz", "f11"), ("fuzz", "f12"), ("fuzz", "f106"), ("fuzz", "e3"), ("long", "lq")])
def test_getsv_host_only_mode_matches_reference(lib, d, s, tmp_path):
    """`getsv -n 0 -D` needs no BAM pass (no insert size, no pairs, no depth): join + merge + filters + formatting of the
    host layer against the reference binary's output - runs without a GPU."""
    import gzip
    clip = str(tmp_path / "clip.gz")
    with gzip.open(clip, "wb") as f:
        f.write(read_text(os.path.join(GOLDEN, d, s + ".clip.txt")).encode("latin-1"))
    out = str(tmp_path / "out.sv")
    r = _cli(["getsv", "-n", "0", "-D", os.path.join(GOLDEN, d, s + ".clip.sam"), os.path.join(GOLDEN, d, s + ".sort.bam"), clip, out,
              str(tmp_path / "unm")])
    assert r.returncode == 0, r.stderr
    assert read_text(out) == read_text(os.path.join(GOLDEN, d, s + ".n0D.sv"))
    assert r.stdout == read_text(os.path.join(GOLDEN, d, s + ".n0D.stdout"))


def test_run_command_chains_segments_and_stops_on_failure(lib, tmp_path):
    """`seeksv run -- a -- b -- c`: seeksv commands run in-process, other segments are executed and waited for; the chain
    stops at the first non-zero status. (Host-only getsv mode, so no GPU is needed here.)"""
    import gzip
    d, s = "micro", "tumor"
    clip = str(tmp_path / "clip.gz")
    with gzip.open(clip, "wb") as f:
        f.write(read_text(os.path.join(GOLDEN, d, s + ".clip.txt")).encode("latin-1"))
    out, marker = str(tmp_path / "out.sv"), str(tmp_path / "marker")
    getsv = ["getsv", "-n", "0", "-D", os.path.join(GOLDEN, d, s + ".clip.sam"), os.path.join(GOLDEN, d, s + ".sort.bam"), clip, out,
             str(tmp_path / "unm")]
    r = _cli(["run", "--", "sh", "-c", "echo one > " + marker, "--", "echo two >> " + marker, "--"] + getsv)
    assert r.returncode == 0, r.stderr
    assert open(marker).read() == "one\ntwo\n"
    assert read_text(out) == read_text(os.path.join(GOLDEN, d, s + ".n0D.sv"))
    os.remove(out)
    r = _cli(["run", "--", "exit 3", "--"] + getsv)
    assert r.returncode == 3 and not os.path.exists(out)
    r = _cli(["run"])
    assert r.returncode == 1


@pytest.mark.parametrize("d,s", [("example", "cancer"), ("micro", "tumor"), ("fuzz", "f11"), ("fuzz", "f12")])
def test_bai_first_offsets_point_at_the_first_record_of_each_reference(lib, d, s):
    """the .bai reader behind svb_bam_open_refs (chromosome shards): its virtual offsets against a walk of the BAM itself"""
    import struct
    import zlib
    import seeksv_b200.lib as L
    path = os.path.join(GOLDEN, d, s + ".sort.bam")
    raw = open(path, "rb").read()
    blocks, o, u = [], 0, 0          # (compressed offset, uncompressed offset, uncompressed length)
    stream = bytearray()
    while o < len(raw):
        bsize = struct.unpack_from("<H", raw, o + 16)[0] + 1
        data = zlib.decompress(raw[o + 18:o + bsize - 8], -15)
        blocks.append((o, u, len(data)))
        stream += data
        u += len(data)
        o += bsize
    l_text = struct.unpack_from("<i", stream, 4)[0]
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", stream, p)[0]
    p += 4
    for _ in range(n_ref):
        p += 8 + struct.unpack_from("<i", stream, p)[0]
    want = [None] * n_ref
    while p + 4 <= len(stream):
        bs, tid = struct.unpack_from("<ii", stream, p)
        if 0 <= tid < n_ref and want[tid] is None:
            co, uo, _ = [b for b in blocks if b[1] <= p < b[1] + b[2]][0]
            want[tid] = co << 16 | (p - uo)
        p += 4 + bs
    assert L.bai_first_offsets(path + ".bai") == want


@pytest.mark.parametrize("d,s", [("micro", "tumor"), ("fuzz", "f11"), ("example", "cancer")])
def test_getsv_seed_file_host_only_matches_reference(lib, d, s, tmp_path):
    """`getsv -B <earlier output> -n 0 -D`: ReadBreakpoint of the host layer against the reference binary's output (no GPU)"""
    import gzip
    clip = str(tmp_path / "clip.gz")
    with gzip.open(clip, "wb") as f:
        f.write(read_text(os.path.join(GOLDEN, d, s + ".clip.txt")).encode("latin-1"))
    out = str(tmp_path / "out.sv")
    r = _cli(["getsv", "-B", os.path.join(GOLDEN, d, s + ".sv"), "-n", "0", "-D", os.path.join(GOLDEN, d, s + ".clip.sam"),
              os.path.join(GOLDEN, d, s + ".sort.bam"), clip, out, str(tmp_path / "unm")])
    assert r.returncode == 0, r.stderr
    assert read_text(out) == read_text(os.path.join(GOLDEN, d, s + ".B.n0D.sv"))
    assert r.stdout == read_text(os.path.join(GOLDEN, d, s + ".B.n0D.stdout"))


@pytest.mark.parametrize("d,s", [("micro", "tumor"), ("fuzz", "f11")])
def test_getsv_connected_reads_host_only_matches_reference(lib, d, s, tmp_path):
    """`getsv -F <connected reads> [-B ... -w 30] -n 0 -D`: FindJunction of the host layer against the reference binary (no GPU)"""
    import gzip
    clip = str(tmp_path / "clip.gz")
    with gzip.open(clip, "wb") as f:
        f.write(read_text(os.path.join(GOLDEN, d, s + ".clip.txt")).encode("latin-1"))
    connect = os.path.join(GOLDEN, d, s + ".connect.sam")
    tail = [os.path.join(GOLDEN, d, s + ".clip.sam"), os.path.join(GOLDEN, d, s + ".sort.bam"), clip]
    for tag, extra in ((".F.n0D", ["-F", connect]), (".FB.n0D", ["-F", connect, "-B", os.path.join(GOLDEN, d, s + ".sv"), "-w", "30"])):
        out = str(tmp_path / (tag + ".sv"))
        r = _cli(["getsv"] + extra + ["-n", "0", "-D"] + tail + [out, str(tmp_path / "unm")])
        assert r.returncode == 0, r.stderr
        assert r.stdout == read_text(os.path.join(GOLDEN, d, s + tag + ".stdout")), tag
        assert read_text(out) == read_text(os.path.join(GOLDEN, d, s + tag + ".sv")), tag
    r = _cli(["getsv", "-F", str(tmp_path / "missing.sam"), "-n", "0", "-D"] + tail + [str(tmp_path / "x.sv"), str(tmp_path / "unm")])
    assert r.returncode == 1 and "fail to open" in r.stderr


@pytest.mark.parametrize("d,s", [("example", "cancer"), ("micro", "tumor"), ("fuzz", "f11"), ("long", "lq")])
def test_getsv_reads_the_realigned_clips_as_bam_too(lib, d, s, tmp_path):
    """The reference's usual hand-off is `samtools view -Sb clip.sam > clip.bam` (README.md:30-31): a file name ending in .bam is
    read as BGZF BAM (getsv.h:437-441). Same host-only run as above with the alignments converted to BAM (the bytes libbam's own
    SAM reader produces, pinned in tests/test_sam_text.py)."""
    import gzip
    from oracle import bamio
    clip_bam = str(tmp_path / "clip.bam")
    with open(clip_bam, "wb") as f:
        f.write(bamio.bgzf_compress(bamio.sam_to_stream(os.path.join(GOLDEN, d, s + ".clip.sam"))))
    clip = str(tmp_path / "clip.gz")
    with gzip.open(clip, "wb") as f:
        f.write(read_text(os.path.join(GOLDEN, d, s + ".clip.txt")).encode("latin-1"))
    out = str(tmp_path / "out.sv")
    r = _cli(["getsv", "-n", "0", "-D", clip_bam, os.path.join(GOLDEN, d, s + ".sort.bam"), clip, out, str(tmp_path / "unm")])
    assert r.returncode == 0, r.stderr
    assert read_text(out) == read_text(os.path.join(GOLDEN, d, s + ".n0D.sv"))
    assert r.stdout == read_text(os.path.join(GOLDEN, d, s + ".n0D.stdout"))


def test_usage_texts_are_byte_identical_to_the_reference():
    """the usage texts and exit codes are part of the CLI contract. tests/golden/usage/* holds what the reference
    binary prints for `seeksv`, `seeksv getclip`, `seeksv getsv`, `seeksv somatic` (oracle/_ref/seeksv, checked live when it is built)."""
    import subprocess
    exe = os.path.join(ROOT, "seeksv_b200", "bin", "seeksv")
    ref = os.path.join(ROOT, "oracle", "_ref", "seeksv")
    gold = os.path.join(ROOT, "tests", "golden", "usage")
    for name, argv in (("top", []), ("getclip", ["getclip"]), ("getsv", ["getsv"]), ("somatic", ["somatic"])):
        r = subprocess.run([exe] + argv, capture_output=True)
        assert r.stdout == open(os.path.join(gold, name + ".stdout"), "rb").read(), name
        assert r.stderr == open(os.path.join(gold, name + ".stderr"), "rb").read(), name
        assert r.returncode == int(open(os.path.join(gold, name + ".rc")).read()), name
        if os.path.exists(ref):
            q = subprocess.run([ref] + argv, capture_output=True)
            assert (q.stdout, q.stderr, q.returncode) == (r.stdout, r.stderr, r.returncode), name


def test_join_struct_sizes_match_the_python_binding(tmp_path):
    """svb_join_line / svb_join_aln / svb_join_cand travel as raw bytes through lib.clip_join_raw and tools/clipjoin_sim --dump:
    the sizes the binding assumes are the C sizes (no padding)"""
    import subprocess
    from seeksv_b200 import lib as L
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "seeksv_b200.h"\nint main(void){printf("%zu %zu %zu\\n", sizeof(svb_join_line), '
                   'sizeof(svb_join_aln), sizeof(svb_join_cand));return 0;}\n')
    exe = str(tmp_path / "sz")
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    assert [int(x) for x in out] == [L.JOIN_LINE_BYTES, L.JOIN_ALN_BYTES, L.JOIN_CAND_BYTES]
