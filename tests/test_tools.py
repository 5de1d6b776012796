"""The analysis tools behind the kernel plans of DESIGN.md section 8 keep running (CPU only, seconds): the host prototype of the
planned inflate tables reproduces zlib on a fixture, the token statistics and the line-granularity analysis produce their reports."""
import os
import subprocess
import sys

from conftest import GOLDEN, ROOT

BAM = os.path.join(GOLDEN, "example", "cancer.sort.bam")


def test_inflate_table_prototype_matches_zlib(tmp_path):
    exe = str(tmp_path / "huff2_proto")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tools", "huff2_proto.cpp"), "-lz"], check=True)
    for bam in (BAM, os.path.join(GOLDEN, "fuzz", "f11.sort.bam")):
        r = subprocess.run([exe, bam, "60"], capture_output=True, text=True)
        assert r.returncode == 0 and "60 blocks decoded, 0 differ from zlib" in r.stdout, r.stdout + r.stderr


def test_deflate_statistics_and_round_simulation():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "deflate_stats.py"), BAM, "4"], capture_output=True, text=True)
    assert r.returncode == 0 and "tokens" in r.stdout, r.stderr
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "inflate_rounds_sim.py"), BAM, "3"], capture_output=True, text=True)
    assert r.returncode == 0 and "rounds: iterations per batch" in r.stdout, r.stderr


def test_head_line_analysis():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "head_lines.py"), BAM, "4"], capture_output=True, text=True)
    assert r.returncode == 0 and "128-byte lines" in r.stdout, r.stderr
