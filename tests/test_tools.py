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


def test_breakpoint2vcf_structure(tmp_path):
    """tools/breakpoint2vcf.py (stand-in for the reference's Python 2 / PyVCF converter; parity unpinned - see its header): two
    mated breakend records per table row, REF / ALT by the strand rules of breakpoint2vcf.py:18-36, INFO ordered by the template"""
    sv = tmp_path / "x.sv"
    cols = ["left_chr", "left_pos", "left_strand", "left_clip_read_NO", "right_chr", "right_pos", "right_strand", "right_clip_read_NO",
            "abnormal_readpair_NO", "left_pos_depth", "right_pos_depth", "left_seq", "right_seq"]
    rows = [["c1", "100", "+", "3", "c2", "200", "+", "4", "7", "30", "40", "ACGT", "GGCA"],
            ["c1", "300", "+", "1", "c1", "900", "-", "2", "0", "10", "11", "TTTC", "ATTT"],
            ["c2", "50", "-", "5", "c3", "60", "+", "6", "1", "12", "13", "CCCA", "TGGG"],
            ["c2", "70", "-", "5", "c3", "80", "-", "6", "1", "12", "13", "CCCA", "TGGG"]]
    sv.write_text("@" + "\t".join(cols) + "\n" + "".join("\t".join(r) + "\n" for r in rows))
    tpl = tmp_path / "t.vcf"
    tpl.write_text('##fileformat=VCFv4.1\n##INFO=<ID=MATEID,Number=1,Type=String,Description="m">\n##INFO=<ID=SVTYPE,Number=1,Type=String,Description="s">\n'
                   "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
    out = tmp_path / "o.vcf"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "breakpoint2vcf.py"), str(sv), str(tpl), str(out)], capture_output=True, text=True)
    assert r.returncode == 0 and "row 4" in r.stderr            # '-' / '-' has no rule in the reference
    lines = out.read_text().split("\n")
    assert lines[:4] == tpl.read_text().split("\n")[:4]
    recs = [l.split("\t") for l in lines[4:] if l]
    assert [x[2] for x in recs] == ["bnd1_U", "bnd1_D", "bnd2_U", "bnd2_D", "bnd3_U", "bnd3_D"]
    assert recs[0][:5] == ["c1", "100", "bnd1_U", "T", "T[<c2>:200["] and recs[1][:5] == ["c2", "200", "bnd1_D", "G", "]<c1>:100]G"]
    assert recs[2][3:5] == ["C", "C]<c1>:900]"] and recs[3][3:5] == ["T", "T]<c1>:300]"]           # right side '-': complement of A
    assert recs[4][3:5] == ["T", "[<c3>:60[T"] and recs[5][3:5] == ["T", "[<c2>:50[T"]             # left side '-': complement of A
    assert recs[0][7] == "MATEID=bnd1_D;SVTYPE=BND;ABNORMAL_READPAIR_NO=7;CLIP_READ_NO=3;DEPTH=30;STRAND=+" and recs[0][5:7] == [".", "PASS"]
