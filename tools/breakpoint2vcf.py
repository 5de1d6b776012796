#!/usr/bin/env python3
"""breakpoint2vcf - seeksv's result table (output.sv.txt) as VCF breakend records.

    python tools/breakpoint2vcf.py output.sv.txt template.vcf output.vcf

Stand-in for the reference's downstream converter (/root/reference/breakpoint2vcf/breakpoint2vcf.py:1-97; SURVEY.md section 2
"OUT OF SCOPE", section 8(f) item 4). The reference script is Python 2 and writes through PyVCF; neither exists in this image, so
this file restates what that script asks PyVCF (0.6.x) to do, from the library's published writer: it is host-only text work and
**parity is unpinned** - there is no reference output to compare with here. What is restated:

  * one pair of records per table row, IDs bnd<k>_U / bnd<k>_D (k counts rows from 1), each naming the other as MATEID;
  * REF = the last base of left_seq / the first base of right_seq, reverse-complemented on a '-' strand side
    (breakpoint2vcf.py:18-36); rows with '-' / '-' strands have no rule in the reference (it stops with an unbound variable): they
    are reported on stderr and skipped here;
  * ALT in PyVCF's breakend notation: the script constructs `_Breakend(chr, pos, orientation, remoteOrientation, base, None)`, and
    with `withinMainAssembly` None the library prints the mate's chromosome in angle brackets - `A[<chr2>:1234[`, `]<chr2>:1234]A`;
  * QUAL '.', FILTER PASS, a FORMAT column holding '.', no sample columns;
  * INFO keys in the order of the template's ##INFO lines, undeclared keys after them in alphabetical order (the writer sorts by
    (position in the template, name));
  * the header: the template's '##' lines and its '#CHROM' line. PyVCF re-serialises the metadata it parsed; this tool copies the
    lines as they are.
"""
import sys

COMPLEMENT = {"A": "T", "T": "A", "C": "G", "G": "C", "a": "T", "t": "A", "c": "G", "g": "C"}


def breakend(chrom, pos, orientation, remote_orientation, base):
    """str(vcf.model._Breakend(chrom, pos, orientation, remote_orientation, base, None))"""
    remote = "<" + chrom + ">"          # withinMainAssembly is None in the reference's calls
    tag = "[%s:%d[" % (remote, pos) if remote_orientation else "]%s:%d]" % (remote, pos)
    return tag + base if orientation else base + tag


def records(row, k):
    """the two records of one table row (breakpoint2vcf.py:17-70), None when the strand pair has no rule"""
    lp, rp = int(row["left_pos"]), int(row["right_pos"])
    ls, rs = row["left_strand"], row["right_strand"]
    if ls == "+" and rs == "+":
        ref1, ref2 = row["left_seq"][-1], row["right_seq"][0]
        alt1, alt2 = breakend(row["right_chr"], rp, False, True, ref1), breakend(row["left_chr"], lp, True, False, ref2)
    elif ls == "+" and rs == "-":
        ref1, ref2 = row["left_seq"][-1], COMPLEMENT[row["right_seq"][0]]
        alt1, alt2 = breakend(row["right_chr"], rp, False, False, ref1), breakend(row["left_chr"], lp, False, False, ref2)
    elif ls == "-" and rs == "+":
        ref1, ref2 = COMPLEMENT[row["left_seq"][-1]], row["right_seq"][0]
        alt1, alt2 = breakend(row["right_chr"], rp, True, True, ref1), breakend(row["left_chr"], lp, True, True, ref2)
    else:
        return None
    up, down = "bnd%d_U" % k, "bnd%d_D" % k
    one = (row["left_chr"], lp, up, ref1, alt1,
           {"SVTYPE": "BND", "MATEID": down, "CLIP_READ_NO": row["left_clip_read_NO"], "STRAND": ls,
            "ABNORMAL_READPAIR_NO": row["abnormal_readpair_NO"], "DEPTH": row["left_pos_depth"]})
    two = (row["right_chr"], rp, down, ref2, alt2,
           {"SVTYPE": "BND", "MATEID": up, "CLIP_READ_NO": row["right_clip_read_NO"], "STRAND": rs,
            "ABNORMAL_READPAIR_NO": row["abnormal_readpair_NO"], "DEPTH": row["right_pos_depth"]})
    return one, two


def info_order(template_lines):
    order = {}
    for line in template_lines:
        if line.startswith("##INFO=<") and "ID=" in line:
            name = line.split("ID=", 1)[1].split(",", 1)[0].rstrip(">\n")
            order.setdefault(name, len(order))
    return order


def convert(table_path, template_path, out_path):
    with open(template_path) as f:
        template = [l for l in f if l.startswith("#")]
    order = info_order(template)
    with open(table_path) as f:
        header = f.readline()
        if not header.startswith("@") or len(header.strip()) < 2:
            sys.stderr.write("Error: breapoint file header should start with '@'\n")     # (the reference's message)
            return 1
        names = header.strip().replace("@", "").split("\t")
        with open(out_path, "w") as out:
            out.writelines(template)
            for k, line in enumerate(f, 1):
                row = dict(zip(names, line.strip().split("\t")))
                pair = records(row, k)
                if pair is None:
                    sys.stderr.write("row %d: no breakend rule for strands %s / %s (the reference stops here)\n"
                                     % (k, row.get("left_strand"), row.get("right_strand")))
                    continue
                for chrom, pos, rid, ref, alt, info in pair:
                    keys = sorted(info, key=lambda x: (order.get(x, len(order)), x))
                    out.write("\t".join([chrom, str(pos), rid, ref, alt, ".", "PASS", ";".join("%s=%s" % (x, info[x]) for x in keys), "."]) + "\n")
    return 0


if __name__ == "__main__":
    if len(sys.argv) != 4:
        sys.stderr.write("usage: breakpoint2vcf.py breakpoint template_vcf vcf_file\n")
        sys.exit(2)
    sys.exit(convert(*sys.argv[1:]))
