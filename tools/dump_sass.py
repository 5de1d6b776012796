"""Write the SASS listing of every hand-written kernel and a resource table under profiles/sass/.

    python tools/dump_sass.py            # after python -m seeksv_b200.build

One file per kernel (`<object>.<kernel>.sass`, `cuobjdump -sass -fun`), CUB's instantiations left out (library code), plus
`profiles/sass/README.md`: registers / shared / stack per kernel from `cuobjdump -res-usage` and the counts of the load/store
forms that matter for an HBM-bound byte kernel (128-bit global loads, non-coherent loads, shuffles/votes, local-memory spills).
Runs without a GPU.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "seeksv_b200", "build")
OUT = os.path.join(ROOT, "profiles", "sass")
CUOBJDUMP = os.environ.get("CUOBJDUMP", "/usr/local/cuda/bin/cuobjdump")
CUFILT = os.environ.get("CUFILT", "/usr/local/cuda/bin/cu++filt")


def sh(*cmd):
    return subprocess.run(cmd, check=True, capture_output=True, text=True).stdout


def short(demangled):
    name = re.sub(r"\(anonymous namespace\)::", "", demangled)
    name = re.sub(r"\((?:int|unsigned int|bool)\)", "", name)               # template arguments print as <(int)32>
    name = re.sub(r"^void ", "", name).split("(")[0]
    name = name.replace("<", "_").replace(">", "").replace(" ", "")
    return name.split("::")[-1]


def main():
    os.makedirs(OUT, exist_ok=True)
    for f in os.listdir(OUT):
        if f.endswith(".sass"):
            os.remove(os.path.join(OUT, f))
    rows = []
    for obj in sorted(os.listdir(OBJ)):
        if not obj.endswith(".cu.o"):
            continue
        path = os.path.join(OBJ, obj)
        res = sh(CUOBJDUMP, "-res-usage", path)
        funcs = re.findall(r"Function (\S+):\n\s*(.*)", res)
        for mangled, usage in funcs:
            if "cub" in mangled and "CUB_" in mangled:
                continue
            dem = sh(CUFILT, mangled).strip()
            name = short(dem)
            sass = sh(CUOBJDUMP, "-sass", "-fun", mangled, path)
            sass = re.sub(r"[ \t]*/\* 0x[0-9a-f]{16} \*/", "", sass)          # drop the encodings: mnemonics are the evidence
            sass = "\n".join(l.rstrip() for l in sass.splitlines() if l.strip()) + "\n"
            fn = f"{obj[:-5]}.{name}.sass"
            with open(os.path.join(OUT, fn), "w") as o:
                o.write(f"// {dem}\n// {usage.strip()}\n// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo ({obj[:-2]})\n")
                o.write(sass)
            ops = collections.Counter()
            n_ins = 0
            for line in sass.splitlines():
                m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
                if not m:
                    continue
                n_ins += 1
                op = m.group(1)
                if op.startswith("LDG"):
                    ops["ldg"] += 1
                    if ".128" in op:
                        ops["ldg128"] += 1
                    if ".CONSTANT" in op:
                        ops["ldg_nc"] += 1
                elif op.startswith("STG"):
                    ops["stg"] += 1
                    if ".128" in op:
                        ops["stg128"] += 1
                elif op.startswith(("LDL", "STL")):
                    ops["local"] += 1
                elif op.startswith(("LDS", "STS")):
                    ops["smem"] += 1
                elif op.startswith(("SHFL", "VOTE", "MATCH", "REDUX")):
                    ops["warp"] += 1
                elif op.startswith(("ATOM", "RED", "ATOMS", "ATOMG")):
                    ops["atom"] += 1
                elif op.startswith("SHF"):
                    ops["shf"] += 1
            u = dict(re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", usage))
            rows.append((obj[:-5], name, u.get("REG"), u.get("SHARED"), u.get("STACK"), n_ins, ops, fn))
    with open(os.path.join(OUT, "README.md"), "w") as o:
        o.write("# SASS listings (sm_100a) of the hand-written kernels\n\n"
                "Made by `python tools/dump_sass.py` from the objects `python -m seeksv_b200.build` leaves in `seeksv_b200/build/`\n"
                "(`cuobjdump -sass -fun <kernel>`; CUB's sort/scan instantiations are library code and left out). No tensor-core\n"
                "instruction appears on purpose: the path is integer/byte work on packed, unaligned records, bound by HBM. The two\n"
                "`*_stream` kernels (the TMA-staged alternative form of the full passes, off by default) show the bulk-copy path:\n"
                "`UBLKCP.S.G` (cp.async.bulk global -> shared) with `SYNCS.ARRIVE.TRANS64` / `SYNCS.PHASECHK.TRANS64.TRYWAIT` (mbarrier).\n"
                "Columns: registers per thread, static shared bytes, stack bytes (`cuobjdump -res-usage`), SASS instruction count,\n"
                "global loads (of which 128-bit / read-only `.CONSTANT`), global stores (of which 128-bit), shared-memory accesses,\n"
                "warp shuffles/votes/matches, funnel shifts (`SHF`, the unaligned-word assembly), atomics, local-memory accesses\n"
                "(`LDL`/`STL`: spills or indexed per-thread arrays).\n\n"
                "| object | kernel | regs | smem | stack | instr | LDG (128 / nc) | STG (128) | LDS/STS | SHFL/VOTE | SHF | atom | LDL/STL | file |\n"
                "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for obj, name, reg, smem, stack, n, ops, fn in rows:
            o.write(f"| {obj} | `{name}` | {reg} | {smem} | {stack} | {n} | {ops['ldg']} ({ops['ldg128']} / {ops['ldg_nc']}) | "
                    f"{ops['stg']} ({ops['stg128']}) | {ops['smem']} | {ops['warp']} | {ops['shf']} | {ops['atom']} | {ops['local']} | `{fn}` |\n")
    print(f"{len(rows)} kernels -> {OUT}")


if __name__ == "__main__":
    sys.exit(main())
