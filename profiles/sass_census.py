"""Per-kernel opcode census of libcvf_sm100.so (cuobjdump -sass): the instructions that prove which hardware path a kernel takes --
UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk, 1-D TMA), UTMALDG
(tensor-map TMA), FFMA2 (packed fp32 FMA), LDGSTS (cp.async), STL / LDL (register spills), DFMA, RED / ATOM.

    python profiles/sass_census.py > profiles/r02_sass_census.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "colvars-finder_b200", "colvarsfinder", "libcvf_sm100.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FFMA", "DFMA", "MUFU", "LDGSTS", "LDS", "STS",
       "LDG", "STG", "RED", "ATOM", "SHFL", "STL", "LDL", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    print(f"# opcode census of {os.path.relpath(LIB, ROOT)} (source hash {g.library_hash()}); counts are static instructions per kernel")
    print("# " + " ".join(f"{o:>8s}" for o in OPS) + "  total  kernel")
    name, counts = None, None
    table = []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                table.append((name, counts))
            name, counts = m.group(1), collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and counts is not None:
            op = m.group(1)
            counts["total"] += 1
            for o in OPS:
                if op == o or (o in ("RED", "ATOM") and op.startswith(o)):
                    counts[o] += 1
    if name:
        table.append((name, counts))
    for name, c in sorted(table, key=lambda t: -t[1]["total"]):
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(.*", "", dem)
        print("  " + " ".join(f"{c[o]:8d}" for o in OPS) + f"  {c['total']:6d}  {dem}")


if __name__ == "__main__":
    main()
