"""Counts the SASS mnemonics that prove which hardware paths a kernel uses (B200_PROFILING.md: UTC*MMA = tcgen05.mma, LDTM/STTM =
tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, HMMA = mma.sync, LDGSTS = cp.async, SYNCS = mbarrier, ACQBULK/PREEXIT = programmatic
dependent launch) per kernel of the built library, from `cuobjdump -sass` (no GPU needed).  Writes profiles/<name>.txt.
Run: python tools/sass_mnemonics.py r1e_sass_mnemonics"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "qwen3-asr-swift_b200", "lib", "libq3asr.so")
COLS = [("UTC*MMA", r"\bUTC\w*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
        ("UBLKCP", r"\bUBLKCP"), ("HMMA", r"\bHMMA"), ("LDGSTS", r"\bLDGSTS"), ("SYNCS", r"\bSYNCS"), ("ACQBULK", r"\bACQBULK"),
        ("PREEXIT", r"\bPREEXIT"), ("FFMA", r"\bFFMA"), ("LDL/STL", r"\b(?:LDL|STL)\b")]


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "sass_mnemonics"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None or "/*" not in line:
            continue
        for col, pat in COLS:
            if re.search(pat, line):
                counts[cur][col] += 1
        counts[cur]["total"] += 1
    rows = []
    for sym in order:
        dem = subprocess.run(["c++filt", sym], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(anonymous namespace\)::", "", dem)
        dem = re.sub(r"^void ", "", dem)
        dem = re.sub(r"\(.*", "", dem)
        rows.append((dem, counts[sym]))
    rows.sort(key=lambda r: r[0])
    out = os.path.join(ROOT, "profiles", name + ".txt")
    with open(out, "w") as f:
        f.write("cuobjdump -sass lib/libq3asr.so (sm_100a): instruction counts per kernel (static, not executed counts)\n")
        f.write("%-84s %7s " % ("kernel", "instrs") + " ".join("%7s" % c for c, _ in COLS) + "\n")
        for dem, c in rows:
            f.write("%-84s %7d " % (dem[:84], c["total"]) + " ".join("%7s" % (c[col] or ".") for col, _ in COLS) + "\n")
        tc = [d for d, c in rows if c["UTC*MMA"]]
        tma = [d for d, c in rows if c["UTMALDG"] or c["UTMASTG"] or c["UBLKCP"]]
        f.write("\n%d kernels; %d issue tcgen05.mma, %d use TMA, %d use mma.sync (HMMA), %d touch local memory\n"
                % (len(rows), len(tc), len(tma), sum(1 for _, c in rows if c["HMMA"]), sum(1 for _, c in rows if c["LDL/STL"])))
    print(out, len(rows), "kernels;", len(tc), "tcgen05,", len(tma), "TMA")


if __name__ == "__main__":
    main()
