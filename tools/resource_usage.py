"""Per-kernel registers / stack / static shared memory / local memory of the built library, from `cuobjdump --dump-resource-usage`
(no GPU needed).  Writes profiles/<name>.txt.  Run: python tools/resource_usage.py r1e_resource_usage"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "qwen3-asr-swift_b200", "lib", "libq3asr.so")


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "resource_usage"
    lines = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True, check=True).stdout.splitlines()
    rows = []
    for i, l in enumerate(lines):
        l = l.strip()
        if not l.startswith("Function"):
            continue
        sym = l.split()[1].rstrip(":")
        dem = subprocess.run(["c++filt", sym], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(anonymous namespace\)::", "", dem)
        dem = re.sub(r"^void ", "", dem)
        dem = re.sub(r"\(.*", "", dem)
        m = dict(x.split(":") for x in lines[i + 1].split() if ":" in x and not x.startswith("CONSTANT"))
        rows.append((dem, int(m["REG"]), int(m["STACK"]), int(m["SHARED"]), int(m["LOCAL"])))
    rows.sort()
    out = os.path.join(ROOT, "profiles", name + ".txt")
    with open(out, "w") as f:
        f.write("cuobjdump --dump-resource-usage lib/libq3asr.so (sm_100a); dynamic shared memory is set at launch and not shown\n")
        f.write("%-92s %5s %6s %12s %6s\n" % ("kernel", "regs", "stack", "smem(static)", "local"))
        for r in rows:
            f.write("%-92s %5d %6d %12d %6d\n" % r)
        spills = [(r[0], r[2]) for r in rows if r[2] > 0 or r[4] > 0]
        f.write("\n%d kernels; with a stack frame or local memory: %s\n" % (len(rows), spills if spills else "none"))
    print(out, len(rows), "kernels;", len(spills), "with stack/local")


if __name__ == "__main__":
    main()
