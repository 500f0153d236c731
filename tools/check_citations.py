"""Checks that every `Something.swift:a-b` citation in the headers, sources, oracle, tests and docs names a file that exists under
/root/reference and has at least b lines.  Development aid (needs the reference checkout; not part of the test suite because the
GPU box has none).  Run: python tools/check_citations.py"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PAT = re.compile(r"([A-Za-z0-9_+/.-]+\.(?:swift|py|md)):(\d+)(?:-(\d+))?")


def reference_files():
    by_name = {}
    for d, _, fs in os.walk(REF):
        if "/.git" in d:
            continue
        for f in fs:
            by_name.setdefault(f, []).append(os.path.join(d, f))
    return by_name


def main():
    if not os.path.isdir(REF):
        print("no reference checkout; nothing to check")
        return 0
    by_name = reference_files()
    lens = {}
    bad = checked = 0
    for d, _, fs in os.walk(ROOT):
        if any(x in d for x in ("/.git", "/gpurun_out", "/build", "/__pycache__", "/profiles", "/.pytest_cache")):
            continue
        for f in fs:
            if not f.endswith((".h", ".cu", ".cuh", ".hpp", ".cpp", ".py", ".md", ".swift", ".c")) or f in ("SURVEY.md", "PAPERS.md", "SNIPPETS.md"):
                continue
            path = os.path.join(d, f)
            for ln, line in enumerate(open(path, errors="replace"), 1):
                for m in PAT.finditer(line):
                    name, a, b = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
                    base = os.path.basename(name)
                    cands = [p for p in by_name.get(base, []) if p.endswith(name.lstrip("./")) or "/" not in name]
                    if not cands:
                        if base in by_name or not base.endswith(".swift"):
                            continue  # a path of ours, or ambiguous: skip
                        print(f"{os.path.relpath(path, ROOT)}:{ln}: no such reference file {name}")
                        bad += 1
                        continue
                    checked += 1
                    n = max(lens.setdefault(p, sum(1 for _ in open(p, errors="replace"))) for p in cands)
                    if b > n or a > b:
                        print(f"{os.path.relpath(path, ROOT)}:{ln}: {name}:{a}-{b} but the file has {n} lines")
                        bad += 1
    print(f"{checked} citations checked, {bad} bad")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
