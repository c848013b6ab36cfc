"""Checks that every reference citation `file.h:line[-line]` in the sources and documents of this repository points at an
existing file of the reference tree with at least that many lines (run where /root/reference exists):
python tools/check_citations.py [--list-bad]"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GCR_REFERENCE_DIR", "/root/reference")
INC = os.path.join(REF, "src", "pygcransac", "include")
SRC = os.path.join(REF, "src", "pygcransac", "src")
PAT = re.compile(r"(?<![\w/.])((?:py/|[\w]+/)*[\w]+\.(?:h|cpp|py)):(\d+)(?:-(\d+))?")


def index_reference():
    by_name = {}
    for base in (INC, SRC, os.path.join(REF, "src", "pygcransac"), REF):
        for d, _, files in os.walk(base):
            for f in files:
                if f.endswith((".h", ".cpp", ".py", ".txt")):
                    by_name.setdefault(f, set()).add(os.path.join(d, f))
    return by_name


def main():
    if not os.path.isdir(INC):
        print("reference tree not present")
        return 0
    by_name = index_reference()
    lines_of = {}
    own = set()
    for d, _, files in os.walk(ROOT):
        if any(p in d for p in (".git", "gpurun_out", "__pycache__", "build")):
            continue
        for f in files:
            own.add(f)
    checked = bad = 0
    bad_list = []
    for d, _, files in os.walk(ROOT):
        if any(p in d for p in (".git", "gpurun_out", "__pycache__", "/build", "profiles")):
            continue
        for f in files:
            if not f.endswith((".h", ".cuh", ".cu", ".cpp", ".py", ".md", ".sh")) or f in ("SURVEY.md", "PAPERS.md", "SNIPPETS.md", "VERDICT.md", "ADVICE.md", "BASELINE.md"):
                continue
            path = os.path.join(d, f)
            text = open(path, errors="replace").read()
            for m in PAT.finditer(text):
                name, lo, hi = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
                base = os.path.basename(name)
                if base not in by_name:
                    continue            # a file of this repository (or something else): not a reference citation
                if base in own and "/" not in name:
                    continue            # ambiguous: this repository has a file of that name too
                cands = [p for p in by_name[base] if p.endswith(name.replace("py/", ""))] or list(by_name[base])
                ok = False
                for c in cands:
                    if c not in lines_of:
                        lines_of[c] = sum(1 for _ in open(c, errors="replace"))
                    if hi <= lines_of[c] and lo <= hi:
                        ok = True
                checked += 1
                if not ok:
                    bad += 1
                    bad_list.append("%s: %s (file has %s lines)" % (os.path.relpath(path, ROOT), m.group(0), sorted(lines_of[c] for c in cands)))
    print("%d reference citations checked, %d out of range" % (checked, bad))
    if "--list-bad" in sys.argv or bad:
        for b in bad_list[:50]:
            print("  " + b)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
