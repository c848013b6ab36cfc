"""Per-source-line stall samples / instruction counts of one kernel from an .ncu-rep (ncu --set full --import-source on).

    python tools/ncu_lines.py report.ncu-rep [top]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur = None
    hdr = None
    lines = []
    for r in csv.reader(out.splitlines()):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
            try:
                lines.append((cur, int(r[0]), r[1].strip(), int(r[4]), int(r[7])))
            except ValueError:
                pass
    tot_s = sum(x[3] for x in lines) or 1
    tot_i = sum(x[4] for x in lines) or 1
    print("total samples %d, warp instructions %d" % (tot_s, tot_i))
    for f, ln, src, s, i in sorted(lines, key=lambda x: -x[3])[:top]:
        print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100.0 * s / tot_s, 100.0 * i / tot_i, f, ln, src[:110]))


if __name__ == "__main__":
    main()
