#!/usr/bin/env python
"""One line per bench log in gpurun_out/ (value, step time, e2e, K3 time/share, roofline fraction)."""
import glob, json, sys
for f in sorted(glob.glob(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            d = json.loads(line); r = d.get("roofline") or {}
            print("%-34s value %.3e ms/step %.3f e2e %.3e | K3 %.3f ms share %.2f frac %.3f | launches %d sm_mhz %s %s" % (
                f.split("/")[-1], d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("k3_ms_per_launch", 0),
                r.get("k3_share_of_step") or 0, r.get("frac") or 0, d.get("gpu_launches", 0), (d.get("clocks") or {}).get("sm_mhz"), (d.get("clocks") or {}).get("reasons")))
            if "cpu_baseline" in d: print("    cpu:", d["cpu_baseline"])
        elif "rror" in line: print(f, line.strip())
