#!/usr/bin/env python
"""Readable summary of one default bench.py line (headline + `extra` block): python scripts/bench_line_summary.py line.json [reference.json]"""
import json
import sys


def last_json(path):
    for line in reversed(open(path).read().strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise SystemExit("no JSON line in %s" % path)


d = last_json(sys.argv[1])
c, r, e = d["config"], d["roofline"], d["e2e"]
print("# %s" % c["workload"])
print("n_gpus %d  steps %d x %d rounds  timed region %.2f s  clocks %s MHz (max %s) reasons %s samples under load %s" % (
    d["n_gpus"], d["steps"], c["rounds_per_step"], c["timed_region_s"], d["clocks"].get("sm_mhz"), d["clocks"].get("sm_max_mhz"),
    d["clocks"].get("reasons"), d["clocks"].get("samples_under_load")))
print("value      %.4g %s   (%.4f ms per round, one stream)" % (d["value"], d["unit"], c["ms_per_round"]))
m = c.get("multi_stream_resident")
if m:
    print("           %.4g %s   (%.4f ms per round, %d streams)" % (m["evals_per_s"], d["unit"], m["ms_per_round"], m["streams"]))
print("e2e        %.4g %s   (%.4f ms per round; H2D %d B, D2H %d B per step; %s)" % (
    e["value"], e["unit"], e["ms_per_round"], e["h2d_bytes_per_step"], e["d2h_bytes_per_step"], e["overlap"]))
print("roofline   K3 %.4f ms per launch, share of step %.3f, achieved %.2f of %.2f TFLOP/s = %.4f (with exact counts %.4f); traffic %s B" % (
    r["k3_ms_per_launch"], r["k3_share_of_step"] or 0, r["achieved"], r["peak"], r["frac"], r.get("frac_with_exact_counts") or 0, r["traffic"]))
print("winner     %s" % c["winner"])
if d.get("cpu_baseline"):
    b = d["cpu_baseline"]
    print("cpu        %.4g %s on %d cores (%s): %s" % (b["value"], b["unit"], b["cores"], b["kind"], b["sample"]))
if len(sys.argv) > 2:
    ref = last_json(sys.argv[2])
    print("reference arm  %.4g %s on %d cores -> e2e ratio %.1f, resident ratio %.1f" % (
        ref["value"], ref["unit"], ref["cpu_baseline"]["cores"], e["value"] / ref["value"], d["value"] / ref["value"]))
print()
print("%-12s %12s %10s %12s %8s %12s %12s %12s %10s" % ("extra", "value", "unit", "e2e", "K3 frac", "pairs/s", "pairs/s e2e", "cpu arm", "x cpu"))
for k, v in d.get("extra", {}).items():
    rr = v.get("roofline") or {}
    cpu = (v.get("cpu_baseline") or {}).get("value")
    print("%-12s %12.4g %10s %12.4g %8s %12s %12s %12s %10s" % (
        k, v["value"], v["unit"].split("/")[0], v.get("e2e") or 0, ("%.3f" % rr["frac"]) if rr.get("frac") else "-",
        ("%.4g" % v["pairs_per_s"]) if v.get("pairs_per_s") else "-", ("%.4g" % v["pairs_per_s_e2e"]) if v.get("pairs_per_s_e2e") else "-",
        ("%.4g" % cpu) if cpu else "-", ("%.1f" % ((v.get("e2e") or v["value"]) / cpu)) if cpu else "-"))
    if v.get("sharded_equals_single_gpu") is not None:
        print("%-12s N-rank winner record == single-GPU record: %s" % ("", v["sharded_equals_single_gpu"]))
