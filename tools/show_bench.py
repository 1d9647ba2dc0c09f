#!/usr/bin/env python
"""Pretty-print one bench.py JSON line (headline numbers + per-kernel table)."""
import json
import sys

for path in sys.argv[1:]:
    j = json.loads(open(path).read().strip().splitlines()[-1])
    print("== %s  %s" % (path, j["config"]["workload"][:60]))
    print("value %.3f M/s  %.3f ms/step  wall %.3f  e2e %.3f M/s (%.3f ms)  launches/step %s  clocks %s" % (
        j["value"] / 1e6, j["ms_per_step"], j.get("wall_ms_per_step", 0), j.get("e2e", {}).get("value", 0) / 1e6,
        j.get("e2e", {}).get("ms_per_step", 0), j.get("launches_per_step"), j.get("clocks")))
    if "cpu_baseline" in j:
        print("cpu", j["cpu_baseline"])
    print("roofline", {k: v for k, v in j["roofline"].items() if k != "note"})
    for k in j.get("kernels", []):
        print("  %-32s %8.2f us  share %.3f  %8s GB/s  frac %s" % (k["kernel"], k["us_per_launch"], k["share"],
                                                                  k["algo_gbs"], k["frac_of_hbm_peak"]))
    for k in j.get("hbm_kernels", []):
        print("  hbm:", k)
    for k in j:
        if k.startswith("act_select"):
            print("  ", k, j[k])
