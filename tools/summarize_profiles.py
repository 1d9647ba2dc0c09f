#!/usr/bin/env python
"""Turn gpurun_out/launches.csv + gpurun_out/prof_round.ncu-rep into the committed summaries under profiles/.

    python tools/summarize_profiles.py r01b      # tag of the round / build

Writes profiles/<tag>_launches.summary.txt (per-kernel launch count, mean device time, share of the step),
profiles/<tag>_full.csv (selected `ncu --set full` raw metrics per captured launch) and profiles/traffic.json
(DRAM bytes read+written per launch, per kernel: the `roofline.traffic` field of bench.py)."""
import csv
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
WL = sys.argv[2] if len(sys.argv) > 2 else "qmix_5v5_b32"
OUT = os.path.join(ROOT, "profiles")


def short(name):
    n = name.replace("void ", "")
    return n.split("(")[0].split("<")[0]


# ---------------------------------------------------------------- launch list
rows = []
with open(os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % WL)) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(io.StringIO("".join(lines)))
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((short(r["Kernel Name"]), us))
agg = OrderedDict()
for k, us in rows:
    agg.setdefault(k, []).append(us)
NOT_STEP = ("k_agent_step", "k_record_copy_tma", "k_max_t_filled", "k_eps_greedy_select")   # bench.py's other legs
tot = sum(us for k, us in rows if k not in NOT_STEP)
with open(os.path.join(OUT, tag + "_launches_%s.summary.txt" % WL), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:... -c 600  python bench.py --steps 4 --warmup 3 --no-cpu-baseline\n")
    f.write("# cold-cache, serialised launch times: compare SHARES with bench.py's `kernels` list (timed live, kernels alone)\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        share = "share_of_learner_step=%.3f" % (sum(v) / tot) if k not in NOT_STEP else "(act-select / replay legs of bench.py)"
        f.write("%-28s n=%4d avg_us=%9.2f %s\n" % (k, len(v), sum(v) / len(v), share))
print(open(os.path.join(OUT, tag + "_launches_%s.summary.txt" % WL)).read())

# ---------------------------------------------------------------- full capture
rawcsv = os.path.join(ROOT, "gpurun_out", "full_raw_%s.csv" % WL)
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
           "sm__inst_executed_pipe_fma.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
raw = open(rawcsv).read()          # `ncu -i <report> --page raw --csv`, exported on the GPU box by tools/profile_round.sh
rd = list(csv.reader(io.StringIO(raw)))
hdr, units = rd[0], rd[1]
idx = {n: i for i, n in enumerate(hdr)}
cols = [m for m in METRICS if m in idx]
traffic = defaultdict(list)
with open(os.path.join(OUT, tag + "_full_%s.csv" % WL), "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel"] + ["%s [%s]" % (c, units[idx[c]]) for c in cols])
    for r in rd[2:]:
        k = short(r[idx["Kernel Name"]])
        w.writerow([k] + [r[idx[c]] for c in cols])

        def val(c):
            x = float(r[idx[c]].replace(",", ""))
            u = units[idx[c]].lower()
            return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        traffic[k].append(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
tj = {k: int(sum(v) / len(v)) for k, v in traffic.items()}
# kernels that exist in several flavours share the name bench.py's profile scopes use
for k in list(tj):
    for pre, name in (("k_gru_fwd", "k_gru_fwd"), ("k_gru_bwd", "k_gru_bwd")):
        if k.startswith(pre) and k != name:
            tj[name] = tj[k]
tj["_source"] = "profiles/%s_full_%s.csv: dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured launches)" % (tag, WL)
tp = os.path.join(OUT, "traffic.json")
try:
    allt = json.load(open(tp))
    if allt and all(not isinstance(v, dict) for v in allt.values()):      # flat round-1 form = the 5v5 capture
        allt = {"qmix_5v5_b32": allt}
except Exception:
    allt = {}
allt[WL] = tj
json.dump(allt, open(tp, "w"), indent=1, sort_keys=True)
print(json.dumps(tj, indent=1, sort_keys=True))
