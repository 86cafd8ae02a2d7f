#!/bin/bash
# Where does the time of the largest molecular-shape config (32 / 140 ADC(3), 8 singlets) go?
# 1. plain run with the solver / matrix timers; 2. cProfile of the host side; 3. ncu launch list.
set -u
mkdir -p gpurun_out
CASE=${1:-methyloxirane_ccpvdz}
SHOW_TIMERS=1 python tools/run_configs.py $CASE $CASE > gpurun_out/diag_c3_timers.txt 2>&1
PROFILE=1 python tools/diag_small.py $CASE > gpurun_out/diag_c3_host.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv \
    --log-file gpurun_out/diag_c3_launches.csv python tools/run_configs.py $CASE > gpurun_out/diag_c3_ncu.log 2>&1
python - <<'PY'
import csv, collections, re
rows = [r for r in csv.reader(open("gpurun_out/diag_c3_launches.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4])[:70]
    agg[name][0] += 1
    agg[name][1] += float(r[-1]) * 1e-6
tot = sum(v[1] for v in agg.values())
with open("gpurun_out/diag_c3_launch_summary.txt", "w") as f:
    f.write(f"# {len(rows)} launches, {tot:.1f} ms device time (ncu serialised, cold cache)\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:72s} {v[0]:6d} {v[1]:9.2f} ms {100*v[1]/tot:5.1f} %\n")
PY
head -40 gpurun_out/diag_c3_launch_summary.txt; cat gpurun_out/diag_c3_timers.txt | head -60
