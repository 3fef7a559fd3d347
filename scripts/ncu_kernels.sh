#!/bin/bash
# Developer helper (GPU box): per-launch ncu metrics of the kernels matching $1 for the python command after it.
#   scripts/ncu_kernels.sh regex:k_pf_ gpurun_out/pf_passes.csv python scripts/sir_probe.py rs78 26
pat=$1; out=$2; shift 2
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sectors.sum \
  --clock-control none -k "$pat" -c 24 --csv --log-file "$out" "$@" > /dev/null 2>&1
python - "$out" <<'PY'
import csv, sys, collections
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
d = collections.OrderedDict()
for r in rows:
    d.setdefault((r["ID"], r["Kernel Name"][:44]), {})[r["Metric Name"]] = f'{r["Metric Value"]} {r["Metric Unit"]}'
for k, v in list(d.items())[-9:]:
    print(k[0], k[1])
    for m, x in v.items():
        print("    ", m, x)
PY
