#!/bin/bash
# GPU box: the round's evidence in one go (run under gpurun; outputs land in gpurun_out/).
#   scripts/capture_round.sh r1            -> bench line, reference arm, ncu launch list, two --set full reports
# Then, back in the container:  python scripts/summarize_profiles.py r1
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
set -x
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
# launch list of the same command (a number printed under ncu is never a bench value)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/ncu_launches.log 2>&1
# one --set full capture per hot kernel, second launch of each (warmed up)
ncu --set full --clock-control none --import-source on -k regex:k_search_workers -s 1 -c 1 -f \
  -o $out/prof_workers_$tag python scripts/profile_step.py > $out/ncu_search.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pf_fused -s 1 -c 1 -f \
  -o $out/prof_pffused_$tag python scripts/profile_step.py > $out/ncu_pf.log 2>&1
# instructions per rollout step: thread instructions and the step counter of the SAME kernel pass (the --set full
# capture replays the kernel dozens of times, every pass grows a different tree, and the counter read afterwards
# belongs to the last one) -- three SM counters fit one pass
ncu --metrics smsp__thread_inst_executed.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:k_search_workers -s 1 -c 1 --csv --log-file $out/workers_inst_$tag.csv \
  python scripts/profile_step.py > $out/workers_inst_$tag.log 2>&1
tail -n 2 $out/ncu_search.log; tail -n 2 $out/ncu_pf.log
cat $out/bench_$tag.json
