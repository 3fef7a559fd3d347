#!/bin/bash
# GPU box: one gpurun call's worth of work; stages are picked by name, outputs land in gpurun_out/<tag>_*.
#   scripts/gpu_session.sh r2a tests bench ref c5 quality episodes ncu ncupf sir multi fuzz
tag=$1; shift
out=gpurun_out
mkdir -p $out
for stage in "$@"; do
  echo "=== $stage $(date +%T)"
  case $stage in
    tests)    python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -n 5 $out/${tag}_pytest.log ;;
    bench)    python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 600 $out/${tag}_bench.err; head -c 400 $out/${tag}_bench.json; echo ;;
    ref)      python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; head -c 400 $out/${tag}_bench_ref.json; echo ;;
    c5)       python bench.py --config c5 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err; tail -c 600 $out/${tag}_bench_c5.err; head -c 400 $out/${tag}_bench_c5.json; echo ;;
    quality)  python scripts/probe.py knob --knob inflight_max --values 64,256,1184 --seeds 12 --out $out/${tag}_knob_inflight.json > $out/${tag}_knob_inflight.log 2>&1
              python scripts/probe.py equal-time --budget-ms 14 --values 64,256,1184 --seeds 12 --out $out/${tag}_equal_time.json > $out/${tag}_equal_time.log 2>&1
              python scripts/probe.py knob --knob n_trees --values 1,2,4,8 --queries 1000000 --seeds 12 --out $out/${tag}_knob_trees.json > $out/${tag}_knob_trees.log 2>&1
              tail -n 4 $out/${tag}_knob_inflight.log $out/${tag}_equal_time.log $out/${tag}_knob_trees.log ;;
    episodes) for q in 10000 100000 1000000; do
                python scripts/episode_table.py --arm gpu --queries $q --episodes 24 --out $out/${tag}_episodes_gpu_$q.json
              done
              python scripts/episode_table.py --arm gpu --queries 1000000 --episodes 24 --inflight-max 256 --out $out/${tag}_episodes_gpu_1000000_infl256.json ;;
    rates)    python scripts/probe.py rates > $out/${tag}_rates.log 2>&1; cat $out/${tag}_rates.log
              python scripts/probe.py cycles > $out/${tag}_cycles.log 2>&1; cat $out/${tag}_cycles.log ;;
    sir)      for variant in default POMCP_PF_NOCOOP POMCP_PF_TWO_BARRIERS; do
                echo "--- $variant" >> $out/${tag}_sir.log
                env $( [ $variant = default ] && echo X_UNUSED=1 || echo $variant=1 ) python scripts/probe.py sir --problem rs78 --log2 12,16,20,24,26 --out $out/${tag}_sir_rs78_$variant.json >> $out/${tag}_sir.log 2>&1
                env $( [ $variant = default ] && echo X_UNUSED=1 || echo $variant=1 ) python scripts/probe.py sir --problem lightdark --log2 1000000,20,24,26 --out $out/${tag}_sir_ld_$variant.json >> $out/${tag}_sir.log 2>&1
              done
              cat $out/${tag}_sir.log ;;
    multi)    # needs gpurun --gpus N; N from $NGPU (default 2)
              N=${NGPU:-2}
              TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
              python -m pytest tests/test_gpu_multirank.py -m gpu -q > $out/${tag}_pytest_multirank.log 2>&1; tail -n 3 $out/${tag}_pytest_multirank.log
              $TR bench.py --gpus $N --steps 10 --warmup 3 > $out/${tag}_bench_n$N.json 2> $out/${tag}_bench_n$N.err; tail -c 400 $out/${tag}_bench_n$N.err; head -c 300 $out/${tag}_bench_n$N.json; echo
              $TR bench.py --gpus $N --config c5 --steps 3 --warmup 3 > $out/${tag}_bench_c5_n$N.json 2> $out/${tag}_bench_c5_n$N.err; tail -c 400 $out/${tag}_bench_c5_n$N.err; head -c 300 $out/${tag}_bench_c5_n$N.json; echo
              python scripts/probe.py c4 > $out/${tag}_c4_n1.json 2>&1; cat $out/${tag}_c4_n1.json
              $TR scripts/probe.py c4 > $out/${tag}_c4_n$N.json 2> $out/${tag}_c4_n$N.err; tail -n 1 $out/${tag}_c4_n$N.json ;;
    ncupf)    ncu --set full --clock-control none --import-source on -k regex:k_pf_ -s 1 -c 1 -f \
                -o $out/prof_pf_$tag python scripts/probe.py step > $out/${tag}_ncu_pf.log 2>&1
              tail -n 2 $out/${tag}_ncu_pf.log ;;
    ncu)      # launch list of the bench command (a number printed under ncu is never a bench value)
              ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
                python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $out/${tag}_ncu_launches.log 2>&1
              # thread instructions + the step counters of the SAME pass, 8 launches that grow different trees
              ncu --metrics smsp__thread_inst_executed.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
                --clock-control none -k regex:k_search_workers -c 8 --csv --log-file $out/workers_inst_$tag.csv \
                python scripts/probe.py step --repeat 8 > $out/workers_inst_$tag.log 2>&1
              # pipe utilisation of the worker kernel (names the binding pipe)
              ncu --metrics sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
                --clock-control none -k regex:k_search_workers -s 1 -c 1 --csv --log-file $out/workers_pipes_$tag.csv \
                python scripts/probe.py step > $out/workers_pipes_$tag.log 2>&1
              ncu --set full --clock-control none --import-source on -k regex:k_search_workers -s 1 -c 1 -f \
                -o $out/prof_workers_$tag python scripts/probe.py step > $out/${tag}_ncu_search.log 2>&1
              tail -n 2 $out/${tag}_ncu_search.log ;;  # the SIR kernel's capture is stage ncupf (gpurun returns <= 64 MiB per call)
    fuzz)     python scripts/fuzz.py --cases ${FUZZ_CASES:-1000} --seed ${FUZZ_SEED:-0} 2>&1 | grep -v "external belief updater" > $out/${tag}_fuzz.log; tail -n 5 $out/${tag}_fuzz.log ;;
    *) echo "unknown stage $stage" ;;
  esac
done
echo "=== done $(date +%T)"
