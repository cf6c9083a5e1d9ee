#!/bin/bash
# Round measurement pass on one B200 (run through gpurun): tests, bench lines, ncu launch list and counters.
# Outputs go to gpurun_out/$TAG_*; summaries are copied to profiles/ by tools/summarise_round.py afterwards.
TAG=${1:-r01c}
O=gpurun_out
set -x
timeout 900 python -m pytest tests -x -q -m gpu > $O/${TAG}_tests.log 2>&1; tail -3 $O/${TAG}_tests.log
timeout 600 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; tail -c 600 $O/${TAG}_bench_n1.json
timeout 600 python bench.py --nn-engine dmma --no-cpu-baseline --steps 3 > $O/${TAG}_bench_n1_dmma.json 2>> $O/${TAG}_bench_n1.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; tail -c 400 $O/${TAG}_bench_ref.json
# launch list of the bench command (per-launch durations under ncu are cold-cache and serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file $O/${TAG}_launches_cfg4.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs --no-screened > $O/${TAG}_ncu1.log 2>&1
# full counters of the three score kernels (first launch of each)
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_score_feas|k_prep_i8|k_mlp_i8" -c 3 -f -o $O/${TAG}_prof_cfg4 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs --no-screened > $O/${TAG}_ncu2.log 2>&1
timeout 300 python tools/cover_bench.py > $O/${TAG}_cover_bench.log 2>&1; cat $O/${TAG}_cover_bench.log
timeout 120 python tools/i8_trace.py -q > $O/${TAG}_i8_trace.log 2>&1; timeout 120 python tools/i8_trace.py -q --screen > $O/${TAG}_i8_trace_screen.log 2>&1; head -3 $O/${TAG}_i8_trace.log $O/${TAG}_i8_trace_screen.log
timeout 300 python bench.py --separate-eig --no-cpu-baseline --no-other-configs --no-screened --steps 4 --warmup 3 > $O/${TAG}_bench_n1_separate_eig.json 2>> $O/${TAG}_bench_n1.err
timeout 300 python tools/latency_profile.py > $O/${TAG}_latency_profile.log 2>&1; grep "^strat" $O/${TAG}_latency_profile.log
