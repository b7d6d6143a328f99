#!/bin/bash
# last pass of round 2 on the committed tree: GPU tests, driver-style bench line, ncu capture + launch list of the c5a step
P=gpurun_out
mkdir -p $P
timeout 150 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 150 python bench.py --steps 20 --warmup 5 > $P/r2_bench_n1_final.json 2> $P/r2_bench_n1_final.err; echo "rc=$? bench"
TAD_DFTD3_B200_NO_GRAPH=1 timeout 120 ncu --set full --clock-control none --import-source on -k regex:system_atm3 -s 2 -c 1 -o $P/prof_r2_atm3_tab -f python bench.py --workload c5a --steps 1 --warmup 3 --no-cpu-baseline > $P/ncu_atm3_tab_full.log 2>&1; tail -1 $P/ncu_atm3_tab_full.log
TAD_DFTD3_B200_NO_GRAPH=1 timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sym_|system_|molecule_|stage_" -c 300 --csv --log-file $P/r2_launches_bench_c5a_tab.csv python bench.py --workload c5a --steps 1 --warmup 3 --no-cpu-baseline > $P/ncu_c5a_tab.log 2>&1
