# Round-2 measurement pass on ONE B200 (every ncu run only after the same command exited 0 without ncu).
set -x
P=gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > $P/r2_bench_n1.json 2> $P/r2_bench_n1.err; echo "rc=$? bench all"
python bench.py --impl reference --steps 20 --warmup 5 > $P/r2_bench_ref.json 2> $P/r2_bench_ref.err; echo "rc=$? ref"
python bench.py --workload c3 --dtype f32 --steps 20 --warmup 3 --no-cpu-baseline > $P/r2_bench_c3_f32.json 2>> $P/r2_bench_n1.err; echo "rc=$? f32"
python bench.py --workload c5a --cutoff 50 --steps 3 --warmup 3 --no-cpu-baseline > $P/r2_bench_c5a_cut50.json 2>> $P/r2_bench_n1.err; echo "rc=$? c5a50"
python bench.py --workload c5b --steps 5 --warmup 3 > $P/r2_bench_c5b.json 2>> $P/r2_bench_n1.err; echo "rc=$? c5b"
# launch lists (per-launch durations are cold-cache and serialised under ncu: shares, not absolutes)
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $P/plain_c3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $P/r2_launches_bench_c3.csv python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $P/ncu_c3.log 2>&1
TAD_DFTD3_B200_NO_GRAPH=1 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > $P/plain_c4.log 2>&1 && TAD_DFTD3_B200_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sym_|system_|molecule_|stage_" -c 300 --csv --log-file $P/r2_launches_bench_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > $P/ncu_c4.log 2>&1
TAD_DFTD3_B200_NO_GRAPH=1 python bench.py --workload c5a --steps 1 --warmup 3 --no-cpu-baseline > $P/plain_c5a.log 2>&1 && TAD_DFTD3_B200_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sym_|system_|molecule_|stage_" -c 300 --csv --log-file $P/r2_launches_bench_c5a.csv python bench.py --workload c5a --steps 1 --warmup 3 --no-cpu-baseline > $P/ncu_c5a.log 2>&1
# full captures of the three dominant kernels
ncu --set full --clock-control none --import-source on -k regex:molecule_kernel_v2 -s 3 -c 1 -o $P/prof_r2_c3 -f python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $P/ncu_c3_full.log 2>&1; tail -1 $P/ncu_c3_full.log
TAD_DFTD3_B200_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:sym_pair_t -s 3 -c 1 -o $P/prof_r2_sym_pair -f python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > $P/ncu_c4_full.log 2>&1; tail -1 $P/ncu_c4_full.log
TAD_DFTD3_B200_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:system_atm3 -s 2 -c 1 -o $P/prof_r2_atm3 -f python bench.py --workload c5a --steps 1 --warmup 3 --no-cpu-baseline > $P/ncu_atm3_full.log 2>&1; tail -1 $P/ncu_atm3_full.log
