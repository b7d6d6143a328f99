set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1h_bench_ref.json 2>gpurun_out/r1h_ref.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r1h_bench_c3.json 2>gpurun_out/r1h_c3.err
python bench.py --steps 20 --warmup 3 --dtype f32 --no-cpu-baseline > gpurun_out/r1h_bench_c3_f32.json 2>>gpurun_out/r1h_c3.err
python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/r1h_bench_c4.json 2>gpurun_out/r1h_c4.err
python bench.py --workload c5a --steps 5 --warmup 3 > gpurun_out/r1h_bench_c5a.json 2>gpurun_out/r1h_c5a.err
python bench.py --workload c5a --cutoff 50 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1h_bench_c5a_cut50.json 2>>gpurun_out/r1h_c5a.err
python bench.py --workload c5b --steps 5 --warmup 3 > gpurun_out/r1h_bench_c5b.json 2>gpurun_out/r1h_c5b.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_c3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1h_launches_bench_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1
python bench.py --workload c5a --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_c5a.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sym_|system_|molecule_|stage_" -c 200 --csv --log-file gpurun_out/r1h_launches_bench_c5a.csv python bench.py --workload c5a --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c5a.log 2>&1
python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_c4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sym_|system_|molecule_|stage_" -c 200 --csv --log-file gpurun_out/r1h_launches_bench_c4.csv python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:molecule_kernel_v2 -s 3 -c 1 -o gpurun_out/prof_r1h_c3 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c3_full.log 2>&1
tail -2 gpurun_out/ncu_c3_full.log
ncu --set full --clock-control none --import-source on -k regex:system_atm3 -s 1 -c 1 -o gpurun_out/prof_r1h_atm3 -f python bench.py --workload c5a --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_atm3_full.log 2>&1
tail -2 gpurun_out/ncu_atm3_full.log
python -c 'import __graft_entry__ as g; g.smoke()'
