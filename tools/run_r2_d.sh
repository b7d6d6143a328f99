set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --workload c3 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_c3.json 2> gpurun_out/r2d_c3.err; echo rc=$?
tail -c 300 gpurun_out/r2d_c3.err
ncu --set full --clock-control none --import-source on -k regex:molecule_kernel_v2 -s 3 -c 1 -o gpurun_out/prof_r2d_c3 -f python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c3_full.log 2>&1
tail -2 gpurun_out/ncu_c3_full.log
