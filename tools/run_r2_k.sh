set -x
timeout 300 python -m pytest tests -m gpu -x -q -k "atm or c5a or engine" 2>&1 | tail -3
timeout 300 python bench.py --workload c5a --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_c5a.json 2> gpurun_out/r2k_c5a.err; echo rc=$?
timeout 100 python bench.py --workload c3 --dtype f32 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_c3_f32.json 2>&1; echo rc=$?
