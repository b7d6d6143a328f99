set -x
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python bench.py --workload c3 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_c3.json 2> gpurun_out/r2r_c3.err; echo rc=$?
timeout 100 python bench.py --workload c3 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_c3b.json 2>> gpurun_out/r2r_c3.err; echo rc=$?
