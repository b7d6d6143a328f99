set -x
timeout 300 python -m pytest tests -m gpu -x -q -k "f32 or fp32 or host or reference_runs" 2>&1 | tail -3
timeout 100 python bench.py --workload c3 --dtype f32 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_c3_f32.json 2>&1; echo rc=$?
timeout 100 python bench.py --workload c3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_c3_f64.json 2>&1; echo rc=$?
