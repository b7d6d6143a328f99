set -x
for c in 0 592 -592 -296 -1184 -148; do
  timeout 200 python bench.py --workload c3 --steps 30 --warmup 3 --no-cpu-baseline --e2e-chunk=$c > gpurun_out/r2g_c3_e2e_$c.json 2> gpurun_out/r2g_c3_e2e_$c.err; echo rc=$?
  tail -c 200 gpurun_out/r2g_c3_e2e_$c.err
done
timeout 300 python -m pytest tests/test_host_entry.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python -m pytest tests -m gpu -x -q -k "atm or c5a" 2>&1 | tail -3
timeout 300 python bench.py --workload c5a --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_c5a.json 2> gpurun_out/r2g_c5a.err; echo rc=$?
