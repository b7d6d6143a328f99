set -x
for c in -9 -8 -10 -7 -1184 -2048; do
  timeout 120 python bench.py --workload c3 --steps 30 --warmup 3 --no-cpu-baseline --e2e-chunk=$c > gpurun_out/r2h_c3_e2e_$c.json 2> gpurun_out/r2h_c3_e2e_$c.err; echo "rc=$? $c"
  tail -c 200 gpurun_out/r2h_c3_e2e_$c.err
done
