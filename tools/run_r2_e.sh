set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "cuda_parity or full_size" 2>&1 | tail -3
for v in main nr257; do
  if [ $v != main ]; then export TAD_DFTD3_B200_LIB=$PWD/build/variants/libd3_$v.so; fi
  for rep in 1 2; do
  timeout 200 python bench.py --workload c3 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_c3_${v}_$rep.json 2> gpurun_out/r2e_c3_$v.err; echo rc=$?
  done
done
