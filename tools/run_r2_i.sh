set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "cuda_parity or full_size or host" 2>&1 | tail -3
for v in main unweighted; do
  if [ $v != main ]; then export TAD_DFTD3_B200_LIB=$PWD/build/variants/libd3_$v.so; fi
  timeout 200 python bench.py --workload c3 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_c3_${v}.json 2> gpurun_out/r2i_c3_$v.err; echo rc=$?
done
