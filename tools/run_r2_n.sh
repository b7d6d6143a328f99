set -x
for v in main al4 al2 cnb2; do
  if [ $v != main ]; then export TAD_DFTD3_B200_LIB=$PWD/build/variants/libd3_$v.so; fi
  timeout 100 python bench.py --workload c3 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_c3_${v}.json 2> gpurun_out/r2n_c3_$v.err; echo rc=$?
done
