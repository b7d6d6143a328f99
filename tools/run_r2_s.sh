set -x
for rep in 1 2; do
for v in main prev; do
  if [ $v != main ]; then export TAD_DFTD3_B200_LIB=$PWD/build/variants/libd3_$v.so; else unset TAD_DFTD3_B200_LIB; fi
  timeout 100 python bench.py --workload c3 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_c3_${v}_$rep.json 2> gpurun_out/r2s_c3_$v.err; echo rc=$?
done
done
