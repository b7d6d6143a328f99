set -x
for v in main ovh8 ovh12 ovh26; do
  if [ $v != main ]; then export TAD_DFTD3_B200_LIB=$PWD/build/variants/libd3_$v.so; fi
  timeout 200 python bench.py --workload c3 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_c3_${v}.json 2> gpurun_out/r2f_c3_$v.err; echo rc=$?
done
