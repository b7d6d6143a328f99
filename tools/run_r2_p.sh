set -x
for v in chunk4 chunk16; do
  export TAD_DFTD3_B200_LIB=$PWD/build/variants/libd3_$v.so
  timeout 200 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_c4_${v}.json 2> gpurun_out/r2p_c4_$v.err; echo rc=$?
done
