set -x
timeout 300 python -m pytest tests -m gpu -x -q -k "c4 or tiled or engine" 2>&1 | tail -3
for v in main batch1 batch2; do
  if [ $v != main ]; then export TAD_DFTD3_B200_LIB=$PWD/build/variants/libd3_$v.so; fi
  timeout 200 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_c4_${v}.json 2> gpurun_out/r2p_c4_$v.err; echo rc=$?
done
