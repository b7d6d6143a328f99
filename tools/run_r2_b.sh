set -x
export D3B200_HANG_DUMP_S=100
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613"
for v in default sym; do
  if [ $v = sym ]; then export TAD_DFTD3_B200_FULLROW=0; fi
  timeout 200 $TR bench.py --gpus 2 --workload c4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_c4_n2_$v.json 2> gpurun_out/r2b_c4_n2_$v.err; echo rc=$?
  tail -c 300 gpurun_out/r2b_c4_n2_$v.err
done
unset TAD_DFTD3_B200_FULLROW
TAD_DFTD3_B200_FULLROW=1 timeout 200 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_c4_n1_fullrow.json 2> gpurun_out/r2b_c4_n1_fullrow.err
timeout 300 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -3
