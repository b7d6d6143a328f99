set -x
export D3B200_HANG_DUMP_S=100
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614"
timeout 400 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 200 $TR bench.py --gpus 2 --workload c5a --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_c5a_n2.json 2> gpurun_out/r2j_c5a_n2.err; echo rc=$?
timeout 200 $TR bench.py --gpus 2 --workload c4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_c4_n2.json 2> gpurun_out/r2j_c4_n2.err; echo rc=$?
