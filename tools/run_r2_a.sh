set -x
export D3B200_HANG_DUMP_S=100
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612"
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo rc=$?
tail -c 600 gpurun_out/r2a_bench_n1.err
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2a_bench_n2.json 2> gpurun_out/r2a_bench_n2.err; echo rc=$?
tail -c 600 gpurun_out/r2a_bench_n2.err
timeout 400 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -5
