# Final scaling pass (defaults of the committed tree): the driver's command at N = 8, 4, 2 + the multi-GPU parity tests at 8.
set -x
export D3B200_HANG_DUMP_S=100
P=gpurun_out
tr() { n=$1; port=$2; shift 2; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@"; }
tr 8 29641 bench.py --gpus 8 --steps 20 --warmup 5 > $P/r2_final_n8.json 2> $P/r2_final_n8.err; echo "rc=$? n8"
tr 4 29642 bench.py --gpus 4 --steps 20 --warmup 5 > $P/r2_final_n4.json 2> $P/r2_final_n4.err; echo "rc=$? n4"
tr 2 29643 bench.py --gpus 2 --steps 20 --warmup 5 > $P/r2_final_n2.json 2> $P/r2_final_n2.err; echo "rc=$? n2"
timeout 400 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -3
