set -x
export D3B200_HANG_DUMP_S=100
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29622"
run() { name=$1; shift; timeout 150 "$@" > gpurun_out/r2n8_$name.json 2> gpurun_out/r2n8_$name.err; echo "rc=$? $name"; tail -c 200 gpurun_out/r2n8_$name.err; }
run c4_n8_fullrow $TR8 bench.py --gpus 8 --workload c4 --steps 20 --warmup 3 --no-cpu-baseline
TAD_DFTD3_B200_FULLROW=0 run c4_n8_sym $TR8 bench.py --gpus 8 --workload c4 --steps 20 --warmup 3 --no-cpu-baseline
TAD_DFTD3_B200_FULLROW=0 TAD_DFTD3_B200_SYM_JSPLIT=12 run c4_n8_sym12 $TR8 bench.py --gpus 8 --workload c4 --steps 20 --warmup 3 --no-cpu-baseline
TAD_DFTD3_B200_FULLROW=0 run c4_n4_sym $TR4 bench.py --gpus 4 --workload c4 --steps 20 --warmup 3 --no-cpu-baseline
run all_n8 $TR8 bench.py --gpus 8 --steps 20 --warmup 3
