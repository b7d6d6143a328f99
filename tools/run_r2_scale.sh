# Round-2 scaling pass on one 8-GPU box: the default bench line at N = 8, 4, 2 (strong scaling of c3, c4, c5a).
set -x
export D3B200_HANG_DUMP_S=100
P=gpurun_out
tr() { n=$1; port=$2; shift 2; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@"; }
tr 8 29631 bench.py --gpus 8 --steps 20 --warmup 5 > $P/r2_scale_n8.json 2> $P/r2_scale_n8.err; echo "rc=$? n8"
tr 4 29632 bench.py --gpus 4 --steps 20 --warmup 5 > $P/r2_scale_n4.json 2> $P/r2_scale_n4.err; echo "rc=$? n4"
tr 2 29633 bench.py --gpus 2 --steps 20 --warmup 5 > $P/r2_scale_n2.json 2> $P/r2_scale_n2.err; echo "rc=$? n2"
TAD_DFTD3_B200_TARGET_CTAS=3552 tr 8 29634 bench.py --gpus 8 --workload c4 --steps 20 --warmup 3 --no-cpu-baseline > $P/r2_scale_c4_n8_t24.json 2> $P/r2_scale_c4_n8_t24.err; echo "rc=$? c4 t24"
TAD_DFTD3_B200_TARGET_CTAS=1184 tr 8 29635 bench.py --gpus 8 --workload c4 --steps 20 --warmup 3 --no-cpu-baseline > $P/r2_scale_c4_n8_t8.json 2> $P/r2_scale_c4_n8_t8.err; echo "rc=$? c4 t8"
tr 8 29636 bench.py --gpus 8 --workload c3 --scaling weak --steps 20 --warmup 3 --no-cpu-baseline > $P/r2_scale_c3_weak_n8.json 2> $P/r2_scale_c3_weak_n8.err; echo "rc=$? c3 weak"
