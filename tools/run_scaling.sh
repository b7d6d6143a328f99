# usage: bash tools/run_scaling.sh N   (inside gpurun --gpus N)
N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $N "${@:3}" > gpurun_out/scale_$1_n$N.json 2> gpurun_out/scale_$1_n$N.err; }
run c3 29601 --steps 20 --warmup 3
run c4 29602 --workload c4 --steps 5 --warmup 3
run c5a 29603 --workload c5a --steps 3 --warmup 3
for w in c3 c4 c5a; do python -c "
import json;d=json.loads([l for l in open('gpurun_out/scale_${w}_n$N.json') if l.startswith('{')][-1]);print('$w', d['n_gpus'], round(d['ms_per_step'],3), '%.4g' % d['value'], round(d['e2e']['ms_per_step'],3), d['config'].get('atm_ms'))" || tail -5 gpurun_out/scale_${w}_n$N.err; done
