# usage: bash tools/run_c4_scaling_variants.sh N   (inside gpurun --gpus N)
N=$1
port=29800
for v in default 4 8 12 16; do
  port=$((port+1))
  if [ "$v" = default ]; then unset TAD_DFTD3_B200_SYM_JSPLIT; else export TAD_DFTD3_B200_SYM_JSPLIT=$v; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --workload c4 --steps 10 --warmup 3 > gpurun_out/c4var_${v}_n$N.json 2> gpurun_out/c4var_${v}_n$N.err
  python -c "
import json;d=json.loads([l for l in open('gpurun_out/c4var_${v}_n$N.json') if l.startswith('{')][-1]);print('$v', d['n_gpus'], round(d['ms_per_step'],3), '%.4g' % d['value'], round(d['e2e']['ms_per_step'],3))" || tail -3 gpurun_out/c4var_${v}_n$N.err
done
