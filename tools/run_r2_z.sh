#!/bin/bash
# two GPUs: sharded parity tests and the c5a step with the pair-table ATM sweep
set -x
export D3B200_HANG_DUMP_S=100
P=gpurun_out
timeout 200 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29653 bench.py --gpus 2 --workload c5a --steps 10 --warmup 3 --no-cpu-baseline > $P/r2z_c5a_n2.json 2> $P/r2z_c5a_n2.err; echo "rc=$? n2"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2z_c5a_n2.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["ms_per_step"], 3), d["details"]["checks"])
PY
