#!/bin/bash
# c4 pair sweep: L1 prefetch of the staged tile's T rows / T loads before the distance test, same box
mkdir -p gpurun_out
for v in $VARIANTS; do
  lib=""; [ $v != base ] && lib=$PWD/tad_dftd3_b200/libd3_$v.so
  TAD_DFTD3_B200_LIB=$lib timeout 100 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c4x_$v.json 2> gpurun_out/r2c4x_$v.err
  python - $v <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2c4x_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["ms_per_step"], 4), d["details"]["phases_ms_eager_rank0"]["sweep_pair"], d["details"]["checks"]["total_energy"], d["details"]["checks"]["max_abs_net_force"])
except Exception as e:
    print(n, "failed", e)
PY
done
