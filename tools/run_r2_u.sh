#!/bin/bash
# ATM pair-table variants, same box: CTAs per SM (3 -> 4, 128 registers) and R_jk from the table
mkdir -p gpurun_out
for v in base tab4 tabr tab4r; do
  lib=""; [ $v != base ] && lib=$PWD/tad_dftd3_b200/libd3_$v.so
  TAD_DFTD3_B200_LIB=$lib timeout 200 python bench.py --workload c5a --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_$v.json 2> gpurun_out/r2u_$v.err
  python - $v <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2u_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["ms_per_step"], 3), d["details"]["phases_ms_eager_rank0"]["sweep_atm"], d["details"]["checks"])
except Exception as e:
    print(n, "failed", e)
PY
done
