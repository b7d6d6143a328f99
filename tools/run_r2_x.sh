#!/bin/bash
# ATM pair-table sweep: table loads around L1 (cg) and next-record prefetch (j), same box
mkdir -p gpurun_out
for v in $VARIANTS; do
  TAD_DFTD3_B200_LIB=$PWD/tad_dftd3_b200/libd3_$v.so timeout 200 python bench.py --workload c5a --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_$v.json 2> gpurun_out/r2x_$v.err
  python - $v <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2x_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["ms_per_step"], 3), d["details"]["phases_ms_eager_rank0"]["sweep_atm"], d["details"]["checks"]["total_energy"], d["details"]["checks"]["max_abs_net_force"])
except Exception as e:
    print(n, "failed", e)
PY
done
