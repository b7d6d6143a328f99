#!/bin/bash
# ATM with per-edge factors in the pair table: parity tests, c5a step
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu -k "atm or c5a or system or tiled" 2>&1 | tail -5
timeout 200 python bench.py --workload c5a --steps 10 --warmup 3 > gpurun_out/r2v_c5a.json 2> gpurun_out/r2v_c5a.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2v_c5a.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"], 3), d["details"]["phases_ms_eager_rank0"]["sweep_atm"], d["details"]["checks"], d["cpu_baseline_port"])
PY
tail -3 gpurun_out/r2v_c5a.err
