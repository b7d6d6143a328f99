#!/bin/bash
# pair-table variant of the ATM sweep: parity tests, then same-box A/B of the c5a step
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu -k "atm or c5a or system or tiled" 2>&1 | tail -5
timeout 200 python bench.py --workload c5a --steps 10 --warmup 3 > gpurun_out/r2t_c5a_tab.json 2> gpurun_out/r2t_c5a_tab.err
D3B200_ATM_PAIRTAB_MAX_BYTES=0 timeout 200 python bench.py --workload c5a --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_c5a_notab.json 2> gpurun_out/r2t_c5a_notab.err
python - <<'PY'
import json
for n in ("tab", "notab"):
    try:
        d = json.loads(open(f"gpurun_out/r2t_c5a_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d.get("e2e", {}).get("value"), d.get("details", {}).get("phases_ms_eager_rank0"), d.get("cpu_baseline", {}).get("max_rel_err_gpu_vs_port"))
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/r2t_c5a_tab.err
