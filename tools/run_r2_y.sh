#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu -k "atm or c5a or system or tiled" 2>&1 | tail -3
VARIANTS="asm0 asm1" bash tools/run_r2_x.sh
