timeout 120 python tools/probe_overlap.py 2>&1 | tail -8
