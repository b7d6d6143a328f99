# compute-sanitizer evidence: memcheck and racecheck over one launch of every kernel family.
# Wrapped in timeouts: a hung sanitizer run must not hold the GPU box.
set -x
timeout 60 python tools/sanitizer_smoke.py > gpurun_out/sanitizer_plain.log 2>&1; echo "plain rc=$?"
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python tools/sanitizer_smoke.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -8 gpurun_out/sanitizer_memcheck.log
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python tools/sanitizer_smoke.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -8 gpurun_out/sanitizer_racecheck.log
