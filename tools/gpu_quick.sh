#!/bin/bash
# quick parity probe + kernel-only bench (+ timing experiments with VIT_DEV_FLAGS: results invalid, timing only)
set -u
TAG=${1:-q}
timeout 300 python tools/gpu_check.py > gpurun_out/check_$TAG.log 2>&1; echo "check rc=$? bad=$(grep -c -E 'False|ERROR' gpurun_out/check_$TAG.log)"
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu"
$B > gpurun_out/benchq_$TAG.json 2>gpurun_out/benchq_$TAG.err; python - <<PY
import json
d=json.load(open('gpurun_out/benchq_$TAG.json'))
print('full   : step %.2f ms fwd %.2f ms frac %.4f parity %s' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['parity_vs_oracle']))
PY
for F in ${2:-}; do
VIT_DEV_FLAGS=$F $B > gpurun_out/benchq_${TAG}_dev$F.json 2>/dev/null; python - <<PY
import json
d=json.load(open('gpurun_out/benchq_${TAG}_dev$F.json'))
print('dev=$F  : step %.2f ms fwd %.2f ms' % (d['ms_per_step'], d['roofline']['kernel_ms']))
PY
done
