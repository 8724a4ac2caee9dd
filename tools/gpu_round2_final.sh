#!/bin/bash
# End-of-round validation in one gpurun call: GPU parity tests, smoke, the default bench line, the reference arm, and
# full-size ncu --set full captures of the forward-backward kernels (DRAM traffic per launch -> profiles/forward_traffic.json).
set -u
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu_$TAG.log)"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke_$TAG.log | cut -c1-100)"
S0=$(date +%s)
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$? wall $(( $(date +%s) - S0 )) s"
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "reference arm rc=$?"
for IMPL in banded tc; do
  K=fb_conv_pass; [ $IMPL = tc ] && K=fb_tc_pass
  FB="python tools/bench_fb.py --impl $IMPL --steps 1 --warmup 0"
  timeout 100 $FB > gpurun_out/fb_plain_${IMPL}_$TAG.log 2>&1 &&
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -s 0 -c 2 -f -o gpurun_out/prof_fullsize_${IMPL}_$TAG $FB > gpurun_out/ncu_fullsize_${IMPL}_$TAG.log 2>&1
  echo "ncu $K (1024 x 3000 x 361) rc=$?"
done
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
fb=d.get('forward_backward',{})
print('value %.1f M  e2e %.1f M  frac %.4f  traffic %s  fb tc %.2f ms  fb structured %.2f ms (frac %.3f)' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['roofline'].get('traffic'), fb.get('ms_per_step',0), fb.get('structured_fast_path',{}).get('ms_per_step',0), fb.get('structured_fast_path',{}).get('roofline_hbm',{}).get('frac',0)))
PY
