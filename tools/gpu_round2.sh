#!/bin/bash
# Round-2 validation in one gpurun call: GPU parity tests, smoke, the default bench line (all legs), the reference arm, and
# ncu --set full captures of the two structured forward-backward kernels (after their command has exited 0 without ncu).
set -u
TAG=${1:-r02d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu_$TAG.log)"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke_$TAG.log | cut -c1-100)"
S0=$(date +%s)
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$? wall $(( $(date +%s) - S0 )) s"
S0=$(date +%s)
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "reference arm rc=$? wall $(( $(date +%s) - S0 )) s"
FB="python tools/bench_fb.py --impl banded --frames 600 --steps 1 --warmup 0"
timeout 100 $FB > gpurun_out/fb_plain_$TAG.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fb_conv_pass -s 0 -c 2 -f -o gpurun_out/prof_fb_conv_$TAG $FB > gpurun_out/ncu_fb_conv_$TAG.log 2>&1
echo "ncu fb_conv rc=$?"
VIT_FB_CONV=0 timeout 100 $FB > gpurun_out/fb_plain2_$TAG.log 2>&1 &&
VIT_FB_CONV=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fb_banded_pass -s 0 -c 2 -f -o gpurun_out/prof_fb_banded_$TAG $FB > gpurun_out/ncu_fb_banded_$TAG.log 2>&1
echo "ncu fb_banded rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
fb=d.get('forward_backward',{})
print('value %.1f M  e2e %.1f M  frac %.4f  traffic %s  fb tc %.2f ms  fb structured %.2f ms (frac %.3f)' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['roofline'].get('traffic'), fb.get('ms_per_step',0), fb.get('structured_fast_path',{}).get('ms_per_step',0), fb.get('structured_fast_path',{}).get('roofline_hbm',{}).get('frac',0)))
PY
