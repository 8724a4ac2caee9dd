#!/bin/bash
# After the forward-backward / emission kernel work: GPU tests, smoke, forward-backward bench + ncu summary, pipeline bench.
set -u
TAG=${1:-r01k}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu_$TAG.log)"
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
timeout 200 python tools/bench_fb.py > gpurun_out/bench_fb_$TAG.json 2> gpurun_out/bench_fb_$TAG.err; echo "fb rc=$?"
timeout 200 python tools/bench_pipeline.py > gpurun_out/bench_pipeline_$TAG.json 2> gpurun_out/bench_pipeline_$TAG.err; echo "pipeline rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:fb_tc_pass -s 0 -c 1 -f -o gpurun_out/prof_fb_$TAG python tools/bench_fb.py --frames 300 --steps 1 --warmup 0 > gpurun_out/ncu_fb_$TAG.log 2>&1
echo "ncu fb rc=$?"
cat gpurun_out/bench_fb_$TAG.json | cut -c1-300
