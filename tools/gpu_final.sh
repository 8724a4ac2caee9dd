#!/bin/bash
# Round-end evidence in one gpurun call: GPU tests, smoke, default bench, reference arm, launch list, ncu summaries of
# the emission kernel, forward-backward and pipeline benches.  Outputs land in gpurun_out/.
set -u
TAG=${1:-r01j}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_gpu_$TAG.log)"
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
timeout 400 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
timeout 200 python tools/bench_fb.py > gpurun_out/bench_fb_$TAG.json 2> gpurun_out/bench_fb_$TAG.err; echo "fb rc=$?"
timeout 200 python tools/bench_pipeline.py > gpurun_out/bench_pipeline_$TAG.json 2> gpurun_out/bench_pipeline_$TAG.err; echo "pipeline rc=$?"
SHORT="python bench.py --steps 2 --warmup 1 --frames 300 --no-e2e --no-cpu"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'vit|forward|backtrace|pack' -c 60 --csv --log-file gpurun_out/launches_$TAG.csv $SHORT > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:emissions_reg -s 1 -c 1 -f -o gpurun_out/prof_emis_$TAG python tools/emis_probe.py softmax 1024 1000 > gpurun_out/ncu_emis_$TAG.log 2>&1
echo "ncu emis rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:fb_tc_pass -s 0 -c 1 -f -o gpurun_out/prof_fb_$TAG python tools/bench_fb.py --frames 300 --steps 1 --warmup 0 > gpurun_out/ncu_fb_$TAG.log 2>&1
echo "ncu fb rc=$?"
tail -c 600 gpurun_out/bench_$TAG.json
