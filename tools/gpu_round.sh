#!/bin/bash
# One gpurun call: GPU parity tests, the default bench, the ncu launch list of a short bench and one --set full
# capture of the forward kernel.  Outputs land in gpurun_out/ (summarised into profiles/ by tools/ncu_summary.py).
set -u
TAG=${1:-r1}
KREGEX=${2:-forward}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
SHORT="python bench.py --steps 2 --warmup 1 --frames 300 --no-e2e --no-cpu"
$SHORT > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'vit|forward|backtrace|pack' -c 60 --csv --log-file gpurun_out/launches_$TAG.csv $SHORT > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
$SHORT > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 1 -c 1 -f -o gpurun_out/prof_fwd_$TAG $SHORT > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
tail -c 1500 gpurun_out/bench_$TAG.json
