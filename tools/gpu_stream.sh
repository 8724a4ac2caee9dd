#!/bin/bash
set -u
mkdir -p gpurun_out
for M in jdc imm; do
timeout 600 python tools/bench_waves.py --config 3 --model $M --algo dense > gpurun_out/waves_cfg3_${M}_stream.json 2> gpurun_out/waves_cfg3_${M}_stream.err; echo "cfg3 $M rc=$?"; cat gpurun_out/waves_cfg3_${M}_stream.json
done
SHORT="python bench.py --steps 2 --warmup 1 --states 722 --clips 2072 --frames 200 --algo stream --no-e2e --no-cpu"
$SHORT > gpurun_out/plain_stream.log 2>&1; tail -c 600 gpurun_out/plain_stream.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_forward -s 1 -c 1 -f -o gpurun_out/prof_stream $SHORT > gpurun_out/ncu_stream.log 2>&1
echo "ncu stream rc=$?"
