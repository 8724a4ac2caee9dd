#!/bin/bash
# One gpurun call: wave tests, configs 5 and 3 at their named sizes, ncu capture of the wide banded kernel.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k wave 2>&1 | tail -3
for A in dense auto; do
  timeout 600 python tools/bench_waves.py --config 5 --algo $A > gpurun_out/waves_cfg5_$A.json 2> gpurun_out/waves_cfg5_$A.err; echo "cfg5 $A rc=$?"; cat gpurun_out/waves_cfg5_$A.json
  timeout 600 python tools/bench_waves.py --config 3 --algo $A > gpurun_out/waves_cfg3_$A.json 2> gpurun_out/waves_cfg3_$A.err; echo "cfg3 $A rc=$?"; cat gpurun_out/waves_cfg3_$A.json
done
timeout 600 python tools/bench_waves.py --config 3 --model imm --algo dense > gpurun_out/waves_cfg3_imm.json 2> gpurun_out/waves_cfg3_imm.err; echo "cfg3 imm rc=$?"; cat gpurun_out/waves_cfg3_imm.json
SHORT="python bench.py --steps 2 --warmup 1 --states 722 --clips 1184 --frames 300 --algo banded --no-e2e --no-cpu"
$SHORT > gpurun_out/plain_wide.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wide_forward -s 1 -c 1 -f -o gpurun_out/prof_wide $SHORT > gpurun_out/ncu_wide.log 2>&1
echo "ncu wide rc=$?"
