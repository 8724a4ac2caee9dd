#!/bin/bash
# Structured forward-backward across batch sizes and state sets -> one JSON line each (gpurun_out/fb_table.jsonl)
mkdir -p gpurun_out; : > gpurun_out/fb_table.jsonl
for S in 321 361; do for B in 256 1024 1184 2368 4096; do
  timeout 120 python tools/bench_fb.py --impl banded --states $S --clips $B --steps 3 --warmup 1 2>/dev/null >> gpurun_out/fb_table.jsonl
done; done
for B in 1024 4096; do timeout 200 python tools/bench_fb.py --impl banded --states 722 --clips $B --steps 2 --warmup 1 2>/dev/null >> gpurun_out/fb_table.jsonl; done
python - <<PY
import json
for l in open('gpurun_out/fb_table.jsonl'):
    d=json.loads(l); print(d['config']['workload'], '%.2f ms' % d['ms_per_step'], '%.0f M frames/s' % (d['value']/1e6), 'hbm frac %.3f' % d['roofline_hbm']['frac'], 'gamma err %.1e' % d['parity']['max_abs_gamma_err'])
PY
