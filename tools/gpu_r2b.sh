#!/bin/bash
# round 2, call B: parity, stress (mixed modes incl. pipelined host frames), bench both arms, sanitizer initcheck
tag=${1:-r02b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_parity.log 2>&1
tail -n 12 gpurun_out/${tag}_parity.log
timeout 300 python tools/stress.py C4 300 > gpurun_out/${tag}_stress_c4.log 2>&1; tail -n 3 gpurun_out/${tag}_stress_c4.log
timeout 300 python tools/stress.py C1 120 > gpurun_out/${tag}_stress_c1.log 2>&1; tail -n 3 gpurun_out/${tag}_stress_c1.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/${tag}_bench_c1.json 2> gpurun_out/${tag}_bench_c1.err
tail -c 600 gpurun_out/${tag}_bench_c1.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/${tag}_bench_c1.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'sync',d['e2e']['one_frame_at_a_time']['value'],'strict',d['e2e']['strict']['value'])
print(d['stages_ms'])
"
for tool in initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/stress.py small435 40 > gpurun_out/${tag}_sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY' gpurun_out/${tag}_sanitize_$tool.log | tail -n 1)"
  grep -E "mismatches" gpurun_out/${tag}_sanitize_$tool.log | tail -n 1
done
