#!/bin/bash
tag=${1:-r02j}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_parity.log 2>&1; tail -n 5 gpurun_out/${tag}_parity.log
timeout 300 python tools/stress.py C1 100 > gpurun_out/${tag}_stress_c1.log 2>&1; tail -n 1 gpurun_out/${tag}_stress_c1.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_c1.json 2> gpurun_out/${tag}_bench_c1.err
python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench_c1.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'sync',d['e2e']['one_frame_at_a_time']['value'],'strict',d['e2e']['strict']['value'], d['clocks'], d['arm']['host_enqueue_us_per_frame'])
"
