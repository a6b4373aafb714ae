#!/bin/bash
tag=${1:-r02e}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_parity.log 2>&1; tail -n 6 gpurun_out/${tag}_parity.log
timeout 300 python tools/stress.py C4 300 > gpurun_out/${tag}_stress_c4.log 2>&1; tail -n 2 gpurun_out/${tag}_stress_c4.log
timeout 300 python tools/stress.py C1 150 > gpurun_out/${tag}_stress_c1.log 2>&1; tail -n 2 gpurun_out/${tag}_stress_c1.log
timeout 900 python bench.py --steps 40 --warmup 6 > gpurun_out/${tag}_bench_c1.json 2> gpurun_out/${tag}_bench_c1.err
tail -c 600 gpurun_out/${tag}_bench_c1.err
python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench_c1.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'sync',d['e2e']['one_frame_at_a_time']['value'],'strict',d['e2e']['strict']['value'])
print(d['stages_ms']); print({k:(v.get('env_frames_per_s') or v.get('frames_per_s')) for k,v in d['batched'].items()})
"
