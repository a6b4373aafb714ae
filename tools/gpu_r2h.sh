#!/bin/bash
# round 2 profile capture: bench (both arms), launch list, full ncu of one frame's kernels, sanitizers, experiments
tag=${1:-r02h}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 50 --warmup 10 > gpurun_out/${tag}_bench_c1.json 2> gpurun_out/${tag}_bench_c1.err
timeout 900 python bench.py --impl reference --steps 50 --warmup 10 > gpurun_out/${tag}_bench_c1_reference.json 2> gpurun_out/${tag}_bench_c1_reference.err
python -c "
import json
for f in ('gpurun_out/${tag}_bench_c1.json','gpurun_out/${tag}_bench_c1_reference.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['e2e']['value'], {k:(v.get('env_frames_per_s') or v.get('frames_per_s')) for k,v in d.get('batched',{}).items()})
"
timeout 600 python tools/cost_bpsm_ab.py > gpurun_out/${tag}_cost_bpsm.txt 2>&1; cat gpurun_out/${tag}_cost_bpsm.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-batched > gpurun_out/${tag}_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:front7|cost_kernel|aggr_|lr_median|dilate' -s 24 -c 8 -o gpurun_out/${tag}_full \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-batched > gpurun_out/${tag}_full.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k 'regex:simInfraredNoise|front7_kernel' --csv --log-file gpurun_out/${tag}_noise_traffic.csv \
  python tools/noise_traffic.py > gpurun_out/${tag}_noise_traffic.log 2>&1
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/stress.py small435 40 > gpurun_out/${tag}_sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${tag}_sanitize_$tool.log | tail -n 1) / $(grep -E 'mismatches' gpurun_out/${tag}_sanitize_$tool.log | tail -n 1)"
done
timeout 300 python tools/ref_stage_times.py C1 20 > gpurun_out/${tag}_ref_stages_c1.md 2>&1
timeout 300 python tools/ref_stage_times.py C4 50 > gpurun_out/${tag}_ref_stages_c4.md 2>&1
for w in C2 C3 C5; do timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-batched > gpurun_out/${tag}_bench_${w}.json 2> gpurun_out/${tag}_bench_${w}.err; done
timeout 600 python bench.py --workload C4 --batch 256 --steps 10 --warmup 3 --no-cpu-baseline --no-batched > gpurun_out/${tag}_bench_C4.json 2> gpurun_out/${tag}_bench_C4.err
for w in C2 C3; do timeout 600 python bench.py --impl reference --workload $w --steps 5 --warmup 2 --no-batched > gpurun_out/${tag}_bench_${w}_reference.json 2> gpurun_out/${tag}_bench_${w}_reference.err; done
ls gpurun_out | grep ${tag}
