#!/bin/bash
# Profile capture: bench (both arms), launch list, full ncu of one frame's kernels, per-path timeline.
#   gpurun -- 'bash tools/gpu_prof.sh r01d'
tag=${1:-cap}
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 10 > gpurun_out/${tag}_bench_c1.json 2> gpurun_out/${tag}_bench_c1.err
python bench.py --impl reference --steps 50 --warmup 10 > gpurun_out/${tag}_bench_c1_reference.json 2> gpurun_out/${tag}_bench_c1_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:front7|cost_kernel|aggr_|lr_median|dilate' -s 24 -c 8 -o gpurun_out/${tag}_full \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_full.log 2>&1
python tools/trace_aggr.py C1 > gpurun_out/${tag}_trace.txt 2>&1
python tools/e2e_breakdown.py > gpurun_out/${tag}_e2e_breakdown.txt 2>&1
for w in C2 C3 C5; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_${w}.json 2> gpurun_out/${tag}_bench_${w}.err; done
python bench.py --workload C4 --batch 256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_C4.json 2> gpurun_out/${tag}_bench_C4.err
ls gpurun_out | grep ${tag} | head -30
