#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/parity41.log 2>&1
tail -n 3 gpurun_out/parity41.log
for w in C1 C3 C5; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench41_$w.json 2> gpurun_out/bench41_$w.err; done
timeout 300 python bench.py --workload C4 --batch 256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench41_C4.json 2> gpurun_out/bench41_C4.err
