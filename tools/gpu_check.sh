#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/parity34.log 2>&1
tail -n 3 gpurun_out/parity34.log
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench34.json 2> gpurun_out/bench34.err
timeout 300 python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench34_C3.json 2> gpurun_out/bench34_C3.err
timeout 300 python bench.py --workload C4 --batch 256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench34_C4.json 2> gpurun_out/bench34_C4.err
timeout 300 python bench.py --workload C5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench34_C5.json 2> gpurun_out/bench34_C5.err
timeout 300 python bench.py --workload C2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench34_C2.json 2> gpurun_out/bench34_C2.err
