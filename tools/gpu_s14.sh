#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/parity14.log 2>&1
tail -n 3 gpurun_out/parity14.log
python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench14_C1.json 2> gpurun_out/bench14_C1.err
python bench.py --workload C2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench14_C2.json 2> gpurun_out/bench14_C2.err
python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench14_C3.json 2> gpurun_out/bench14_C3.err
python bench.py --workload C4 --batch 256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench14_C4.json 2> gpurun_out/bench14_C4.err
python bench.py --workload C5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench14_C5.json 2> gpurun_out/bench14_C5.err
