#!/bin/bash
# round 2, call A: parity (all GPU tests), bench both arms with the batched block, pipeline experiment, reference stage times
tag=${1:-r02a}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_parity.log 2>&1
tail -n 15 gpurun_out/${tag}_parity.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/${tag}_bench_c1.json 2> gpurun_out/${tag}_bench_c1.err
tail -c 1500 gpurun_out/${tag}_bench_c1.err
timeout 600 python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/${tag}_bench_c1_reference.json 2> gpurun_out/${tag}_bench_c1_reference.err
tail -c 800 gpurun_out/${tag}_bench_c1_reference.err
timeout 400 python tools/pipe_experiment.py C4 1024 1,2,3,4 6 > gpurun_out/${tag}_pipes_c4.txt 2>&1
timeout 300 python tools/pipe_experiment.py C3 64 1,2,4 4 > gpurun_out/${tag}_pipes_c3.txt 2>&1
timeout 300 python tools/pipe_experiment.py C5 4 1,2 6 > gpurun_out/${tag}_pipes_c5.txt 2>&1
cat gpurun_out/${tag}_pipes_c*.txt
timeout 300 python tools/ref_stage_times.py C1 20 > gpurun_out/${tag}_ref_stages_c1.md 2>&1
timeout 300 python tools/ref_stage_times.py C4 50 > gpurun_out/${tag}_ref_stages_c4.md 2>&1
