#!/bin/bash
tag=${1:-r02d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_parity.log 2>&1; tail -n 4 gpurun_out/${tag}_parity.log
timeout 300 python tools/pipe_experiment.py C1 2 1,2 20 > gpurun_out/${tag}_pipes_c1.txt 2>&1
timeout 300 python tools/pipe_experiment.py C1 4 2,4 10 >> gpurun_out/${tag}_pipes_c1.txt 2>&1
timeout 300 python tools/pipe_experiment.py C5 2 1,2 10 >> gpurun_out/${tag}_pipes_c1.txt 2>&1
cat gpurun_out/${tag}_pipes_c1.txt
