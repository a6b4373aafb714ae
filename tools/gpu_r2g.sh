#!/bin/bash
tag=${1:-r02g}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_parity.log 2>&1; tail -n 6 gpurun_out/${tag}_parity.log
timeout 600 python tools/pack_ab.py > gpurun_out/${tag}_pack_ab.txt 2>&1; cat gpurun_out/${tag}_pack_ab.txt | cut -c1-420
