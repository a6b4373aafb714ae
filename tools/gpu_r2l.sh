#!/bin/bash
tag=${1:-r02l}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_parity.log 2>&1; tail -n 5 gpurun_out/${tag}_parity.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 2
timeout 300 python tools/noise_fps.py 2>&1 | tail -n 3
timeout 600 ncu --metrics gpu__time_duration.sum -k 'regex:front7_kernel' --csv --log-file gpurun_out/${tag}_noise_kernels.csv python tools/noise_traffic.py > /dev/null 2>&1
grep front7 gpurun_out/${tag}_noise_kernels.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '
echo
