#!/bin/bash
mkdir -p gpurun_out
for on in 0 1; do
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:cost_kernel' -s 2 -c 1 -f -o gpurun_out/r02q_cost_$on python tools/cost_prof.py $on > gpurun_out/r02q_cost_$on.log 2>&1
tail -n 2 gpurun_out/r02q_cost_$on.log
done
