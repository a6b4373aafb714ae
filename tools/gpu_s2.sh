#!/bin/bash
# GPU session 2: instruction probes, cost-kernel variants x fork, parity under the variants
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench tools/ubench.cu && /tmp/ubench > gpurun_out/ubench.txt 2>&1
for tx in 0 24 32; do for fk in 0 2; do
  SSB_COST_TX=$tx SSB_AGGR_FORK=$fk python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_tx${tx}_fk${fk}.json 2> gpurun_out/bench_tx${tx}_fk${fk}.err
done; done
SSB_COST_TX=32 SSB_AGGR_FORK=2 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "small_all_stages or c1_full or c5 or golden" > gpurun_out/parity_tx32_fk2.log 2>&1
SSB_COST_TX=24 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "c1_full" > gpurun_out/parity_tx24.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/parity_default.log 2>&1
tail -3 gpurun_out/parity_*.log
