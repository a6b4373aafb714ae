#!/bin/bash
tag=${1:-r02k}
mkdir -p gpurun_out
timeout 600 python tools/cost_tx_ab.py > gpurun_out/${tag}_cost_tx.txt 2>&1; cat gpurun_out/${tag}_cost_tx.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/e2e_diag.py > gpurun_out/${tag}_e2e_diag_8.txt 2> gpurun_out/${tag}_e2e_diag_8.err
grep -A12 "^# e2e diagnosis" gpurun_out/${tag}_e2e_diag_8.err gpurun_out/${tag}_e2e_diag_8.txt | cut -c1-200
