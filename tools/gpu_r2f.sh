#!/bin/bash
# N=2: full GPU test suite on GPU 0, then both bench arms under torchrun on 2 GPUs
tag=${1:-r02f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_parity.log 2>&1; tail -n 4 gpurun_out/${tag}_parity.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/${tag}_bench_2gpu.json 2> gpurun_out/${tag}_bench_2gpu.err
grep -c "NCCL INFO" gpurun_out/${tag}_bench_2gpu.err; grep -m2 "nranks" gpurun_out/${tag}_bench_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 30 --warmup 5 > gpurun_out/${tag}_bench_2gpu_reference.json 2> gpurun_out/${tag}_bench_2gpu_reference.err
python - <<PY
import json
for f in ("gpurun_out/${tag}_bench_2gpu.json", "gpurun_out/${tag}_bench_2gpu_reference.json"):
    try:
        d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
    except Exception as ex:
        print(f, "ERR", ex); print(open(f.replace(".json", ".err")).read()[-2000:]); continue
    print(f, "value", d["value"], "e2e", d["e2e"]["value"], {k: (v.get("env_frames_per_s") or v.get("frames_per_s")) for k, v in d.get("batched", {}).items()})
    print(json.dumps(d.get("batched", {}).get("C4", {}).get("gather")))
PY
