#!/bin/bash
# 8-GPU box: both bench arms at N = 8 and N = 4 (torchrun), topology record
tag=${1:-r02i}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/${tag}_topo.txt 2>&1
for n in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 40 --warmup 6 > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --impl reference --gpus $n --steps 40 --warmup 6 > gpurun_out/${tag}_bench_${n}gpu_reference.json 2> gpurun_out/${tag}_bench_${n}gpu_reference.err
done
grep -m3 "nranks" gpurun_out/${tag}_bench_8gpu.err | cut -c1-200
python - <<PY
import json
for n in (8, 4):
    for suf in ("", "_reference"):
        f = f"gpurun_out/${tag}_bench_{n}gpu{suf}.json"
        try:
            d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        except Exception as ex:
            print(f, "ERR", ex); print(open(f.replace(".json", ".err")).read()[-1500:]); continue
        e = d["e2e"]
        print(n, suf, "value", round(d["value"], 1), "e2e", round(e["value"], 1), {k: round(v["value"], 1) for k, v in e.items() if isinstance(v, dict)},
              {k: round(v.get("env_frames_per_s") or v.get("frames_per_s"), 1) for k, v in d.get("batched", {}).items()}, json.dumps(d.get("batched", {}).get("C4", {}).get("gather")))
PY
