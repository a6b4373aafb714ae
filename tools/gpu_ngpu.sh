#!/bin/bash
# both bench arms at N GPUs of one box, launched the way the driver does (torchrun):  bash tools/gpu_ngpu.sh <tag> <N> [N ...]
tag=$1; shift
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
for n in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 40 --warmup 6 > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --impl reference --gpus $n --steps 40 --warmup 6 > gpurun_out/${tag}_bench_${n}gpu_reference.json 2> gpurun_out/${tag}_bench_${n}gpu_reference.err
  python - <<PY
import json
n = $n
for suf in ("", "_reference"):
    f = f"gpurun_out/${tag}_bench_{n}gpu{suf}.json"
    try:
        d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
    except Exception as ex:
        print(f, "ERR", ex); print(open(f.replace(".json", ".err")).read()[-1500:]); continue
    e = d["e2e"]
    print(n, suf, "value", round(d["value"], 1), "e2e", round(e["value"], 1), {k: round(v["value"], 1) for k, v in e.items() if isinstance(v, dict)},
          {k: round(v.get("env_frames_per_s") or v.get("frames_per_s"), 1) for k, v in d.get("batched", {}).items()}, json.dumps(d.get("batched", {}).get("C4", {}).get("gather")), d.get("clocks"))
PY
done
