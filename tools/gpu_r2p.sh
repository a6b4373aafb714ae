#!/bin/bash
# final capture of round 2: all GPU tests, stress, smoke, both bench arms (C1 line with the batched block), C2 / C3 lines
tag=${1:-r02p}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_parity.log 2>&1; tail -n 4 gpurun_out/${tag}_parity.log
timeout 300 python tools/stress.py C1 150 2>&1 | tail -n 1
timeout 300 python tools/stress.py C4 300 2>&1 | tail -n 1
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 1
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_bench_c1_reference.json 2> gpurun_out/${tag}_bench_c1_reference.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_c1.json 2> gpurun_out/${tag}_bench_c1.err
for w in C2 C3; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-batched > gpurun_out/${tag}_bench_${w}.json 2> gpurun_out/${tag}_bench_${w}.err
done
python - <<PY
import json
for f in ("c1_reference", "c1", "C2", "C3"):
    p = "gpurun_out/${tag}_bench_%s.json" % f
    try:
        d = json.loads([l for l in open(p).read().strip().splitlines() if l.startswith("{")][-1])
    except Exception as ex:
        print(p, "ERR", ex); print(open(p.replace(".json", ".err")).read()[-1200:]); continue
    e = d["e2e"]
    print(f, "value", round(d["value"], 1), "e2e", round(e["value"], 1), {k: round(v["value"], 1) for k, v in e.items() if isinstance(v, dict)},
          {k: round(v.get("env_frames_per_s") or v.get("frames_per_s"), 1) for k, v in d.get("batched", {}).items()}, d.get("clocks"))
PY
