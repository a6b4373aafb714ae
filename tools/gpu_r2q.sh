#!/bin/bash
# capture after the streaming cost kernel, the paired final pass (D = 64 batches) and the third lane:
# all GPU tests, stress, smoke, both bench arms, C2 / C3 lines, ncu launch list + full-set captures (C1 frame, C4 batch)
tag=${1:-r02q}
mkdir -p gpurun_out
if [ "$2" != "noparity" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_parity.log 2>&1; tail -n 3 gpurun_out/${tag}_parity.log
timeout 300 python tools/stress.py C1 150 2>&1 | tail -n 1
timeout 300 python tools/stress.py C4 300 2>&1 | tail -n 1
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 1
fi
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_bench_c1_reference.json 2> gpurun_out/${tag}_bench_c1_reference.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_c1.json 2> gpurun_out/${tag}_bench_c1.err
for w in C2 C3; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-batched > gpurun_out/${tag}_bench_${w}.json 2> gpurun_out/${tag}_bench_${w}.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-batched > gpurun_out/${tag}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:front7|cost_kernel|aggr_|lr_median|dilate' -s 16 -c 8 -f -o gpurun_out/${tag}_full \
  python tools/frame_prof.py C1 > gpurun_out/${tag}_full.log 2>&1
# (a second full-set report would push gpurun_out/ past the 64 MiB that travel back: the C4 batch is captured by a
#  separate call: ncu --set full --clock-control none -k 'regex:aggr_wta|cost_kernel|lr_median|front7' -s 5 -c 4 -o gpurun_out/r02q_full_c4 python tools/batch_prof.py C4 256)
python tools/trace_cost.py C1 > gpurun_out/${tag}_trace_cost.txt 2>&1
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
          {k: round(v.get("env_frames_per_s") or v.get("frames_per_s"), 1) for k, v in d.get("batched", {}).items()}, d.get("clocks"), d.get("roofline"))
PY
