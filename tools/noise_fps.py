#!/usr/bin/env python
"""C1 frames/s with the IR noise stage on (the stock StereoDepthSensorConfig: speckle 1.0, thermal 1.0) and off."""
import dataclasses, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import configs
from sapien_b200 import simsense, synth

base = configs.params("C1")
sets = []
for s in range(4):
    l, r = synth.make_pair(base.rows, base.cols, base.max_disp, s)[:2]
    sets.append((torch.from_numpy(synth.to_rgba(l)).cuda(), torch.from_numpy(synth.to_rgba(r)).cuda()))
for name, prm in (("noise off", base), ("noise on (stock)", dataclasses.replace(base, speckle_shape=1333.33, speckle_scale=1 / 1333.33, gaussian_mu=0.0, gaussian_sigma=0.25))):
    eng = simsense.DepthSensorEngine(*prm.engine_args())
    es = torch.cuda.ExternalStream(eng.cuda_stream)
    for i in range(6):
        eng.compute(*sets[i % 4], stream=eng.cuda_stream, sync=False)
    eng.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(es)
    n = 60
    for i in range(n):
        eng.compute(*sets[i % 4], stream=eng.cuda_stream, sync=False)
    e1.record(es)
    es.synchronize()
    ms = e0.elapsed_time(e1) / n
    eng.set_profiling(True); eng.get_stage_times()
    for i in range(10):
        eng.compute(*sets[i % 4], stream=eng.cuda_stream, sync=False)
    eng.synchronize()
    st = dict(eng.get_stage_times()); st.pop("frames", None)
    print(json.dumps({"config": name, "ms_per_frame": ms, "fps": 1e3 / ms, "front_ms_one_lane": st.get("front")}), flush=True)
    del eng
