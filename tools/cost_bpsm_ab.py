#!/usr/bin/env python
"""Experiment: size the cost kernel's grid for 1 block per SM (half the register file) so that the other lane's
HBM-bound pass can be resident beside it.  C1, device inputs, frames back to back on two lanes."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import configs
from sapien_b200 import simsense, synth

lib = ctypes.CDLL(os.path.join(ROOT, "sapien_b200", "csrc", "libss_b200.so"))
prm = configs.params("C1")
sets = []
for s in range(4):
    l, r = synth.make_pair(prm.rows, prm.cols, prm.max_disp, s)[:2]
    sets.append((torch.from_numpy(synth.to_rgba(l)).cuda(), torch.from_numpy(synth.to_rgba(r)).cuda()))
for bpsm in (0, 1, 0, 1):
    lib.ssb_debug_set_cost_bpsm(bpsm)
    for lanes in (1, 2):
        eng = simsense.DepthSensorEngine(*prm.engine_args(), lanes=lanes)
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
        print(json.dumps({"cost_blocks_per_sm": bpsm or "occupancy (2)", "lanes": lanes, "ms_per_frame": ms, "fps": 1e3 / ms}), flush=True)
        del eng
