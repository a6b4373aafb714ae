#!/usr/bin/env python
"""Runs a few frames of one workload with the cost-kernel schedule given on the command line (for ncu)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import configs
from sapien_b200 import simsense, synth

on, key = int(sys.argv[1]), (sys.argv[2] if len(sys.argv) > 2 else "C1")
lib = ctypes.CDLL(os.path.join(ROOT, "sapien_b200", "csrc", "libss_b200.so"))
lib.ssb_debug_set_cost_stream(on)
prm = configs.params(key)
l, r = synth.make_pair(prm.rows, prm.cols, prm.max_disp, 0)[:2]
tl, tr = torch.from_numpy(synth.to_rgba(l)).cuda(), torch.from_numpy(synth.to_rgba(r)).cuda()
eng = simsense.DepthSensorEngine(*prm.engine_args(), lanes=1) if "lanes" in (simsense.DepthSensorEngine.__init__.__doc__ or "") else simsense.DepthSensorEngine(*prm.engine_args())
for _ in range(4):
    eng.compute(tl, tr)
eng.synchronize()
