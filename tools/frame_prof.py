#!/usr/bin/env python
"""Runs a few single frames of one workload (for ncu): python tools/cost_prof.py C1"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import configs
from sapien_b200 import simsense, synth

key = sys.argv[1] if len(sys.argv) > 1 else "C1"
lib = ctypes.CDLL(os.path.join(ROOT, "sapien_b200", "csrc", "libss_b200.so"))
prm = configs.params(key)
l, r = synth.make_pair(prm.rows, prm.cols, prm.max_disp, 0)[:2]
tl, tr = torch.from_numpy(synth.to_rgba(l)).cuda(), torch.from_numpy(synth.to_rgba(r)).cuda()
eng = simsense.DepthSensorEngine(*prm.engine_args())
for _ in range(4):
    eng.compute(tl, tr)
eng.synchronize()
