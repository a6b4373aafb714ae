#!/usr/bin/env python
"""Runs a few batched frames of one workload (for ncu): python tools/batch_prof.py C4 64"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import configs
from sapien_b200 import simsense, synth

key, batch = sys.argv[1], int(sys.argv[2])
prm = configs.params(key)
pairs = [synth.make_pair(prm.rows, prm.cols, prm.max_disp, s)[:2] for s in range(2)]
l = np.stack([pairs[i % 2][0] for i in range(batch)])
r = np.stack([pairs[i % 2][1] for i in range(batch)])
tl, tr = torch.from_numpy(synth.to_rgba(l)).cuda(), torch.from_numpy(synth.to_rgba(r)).cuda()
eng = simsense.DepthSensorEngine(*prm.engine_args(), batch=batch)
for _ in range(3):
    eng.compute(tl, tr)
eng.synchronize()
