#!/usr/bin/env python
"""IR noise stage: one noisy C1 frame on the reference (per-pixel XORWOW states read and written back) and on this
engine (stateless Philox inside the front-end kernel).  Run under
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:simInfraredNoise\\|front7_kernel --csv
to get the DRAM traffic of the noise kernels (profiles/r02_noise_traffic.md)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dataclasses
import torch
from oracle import RefEngine, configs
from sapien_b200 import simsense

prm = dataclasses.replace(configs.params("C1"), speckle_shape=1333.33, speckle_scale=1 / 1333.33, gaussian_mu=0.0, gaussian_sigma=0.25)
l, r = configs.pair(prm, 0)
ref = RefEngine(prm)
for _ in range(2):
    ref.compute_host(l, r)
ref.close()
eng = simsense.DepthSensorEngine(*prm.engine_args(), lanes=1)
for _ in range(2):
    eng.compute(l, r)
quiet = simsense.DepthSensorEngine(*dataclasses.replace(prm, speckle_shape=0.0).engine_args(), lanes=1)
for _ in range(2):
    quiet.compute(l, r)
torch.cuda.synchronize()
print("done")
