#!/usr/bin/env python
"""A/B of a debug switch on the C5 sweep exactly as bench.py runs it (4-frame batches, N sub-batch pipelines):
python tools/c5_ab.py ssb_debug_set_cost_blocks 3,2,3,2 [frames] [pipelines]"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from oracle import configs
from sapien_b200 import sharding

lib = ctypes.CDLL(os.path.join(ROOT, "sapien_b200", "csrc", "libss_b200.so"))
hook = getattr(lib, sys.argv[1])
values = [int(v) for v in sys.argv[2].split(",")]
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 512
pipes = int(sys.argv[4]) if len(sys.argv) > 4 else 2
prm = configs.params("C5")
bl, br = bench.c5_inputs(prm, torch)
sets5 = [(bl[0:4], br[0:4]), (bl[4:8], br[4:8])]
for v in values:
    hook(v)
    sh5 = sharding.ShardedStereoDepth(prm.engine_args(), 4, 0, 1, device=0, pipelines=pipes)
    ms = bench.time_sharded(sh5, sets5, frames // 4, 2, torch, 1)
    print(json.dumps({"switch": v, "pipelines": len(sh5.engines), "frames": frames, "frames_per_s": round(frames / ms * 1e3, 1)}), flush=True)
    del sh5
    torch.cuda.empty_cache()
