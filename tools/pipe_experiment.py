#!/usr/bin/env python
"""How many engines (sub-block pipelines) per GPU?  Splits a batched workload into K contiguous sub-blocks with
one engine each (ShardedStereoDepth(pipelines=K)) and times back-to-back passes on the device.

  python tools/pipe_experiment.py C4 1024 1,2,3,4 [steps]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench
    from oracle import configs
    from sapien_b200 import sharding

    key = sys.argv[1]
    n = int(sys.argv[2])
    ks = [int(x) for x in sys.argv[3].split(",")]
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 6
    prm = configs.params(key)
    if key == "C4":
        sets = bench.c4_inputs(prm, 0, n, 2, torch)
    else:
        import numpy as np

        from sapien_b200 import synth

        base = [synth.make_pair(prm.rows, prm.cols, prm.max_disp, s)[:2] for s in range(min(n, 8))]
        l = torch.from_numpy(synth.to_rgba(np.stack([base[i % len(base)][0] for i in range(n)]))).cuda()
        r = torch.from_numpy(synth.to_rgba(np.stack([base[i % len(base)][1] for i in range(n)]))).cuda()
        sets = [(l, r), (r, l)]
    alg = configs.algorithmic_bytes(prm, rgba_input=True)
    for k in ks:
        sh = sharding.ShardedStereoDepth(prm.engine_args(), n, 0, 1, device=0, pipelines=k)
        ms = bench.time_sharded(sh, sets, steps, 2, torch, 1)
        rate = n * steps / (ms / 1e3)
        print(json.dumps({"workload": key, "envs": n, "pipelines": k, "ms_per_pass": ms / steps, "env_frames_per_s": rate,
                          "frac_hbm": rate * alg / 1e9 / bench.hbm_peak()[0]}), flush=True)
        del sh
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
