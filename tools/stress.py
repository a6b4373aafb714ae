#!/usr/bin/env python
"""Long mixed-mode run on one engine (needs a B200): device async bursts, host frames, banded output, ROI frames in
random order; every observable result must equal the result of the same inputs computed alone.  Catches ordering
bugs between the engine's streams that a short test can miss."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import configs
from sapien_b200 import simsense

cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 400
prm = configs.params(cfg)
n = 8
pairs = [configs.pair(prm, seed=200 + i) for i in range(n)]
solo = simsense.DepthSensorEngine(*prm.engine_args(), device=0)
want = []
for l, r in pairs:
    solo.compute(l, r); want.append(solo.get_ndarray())
bbox = (16, 8, prm.cols // 2, prm.rows // 2)
want_bbox = []
for l, r in pairs:
    solo.compute(l, r, True, *bbox); want_bbox.append(solo.get_ndarray())
eng = simsense.DepthSensorEngine(*prm.engine_args(), device=0)
es = torch.cuda.ExternalStream(eng.cuda_stream)
dl = [torch.from_numpy(l).cuda() for l, _ in pairs]
dr = [torch.from_numpy(r).cuda() for _, r in pairs]
out = torch.empty((prm.rgb_rows, prm.rgb_cols), dtype=torch.float32).pin_memory().numpy()
snaps = [torch.empty((prm.rgb_rows, prm.rgb_cols), dtype=torch.float32, device="cuda") for _ in range(n)]
pouts = [torch.empty((prm.rgb_rows, prm.rgb_cols), dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
ppairs = [(torch.from_numpy(l).pin_memory().numpy(), torch.from_numpy(r).pin_memory().numpy()) for l, r in pairs]
rng = np.random.default_rng(7)
bound = False
t0 = time.time()
bad = 0
for it in range(iters):
    mode = rng.choice(["burst", "host", "bind", "unbind", "roi", "dev_sync", "pipe"])
    i = int(rng.integers(0, n))
    if mode == "burst":
        k = int(rng.integers(2, n + 1))
        for j in range(k):
            eng.compute(dl[j], dr[j], sync=False)
            with torch.cuda.stream(es):
                snaps[j].copy_(eng.get_cuda().torch(), non_blocking=True)
        es.synchronize()
        for j in range(k):
            if not np.array_equal(snaps[j].cpu().numpy().view(np.uint32), want[j].view(np.uint32)):
                bad += 1; print("MISMATCH burst", it, j)
    elif mode == "pipe":  # asynchronous host frames, two in flight
        k = int(rng.integers(2, n + 1))
        tk = []
        for j in range(k):
            if j >= 2:
                eng.wait(tk[j - 2])
                if not np.array_equal(pouts[j % 2].view(np.uint32), want[j - 2].view(np.uint32)):
                    bad += 1; print("MISMATCH pipe", it, j - 2)
            tk.append(eng.submit(ppairs[j][0], ppairs[j][1], out=pouts[j % 2]))
        for j in range(max(0, k - 2), k):
            eng.wait(tk[j])
            if not np.array_equal(pouts[j % 2].view(np.uint32), want[j].view(np.uint32)):
                bad += 1; print("MISMATCH pipe tail", it, j)
    elif mode == "host":
        eng.compute(*pairs[i])
        got = eng.get_ndarray(out=out) if bound else eng.get_ndarray()
        if not np.array_equal(got.view(np.uint32), want[i].view(np.uint32)):
            bad += 1; print("MISMATCH host", it, i, bound)
    elif mode == "bind":
        eng.bind_output(out); bound = True
    elif mode == "unbind":
        eng.bind_output(None); bound = False
    elif mode == "roi":
        eng.compute(*pairs[i], True, *bbox)
        if not np.array_equal(eng.get_ndarray().view(np.uint32), want_bbox[i].view(np.uint32)):
            bad += 1; print("MISMATCH roi", it, i)
    else:
        eng.compute(dl[i], dr[i])
        if not np.array_equal(eng.get_cuda().torch().cpu().numpy().view(np.uint32), want[i].view(np.uint32)):
            bad += 1; print("MISMATCH dev", it, i)
print(f"{cfg}: {iters} mixed operations in {time.time() - t0:.1f} s, {bad} mismatches")
sys.exit(1 if bad else 0)
