#!/usr/bin/env python
"""Per-path timeline of the four aggregation kernels (debug aid; needs a B200).
Registers a device buffer through ssb_debug_set_aggr_trace, runs C1 frames, and prints, per kernel,
when paths start, how long they take, and how that depends on the SM they landed on."""
import ctypes, os, sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import configs
from sapien_b200 import simsense, synth, _build

STRIDE = 8192
lib = ctypes.CDLL(_build.LIB)
prm = configs.params(sys.argv[1] if len(sys.argv) > 1 else "C1")
eng = simsense.DepthSensorEngine(*prm.engine_args(), device=0)
l, r = configs.pair(prm, 0)
tl = torch.from_numpy(synth.to_rgba(l)).cuda(); tr_ = torch.from_numpy(synth.to_rgba(r)).cuda()
for _ in range(3):
    eng.compute(tl, tr_)
cbuf = torch.zeros(4 * STRIDE, dtype=torch.int64, device="cuda")
assert lib.ssb_debug_set_cost_trace(ctypes.c_void_p(cbuf.data_ptr())) == 0
buf = torch.zeros(4 * 4 * STRIDE, dtype=torch.int64, device="cuda")
assert lib.ssb_debug_set_aggr_trace(ctypes.c_void_p(buf.data_ptr())) == 0
eng.compute(tl, tr_)
torch.cuda.synchronize()
lib.ssb_debug_set_aggr_trace(ctypes.c_void_p(0))
lib.ssb_debug_set_cost_trace(ctypes.c_void_p(0))
c = cbuf.cpu().numpy().reshape(STRIDE, 4)
m = c[:, 0] > 0
if m.any():
    c0 = int(c[m, 0].min())
    st, en, sm = c[m, 0] - c0, c[m, 1] - c0, c[m, 2]
    dur = en - st
    print(f"cost: {m.sum()} blocks; start {st.min()/1e3:.1f}..{st.max()/1e3:.1f} us, end {en.min()/1e3:.1f}..{en.max()/1e3:.1f} us; "
          f"duration min/med/max {dur.min()/1e3:.1f}/{np.median(dur)/1e3:.1f}/{dur.max()/1e3:.1f} us")
    cnt = np.bincount(sm.astype(int), minlength=148)
    for cc in sorted(set(cnt)):
        sel = np.isin(sm, np.where(cnt == cc)[0])
        print(f"    SMs hosting {cc} blocks: {int((cnt == cc).sum())} SMs; block duration med {np.median(dur[sel])/1e3:.1f} us, last end {en[sel].max()/1e3:.1f} us")
    h, e = np.histogram(st / 1e3, bins=8)
    print("    start histogram:", " ".join(f"{e[i]:.0f}-{e[i+1]:.0f}us:{h[i]}" for i in range(len(h))))
    h, e = np.histogram(dur / 1e3, bins=8)
    print("    duration histogram:", " ".join(f"{e[i]:.0f}-{e[i+1]:.0f}us:{h[i]}" for i in range(len(h))))
t = buf.cpu().numpy().reshape(4, STRIDE, 4)
names = ["left (R->L)", "down (T->B)", "up (B->T)", "right+wta producer"]
t0 = min(int(t[k][t[k][:, 0] > 0][:, 0].min()) for k in range(4) if (t[k][:, 0] > 0).any())
for k in range(4):
    m = t[k][:, 0] > 0
    if not m.any():
        continue
    st, en, sm = t[k][m, 0] - t0, t[k][m, 1] - t0, t[k][m, 2]
    dur = en - st
    print(f"{names[k]}: {m.sum()} paths; start {st.min()/1e3:.1f}..{st.max()/1e3:.1f} us, end {en.min()/1e3:.1f}..{en.max()/1e3:.1f} us; "
          f"duration min/med/max {dur.min()/1e3:.1f}/{np.median(dur)/1e3:.1f}/{dur.max()/1e3:.1f} us")
    cnt = np.bincount(sm.astype(int), minlength=148)
    for c in sorted(set(cnt)):
        sel = np.isin(sm, np.where(cnt == c)[0])
        if sel.any():
            print(f"    SMs hosting {c} paths: {int((cnt == c).sum())} SMs; path duration med {np.median(dur[sel])/1e3:.1f} us, last end {en[sel].max()/1e3:.1f} us")
    if k == 3:
        cen = t[k][m, 3] - t0
        print(f"    consumer ends {cen.min()/1e3:.1f}..{cen.max()/1e3:.1f} us (lag behind producer med {np.median(cen - en)/1e3:.2f} us)")
    # histogram of durations
    h, e = np.histogram(dur / 1e3, bins=8)
    print("    duration histogram:", " ".join(f"{e[i]:.0f}-{e[i+1]:.0f}us:{h[i]}" for i in range(len(h))))
