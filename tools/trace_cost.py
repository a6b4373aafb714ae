#!/usr/bin/env python
"""Per-block timeline of the cost kernel (debug aid; needs a B200): start / end of warp 0 of every block,
by SM, by column block and by row band."""
import ctypes, os, sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import configs
from sapien_b200 import simsense, synth, _build

STRIDE = 8192
lib = ctypes.CDLL(_build.LIB)
key = sys.argv[1] if len(sys.argv) > 1 else "C1"
prm = configs.params(key)
l, r = configs.pair(prm, 0)
tl = torch.from_numpy(synth.to_rgba(l)).cuda(); tr_ = torch.from_numpy(synth.to_rgba(r)).cuda()
for variant in [0]:
    eng = simsense.DepthSensorEngine(*prm.engine_args(), device=0)
    eng.set_profiling(True)  # one lane, stages back to back
    for _ in range(3):
        eng.compute(tl, tr_)
    cbuf = torch.zeros(4 * STRIDE, dtype=torch.int64, device="cuda")
    assert lib.ssb_debug_set_cost_trace(ctypes.c_void_p(cbuf.data_ptr())) == 0
    eng.compute(tl, tr_)
    torch.cuda.synchronize()
    lib.ssb_debug_set_cost_trace(ctypes.c_void_p(0))
    c = cbuf.cpu().numpy().reshape(STRIDE, 4)
    m = c[:, 0] > 0
    idx = np.nonzero(m)[0]
    c0 = int(c[m, 0].min())
    st, en, sm = (c[m, 0] - c0) / 1e3, (c[m, 1] - c0) / 1e3, c[m, 2].astype(int)
    dur = en - st
    print(f"variant {variant}: {m.sum()} blocks; start {st.min():.1f}..{st.max():.1f} us, end {en.min():.1f}..{en.max():.1f} us; "
          f"duration min/med/max {dur.min():.1f}/{np.median(dur):.1f}/{dur.max():.1f} us")
    cnt = np.bincount(sm, minlength=148)
    for cc in sorted(set(cnt)):
        sel = np.isin(sm, np.where(cnt == cc)[0])
        if sel.any():
            print(f"    SMs hosting {cc} blocks: {int((cnt == cc).sum())} SMs; duration med {np.median(dur[sel]):.1f} us, ends {en[sel].min():.1f}..{en[sel].max():.1f} us")
    nxb = (prm.cols + 63) // 64 * max(1, prm.max_disp // 128)
    xb, band = idx % nxb, idx // nxb
    print("    end by column block:", " ".join(f"{b}:{en[xb == b].mean():.0f}" for b in sorted(set(xb))))
    print("    end by row band:    ", " ".join(f"{b}:{en[band == b].mean():.0f}" for b in sorted(set(band))))
    h, e = np.histogram(en, bins=10)
    print("    end histogram:", " ".join(f"{e[i]:.0f}-{e[i+1]:.0f}us:{h[i]}" for i in range(len(h))))
    h, e = np.histogram(st, bins=6)
    print("    start histogram:", " ".join(f"{e[i]:.0f}-{e[i+1]:.0f}us:{h[i]}" for i in range(len(h))))
    del eng
