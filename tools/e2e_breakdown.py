#!/usr/bin/env python
"""Where the end-to-end time of compute(host u8) + get_ndarray(pinned) goes (needs a B200)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import configs
from sapien_b200 import simsense, synth

prm = configs.params("C1")
eng = simsense.DepthSensorEngine(*prm.engine_args(), device=0)
l, r = configs.pair(prm, 0)
pl, pr_ = torch.from_numpy(l).pin_memory(), torch.from_numpy(r).pin_memory()
out = torch.empty((prm.rgb_rows, prm.rgb_cols), dtype=torch.float32).pin_memory()
ln, rn, on = pl.numpy(), pr_.numpy(), out.numpy()
dl, dr = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
def both():
    eng.compute(ln, rn); eng.get_ndarray(out=on)
print(f"compute(host u8) + get_ndarray(pinned): {t(both):.1f} us")
print(f"compute(host u8) only:                  {t(lambda: eng.compute(ln, rn)):.1f} us")
eng.compute(ln, rn)
print(f"get_ndarray(pinned) only:               {t(lambda: eng.get_ndarray(out=on)):.1f} us")
print(f"compute(device u8, sync):               {t(lambda: eng.compute(dl, dr)):.1f} us")
h2d = torch.empty_like(dl)
print(f"torch H2D of one image (pinned):        {t(lambda: h2d.copy_(pl, non_blocking=True)):.1f} us")
dout = eng.get_cuda().torch()
print(f"torch D2H of the depth map (pinned):    {t(lambda: out.copy_(dout, non_blocking=True)):.1f} us")

# ---- banded output ----------------------------------------------------------------------------
eng.bind_output(on)
def banded():
    eng.compute(ln, rn); eng.get_ndarray(out=on)
print(f"bind_output: compute + get_ndarray:     {t(banded):.1f} us")
print(f"bind_output: compute(host u8) only:     {t(lambda: eng.compute(ln, rn)):.1f} us")
eng.bind_output(None)
import ctypes
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
W, Hh = prm.rgb_cols, prm.rgb_rows
for a, b in ((0, 236), (236, 716), (716, 1196), (1196, 1920), (0, 1920), (0, 480), (0, 960)):
    fn = lambda: rt.cudaMemcpy2DAsync(out.data_ptr() + 4 * a, 4 * W, dout.data_ptr() + 4 * a, 4 * W, 4 * (b - a), Hh, 2, None)
    us = t(fn, 30)
    print(f"2-D D2H of columns [{a},{b}): {us:.1f} us = {4*(b-a)*Hh/us/1e3:.1f} GB/s")
