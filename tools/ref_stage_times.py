#!/usr/bin/env python
"""Per-stage breakdown of the REFERENCE simsense on this GPU, from its own PRINT_RUNTIME timing
(include/simsense/config.h:26; core.cu:549-780), next to ours (ss_set_profiling).  SURVEY 8(d) "Baseline 1".

  python tools/ref_stage_times.py [C1|C3|C4|C5] [frames]  ->  markdown table on stdout

oracle/_ref/libsimsense_ref_timed.so = the unmodified reference sources compiled with -DPRINT_RUNTIME
(oracle/Makefile).  The reference prints with printf: fd 1 is redirected into a file around the run.
Note that with PRINT_RUNTIME every stage is followed by a device-wide sync, so the sum of its stages is
larger than its untimed frame; both are printed."""
import os
import re
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from oracle import REF_TIMED_SO, RefEngine, configs
    from sapien_b200 import synth

    key = sys.argv[1] if len(sys.argv) > 1 else "C1"
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    prm = configs.params(key)
    pairs = [configs.pair(prm, s) for s in range(4)]
    dev = [(torch.from_numpy(synth.to_rgba(l)).cuda(), torch.from_numpy(synth.to_rgba(r)).cuda()) for l, r in pairs]
    ref = RefEngine(prm, so=REF_TIMED_SO)
    for i in range(3):
        ref.compute_device(dev[i % 4][0].data_ptr(), dev[i % 4][1].data_ptr(), None)
    torch.cuda.synchronize()
    sys.stdout.flush()
    import ctypes

    libc = ctypes.CDLL(None)
    devnull = os.open(os.devnull, os.O_WRONLY)  # drop the warm-up frames' printf output
    keep = os.dup(1)
    os.dup2(devnull, 1)
    libc.fflush(None)
    os.dup2(keep, 1)
    os.close(keep)
    os.close(devnull)
    tmp = tempfile.NamedTemporaryFile("w+", delete=False)
    saved = os.dup(1)
    os.dup2(tmp.fileno(), 1)
    t0 = time.perf_counter()
    try:
        for i in range(frames):
            ref.compute_device(dev[i % 4][0].data_ptr(), dev[i % 4][1].data_ptr(), None)
        torch.cuda.synchronize()
        libc.fflush(None)
    finally:
        wall = time.perf_counter() - t0
        os.dup2(saved, 1)
        os.close(saved)
    tmp.seek(0)
    tot = {}
    order = []
    for line in tmp:
        m = re.match(r"Runtime of (.*): ([0-9.eE+-]+) ms", line.strip())
        if m:
            if m.group(1) not in tot:
                order.append(m.group(1))
            tot[m.group(1)] = tot.get(m.group(1), 0.0) + float(m.group(2))
    os.unlink(tmp.name)
    ref.close()
    # ours, same inputs
    from sapien_b200 import simsense

    eng = simsense.DepthSensorEngine(*prm.engine_args())
    for i in range(3):
        eng.compute(*dev[i % 4], stream=eng.cuda_stream)
    eng.set_profiling(True)
    eng.get_stage_times()
    for i in range(frames):
        eng.compute(*dev[i % 4], stream=eng.cuda_stream, sync=False)
    eng.synchronize()
    ours = dict(eng.get_stage_times())
    ours.pop("frames", None)
    print(f"## {key}: reference simsense per-stage device time (its own PRINT_RUNTIME build), {frames} frames, {torch.cuda.get_device_name(0)}\n")
    print("| reference stage (core.cu printf label) | ms / frame |")
    print("|---|---:|")
    for k in order:
        print(f"| {k} | {tot[k] / frames:.4f} |")
    print(f"| **sum of stages** | **{sum(tot.values()) / frames:.4f}** |")
    print(f"| wall per frame of this (per-stage synchronised) build | {wall / frames * 1e3:.4f} |\n")
    print("| this engine (ss_set_profiling) | ms / frame |")
    print("|---|---:|")
    for k, v in ours.items():
        print(f"| {k} | {v:.4f} |")
    print(f"| **sum of stages** | **{sum(ours.values()):.4f}** |")


if __name__ == "__main__":
    main()
