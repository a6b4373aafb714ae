#!/usr/bin/env python
"""A/B of a debug switch of the library: python tools/hook_ab.py ssb_debug_set_wta_pairs 0,1,0,1 C1 C4 ...
Stage times with one lane, frame rate with the engine's default lanes, final map compared bit by bit."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    from oracle import configs
    from sapien_b200 import simsense, synth

    lib = ctypes.CDLL(os.path.join(ROOT, "sapien_b200", "csrc", "libss_b200.so"))
    hook = getattr(lib, sys.argv[1])
    values = [int(v) for v in sys.argv[2].split(",")]
    which = sys.argv[3:] or ["C1", "C4", "C3", "C5"]
    batches = {"C1": 1, "C2": 1, "C4": 256, "C3": 64, "C5": 2}
    for key in which:
        batch = batches[key]
        prm = configs.params(key)
        pairs = [synth.make_pair(prm.rows, prm.cols, prm.max_disp, s)[:2] for s in range(2)]
        l = np.stack([pairs[i % 2][0] for i in range(batch)])
        r = np.stack([pairs[i % 2][1] for i in range(batch)])
        if batch == 1:
            l, r = l[0], r[0]
        tl, tr = torch.from_numpy(synth.to_rgba(l)).cuda(), torch.from_numpy(synth.to_rgba(r)).cuda()
        bb = (True, *configs.BBOX_C2) if key == "C2" else ()
        ref = None
        for on in values:
            hook(on)
            eng = simsense.DepthSensorEngine(*prm.engine_args(), batch=batch)
            es = torch.cuda.ExternalStream(eng.cuda_stream)
            n = 40 if batch == 1 else 8
            for _ in range(5):
                eng.compute(tl, tr, *bb, stream=eng.cuda_stream, sync=False)
            eng.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(es)
            for _ in range(n):
                eng.compute(tl, tr, *bb, stream=eng.cuda_stream, sync=False)
            e1.record(es)
            es.synchronize()
            ms = e0.elapsed_time(e1) / n
            eng.set_profiling(True)
            eng.get_stage_times()
            for _ in range(n):
                eng.compute(tl, tr, *bb, stream=eng.cuda_stream, sync=False)
            eng.synchronize()
            st = dict(eng.get_stage_times())
            st.pop("frames", None)
            out = eng.get_cuda().torch().clone()
            if ref is None:
                ref = out
            print(json.dumps({"workload": key, "batch": batch, "switch": on, "ms_per_step": round(ms, 4), "per_s": round(batch / ms * 1e3, 1),
                              "stages_ms": {k: round(v, 4) for k, v in st.items()},
                              "same_bits": bool(torch.equal(out.view(torch.int32), ref.view(torch.int32)))}), flush=True)
            del eng


if __name__ == "__main__":
    main()
