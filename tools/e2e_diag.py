#!/usr/bin/env python
"""What bounds the end-to-end (host buffers) frame rate when N ranks share one host?  Run under torchrun:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/e2e_diag.py

Per rank, C1 sizes (2 x 0.92 MB up, 8.3 MB down per frame), every phase first on rank 0 ALONE and then on ALL ranks at once:
  copies   -- the frame's transfers only (pinned buffers, H2D on one stream, D2H on another), no compute
  device   -- the engine with device-resident inputs (no transfers)
  piped    -- submit()/wait() with four frames in flight (bench.py's e2e.value)
  sync     -- compute(pinned) + get_ndarray(out=bound pinned), one frame at a time
and the host time spent inside submit() / wait().  Prints one table on rank 0."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import torch.distributed as dist

    import bench
    from oracle import configs
    from sapien_b200 import simsense

    rank, world, local = bench.dist_setup(0)
    prm = configs.params("C1")
    pairs = [configs.pair(prm, s) for s in range(4)]
    pin = [(torch.from_numpy(l).pin_memory(), torch.from_numpy(r).pin_memory()) for l, r in pairs]
    dev = [(torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()) for l, r in pairs]
    outs = [torch.empty((prm.rgb_rows, prm.rgb_cols), dtype=torch.float32).pin_memory() for _ in range(4)]
    dmap = torch.zeros((prm.rgb_rows, prm.rgb_cols), dtype=torch.float32, device="cuda")
    dl, dr = torch.empty_like(dev[0][0]), torch.empty_like(dev[0][1])
    eng = simsense.DepthSensorEngine(*prm.engine_args(), device=local)
    s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
    n = 300

    def copies():
        for i in range(n):
            with torch.cuda.stream(s_up):
                dl.copy_(pin[i % 4][0], non_blocking=True)
                dr.copy_(pin[i % 4][1], non_blocking=True)
            with torch.cuda.stream(s_down):
                outs[i % 4].copy_(dmap, non_blocking=True)
        torch.cuda.synchronize()

    def device():
        for i in range(n):
            eng.compute(dev[i % 4][0], dev[i % 4][1], stream=eng.cuda_stream, sync=False)
        eng.synchronize()

    host = {"submit_us": 0.0, "wait_us": 0.0}

    def piped():
        tk = [None] * 4
        ts = tw = 0.0
        for i in range(n):
            if tk[i % 4] is not None:
                t0 = time.perf_counter()
                eng.wait(tk[i % 4])
                tw += time.perf_counter() - t0
            t0 = time.perf_counter()
            tk[i % 4] = eng.submit(pin[i % 4][0].numpy(), pin[i % 4][1].numpy(), out=outs[i % 4].numpy())
            ts += time.perf_counter() - t0
        for t in tk:
            eng.wait(t)
        host["submit_us"], host["wait_us"] = ts / n * 1e6, tw / n * 1e6

    def sync():
        out = outs[0].numpy()
        eng.bind_output(out)
        for i in range(n):
            eng.compute(pin[i % 4][0].numpy(), pin[i % 4][1].numpy())
            eng.get_ndarray(out=out)
        eng.bind_output(None)

    rows = []
    for name, fn in (("copies", copies), ("device", device), ("piped", piped), ("sync", sync)):
        res = {}
        for mode in ("alone", "all"):
            fn() if (mode == "all" or rank == 0) else None  # warm-up
            bench.barrier(world)
            rate = 0.0
            if mode == "all" or rank == 0:
                t0 = time.perf_counter()
                fn()
                rate = n / (time.perf_counter() - t0)
            bench.barrier(world)
            allr = [None] * world
            dist.all_gather_object(allr, rate) if world > 1 else allr.__setitem__(0, rate)
            res[mode] = allr
        rows.append((name, res, dict(host)))
    if rank == 0:
        mb = (2 * prm.rows * prm.cols + prm.rgb_rows * prm.rgb_cols * 4) / 1e6
        print(f"# e2e diagnosis, {world} ranks on one host, C1 ({mb:.1f} MB of PCIe traffic per frame), frames/s per rank\n")
        print("| phase | rank 0 alone | all ranks: min / mean / max per rank | aggregate | aggregate PCIe GB/s |")
        print("|---|---:|---|---:|---:|")
        for name, res, h in rows:
            a = res["alone"][0]
            al = [x for x in res["all"] if x]
            agg = sum(al)
            pcie = "-" if name == "device" else f"{agg * mb / 1e3:.1f}"
            print(f"| {name} | {a:.0f} | {min(al):.0f} / {np.mean(al):.0f} / {max(al):.0f} | {agg:.0f} | {pcie} |")
        print(f"\nhost time per frame in the pipelined loop (rank 0, all ranks running): submit() {rows[2][2]['submit_us']:.0f} us, wait() {rows[2][2]['wait_us']:.0f} us")
        print(json.dumps({"cpus": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
