#!/usr/bin/env python
"""bench.py -- stereo depth frames/s on the BASELINE.json workloads.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C1|C2|C3|C4|C5]
  python bench.py --impl reference ...        the unmodified reference simsense CUDA build
                                              (oracle/_ref) on the same GPU
  python bench.py --impl reference-cpu ...    the scalar C/OpenMP oracle on the host cores

One "step" = one pass of the hot path (DepthSensorEngine compute: front-end, cost volume, 4-path
SGM, WTA, LR, median, depth, registration) over one synthetic batch.  `value` is measured with the
inputs resident in HBM (device RGBA float32, the reference's CUDA input format), CUDA-event timed
on the stream the work is ordered on; `e2e` goes through the public Python API with HOST
buffers, uploads and the depth read-back inside the timed region (two figures: the extension path
with a bound pinned output, and the strict reference signature compute(l, r) + get_ndarray()).
N>1: one process per GPU (torchrun), the same per-GPU workload on every rank (environments/frames
are independent: no collective on the data path), barrier + max-over-ranks timing, `scaling` =
weak.  The same line carries a `batched` block measured in the same process: north_star's batched,
sharded workloads -- C4 (1 024 envs x 256x256, D=64) split env_range(1024, rank, world) through
ShardedStereoDepth (STRONG scaling: 1024/N envs per rank) with the optional NCCL all-gather timed
separately, and the C5 sweep (1920x1080, D=256, 4-frame batches, frames split contiguously).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "stereo depth frames/sec at 1280x720/128 disp per B200 and batched env-frames/sec 1-8 GPU"
WORKLOADS = {
    # name: (params key, batch, bbox, point cloud, description)
    "C1": ("C1", 1, None, "", "single D415-like IR pair 1280x720, 128 disp, remap + 7x7 census + 7x7 block + 4-path SGM + LR + median3 + registration/dilation to 1920x1080"),
    "C2": ("C2", 1, "bbox", "xyzrgb", "C1 with bbox ROI (100,100) 640x360 + RGB point cloud"),
    "C3": ("C3", 64, None, "", "batched 64 envs x 848x480, 96 disp (D435)"),
    "C4": ("C4", 1024, None, "", "1024 envs x 256x256, 64 disp, env-sharded"),
    "C5": ("C5", 4, None, "", "1920x1080, 256 disp, 4-frame batches"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cpu"])
    ap.add_argument("--workload", default="C1", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the env batch of the workload (per GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batched", action="store_true", help="skip the batched block (C4 1024-env shard + C5 sweep)")
    ap.add_argument("--c4-envs", type=int, default=1024)
    ap.add_argument("--c5-frames", type=int, default=4096, help="frames of the C5 sweep over all ranks (north_star: 4096)")
    ap.add_argument("--pipelines", type=int, default=0, help="engines per rank for the batched block (0 = default per workload)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.f:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            try:
                pw.append(float(c[3]))
            except ValueError:
                pw.append(0.0)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            # the sampled span also holds host-side set-up between the timed regions: "under load" = the samples whose
            # power draw is at least half way between the lowest and the highest one seen
            lo, hi = min(pw), max(pw)
            load = [s for s, p in zip(sm, pw) if p >= lo + 0.5 * (hi - lo)] or sm
            out.update(sm_mhz=float(np.median(load)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       samples_under_load=len(load), power_w_max=hi)
        return out


def pin_to_gpu_numa_node(torch, local):
    """Run this rank on the CPUs next to its GPU (NVML's ideal affinity), before any pinned host buffer is
    allocated: with 8 ranks the uploads and the 8.3 MB read-back of every frame otherwise cross the socket
    interconnect for half of the GPUs.  Both arms do this.  Best effort: silently skipped if NVML is unavailable."""
    try:
        import pynvml

        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode()))
    except Exception:
        pass


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line, on the real stdout."""
    text = json.dumps(line) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, text.encode())


def dist_setup(n_gpus):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    pin_to_gpu_numa_node(torch, local)
    if world > 1:
        import torch.distributed as dist

        # NCCL's INFO log (communicator size, transports) is wanted, but it goes to the C-level stdout and stdout must
        # carry ONE JSON line: fd 1 is pointed at stderr for the life of the process and the JSON line is written to
        # the saved original stdout (emit()).
        os.environ["NCCL_DEBUG"] = os.environ.get("SSB_NCCL_DEBUG", "INFO")  # (the image presets NCCL_DEBUG=VERSION)
        global _REAL_STDOUT
        if _REAL_STDOUT is None:
            sys.stdout.flush()
            _REAL_STDOUT = os.dup(1)
            os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier(world):
    import torch

    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()


def sum_over_ranks(v: float, world) -> float:
    if world == 1:
        return v
    import torch
    import torch.distributed as dist

    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def max_over_ranks(v: float, world) -> float:
    if world == 1:
        return v
    import torch
    import torch.distributed as dist

    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def make_inputs(prm, batch, n_sets, torch):
    """n_sets distinct batches of device RGBA float32 pairs (+ the u8 originals on the host)."""
    from sapien_b200 import synth

    sets = []
    base = [synth.make_pair(prm.rows, prm.cols, prm.max_disp, s)[:2] for s in range(min(8, max(n_sets, batch)))]
    for k in range(n_sets):
        l = np.stack([base[(k + i) % len(base)][0] for i in range(batch)])
        r = np.stack([base[(k + i) % len(base)][1] for i in range(batch)])
        if batch == 1:
            l, r = l[0], r[0]
        sets.append((l, r))
    dev = [(torch.from_numpy(synth.to_rgba(l)).cuda(), torch.from_numpy(synth.to_rgba(r)).cuda()) for l, r in sets]
    return sets, dev


def cpu_baseline(prm, bbox, seconds, torch_threads=None):
    """The scalar C/OpenMP oracle timed on the host cores over a bounded sample of the workload."""
    from oracle import Oracle, configs

    orc = Oracle()
    l, r = configs.pair(prm, 0)
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    orc.pipeline(prm, l, r, bbox=bbox, stages=False)
    first = time.perf_counter() - t0
    n = int(max(1, min(50, seconds / max(first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(n):
        orc.pipeline(prm, l, r, bbox=bbox, stages=False)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{n} frames of the workload, one frame at a time, OpenMP over rows/columns on {cores} threads (after 1 warm-up frame)"}


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def shared_config(args, desc, batch, world):
    """`config` is identical in both arms (what differs between the arms is in `arm`)."""
    return {"workload": f"{args.workload}: {desc}", "batch_per_gpu": batch, "input": "device float32 RGBA pairs (the reference's CUDA input format)",
            "l2": "8 distinct input sets rotate; per-step intermediate traffic (u16 cost volumes) exceeds the 126 MB L2",
            "parallelism": f"env-sharded x{world}, no collective"}


def c4_inputs(prm, start, stop, n_sets, torch):
    """Device RGBA batches [n_local,H,W,4] for envs [start, stop): env e uses seed e % 64 (shifted per set)."""
    from sapien_b200 import synth

    base = [synth.make_pair(prm.rows, prm.cols, prm.max_disp, s)[:2] for s in range(64)]
    bl = torch.from_numpy(synth.to_rgba(np.stack([b[0] for b in base]))).cuda()
    br = torch.from_numpy(synth.to_rgba(np.stack([b[1] for b in base]))).cuda()
    sets = []
    for k in range(n_sets):
        idx = torch.tensor([(e + 7 * k) % 64 for e in range(start, stop)], device="cuda", dtype=torch.long)
        sets.append((bl.index_select(0, idx).contiguous(), br.index_select(0, idx).contiguous()))
    return sets


def c5_inputs(prm, torch, n_base=8):
    from sapien_b200 import synth

    base = [synth.make_pair(prm.rows, prm.cols, prm.max_disp, s)[:2] for s in range(n_base)]
    l = torch.from_numpy(synth.to_rgba(np.stack([b[0] for b in base]))).cuda()
    r = torch.from_numpy(synth.to_rgba(np.stack([b[1] for b in base]))).cuda()
    return l, r


def time_sharded(sh, sets, steps, warmup, torch, world):
    """`steps` passes of the local block through every pipeline of `sh`, device-timed: a timing stream opens the
    region, every engine stream is ordered after it, and it closes behind all of them (max over ranks)."""
    if not sh.engines:
        barrier(world)
        barrier(world)
        return max_over_ranks(0.0, world)
    ts = torch.cuda.Stream()
    ext = [torch.cuda.ExternalStream(e.cuda_stream) for e in sh.engines]
    for i in range(warmup):
        sh.enqueue(*sets[i % len(sets)], inputs_ready=True)
    sh.synchronize()
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    for e in sh.engines:
        e.wait_stream(ts.cuda_stream)
    for i in range(steps):
        sh.enqueue(*sets[i % len(sets)], inputs_ready=True)
    for x in ext:
        ts.wait_stream(x)
    e1.record(ts)
    ts.synchronize()
    barrier(world)
    return max_over_ranks(e0.elapsed_time(e1), world)


def batched_ours(args, rank, world, local):
    """north_star's batched, sharded workloads in the same process (see the module docstring)."""
    import torch

    from oracle import configs
    from sapien_b200 import sharding

    peak, _ = hbm_peak()
    out = {}
    steps = max(3, min(args.steps, 10))
    # ---- C4: 1024 envs x 256x256, D=64, env_range(1024, rank, world) per rank: strong scaling ----
    prm = configs.params("C4")
    pipes = args.pipelines or 1  # (2-4 concurrent sub-batches measured no faster on C3 / C4: batched kernels fill the GPU)
    sh = sharding.ShardedStereoDepth(prm.engine_args(), args.c4_envs, rank, world, device=local, pipelines=pipes)
    sets = c4_inputs(prm, sh.start, sh.stop, 2, torch)
    ms = time_sharded(sh, sets, steps, 2, torch, world)
    alg = configs.algorithmic_bytes(prm, rgba_input=True)
    rate = args.c4_envs * steps / (ms / 1e3)
    launches = sum(e.get_launches_per_compute() for e in sh.engines)
    c4 = {"workload": f"C4: {args.c4_envs} envs x 256x256, 64 disp, split env_range({args.c4_envs}, rank, {world}) through ShardedStereoDepth",
          "env_frames_per_s": rate, "envs_per_gpu": sh.local, "n_gpus": world, "scaling": "strong", "steps": steps, "ms_per_step": ms / steps,
          "pipelines_per_gpu": len(sh.engines), "input": "device float32 RGBA [n,256,256,4] batches, 2 sets alternate (each 268 MB per image at 1024 envs)",
          "algorithmic_bytes_per_env_frame": int(alg), "hbm_roofline_env_frames_per_s_per_gpu": peak * 1e9 / alg,
          "frac_of_hbm_roofline": rate / world * alg / 1e9 / peak, "gpu_launches_per_step": launches}
    # optional epilogue, outside the metric: NCCL all-gather of the depth maps for one host-side consumer
    depth = sh.depth()
    like = ((prm.rgb_rows, prm.rgb_cols), torch.float32, torch.device("cuda", local))
    if world > 1:
        for _ in range(3):  # warm-up: NCCL channels, and both result blocks of the caching allocator (a cudaMalloc of 268 MB inside
            full = sh.gather_depth(depth, like=like)  # the timed region measured 14 ms per gather instead of 0.33 ms on two GPUs)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(world)
        g0.record()
        for _ in range(5):
            full = sh.gather_depth(depth, like=like)
        g1.record()
        torch.cuda.synchronize()
        gms = max_over_ranks(g0.elapsed_time(g1) / 5, world)
        nbytes = sh.local * prm.rgb_rows * prm.rgb_cols * 4
        c4["gather"] = {"op": "all_gather_into_tensor (NCCL over NVLink), outside the metric", "bytes_per_rank": int(nbytes), "ms": gms,
                        "algbw_gbs": nbytes * world / (gms * 1e-3) / 1e9, "result_shape": list(full.shape)}
        del full
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        cb = cpu_baseline(prm, None, 2.0)
        c4["cpu_oracle"] = {"env_frames_per_s": cb["value"], "cores": cb["cores"], "sample": cb["sample"]}
    out["C4"] = c4
    del sh, sets, depth
    torch.cuda.empty_cache()
    # ---- C5: 1920x1080, D=256, sweep of F frames in 4-frame batches, frames split contiguously over the ranks ----
    prm = configs.params("C5")
    f0, f1 = sharding.env_range(args.c5_frames, rank, world)
    nb = (f1 - f0) // 4  # 4-frame batches of this rank
    pipes5 = args.pipelines or 2
    sh5 = sharding.ShardedStereoDepth(prm.engine_args(), 4, 0, 1, device=local, pipelines=pipes5)
    bl, br = c5_inputs(prm, torch)
    sets5 = [(bl[0:4], br[0:4]), (bl[4:8], br[4:8])]
    frames = int(sum_over_ranks(4.0 * nb, world))  # (a remainder of < 4 frames per rank is left out)
    ms5 = time_sharded(sh5, sets5, max(nb, 1), 1, torch, world) if frames else 0.0
    alg5 = configs.algorithmic_bytes(prm, rgba_input=True)
    rate5 = frames / (ms5 / 1e3) if ms5 else 0.0
    out["C5"] = {"workload": f"C5: 1920x1080, 256 disp, sweep of {args.c5_frames} frames (--c5-frames), 4-frame batches, frames split contiguously over {world} ranks",
                 "frames_per_s": rate5, "frames": frames, "frames_per_gpu": 4 * nb, "n_gpus": world, "scaling": "strong", "ms_total": ms5,
                 "pipelines_per_gpu": len(sh5.engines), "input": "device float32 RGBA [4,1080,1920,4] batches, 8 base pairs cycle",
                 "algorithmic_bytes_per_frame": int(alg5), "hbm_roofline_frames_per_s_per_gpu": peak * 1e9 / alg5,
                 "frac_of_hbm_roofline": rate5 / world * alg5 / 1e9 / peak}
    return out


def batched_reference(args, rank, world, local):
    """The same split on the reference: one engine per rank, a sequential loop over the rank's envs / frames (the reference
    has no batch API).  Bounded samples: every local C4 env once; up to 16 local C5 frames."""
    import torch

    from oracle import RefEngine, configs
    from sapien_b200 import sharding

    out = {}
    prm = configs.params("C4")
    a, b = sharding.env_range(args.c4_envs, rank, world)
    (l, r), = c4_inputs(prm, a, b, 1, torch)
    ref = RefEngine(prm)
    stride = prm.rows * prm.cols * 16
    for i in range(min(8, b - a)):
        ref.compute_device(l.data_ptr() + i * stride, r.data_ptr() + i * stride, None)
    torch.cuda.synchronize()
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(b - a):
        ref.compute_device(l.data_ptr() + i * stride, r.data_ptr() + i * stride, None)
    e1.record()
    torch.cuda.synchronize()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    out["C4"] = {"workload": f"C4: {args.c4_envs} envs x 256x256, 64 disp, split env_range({args.c4_envs}, rank, {world}); sequential loop per rank",
                 "env_frames_per_s": args.c4_envs / (ms / 1e3), "envs_per_gpu": b - a, "n_gpus": world, "scaling": "strong", "steps": 1, "ms_per_step": ms}
    ref.close()
    del l, r
    torch.cuda.empty_cache()
    prm = configs.params("C5")
    f0, f1 = sharding.env_range(args.c5_frames, rank, world)
    n = min(16, f1 - f0)
    bl, br = c5_inputs(prm, torch)
    ref = RefEngine(prm)
    stride = prm.rows * prm.cols * 16
    for i in range(2):
        ref.compute_device(bl.data_ptr() + i * stride, br.data_ptr() + i * stride, None)
    torch.cuda.synchronize()
    barrier(world)
    e0.record()
    for i in range(n):
        ref.compute_device(bl.data_ptr() + (i % 8) * stride, br.data_ptr() + (i % 8) * stride, None)
    e1.record()
    torch.cuda.synchronize()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    out["C5"] = {"workload": f"C5: 1920x1080, 256 disp, frames split contiguously over {world} ranks; sequential loop per rank",
                 "frames_per_s": n * world / (ms / 1e3) if n else 0.0, "sample": f"{n} frames per rank (of {f1 - f0})", "n_gpus": world, "scaling": "strong"}
    ref.close()
    return out


def run_ours(args, rank, world, local):
    import torch

    from oracle import configs
    from sapien_b200 import simsense

    key, batch, bbox, pc, desc = WORKLOADS[args.workload]
    if args.batch:
        batch = args.batch
    prm = configs.params(key)
    bbox_t = configs.BBOX_C2 if bbox else None
    eng = simsense.DepthSensorEngine(*prm.engine_args(), device=local, batch=batch)
    n_sets = 8 if batch == 1 else 2
    host_sets, dev_sets = make_inputs(prm, batch, n_sets, torch)
    rgba = None
    if pc:
        from sapien_b200 import synth

        rgba = torch.from_numpy(synth.make_rgb(prm.rgb_rows, prm.rgb_cols, 0)).cuda()
    # Frames are enqueued back to back on the ENGINE's stream (inputs are resident and ready: no caller stream
    # to order against), and the CUDA events that time them are recorded on that same stream.
    stream = torch.cuda.ExternalStream(eng.cuda_stream, device=torch.device("cuda", local))
    bb = (True, *bbox_t) if bbox_t else (False, 0, 0, 0, 0)

    own = eng.cuda_stream  # stream=own: the inputs are resident and complete, no ordering against a caller stream

    def step(i):
        l, r = dev_sets[i % n_sets]
        eng.compute(l, r, *bb, stream=own, sync=False)
        if pc:
            eng.get_rgb_point_cloud_cuda(rgba, sync=False)  # stream-ordered: the kernel follows the frame on its lane

    torch.cuda.synchronize()
    sampler = ClockSampler(local)  # 100 ms samples from the warm-up to the end of the batched block (each timed region is short)
    sampler.start()
    for i in range(args.warmup):
        step(i)
    stream.synchronize()
    barrier(world)
    eng.set_profiling(True)
    eng.get_stage_times()
    for i in range(min(args.steps, 20)):  # per-kernel stage times: a separate, profiled run (event marks between the stages)
        step(i)
    stream.synchronize()
    stages = dict(eng.get_stage_times())
    eng.set_profiling(False)
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    th0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    host_us = (time.perf_counter() - th0) / args.steps * 1e6  # CPU time to enqueue one frame (the loop never waits)
    e1.record(stream)
    stream.synchronize()
    barrier(world)
    ms = e0.elapsed_time(e1)
    launches = eng.get_launches_per_compute() + (1 if pc else 0)
    eng_lanes = eng.lanes
    ms = max_over_ranks(ms, world)
    frames = batch * args.steps * world
    value = frames / (ms / 1e3)

    # ---- e2e: public API, HOST buffers, H2D + D2H inside the timed region -------------------------
    # (a) extension path: pinned inputs, depth map streamed into a bound pinned buffer;
    # (b) strict reference signature: plain (pageable) numpy inputs, compute(l, r) + get_ndarray() returning a new array.
    pin_l = [torch.from_numpy(l).pin_memory() for l, _ in host_sets]
    pin_r = [torch.from_numpy(r).pin_memory() for _, r in host_sets]
    out_shape = ((batch,) if batch > 1 else ()) + (prm.rgb_rows, prm.rgb_cols)
    pin_out = torch.empty(out_shape, dtype=torch.float32).pin_memory()
    out_np = pin_out.numpy()
    e2e_steps = max(60, min(args.steps, 200))  # (its own step count: 20 frames would mostly time the fill and drain of the pipeline)

    def e2e_loop(fn):
        for i in range(3):
            fn(i)
        barrier(world)
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            fn(i)
        torch.cuda.synchronize()
        return max_over_ranks(time.perf_counter() - t0, world)

    def ext_step(i):
        eng.compute(pin_l[i % n_sets].numpy(), pin_r[i % n_sets].numpy(), *bb)
        eng.get_ndarray(out=out_np)

    keep = [None]

    def strict_step(i):
        eng.compute(host_sets[i % n_sets][0], host_sets[i % n_sets][1], *bb)
        keep[0] = eng.get_ndarray()

    eng.bind_output(out_np)  # the depth map streams into the pinned buffer behind the last aggregation pass
    ext_s = e2e_loop(ext_step)
    eng.bind_output(None)
    strict_s = e2e_loop(strict_step)

    # (c) pipelined extension path: submit()/wait(), two frames in flight per lane.  Every step still uploads its own inputs from
    # pinned host memory and has its own depth map delivered into pinned host memory; the transfers of neighbouring
    # frames overlap the compute (what a simulator loop that owns the sensor does).
    depth = max(4, 2 * eng_lanes)  # frames in flight (two per lane); one pinned output buffer per frame in flight
    outs = [out_np] + [torch.empty(out_shape, dtype=torch.float32).pin_memory().numpy() for _ in range(depth - 1)]
    bbk = dict(zip(("bbox", "bbox_start_x", "bbox_start_y", "bbox_width", "bbox_height"), bb))

    def piped(nsteps):
        tk = [None] * depth
        for i in range(nsteps):
            if tk[i % depth] is not None:
                eng.wait(tk[i % depth])  # frame i-depth delivered: its output buffer may be reused
            tk[i % depth] = eng.submit(pin_l[i % n_sets].numpy(), pin_r[i % n_sets].numpy(), out=outs[i % depth], **bbk)
        for t in tk:
            if t is not None:
                eng.wait(t)

    piped(2 * depth)
    barrier(world)
    t0 = time.perf_counter()
    piped(e2e_steps)
    torch.cuda.synchronize()
    piped_s = max_over_ranks(time.perf_counter() - t0, world)
    h2d, d2h = int(2 * batch * prm.rows * prm.cols), int(out_np.nbytes)
    e2e = {"value": batch * e2e_steps * world / piped_s, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "api": "extension path, pipelined: per step ticket = submit(left_u8 pinned, right_u8 pinned, out=pinned[, bbox]); wait(ticket of step-%d): %d frames in flight (two per lane), " % (depth, depth) +
                  "every step's inputs uploaded and depth map delivered inside the timed region",
           "steps": e2e_steps,
           "one_frame_at_a_time": {"value": batch * e2e_steps * world / ext_s, "unit": "frames/s",
                                   "api": "extension path, synchronous: bind_output(pinned) once; per step compute(left_u8 pinned, right_u8 pinned[, bbox]) + get_ndarray(out=pinned)"},
           "strict": {"value": batch * e2e_steps * world / strict_s, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                      "api": "reference signature only: compute(left_u8 ndarray, right_u8 ndarray[, bbox]) + get_ndarray() (pageable inputs; the map is delivered "
                             "into a page-locked array of a small pool that get_ndarray() hands out -- reused once the caller has dropped it)"}}

    batched = None
    del dev_sets
    if not args.no_batched:
        del eng
        torch.cuda.empty_cache()
        batched = batched_ours(args, rank, world, local)
    clocks = sampler.stop()

    if rank != 0:
        return
    # ---- roofline of the dominant kernel ---------------------------------------------------------
    peak, peak_src = hbm_peak()
    s = (prm.rows * prm.cols) if bbox_t is None else bbox_t[2] * bbox_t[3]
    V = 2 * s * prm.max_disp * batch
    kernel_bytes = {"cost": 8 * s * batch + V, "aggr_left": 2 * V, "aggr_down": 2 * V, "aggr_left_down": 4 * V, "aggr_up": 4 * V, "aggr_right_wta": 2 * V + 6 * s * batch}
    st_ms = {k: v for k, v in stages.items() if k != "frames"}
    dom = max((k for k in kernel_bytes if k in st_ms), key=lambda k: st_ms[k], default=None)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_c1.json")
    traffic_src = None
    if args.workload == "C1" and batch == 1 and os.path.exists(tpath):  # measured DRAM bytes per launch from the committed ncu capture
        tj = json.load(open(tpath))
        traffic, traffic_src = tj["dram_bytes_per_launch"], "profiles/traffic_c1.json (" + tj.get("source", "ncu --set full capture") + ")"
    roofline = None
    if dom:
        ach = kernel_bytes[dom] / (st_ms[dom] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": {"cost": "cost_kernel", "aggr_left": "aggr_kernel<MODE 0> (right->left)", "aggr_down": "aggr_kernel<MODE 0> (top->bottom)",
                                               "aggr_left_down": "aggr_kernel<MODE 0> x2 (right->left || top->bottom, two streams)",
                                               "aggr_up": "aggr_kernel<MODE 1> (bottom->top + L1 + L2)", "aggr_right_wta": "aggr_wta_kernel (left->right + blend + WTA)"}[dom],
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": (traffic or {}).get(dom), "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": int(kernel_bytes[dom]), "avg_launch_ms": st_ms[dom], "peak_source": peak_src,
                    "per_kernel_gbs": {k: kernel_bytes[k] / (st_ms[k] * 1e-3) / 1e9 for k in kernel_bytes if k in st_ms}}
    alg = configs.algorithmic_bytes(prm, rgba_input=True, bbox=bbox_t, point_cloud=pc) * batch
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u16", "data": "synthetic",
        "config": shared_config(args, desc, batch, world),
        "arm": {"impl": "sapien_b200 (this repo)", "volumes_mb": 4 * V / 1e6, "lanes": eng_lanes, "host_enqueue_us_per_frame": host_us,
                "one_lane": {"ms_per_step": sum(st_ms.values()), "value": batch * 1e3 / max(sum(st_ms.values()), 1e-9),
                             "note": "the same frames on ONE lane (the profiled loop: sum of the stage times), i.e. without the overlap of consecutive frames"},
                "pipelining": "frames enqueued back to back; consecutive frames alternate between the engine's lanes (independent stream / buffer sets) and overlap on the GPU, "
                              "results in submission order on the engine's public stream; the front-end of a frame runs under the previous frame of its lane"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches * args.steps,
        "roofline": roofline,
        "frame_roofline": {"algorithmic_bytes_per_step": int(alg), "achieved_gbs": alg / (ms / args.steps * 1e-3) / 1e9 ,
                           "frac_of_peak": alg / (ms / args.steps * 1e-3) / 1e9 / peak},
        "stages_ms": st_ms,
    }
    if batched is not None:
        line["batched"] = batched
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(prm, bbox_t, args.cpu_seconds)
    emit(line)


def run_reference(args, rank, world, local):
    """The unmodified reference simsense (oracle/_ref) through its own compute()/getMat2d() API.  Nothing of this
    repo's engine is imported on this arm (the camera presets and the synthetic images are plain Python/numpy)."""
    import torch

    from oracle import REF_SO, RefEngine, configs

    if not os.path.exists(REF_SO):
        if rank == 0:
            emit({"impl": "reference", "unavailable": "oracle/_ref/libsimsense_ref.so not built (needs /root/reference at build time)"})
        return
    key, batch, bbox, pc, desc = WORKLOADS[args.workload]
    if args.batch:
        batch = args.batch
    prm = configs.params(key)
    bbox_t = configs.BBOX_C2 if bbox else None
    ref = RefEngine(prm)  # the reference has no batch API: a batch is a sequential loop over its environments
    n_sets = 8
    host_sets, dev_sets = make_inputs(prm, 1, n_sets, torch)
    rgba = None
    if pc:
        from sapien_b200 import synth

        rgba = torch.from_numpy(synth.make_rgb(prm.rgb_rows, prm.rgb_cols, 0)).cuda()
    steps = args.steps

    def step(i):
        for b in range(batch):
            l, r = dev_sets[(i + b) % n_sets]
            ref.compute_device(l.data_ptr(), r.data_ptr(), bbox_t)
            if pc:
                ref.rgb_point_cloud_device(rgba.data_ptr())

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    barrier(world)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    value = batch * steps * world / (ms / 1e3)
    e2e_steps = max(20, min(steps, 60))
    out = None
    for i in range(2):
        ref.compute_host(*host_sets[i % n_sets], bbox_t)
        out = ref.depth()
    barrier(world)
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        ref.compute_host(*host_sets[i % n_sets], bbox_t)
        out = ref.depth()
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)
    ref.close()
    del dev_sets
    torch.cuda.empty_cache()
    batched = None if args.no_batched else batched_reference(args, rank, world, local)
    clocks = sampler.stop()
    if rank != 0:
        return
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": shared_config(args, desc, batch, world),
        "arm": {"impl": "unmodified reference simsense CUDA sources recompiled for sm_100a (oracle/_ref), through simsense::DepthSensorEngine::compute(void*, void*, ...)",
                "note": "the reference has no CPU implementation of this path and no batch API (a batch is a sequential loop)"},
        "clocks": clocks,
        "e2e": {"value": e2e_steps * world / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": int(2 * prm.rows * prm.cols),
                "d2h_bytes_per_step": int(out.nbytes), "api": "DepthSensorEngine::compute(Mat2d<u8>, Mat2d<u8>) + getMat2d()", "steps": e2e_steps},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": 0, "kind": "reference",
                         "sample": "reference simsense is CUDA-only: this arm runs its own kernels on the same B200 (0 host compute threads)"},
    }
    if batched is not None:
        line["batched"] = batched
    emit(line)


def run_reference_cpu(args, rank):
    if rank != 0:
        return
    from oracle import configs

    key, batch, bbox, pc, desc = WORKLOADS[args.workload]
    prm = configs.params(key)
    bbox_t = configs.BBOX_C2 if bbox else None
    cb = cpu_baseline(prm, bbox_t, max(args.cpu_seconds, 10.0))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "frames/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / cb["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "note": "scalar C/OpenMP port of the reference kernels (oracle/simsense_oracle.c)"},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    args = parse()
    if args.impl == "reference-cpu":
        run_reference_cpu(args, int(os.environ.get("RANK", "0")))
        return
    rank, world, local = dist_setup(args.gpus)
    try:
        if args.impl == "reference":
            run_reference(args, rank, world, local)
        else:
            run_ours(args, rank, world, local)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
