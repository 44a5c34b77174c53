#!/usr/bin/env python
"""Headline benchmark: spatial-feature-memory write+read throughput (frames/s) on B200.

Workload (BASELINE.json configs[1]): 64 synthetic MP3D-shaped episodes x 20 frames advanced in lock step on one
GPU, 480x640 RGB-D, per-pixel C=256 fp32 features (CHW, the reference's image_features layout), 500x500 grid of
0.2 m cells.  One *step* = one pass over the whole workload = 64 x 20 = 1280 frames; per frame-step the hot path
is  back-project+quantise -> read (normalise, fp16, gather, pool x3) -> write (count, scatter-mean, finalise).

  value        frames/s, inputs resident in HBM, CUDA events, max over ranks
  e2e          the same metric through the Python plugin API with HOST (pinned) inputs: H2D of depth, pose and
               features and D2H of the pooled levels inside the timed region
  roofline     dominant kernel (eod_write_mean): algorithmic bytes per launch / mean CUDA-event duration of that
               launch inside the timed region, vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline the oracle port of the same path (torch-CPU restatement of the reference ops) on the host cores
  extras       (not part of `value`) two more stages of the path on the same batch: the tcgen05 projection + fusion of the
               three levels, and the object regime (detections as mask probabilities + boxes) through step_detections

`--impl reference` times that CPU implementation alone (the reference cannot run as committed: hard-coded
.cuda() and detectron2 imports), on a bounded sample of the same workload per step.

Multi-GPU (torchrun, one process per GPU): episodes are sharded `episode % world == rank`, per-GPU work fixed
(64 episodes each, "weak"), no collective on the hot path; one all_reduce(MAX) for the time and one
all_reduce(SUM) for the frame counters after the loop.
"""
from __future__ import annotations

import argparse
import importlib
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "memory write+read frames/sec (480x640, C=256)"
UNIT = "frames/s"
H, W, C = 480, 640, 256
MAP_W = MAP_H = 500
CELL = np.float32(0.02 * 10)
N_EPISODES, N_FRAMES = 64, 20
N_PIX = H * W


def _gen_episode(args):
    seed, n_frames, map_w, map_h, cell = args
    episodes = importlib.import_module("embodied-object-detection_b200.episodes")
    ep = episodes.make_episode(seed, n_frames, H, W, map_w, map_h, float(cell))
    return ep.depth, ep.xyzhe, ep.map_world_shift


def make_inputs(episode_ids, n_frames=N_FRAMES, map_w=MAP_W, map_h=MAP_H, cell=CELL):
    """Host-side synthetic episodes (seed 1234 + episode id), generated in parallel on the host cores."""
    jobs = [(1234 + int(e), n_frames, map_w, map_h, cell) for e in episode_ids]
    workers = max(1, min(len(jobs), (os.cpu_count() or 2) - 1, 32))
    if workers > 1:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(workers) as pool:
            out = pool.map(_gen_episode, jobs)
    else:
        out = [_gen_episode(j) for j in jobs]
    depth = np.stack([o[0] for o in out])            # (E, T, H, W)
    xyzhe = np.stack([o[1] for o in out])            # (E, T, 5)
    shift = np.stack([o[2] for o in out])            # (E, 3)
    return depth, xyzhe, shift


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(index: int):
    """Pin this rank's CPU threads - and with them the first-touch placement of its pinned staging buffers - to the NUMA node
    its GPU hangs off.  Unbound, the 8 ranks of a box stream most of their H2D traffic across the socket interconnect and the
    e2e arm is host-memory bound at less than half of 8 x PCIe.  Best effort: returns the node, or None when sysfs says nothing."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)], capture_output=True,
                             text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return None
        dom, rest = bus.split(":", 1)
        node = int(open(f"/sys/bus/pci/devices/{dom[-4:]}:{rest}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(E: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of one write-kernel launch, from the committed ncu pass
    (profiles/write_kernel_traffic.json, written by profiles/summarize.py); None if it was taken at another batch size."""
    p = os.path.join(ROOT, "profiles", "write_kernel_traffic.json")
    try:
        d = json.load(open(p))
        return float(d["traffic_bytes_per_launch"]) if int(d["episodes_per_launch"]) == E else None
    except Exception:
        return None


def write_launch_bytes(E: int, touched_per_launch: float) -> float:
    """Algorithmic bytes of ONE eod_write_mean launch over E episodes (DESIGN.md 'Roofline accounting'):
    features read once + index plane + touched grid rows read-modify-written + per-cell sample counts."""
    return E * (N_PIX * C * 4 + N_PIX * 4) + touched_per_launch * (C * 4 * 2 + 4)


def frame_bytes(visible_per_frame: float) -> float:
    """Whole-path algorithmic bytes per frame (SURVEY 8d / BASELINE.md 5), M = V (every pixel sampled)."""
    v = visible_per_frame
    write = N_PIX * 4 + 64 + N_PIX * C * 4 + v * (C * 4 * 2 + 8)
    read = N_PIX * 4 + v * (C * 4 + 4) + 6300 * C * 2
    return write + read


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the path (reference ops restated on the CPU)
# ---------------------------------------------------------------------------------------------------------
def cpu_frames(depth, xyzhe, shift, n_frames, feats):
    """Run n_frames of ONE episode through the CPU port: C back-projection, torch read chain, sparse-equivalent
    (index_add_) write of the dense per-pixel features.  Returns seconds."""
    import oracle
    from oracle import reference_ops as R
    intr = R.intrinsics(W, H, math.radians(67.5))
    T = R.transform3d(torch.from_numpy(xyzhe)).numpy()
    sums = torch.zeros(MAP_W * MAP_H, C)
    counts = torch.zeros(MAP_W * MAP_H)
    observed = torch.ones(H, W, dtype=torch.bool)
    t0 = time.perf_counter()
    for t in range(n_frames):
        idx = oracle.backproject_quantize(depth[t], T[t], intr, np.zeros(3, np.float32), shift, CELL, MAP_W, MAP_H, 0, 0.5, want=("idx",))["idx"]
        proj = torch.from_numpy(idx).long()
        R.read_frame(sums, counts, proj)
        sums, counts = R.write_mean_frame(sums, counts, feats[t % len(feats)], observed, proj, stride=1)
    return time.perf_counter() - t0


def run_cpu_sample(n_frames: int):
    torch.set_num_threads(os.cpu_count() or 1)
    depth, xyzhe, shift = make_inputs([0], n_frames=max(n_frames, 2))
    g = torch.Generator().manual_seed(0)
    feats = [torch.randn(1, C, H, W, generator=g) for _ in range(2)]
    cpu_frames(depth[0], xyzhe[0], shift[0], 1, feats)          # warm-up
    sec = cpu_frames(depth[0], xyzhe[0], shift[0], n_frames, feats)
    return n_frames / sec, sec


def main_reference(args, rank, world):
    if rank != 0:
        return
    n_frames = 4
    torch.set_num_threads(os.cpu_count() or 1)
    depth, xyzhe, shift = make_inputs([0], n_frames=n_frames)
    g = torch.Generator().manual_seed(0)
    feats = [torch.randn(1, C, H, W, generator=g) for _ in range(2)]
    for _ in range(max(args.warmup, 1)):
        cpu_frames(depth[0], xyzhe[0], shift[0], 1, feats)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_frames(depth[0], xyzhe[0], shift[0], n_frames, feats)
    sec = time.perf_counter() - t0
    fps = args.steps * n_frames / sec
    sample = f"{n_frames} frames of episode 0 per step (of the 64x20-frame workload); sparse-equivalent index_add_ write"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(world):
    return {"workload": f"configs[1]: {N_EPISODES} episodes x {N_FRAMES} frames per GPU, 480x640 RGB-D, dense per-pixel C=256 fp32 "
                        f"features (CHW), 500x500 grid @0.2 m, write=per-cell mean (stride 1) + visibility counts, "
                        f"read=normalise+fp16+gather+pool to 60x80/30x40/15x20",
            "episodes_per_gpu": N_EPISODES, "frames_per_episode": N_FRAMES, "grid": [MAP_W, MAP_H], "channels": C,
            "parallelism": f"episode-sharded x{world}, no hot-path collective",
            "l2_policy": "inputs larger than L2 (20.1 GB of features per frame-step, two alternating slabs)"}


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def main_gpu(args, rank, local_rank, world):
    eod = importlib.import_module("embodied-object-detection_b200")
    sharding = eod.sharding
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    eod._lib.lib()                                                   # fail loudly when the extension is missing
    numa_node = None if args.no_numa else bind_to_gpu_numa_node(local_rank)

    E = N_EPISODES
    my_eps = [rank + world * i for i in range(E)]                    # weak scaling: 64 episodes per GPU
    depth_h, xyzhe_h, shift_h = make_inputs(my_eps)
    intr = eod.compute_intrinsics(W, H, math.radians(67.5))
    T = eod.transform3d(torch.from_numpy(xyzhe_h.reshape(-1, 5))).reshape(E, N_FRAMES, 4, 4)
    pose = T[:, :, :3, :].reshape(E, N_FRAMES, 12).permute(1, 0, 2).contiguous().to(dev)       # (T, E, 12)
    depth = torch.from_numpy(depth_h).permute(1, 0, 2, 3).contiguous().to(dev)                 # (T, E, H, W)
    shifts = torch.from_numpy(np.concatenate([np.zeros_like(shift_h), shift_h], 1)).to(dev)   # (E, 6)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    slabs = [torch.randn((E, C, H, W), device=dev, generator=gen) for _ in range(2)]          # 2 x 20.1 GB
    # inputs are resident and static for the whole run, which is what pipeline=True asks of the caller
    batch = eod.EpisodeBatch(E, MAP_W, MAP_H, C, H, W, dev, variant=args.write_variant, pipeline=not args.no_pipeline)

    def one_step():
        batch.reset()
        for t in range(N_FRAMES):
            batch.step(depth[t], pose[t], shifts, intr, float(CELL), slabs[t & 1])
        batch.join()                                                 # the timing events sit on the caller's stream

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()                                                  # samples cover warm-up + timed steps (all under load)
    for _ in range(args.warmup):
        one_step()
    barrier()
    launches0 = eod.ops.launch_count
    batch.profile(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        one_step()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = sharding.max_over_ranks(ev0.elapsed_time(ev1), dev)
    launches = eod.ops.launch_count - launches0
    stage_ms = batch.stage_ms()
    batch.profile(False)
    visible_last_step = float(batch.counts.sum().item())            # +1 per visible cell per frame since the reset
    totals = sharding.gather_counters({"frames": float(args.steps * E * N_FRAMES), "visible": visible_last_step}, dev)

    ms_per_step = ms_total / args.steps
    value = totals["frames"] / (ms_total / 1e3)
    vis_per_frame = visible_last_step / (E * N_FRAMES)
    peak, peak_src = measured_peak_gbs()
    wbytes = write_launch_bytes(E, vis_per_frame * E)
    achieved = wbytes / (stage_ms["write"] * 1e-3) / 1e9 if stage_ms.get("write") else None
    path_gbs = frame_bytes(vis_per_frame) * (value / world) / 1e9

    extras = run_extras(eod, batch, dev, depth, pose, shifts, intr) if (rank == 0 and not args.no_extras) else None
    if world > 1:
        torch.distributed.barrier()
    # ---- e2e through the plugin API with host buffers (rank-local, then max over ranks) ----
    e2e = None if args.no_e2e else run_e2e(eod, batch, dev, depth_h, pose.cpu(), shifts, intr, args, world, sharding, slabs)

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": dict(workload_config(world), host_numa_node=numa_node), "clocks": clocks, "gpu_launches": int(launches),
        "e2e": e2e, "extras": extras,
        "roofline": {"bound": "hbm", "kernel": "write_mean_chw_tma_kernel<256>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "frac_of_nominal_8TBs": (achieved / 8000.0) if achieved else None,
                     "traffic": measured_traffic(E), "peak_source": peak_src, "bytes_per_launch": wbytes, "launch_ms": stage_ms.get("write"),
                     "stage_ms": stage_ms, "visible_cells_per_frame": vis_per_frame,
                     "whole_path": {"bytes_per_frame": frame_bytes(vis_per_frame), "achieved_gbs_per_gpu": path_gbs,
                                    "frac": path_gbs / peak}},
    }
    if rank == 0:
        if world == 1 and not args.no_cpu:
            fps, sec = run_cpu_sample(args.cpu_frames)
            out["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                   "sample": f"{args.cpu_frames} frames of episode 0 ({sec:.1f} s): C back-projection + torch-CPU read chain "
                                             f"+ sparse-equivalent (index_add_) write; the literal one-hot matmul of the reference needs "
                                             f"{N_PIX * MAP_W * MAP_H / 1e9:.0f} GB at stride 1"}
        print(json.dumps(out))


def run_extras(eod, batch, dev, depth, pose, shifts, intr):
    """Two more stages of the same path, timed on the bench batch after the headline loop (CUDA events, resident inputs; not part
    of `value`): the tensor-core projection + fusion of the three read levels and the reference's live object
    regime (kept detections as 28x28 mask probabilities + boxes; mask pasting folded into the write).  Best effort."""
    out = {}
    try:
        E = batch.E
        ops = eod.ops
        g = torch.Generator(device=dev).manual_seed(5)
        levels = [l.permute(0, 3, 1, 2) for l in batch.levels]                        # the last frame's pooled levels, channels-last memory
        N = 256
        ws = [ops.project_split_weights(torch.randn((N, C), device=dev, generator=g) / C ** 0.5) for _ in range(3)]
        bias = [torch.randn((N,), device=dev, generator=g) for _ in range(3)]
        res = [torch.randn((E, N, l.shape[2], l.shape[3]), device=dev, generator=g) for l in levels]
        outs = [torch.empty_like(r) for r in res]
        for _ in range(3):
            ops.project_fuse_levels(levels, ws, bias, res, 5.0, 0, outs)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            ops.project_fuse_levels(levels, ws, bias, res, 5.0, 0, outs)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        nbytes = sum(l.numel() * 2 for l in levels) + 2 * sum(r.numel() * 4 for r in res)
        out["project_fuse"] = {"ms_per_frame_step": ms, "GBps": nbytes / ms / 1e6, "kernel": "project_fuse_persistent_kernel (tcgen05)",
                               "shape": f"E={E}, K={C} -> N={N}, levels 60x80/30x40/15x20"}
        del res, outs
        rng = np.random.default_rng(0)
        Kmax = 16
        bf = np.zeros((E, Kmax, C), np.float32); pr = np.zeros((E, Kmax, 28, 28), np.float32); bx = np.zeros((E, Kmax, 4), np.float32)
        n = np.zeros(E, np.int32)
        for e in range(E):
            f, p_, b_ = eod.episodes.make_mask_head_detections(rng, H, W, C, (4, Kmax), 28)
            n[e] = f.shape[0]; bf[e, : n[e]], pr[e, : n[e]], bx[e, : n[e]] = f, p_, b_
        det = [torch.from_numpy(x).to(dev) for x in (bf, pr, bx, n)]
        batch.join()
        for t in range(3):
            batch.step_detections(depth[t], pose[t], shifts, intr, float(CELL), *det)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for t in range(N_FRAMES):
            batch.step_detections(depth[t], pose[t], shifts, intr, float(CELL), *det)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / N_FRAMES
        out["object_regime"] = {"ms_per_frame_step": ms, "frames_per_s": E / ms * 1e3,
                                "shape": f"E={E}, C={C}, <= {Kmax} detections per frame, 28x28 mask probabilities + boxes, every 8th observed pixel"}
    except Exception as exc:                                                          # never lose the headline line to an extra
        out["error"] = repr(exc)[:200]
    return out


def run_e2e(eod, batch, dev, depth_h, pose_h, shifts, intr, args, world, sharding, slabs):
    """frames/s through EpisodeBatch.step with HOST inputs: per frame-step, H2D of depth (E,H,W), pose (E,12) and
    features (E,C,H,W) from pinned memory on a copy stream (double-buffered against compute), and D2H of the three
    pooled fp16 levels.  The feature H2D (20.1 GB per frame-step) makes this PCIe-bound by construction."""
    E = batch.E
    batch.join()
    batch.pipeline = False                                           # host-fed inputs: every step is ordered on the caller's stream
    n_slots = 4                                                      # pinned staging: 4 episode-frames of features (1.26 GB)
    g = torch.Generator().manual_seed(7)
    pin_feat = torch.randn((n_slots, C, H, W), generator=g).pin_memory()
    pin_depth = torch.from_numpy(depth_h).permute(1, 0, 2, 3).contiguous().pin_memory()      # (T, E, H, W)
    pin_pose = pose_h.contiguous().pin_memory()
    dev_feat = slabs                                                 # reuse the two resident 20.1 GB slabs as H2D targets
    dev_depth = [torch.empty((E, H, W), device=dev) for _ in range(2)]
    dev_pose = [torch.empty((E, 12), device=dev) for _ in range(2)]
    host_levels = [torch.empty(l.shape, dtype=l.dtype).pin_memory() for l in batch.levels]
    copy_s = torch.cuda.Stream(device=dev)
    comp_s = torch.cuda.current_stream(dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    h2d = E * (C * H * W * 4 + H * W * 4 + 12 * 4)
    d2h = sum(l.numel() * 2 for l in batch.levels)
    frames_per_step = E * args.e2e_frames

    def upload(t, buf):
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(freed[buf])
            for e in range(E):
                dev_feat[buf][e].copy_(pin_feat[(e + t) % n_slots], non_blocking=True)
            dev_depth[buf].copy_(pin_depth[t % N_FRAMES], non_blocking=True)
            dev_pose[buf].copy_(pin_pose[t % N_FRAMES], non_blocking=True)
            ready[buf].record(copy_s)

    def one_step():
        batch.reset()
        upload(0, 0)
        for t in range(args.e2e_frames):
            b = t & 1
            if t + 1 < args.e2e_frames:
                upload(t + 1, b ^ 1)
            comp_s.wait_event(ready[b])
            levels = batch.step(dev_depth[b], dev_pose[b], shifts, intr, float(CELL), dev_feat[b])
            for hl, l in zip(host_levels, batch.levels):
                hl.copy_(l, non_blocking=True)
            freed[b].record(comp_s)
        torch.cuda.synchronize()
        return float(host_levels[2].float().abs().sum())             # device->host result actually read on the host

    for f in freed:
        f.record(comp_s)
    one_step()                                                        # warm-up
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        one_step()
    sec = time.perf_counter() - t0
    sec = sharding.max_over_ranks(sec, dev)
    return {"value": world * frames_per_step * args.e2e_steps / sec, "unit": UNIT,
            "h2d_bytes_per_step": int(h2d * args.e2e_frames), "d2h_bytes_per_step": int(d2h * args.e2e_frames),
            "frames_per_step": frames_per_step, "steps": args.e2e_steps,
            "note": "bounded: e2e step = %d frame-steps of the 64-episode batch; PCIe-bound (features cross the bus, which the "
                    "reference never does: its features are produced on the device)" % args.e2e_frames}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-frames", type=int, default=8)
    ap.add_argument("--e2e-frames", type=int, default=4)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--write-variant", type=int, default=0, help="diagnostics: 0 auto, 1 LDG, 2 TMA, 3 TMA dry (no result)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the projection+fusion and object-regime stage timings")
    ap.add_argument("--no-numa", action="store_true", help="diagnostics: do not bind the rank to its GPU's NUMA node")
    ap.add_argument("--no-pipeline", action="store_true", help="diagnostics: stream-ordered EpisodeBatch.step (no cross-frame overlap)")
    args = ap.parse_args()
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        return main_reference(args, rank, world)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        main_gpu(args, rank, local_rank, world)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
