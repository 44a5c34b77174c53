"""Single-episode online latency (the robot loop, robot_demo.py:559-566): EpisodeBatch(E=1).step_detections per frame, eager vs one
CUDA graph per frame (capture_step_detections), C=512, 480x640, 200x200 robot map and 500x500; CUDA events over 200 frames after
warm-up.  Inputs resident on the device (the demo's depth / detections are produced there)."""
import importlib, json, math, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
eod = importlib.import_module("embodied-object-detection_b200")
dev = torch.device("cuda:0")
H, W, C, Kmax, cell, N = 480, 640, 512, 16, 0.2, 200
out = {}
for mw in (200, 500):
    mh = mw
    ep = eod.episodes.make_episode(1234, 8, H, W, mw, mh, cell)
    Tm = eod.transform3d(torch.from_numpy(ep.xyzhe))
    pose = Tm[:, :3, :].reshape(-1, 1, 12).to(dev)
    depth = torch.from_numpy(ep.depth)[:, None].contiguous().to(dev)
    shifts = torch.from_numpy(np.concatenate([np.zeros(3, np.float32), ep.map_world_shift])[None]).to(dev)
    intr = eod.compute_intrinsics(W, H, math.radians(67.5))
    rng = np.random.default_rng(0)
    dets = []
    for s in range(4):
        f, p, b = eod.episodes.make_mask_head_detections(rng, H, W, C, (Kmax, Kmax), 28)
        dets.append(tuple(torch.from_numpy(a[None]).to(dev) for a in (f, p, b)) + (torch.full((1,), Kmax, dtype=torch.int32, device=dev),))
    res = {}
    for mode in ("eager", "graph"):
        batch = eod.EpisodeBatch(1, mw, mh, C, H, W, dev)
        fn = None
        if mode == "graph":
            fn = batch.capture_step_detections(depth[0], pose[0], shifts, intr, cell, *dets[0])
        def frame(t):
            d = dets[t & 3]
            if fn is None:
                return batch.step_detections(depth[t & 7], pose[t & 7], shifts, intr, cell, *d)
            return fn(depth[t & 7], pose[t & 7], shifts, *d)
        for t in range(20):
            frame(t)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for t in range(N):
            frame(t)
        b.record()
        torch.cuda.synchronize()
        res[mode] = {"ms_per_frame_device": a.elapsed_time(b) / N, "ms_per_frame_wall": (time.perf_counter() - t0) * 1e3 / N,
                     "checksum": float(batch.sums.abs().sum()), "cells_seen": int((batch.counts > 0).sum())}
    out[f"map{mw}"] = res
print(json.dumps(out))
