"""Timeline of the object-regime e2e loop (bench.py run_e2e_variants) from CUDA events: where a frame-step's wall time goes."""
import importlib, json, math, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
eod = importlib.import_module("embodied-object-detection_b200")
dev = torch.device("cuda:0")
H, W, C, E, Kmax, mw, cell, T = 480, 640, 256, 64, 16, 500, 0.2, 4
eps = [eod.episodes.make_episode(1234 + e, T, H, W, mw, mw, cell) for e in range(E)]
Tm = eod.transform3d(torch.from_numpy(np.stack([ep.xyzhe for ep in eps]).reshape(-1, 5))).reshape(E, T, 4, 4)
pin_pose = Tm[:, :, :3, :].reshape(E, T, 12).permute(1, 0, 2).contiguous().pin_memory()
pin_depth = (torch.from_numpy(np.stack([ep.depth for ep in eps])).permute(1, 0, 2, 3) * 1000).round().clamp_(0, 65535).to(torch.uint16).contiguous().pin_memory()
shifts = torch.from_numpy(np.stack([np.concatenate([np.zeros(3, np.float32), ep.map_world_shift]) for ep in eps])).to(dev)
intr = eod.compute_intrinsics(W, H, math.radians(67.5))
batch = eod.EpisodeBatch(E, mw, mw, C, H, W, dev)
rng = np.random.default_rng(3)
bf = np.zeros((E, Kmax, C), np.float32); pr = np.zeros((E, Kmax, 28, 28), np.float32); bx = np.zeros((E, Kmax, 4), np.float32); n = np.zeros(E, np.int32)
for e in range(E):
    f, p_, b_ = eod.episodes.make_mask_head_detections(rng, H, W, C, (4, Kmax), 28)
    n[e] = f.shape[0]; bf[e, : n[e]], pr[e, : n[e]], bx[e, : n[e]] = f, p_, b_
pin_det = [torch.from_numpy(x).pin_memory() for x in (bf, pr, bx, n)]
dev_det = [[torch.empty(x.shape, dtype=x.dtype, device=dev) for x in pin_det] for _ in range(2)]
dev_depth = [torch.empty((E, H, W), dtype=torch.uint16, device=dev) for _ in range(2)]
dev_pose = [torch.empty((E, 12), device=dev) for _ in range(2)]
host_l2 = torch.empty(batch.levels[2].shape, dtype=torch.float16).pin_memory()
comp_s = torch.cuda.current_stream(dev); copy_s = torch.cuda.Stream(device=dev)
ready = [torch.cuda.Event() for _ in range(2)]; freed = [torch.cuda.Event() for _ in range(2)]
ev = lambda: torch.cuda.Event(enable_timing=True)
N = 40
marks = []
def upload(t, b, m):
    with torch.cuda.stream(copy_s):
        copy_s.wait_event(freed[b])
        m["u0"].record(copy_s)
        dev_depth[b].copy_(pin_depth[t % T], non_blocking=True)
        dev_pose[b].copy_(pin_pose[t % T], non_blocking=True)
        for d, p_ in zip(dev_det[b], pin_det):
            d.copy_(p_, non_blocking=True)
        m["u1"].record(copy_s)
        ready[b].record(copy_s)
for mode in ("d2h_inline", "no_d2h", "no_upload"):
    for f in freed:
        f.record(comp_s)
    torch.cuda.synchronize()
    marks = [{k: ev() for k in ("u0", "u1", "s0", "s1", "d1")} for _ in range(N + 1)]
    base = ev(); base.record(comp_s)
    host = []
    t0 = time.perf_counter()
    upload(0, 0, marks[0])
    for t in range(N):
        b = t & 1
        h0 = time.perf_counter()
        if t + 1 < N and mode != "no_upload":
            upload(t + 1, b ^ 1, marks[t + 1])
        comp_s.wait_event(ready[b])
        marks[t]["s0"].record(comp_s)
        batch.step_detections(dev_depth[b], dev_pose[b], shifts, intr, cell, *dev_det[b])
        marks[t]["s1"].record(comp_s)
        if mode != "no_d2h":
            host_l2.copy_(batch.levels[2], non_blocking=True)
        marks[t]["d1"].record(comp_s)
        freed[b].record(comp_s)
        host.append(time.perf_counter() - h0)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / N * 1e3
    rel = lambda e_: base.elapsed_time(e_)
    mid = range(5, N - 2)
    out = {"mode": mode, "wall_ms_per_step": wall, "host_issue_ms_per_step": float(np.mean(host)) * 1e3,
           "step_ms": float(np.mean([marks[t]["s0"].elapsed_time(marks[t]["s1"]) for t in mid])),
           "d2h_ms": float(np.mean([marks[t]["s1"].elapsed_time(marks[t]["d1"]) for t in mid])),
           "gap_to_next_step_ms": float(np.mean([marks[t]["d1"].elapsed_time(marks[t + 1]["s0"]) for t in mid]))}
    if mode != "no_upload":
        out["upload_ms"] = float(np.mean([marks[t]["u0"].elapsed_time(marks[t]["u1"]) for t in mid]))
        out["upload_start_after_step_start_ms"] = float(np.mean([rel(marks[t + 1]["u0"]) - rel(marks[t]["s0"]) for t in mid]))
    print(json.dumps(out), flush=True)

# back to back like bench.py's extras.object_regime: no waits on uploads, no D2H, one event pair around the loop
for rep in range(2):
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for t in range(N):
        batch.step_detections(dev_depth[t & 1], dev_pose[t & 1], shifts, intr, cell, *dev_det[t & 1])
    b.record()
    torch.cuda.synchronize()
    print(json.dumps({"mode": "back_to_back", "ms_per_step": a.elapsed_time(b) / N}), flush=True)
depth_f32 = (dev_depth[0].float() / 1000).contiguous()
for rep in range(2):
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for t in range(N):
        batch.step_detections(depth_f32, dev_pose[t & 1], shifts, intr, cell, *dev_det[t & 1])
    b.record()
    torch.cuda.synchronize()
    print(json.dumps({"mode": "back_to_back_f32_depth", "ms_per_step": a.elapsed_time(b) / N}), flush=True)
batch.profile(True)
for t in range(10):
    batch.step_detections(depth_f32, dev_pose[t & 1], shifts, intr, cell, *dev_det[t & 1])
torch.cuda.synchronize()
print(json.dumps({"stage_ms_f32": batch.stage_ms()}), flush=True)
batch.profile(True)
for t in range(10):
    batch.step_detections(dev_depth[0], dev_pose[t & 1], shifts, intr, cell, *dev_det[t & 1])
torch.cuda.synchronize()
print(json.dumps({"stage_ms_u16": batch.stage_ms()}), flush=True)
