"""Stage breakdown of the object-regime frame (the reference's live path) at E=64, C=512, 500x500: CUDA events around every
stage launch, mean over frames.  PROF_E / PROF_K override the batch and the max detections per frame."""
import importlib, json, math, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
eod = importlib.import_module("embodied-object-detection_b200")
ops = eod.ops
dev = torch.device("cuda:0")
H, W = 480, 640
E = int(os.environ.get("PROF_E", 64)); Kmax = int(os.environ.get("PROF_K", 16)); C = 512; mw = mh = 500; cell = 0.2; T = int(os.environ.get("PROF_T", 24))
eps = [eod.episodes.make_episode(1234 + e, 4, H, W, mw, mh, cell) for e in range(E)]
Tm = eod.transform3d(torch.from_numpy(np.stack([ep.xyzhe for ep in eps]).reshape(-1, 5))).reshape(E, 4, 4, 4)
pose = Tm[:, :, :3, :].reshape(E, 4, 12).permute(1, 0, 2).contiguous().to(dev)
depth = torch.from_numpy(np.stack([ep.depth for ep in eps])).permute(1, 0, 2, 3).contiguous().to(dev)
shifts = torch.from_numpy(np.stack([np.concatenate([np.zeros(3, np.float32), ep.map_world_shift]) for ep in eps])).to(dev)
intr = eod.compute_intrinsics(W, H, math.radians(67.5))
batch = eod.EpisodeBatch(E, mw, mh, C, H, W, dev)
rng = np.random.default_rng(0)
dets = []
for s in range(2):
    bf = np.zeros((E, Kmax, C), np.float32); pr = np.zeros((E, Kmax, 28, 28), np.float32); bx = np.zeros((E, Kmax, 4), np.float32)
    n = np.zeros(E, np.int32)
    for e in range(E):
        f, p, b = eod.episodes.make_mask_head_detections(rng, H, W, C, (4, Kmax), 28)
        n[e] = f.shape[0]; bf[e, : n[e]], pr[e, : n[e]], bx[e, : n[e]] = f, p, b
    dets.append(tuple(torch.from_numpy(a).to(dev) for a in (bf, pr, bx, n)))
S = -(-H * W // 8)
slots = ops.ObjectSlots(E, mw * mh, C, S, dev)
stages = {}


def timed(name, fn, *a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = fn(*a, **k); e1.record()
    stages.setdefault(name, []).append((e0, e1))
    return out


tot = []
for t in range(T):
    bf, pr, bx, n = dets[t & 1]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    timed("project", batch.project, depth[t % 4], pose[t % 4], shifts, intr, cell)
    timed("read", batch.read)
    _, observed = timed("paste_observed", ops.paste_masks, pr, bx, (H, W), 0.5, n, want_masks=False, want_observed=True)
    samp = timed("sample_mask", ops.sample_mask, observed, 8)
    timed("frame_count", ops.frame_count, batch.idx, samp, batch.frame_cnt, n, slots)
    timed("write_objects_pasted", ops.write_objects_pasted, bf, pr, bx, n, batch.idx, samp, slots, 0.5)
    timed("flush_slots", ops.flush_slots, batch.frame_cnt, slots, batch.sums)
    timed("finalize", batch._finalize)
    b.record()
    tot.append((a, b))
torch.cuda.synchronize()
res = {k: sum(x.elapsed_time(y) for x, y in v[2:]) / len(v[2:]) for k, v in stages.items()}
res["frame_step_ms"] = sum(x.elapsed_time(y) for x, y in tot[2:]) / len(tot[2:])
res["frames_per_s"] = E / res["frame_step_ms"] * 1e3
res["sampled_px_per_episode"] = float(samp.sum().item()) / E
# the same frame through EpisodeBatch.step_detections (two streams)
ov = []
for t in range(T):
    bf, pr, bx, n = dets[t & 1]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    batch.step_detections(depth[t % 4], pose[t % 4], shifts, intr, cell, bf, pr, bx, n)
    b.record()
    ov.append((a, b))
torch.cuda.synchronize()
ms_sorted = sorted(x.elapsed_time(y) for x, y in ov[2:])
res["step_detections_ms"] = ms_sorted[len(ms_sorted) // 2]                    # median: a single slow step (allocator, clock ramp) does not move it
res["step_detections_ms_mean"] = sum(ms_sorted) / len(ms_sorted)
res["step_detections_frames_per_s"] = E / res["step_detections_ms"] * 1e3
# pipelined: frame t's write side under frame t+1's projection / paste / sample (inputs resident, as the mode requires); throughput over the loop
pb = eod.EpisodeBatch(E, mw, mh, C, H, W, dev, pipeline=True)
for t in range(4):
    bf, pr, bx, n = dets[t & 1]
    pb.step_detections(depth[t % 4], pose[t % 4], shifts, intr, cell, bf, pr, bx, n)
pb.join()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for t in range(2 * T):
    bf, pr, bx, n = dets[t & 1]
    pb.step_detections(depth[t % 4], pose[t % 4], shifts, intr, cell, bf, pr, bx, n)
pb.join()
b.record()
torch.cuda.synchronize()
res["step_detections_pipelined_ms"] = a.elapsed_time(b) / (2 * T)
res["step_detections_pipelined_frames_per_s"] = E / res["step_detections_pipelined_ms"] * 1e3
print(json.dumps({"E": E, "Kmax": Kmax, "C": C, **res}))
