"""Read kernel alone (eod_read_pool) at the bench batch: E=64 episodes, 480x640, 500x500 grid populated by a few frames,
C=256 and C=512; CUDA events over 20 launches after 3 warm-ups."""
import importlib, json, math, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
eod = importlib.import_module("embodied-object-detection_b200")
ops = eod.ops
dev = torch.device("cuda:0")
H, W, mw, mh, cell = 480, 640, 500, 500, 0.2
E = int(os.environ.get("PROF_E", 64))
eps = [eod.episodes.make_episode(1234 + e, 4, H, W, mw, mh, cell) for e in range(E)]
Tm = eod.transform3d(torch.from_numpy(np.stack([ep.xyzhe for ep in eps]).reshape(-1, 5))).reshape(E, 4, 4, 4)
pose = Tm[:, :, :3, :].reshape(E, 4, 12).permute(1, 0, 2).contiguous().to(dev)
depth = torch.from_numpy(np.stack([ep.depth for ep in eps])).permute(1, 0, 2, 3).contiguous().to(dev)
shifts = torch.from_numpy(np.stack([np.concatenate([np.zeros(3, np.float32), ep.map_world_shift]) for ep in eps])).to(dev)
intr = eod.compute_intrinsics(W, H, math.radians(67.5))
out = {"E": E}
for C in (256, 512):
    idx = [ops.backproject_quantize(depth[t], pose[t], shifts, intr, cell, mw, mh)["idx"] for t in range(4)]
    table = (torch.randn((E, mw * mh, C), device=dev) * 3).half()
    levels = [torch.empty((E, H >> s, W >> s, C), dtype=torch.float16, device=dev) for s in (3, 4, 5)]
    for t in range(3):
        ops.read_pool(table, None, idx[t % 4], out=levels)
    torch.cuda.synchronize()
    evs = []
    for t in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.read_pool(table, None, idx[t % 4], out=levels); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = sorted(x.elapsed_time(y) for x, y in evs)
    out[f"C{C}_ms_median"] = ms[10]
    out[f"C{C}_checksum"] = float(levels[0].float().abs().sum())
    del table, levels
print(json.dumps(out))
