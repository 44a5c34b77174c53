"""Small, ncu-friendly run of the hot path: E episodes x T frame-steps of the bench workload shape
(480x640, C=256, 500x500 grid).  Used for `ncu --set full` captures; bench.py is the timing authority."""
import importlib, math, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
eod = importlib.import_module("embodied-object-detection_b200")

E = int(os.environ.get("PROF_E", 16)); T = int(os.environ.get("PROF_T", 3)); C = int(os.environ.get("PROF_C", 256))
H, W, MW, MH = 480, 640, 500, 500
dev = torch.device("cuda:0")
eps = [eod.episodes.make_episode(1234 + e, T, H, W, MW, MH, 0.2) for e in range(E)]
intr = eod.compute_intrinsics(W, H, math.radians(67.5))
batch = eod.EpisodeBatch(E, MW, MH, C, H, W, dev, variant=int(os.environ.get('PROF_VARIANT', 0)))
shifts = torch.from_numpy(np.stack([np.concatenate([np.zeros(3, np.float32), ep.map_world_shift]) for ep in eps])).to(dev)
feat = [torch.randn((E, C, H, W), device=dev) for _ in range(2)]
for t in range(T):
    Tm = eod.transform3d(torch.from_numpy(np.stack([ep.xyzhe[t] for ep in eps])))
    depth = torch.from_numpy(np.stack([ep.depth[t] for ep in eps])).to(dev)
    batch.step(depth, Tm[:, :3].reshape(E, 12).to(dev), shifts, intr, 0.2, feat[t & 1])
torch.cuda.synchronize()
print("ok", float(batch.counts.sum()), float(batch.sums.abs().sum()))
