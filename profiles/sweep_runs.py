"""Run-length sweep of the dense mean write: how the write launch behaves when the runs of equal cell id get
short - finer map cells (0.2 / 0.1 / 0.05 / 0.02 m) and depth noise - i.e. when one L2 reduction per run and channel block stops
being negligible next to the feature stream.  E episodes x 480x640, C=256, fp32 CHW (TMA kernel) and HWC; a 2000x2000 grid so that
0.02 m cells are not all clipped to the border; CUDA events around 10 write launches alone (count / expand / finalize outside).
Output: one JSON line per configuration with the mean run length inside 32-px tiles, runs per frame, launch ms, feature GB/s and
the number of red.global.add.f32 warp instructions (4 per run and channel block)."""
import importlib, json, math, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
eod = importlib.import_module("embodied-object-detection_b200")
ops = eod.ops
dev = torch.device("cuda:0")
H, W, C = 480, 640, 256
E = int(os.environ.get("SWEEP_E", 8))
MW = int(os.environ.get("SWEEP_MAP", 2000))
N = int(os.environ.get("SWEEP_N", 10))
only = os.environ.get("SWEEP_ONLY")
eps = [eod.episodes.make_episode(1234 + e, 2, H, W, 500, 500, 0.2) for e in range(E)]
Tm = eod.transform3d(torch.from_numpy(np.stack([ep.xyzhe for ep in eps]).reshape(-1, 5))).reshape(E, 2, 4, 4)
pose = Tm[:, 1, :3, :].reshape(E, 12).contiguous().to(dev)
depth0 = torch.from_numpy(np.stack([ep.depth[1] for ep in eps])).to(dev)
intr = eod.compute_intrinsics(W, H, math.radians(67.5))
gen = torch.Generator(device=dev).manual_seed(3)
feat_chw = [torch.randn((E, C, H, W), device=dev, generator=gen) for _ in range(2)]
feat_hwc = [f.permute(0, 2, 3, 1).contiguous() for f in feat_chw]
peak = 6540.8
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
rows = []
for cell in (0.2, 0.1, 0.05, 0.02):
    for noise in (0.0, 0.01, 0.03):
        tag = f"cell{cell}_noise{noise}"
        if only and tag != only:
            continue
        # the episode's world origin sits in the middle of the (larger) map
        world_min = np.stack([ep.map_world_shift for ep in eps])                      # (E,3) of the 500x500 @0.2 map
        centre = world_min + np.array([50.0, 0.0, 50.0], np.float32)
        shift = centre - np.array([MW * cell / 2, 0.0, MW * cell / 2], np.float32)
        shifts = torch.from_numpy(np.concatenate([np.zeros((E, 3), np.float32), shift.astype(np.float32)], 1)).to(dev)
        depth = depth0 if noise == 0.0 else (depth0 + noise * torch.randn(depth0.shape, device=dev, generator=gen)).clamp_min(0.05)
        for layout, feats in ((eod._lib.LAYOUT_CHW, feat_chw), (eod._lib.LAYOUT_HWC, feat_hwc)):
            batch = eod.EpisodeBatch(E, MW, MW, C, H, W, dev, layout=layout)
            batch.project(depth, pose, shifts, intr, cell)
            idx = batch.idx.reshape(E, -1, 32)
            heads = int((idx[:, :, 1:] != idx[:, :, :-1]).sum().item()) + idx.shape[0] * idx.shape[1]      # runs inside 32-px tiles
            clipped = float(((batch.idx % MW == 0) | (batch.idx % MW == MW - 1) | (batch.idx < MW) | (batch.idx >= MW * (MW - 1))).float().mean())
            batch._count(None)
            batch._write(feats[1], None)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(N):
                batch._write(feats[i & 1], None)
            b.record()
            torch.cuda.synchronize()
            batch._finalize()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / N
            touched = int((batch.counts > 0).sum().item())
            fbytes = E * (H * W * C * 4 + H * W * 4)
            row = {"cell_m": cell, "depth_noise_m": noise, "layout": "chw_tma" if layout == eod._lib.LAYOUT_CHW else "hwc", "E": E,
                   "runs_per_frame": heads / E, "mean_run_px": H * W * E / heads, "cells_per_frame": touched / E, "border_px_frac": clipped,
                   "launch_ms": ms, "feature_GBps": fbytes / ms / 1e6, "frac_of_peak": fbytes / ms / 1e6 / peak,
                   "red_warp_instr_per_launch": heads * (C // 32), "red_bytes_per_launch": heads * C * 4}
            rows.append(row)
            print(json.dumps(row), flush=True)
            del batch
            torch.cuda.empty_cache()
