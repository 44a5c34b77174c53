"""Adversarial sweep of the quantisation (build_memory_data.py:135-143 semantics) - world coordinates that land within an ulp of a
half-cell boundary, negative / huge / out-of-map coordinates, several cell sizes - GPU (eod_quantize_world) against the torch-CPU
restatement (true division, round-half-even, clip), and the back-projection kernels (vec4 and scalar) against the C oracle on
random poses incl. depth 0 and far depths."""
import importlib, math, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from oracle import reference_ops as R
eod = importlib.import_module("embodied-object-detection_b200")
dev = torch.device("cuda:0")
rng = np.random.default_rng(11)
bad = 0
for cell in (0.2, 0.02, 0.05, 0.3):
    for mw, mh in ((500, 500), (1000, 731), (37, 91)):
        k = rng.integers(-50, max(mw, mh) + 50, (200000, 3)).astype(np.float64)
        half = (k + 0.5) * np.float32(cell).astype(np.float64)                     # on the half-cell boundaries (before fp32 rounding)
        world = half.astype(np.float32)
        world[::3] = np.nextafter(world[::3], np.float32(np.inf))                  # one ulp above
        world[1::3] = np.nextafter(world[1::3], np.float32(-np.inf))               # one ulp below
        world[:50] = [[1e30, 0, -1e30]]                                            # far outside: clipped
        shift = rng.uniform(-3, 3, 3).astype(np.float32)
        ref = R.quantize_flat_index(torch.from_numpy(world + shift).reshape(1, 1, -1, 3), torch.from_numpy(shift), cell, mw, mh).numpy().reshape(-1)
        got = eod.ops.quantize_world(torch.from_numpy(world + shift).to(dev), shift, cell, mw, mh).cpu().numpy().reshape(-1)
        d = int((got != ref).sum())
        if d:
            bad += 1
            print("MISMATCH quantize cell", cell, mw, mh, d)
intr_cache = {}
for case in range(12):
    H, W = (96, 128) if case % 2 == 0 else (67, 93)                                # vec4 kernel / scalar kernel (W % 4 != 0)
    mw, mh, cell = 300, 200, np.float32(0.1)
    depth = rng.uniform(0.0, 12.0, (H, W)).astype(np.float32)
    depth[rng.uniform(size=(H, W)) < 0.1] = 0.0
    depth[0, :8] = [1e-30, 1e6, 65504.0, 0.3, 10.0, np.inf, np.nan, 3e38]
    xyzhe = np.array([[rng.uniform(-5, 5), 1.25, rng.uniform(-5, 5), rng.uniform(0, 6.28), math.pi + rng.uniform(-0.3, 0.3)]], np.float32)
    T = eod.transform3d(torch.from_numpy(xyzhe))
    intr = eod.compute_intrinsics(W, H, math.radians(67.5))
    s0, s1 = rng.uniform(-1, 1, 3).astype(np.float32), rng.uniform(-10, 0, 3).astype(np.float32)
    ref = oracle.backproject_quantize(depth, T[0].numpy(), intr, s0, s1, cell, mw, mh, case % 2, 0.5)
    got = eod.ops.backproject_quantize(torch.from_numpy(depth[None]).to(dev), T[:, :3].reshape(1, 12).to(dev),
                                       torch.from_numpy(np.concatenate([s0, s1])[None]).to(dev), intr, float(cell), mw, mh, case % 2, 0.5,
                                       want_q2=True, want_outlier=True, want_height=True, want_world=True)
    finite = np.isfinite(ref["world"]).all(-1) & (np.abs(ref["world"]).max(-1) < 1e8)     # q2 is int32: only defined while |q| < 2^31
    for key in ("idx", "q2", "outlier", "height", "world"):
        a, b = got[key][0].cpu().numpy(), ref[key]
        if key == "q2":
            a, b = a[finite], b[finite]
        if a.dtype == np.float32:                                                          # NaN payload / sign may differ between CPU and GPU
            same = np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)].view(np.uint32), b[~np.isnan(b)].view(np.uint32))
        else:
            same = np.array_equal(a, b)                                                    # idx and outlier: strict, non-finite depths included
        if not same:
            bad += 1
            print("MISMATCH backproject case", case, key)
    # the C oracle itself against the torch-CPU restatement of Projector.forward + the flat-index lines (non-finite depths included)
    vf = math.radians(67.5)
    q2_t, out_t, h_t = R.projector_forward(torch.from_numpy(depth[None, None]), T, vf, mh, mw, float(cell), torch.from_numpy(s0 + s1), 0.5)
    ref2 = oracle.backproject_quantize(depth, T[0].numpy(), intr, s0 + s1, np.zeros(3, np.float32), cell, mw, mh, 0, 0.5)
    fin = np.isfinite(ref2["world"]).all(-1)
    if not np.array_equal(out_t[0].numpy(), ref2["outlier"].astype(bool)):
        bad += 1
        print("MISMATCH oracle-vs-torch outlier case", case, int((out_t[0].numpy() != ref2["outlier"].astype(bool)).sum()))
    if not np.array_equal(q2_t[0].numpy()[fin & (np.abs(ref2["world"]).max(-1) < 1e8)], ref2["q2"][fin & (np.abs(ref2["world"]).max(-1) < 1e8)]):
        bad += 1
        print("MISMATCH oracle-vs-torch q2 case", case)
    world_t = torch.from_numpy(ref2["world"])[None]
    flat_t = R.quantize_flat_index(world_t, torch.zeros(3), float(cell), mw, mh).numpy().reshape(H, W)
    if not np.array_equal(flat_t, ref2["idx"]):
        bad += 1
        print("MISMATCH oracle-vs-torch flat idx case", case, int((flat_t != ref2["idx"]).sum()))
print("geometry stress:", bad, "mismatches")
sys.exit(1 if bad else 0)
