"""Adversarial sweep of the dense write (all kernels: TMA ring, LDG, channels-last fp32 / bf16, deterministic) against the C oracle:
run structures random data does not produce (one cell for the whole frame, a new cell at every pixel, cells that change exactly at
the 32-pixel tile boundaries, two interleaved cells), sample masks with isolated pixels / empty tiles / only tile edges, feature
magnitudes from 1e-18 to 1e18 in neighbouring channels.  Per-cell means within 1e-5 of the per-channel feature scale; touched sets, per-cell sample
counts (through the means) and visibility counts exact."""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
eod = importlib.import_module("embodied-object-detection_b200")
ops, L = eod.ops, eod._lib
dev = torch.device("cuda:0")
rng = np.random.default_rng(77)
H, W, cells, E = 64, 96, 97, 2
HW = H * W
bad = n = 0
pix = np.arange(HW)
patterns = {
    "one_cell": np.full(HW, 5),
    "every_pixel": pix % cells,
    "tile_boundaries": (pix // 32) % cells,
    "two_interleaved": np.where(pix % 2 == 0, 3, 4),
    "runs_of_31_33": (np.cumsum(np.where(pix % 64 < 31, 0, 1) != np.roll(np.where(pix % 64 < 31, 0, 1), 1)) % cells),
    "noise": rng.integers(0, cells, HW),
}
samps = {
    "all": None,
    "isolated": (pix % 37 == 0).astype(np.uint8),
    "tile_edges": ((pix % 32 == 0) | (pix % 32 == 31)).astype(np.uint8),
    "first_half_empty": (pix >= HW // 2).astype(np.uint8),
    "none": np.zeros(HW, np.uint8),
}
configs = [("tma", L.LAYOUT_CHW, L.WRITE_TMA), ("ldg", L.LAYOUT_CHW, L.WRITE_LDG), ("hwc", L.LAYOUT_HWC, 0), ("hwc_bf16", L.LAYOUT_HWC_BF16, 0), ("det", L.LAYOUT_CHW, L.WRITE_DET)]
for C in (128, 256):
    scale = np.where(np.arange(C) % 3 == 0, 1e-18, np.where(np.arange(C) % 3 == 1, 1.0, 1e18)).astype(np.float32)
    for pname, pat in patterns.items():
        for sname, sm in samps.items():
            feat = (rng.standard_normal((E, C, HW)) * scale[None, :, None]).astype(np.float32)
            idx = np.stack([pat, np.roll(pat, 7)]).astype(np.int32)
            samp = None if sm is None else np.stack([sm, np.roll(sm, 3)])
            for name, layout, variant in configs:
                f = feat
                if layout == L.LAYOUT_HWC_BF16:
                    f = torch.from_numpy(feat).to(torch.bfloat16).float().numpy()
                d_idx = torch.from_numpy(idx).to(dev).view(E, H, W)
                d_samp = None if samp is None else torch.from_numpy(samp).to(dev)
                d_cnt = torch.zeros((E, cells), dtype=torch.int32, device=dev)
                d_sums = torch.zeros((E, cells, C), device=dev)
                d_counts = torch.zeros((E, cells), device=dev)
                d_touched = torch.zeros((E, cells), dtype=torch.uint8, device=dev)
                ops.frame_count(d_idx, d_samp, d_cnt)
                if variant == L.WRITE_DET:
                    ws = ops.DetWorkspace(E, C, HW, cells, dev, HW)
                    ops.write_mean_det(torch.from_numpy(f).to(dev), d_idx, d_samp, d_cnt, d_sums, ws)
                else:
                    ff = torch.from_numpy(f if layout == L.LAYOUT_CHW else np.ascontiguousarray(f.transpose(0, 2, 1))).to(dev)
                    if layout == L.LAYOUT_HWC_BF16:
                        ff = ff.to(torch.bfloat16)
                    ops.write_mean(ff, d_idx, d_samp, d_cnt, d_sums, layout, variant)
                ops.finalize_counts(d_idx, d_cnt, d_counts, d_touched)
                torch.cuda.synchronize()
                n += 1
                for e in range(E):
                    s_, cnt = oracle.cell_sums_seq(f[e].reshape(C, H, W), idx[e].reshape(H, W), None if samp is None else samp[e].reshape(H, W), cells)
                    ref = np.where(cnt[:, None] > 0, s_ / np.maximum(cnt, 1)[:, None].astype(np.float32), 0).astype(np.float32)
                    got = d_sums[e].cpu().numpy()
                    err = np.abs(got - ref).max(0) / np.maximum(np.abs(f[e]).max(1), 1e-30)             # per channel, relative to the FEATURE scale (the scales differ by 1e36)
                    vis = np.zeros(cells, np.float32); vis[np.unique(idx[e])] = 1
                    ok = err.max() <= 1e-5 and np.array_equal(d_touched[e].cpu().numpy().astype(bool), cnt > 0) and \
                        np.array_equal(d_counts[e].cpu().numpy(), vis) and not got[cnt == 0].any() and int(d_cnt.abs().sum()) == 0
                    if not ok:
                        bad += 1
                        print("MISMATCH", name, "C", C, pname, sname, "episode", e, "max rel err", float(err.max()))
print("dense write stress:", n, "cases,", bad, "mismatches")
sys.exit(1 if bad else 0)
