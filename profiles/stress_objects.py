"""Adversarial sweep of the object-regime write (byte-mask and pasted variants) against the oracle chain paste -> box_to_image_features
-> project_image_features (sparse restatement) -> accumulate: every object on the same box (3, 5, 6, 7 ... objects per pixel: the
non-power-of-two divides), up to 140 objects (beyond the 128-object bitmask path), single-pixel and full-image masks, every sampled
pixel in one cell / in its own cell, frames without detections.  Touched sets and visibility counts exact, sums within 1e-5 of scale."""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from oracle import reference_ops as R
eod = importlib.import_module("embodied-object-detection_b200")
dev = torch.device("cuda:0")
rng = np.random.default_rng(4)
H, W, C, mw, mh = 64, 96, 128, 13, 11
cells = mw * mh
bad = n = 0
for case in range(int(os.environ.get("CASES", 30))):
    kind = case % 6
    K = [3, 7, 140, 16, 5, 1][kind]
    f = rng.standard_normal((K, C)).astype(np.float32)
    f = (50.0 * f / np.linalg.norm(f, axis=1, keepdims=True)).astype(np.float32)
    probs = rng.uniform(0.3, 1.0, (K, 28, 28)).astype(np.float32)
    if kind in (0, 1, 4):                      # all objects on (almost) the same box: K objects per pixel
        b0 = np.array([10.3, 8.7, 70.2, 50.9], np.float32)
        boxes = np.tile(b0, (K, 1)) + rng.uniform(-0.5, 0.5, (K, 4)).astype(np.float32)
        probs[:] = 1.0
    elif kind == 2:
        x0, y0 = rng.uniform(0, W - 20, K), rng.uniform(0, H - 20, K)
        boxes = np.stack([x0, y0, x0 + rng.uniform(4, 20, K), y0 + rng.uniform(4, 20, K)], 1).astype(np.float32)
    elif kind == 3:
        boxes = np.tile(np.array([0, 0, W, H], np.float32), (K, 1)); probs[:] = 1.0    # full-image masks
        boxes[::2] = [[40.2, 30.2, 41.4, 31.4]]                                       # and ~single-pixel ones
    else:
        boxes = np.array([[20.5, 10.5, 60.5, 40.5]], np.float32)
    idx_kind = case % 3
    yy, xx = np.mgrid[0:H, 0:W]
    idx = [np.full((H, W), 7), (yy * W + xx) % cells, (yy // 8) * mw % cells + xx // 8][idx_kind].astype(np.int32)
    masks = oracle.paste_masks(probs, boxes, H, W, 0.5)
    img, obs = R.box_to_image_features(torch.from_numpy(f), torch.from_numpy(masks))
    sums0 = torch.from_numpy(rng.standard_normal((cells, C)).astype(np.float32))
    counts0 = torch.from_numpy(rng.integers(0, 3, cells).astype(np.float32))
    if obs.any():
        ref_s, ref_c = R.write_mean_frame(sums0.clone(), counts0.clone(), img, obs, torch.from_numpy(idx).long(), stride=8)
    else:
        ref_s, ref_c = sums0.clone(), counts0.clone()
    for pasted in (False, True):
        batch = eod.EpisodeBatch(1, mw, mh, C, H, W, dev)
        batch.sums.copy_(sums0[None]); batch.counts.copy_(counts0[None])
        batch.set_indices(torch.from_numpy(idx[None]).to(dev))
        n_obj = torch.tensor([K if obs.any() else 0], dtype=torch.int32, device=dev)
        if pasted:
            batch.write_detections(torch.from_numpy(f[None]).to(dev), torch.from_numpy(probs[None]).to(dev), torch.from_numpy(boxes[None]).to(dev), n_obj)
        else:
            batch.write_objects(torch.from_numpy(f[None]).to(dev), torch.from_numpy(masks[None]).to(dev), n_obj)
        torch.cuda.synchronize()
        n += 1
        got_s, got_c = batch.sums[0].cpu(), batch.counts[0].cpu()
        changed_ref, changed_got = (ref_s != sums0).any(1), (got_s != sums0).any(1)
        ok = (got_s - ref_s).abs().max().item() <= 1e-5 * ref_s.abs().max().item() and torch.equal(got_c, ref_c) and torch.equal(changed_ref, changed_got)
        if not ok:
            bad += 1
            print("MISMATCH case", case, "kind", kind, "K", K, "idx", idx_kind, "pasted", pasted, (got_s - ref_s).abs().max().item() / ref_s.abs().max().item(),
                  int((changed_ref != changed_got).sum()), bool(torch.equal(got_c, ref_c)))
print("objects stress:", n, "cases,", bad, "mismatches")
sys.exit(1 if bad else 0)
