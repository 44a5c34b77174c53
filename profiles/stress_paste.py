"""Adversarial sweep of the mask-pasting kernels against oracle/paste.c (which tests/ pin to torch-CPU grid_sample): masks that sit
exactly on the threshold (constant 0.5, {0, 0.5, 1} lattices, +-1e-5 around 0.5), NaN / +-Inf probabilities, boxes with integer and
half-integer corners, sub-pixel boxes, boxes larger than or outside the image, odd mask sizes, thresholds 0.5 and 0.25.  Checks
eod_paste_masks (masks + observed) and the pasted object write's touched-cell set against the two-step path."""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
eod = importlib.import_module("embodied-object-detection_b200")
dev = torch.device("cuda:0")
rng = np.random.default_rng(1)
H, W = 120, 160
bad = n = 0
for case in range(int(os.environ.get("CASES", 60))):
    K = 6
    S = int(rng.choice([28, 14, 7, 28, 28]))
    kind = case % 6
    if kind == 0:   probs = np.full((K, S, S), 0.5, np.float32)
    elif kind == 1: probs = rng.choice([0.0, 0.5, 1.0], (K, S, S)).astype(np.float32)
    elif kind == 2: probs = (rng.integers(0, 3, (K, S, S)) * 0.25 + 0.25).astype(np.float32)
    elif kind == 3: probs = rng.uniform(0.49999, 0.50001, (K, S, S)).astype(np.float32)
    elif kind == 4: probs = rng.uniform(0, 1, (K, S, S)).astype(np.float32)
    else:
        probs = rng.uniform(0, 1, (K, S, S)).astype(np.float32)
        probs[0, 3, 3] = np.nan; probs[1, 2, 2] = np.inf; probs[2, 1, 1] = -np.inf
    boxes = np.zeros((K, 4), np.float32)
    for k in range(K):
        t = rng.integers(0, 5)
        if t == 0:   x0, y0 = rng.integers(0, W - 30), rng.integers(0, H - 30); boxes[k] = (x0, y0, x0 + rng.integers(1, 30), y0 + rng.integers(1, 30))
        elif t == 1: x0, y0 = rng.integers(0, W - 30) + 0.5, rng.integers(0, H - 30) + 0.5; boxes[k] = (x0, y0, x0 + 28, y0 + 28)
        elif t == 2: x0, y0 = rng.uniform(-20, W + 20), rng.uniform(-20, H + 20); boxes[k] = (x0, y0, x0 + rng.uniform(0.01, 3), y0 + rng.uniform(0.01, 3))
        elif t == 3: boxes[k] = (rng.uniform(-50, 0), rng.uniform(-50, 0), W + rng.uniform(0, 50), H + rng.uniform(0, 50))
        else:        x0, y0 = rng.uniform(0, W - 40), rng.uniform(0, H - 40); boxes[k] = (x0, y0, x0 + rng.uniform(5, 40), y0 + rng.uniform(5, 40))
    d_p, d_b = torch.from_numpy(probs[None]).to(dev), torch.from_numpy(boxes[None]).to(dev)
    for thr in (0.5, 0.25):
        ref = oracle.paste_masks(probs, boxes, H, W, thr)
        masks, observed = eod.ops.paste_masks(d_p, d_b, (H, W), thr, want_observed=True)
        got = masks[0].cpu().numpy()
        n += 1
        d = int((got != ref).sum()) + int((observed[0].cpu().numpy().astype(bool) != ref.any(0).reshape(-1)).sum())
        if d:
            bad += 1
            print("MISMATCH case", case, "kind", kind, "S", S, "thr", thr, "diff", d)
    # pasted write == paste then write (touched-cell sets; the sums are unordered fp32 reductions)
    C, mw, mh = 128, 16, 12
    idx = torch.from_numpy(rng.integers(0, mw * mh, (1, H // 8, W // 8)).repeat(8, 1).repeat(8, 2).astype(np.int32)).to(dev)
    bf = torch.from_numpy(rng.standard_normal((1, K, C)).astype(np.float32)).to(dev)
    a = eod.EpisodeBatch(1, mw, mh, C, H, W, dev); b = eod.EpisodeBatch(1, mw, mh, C, H, W, dev)
    a.set_indices(idx); b.set_indices(idx)
    a.write_detections(bf, d_p, d_b)
    b.write_objects(bf, eod.ops.paste_masks(d_p, d_b, (H, W), 0.5)[0])
    torch.cuda.synchronize()
    # NaN / Inf probabilities only decide cover bits; features are finite
    if not (torch.equal(a.sums == 0, b.sums == 0) and torch.equal(a.counts, b.counts)):
        bad += 1
        print("MISMATCH write case", case)
print("paste stress:", n, "cases,", bad, "mismatches")
sys.exit(1 if bad else 0)
