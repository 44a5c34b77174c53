"""Adversarial / randomised sweeps of the kernels against the CPU oracle (and, for the tcgen05 projection, kernel against
kernel): inputs random data never produces.  Each ``stress_*`` returns (cases run, mismatches, messages); tests/test_gpu_stress.py
runs them with bounded case counts under `-m gpu` (so the driver's GPU test tier carries them), profiles/stress.py runs larger
sweeps from the command line.  Test infrastructure: imports the oracle, never imported by the product."""
import importlib
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import reference_ops as R  # noqa: E402

eod = importlib.import_module("embodied-object-detection_b200")
ops, L = eod.ops, eod._lib

def stress_read(dev, cases):
    """Randomised sweep of the read kernels against the C oracle (bit-exact fp16 levels): image sizes, channel counts, batch sizes
and adversarial cell patterns (checkerboards = 16 runs per window, 1-px stripes both ways, noise, large uniform areas, int64
indices, fp32 table + counts).  One-off stress, complements tests/test_gpu_parity.py::test_read_pool_*."""
    msgs = []
    rng = np.random.default_rng(2026)
    n_cases = cases
    bad = 0
    for case in range(n_cases):
        C = int(rng.choice([128, 256, 512]))
        E = int(rng.integers(1, 4))
        H = 32 * int(rng.integers(1, 5)); W = 32 * int(rng.integers(1, 6))
        cells = int(rng.integers(3, 400))
        kind = case % 6
        yy, xx = np.mgrid[0:H, 0:W]
        if kind == 0:   idx = ((yy + xx) % 2) * (cells - 1)                                  # checkerboard: every window 16 runs
        elif kind == 1: idx = (xx % cells)                                                   # 1-px vertical stripes
        elif kind == 2: idx = (yy % cells)                                                   # 1-px horizontal stripes
        elif kind == 3: idx = rng.integers(0, cells, (H, W))                                 # noise
        elif kind == 4: idx = (yy // 24) * 7 % cells + (xx // 40) % 3                        # large blocks, edges off the 4/8/16 lattice
        else:           idx = np.where(rng.uniform(size=(H, W)) < 0.02, rng.integers(0, cells, (H, W)), (yy // 9 + xx // 13) % cells)
        idx = np.stack([np.roll(idx, e, axis=1) for e in range(E)]).astype(np.int32) % cells
        table = (rng.standard_normal((E, cells, C)) * rng.choice([1e-3, 1.0, 300.0])).astype(np.float16)
        table[:, 0, :8] = [0.0, -0.0, 6e-8, -6e-8, 65504.0, -65504.0, 1.0, -1.0]
        use_i64 = bool(case % 2)
        d_idx = torch.from_numpy(idx.astype(np.int64) if use_i64 else idx).to(dev)
        got = eod.ops.read_pool(torch.from_numpy(table).to(dev), None, d_idx)
        torch.cuda.synchronize()
        for e in range(E):
            ref = oracle.read_pool_f16(table[e], idx[e])
            for k in range(3):
                g = got[k][e].contiguous().cpu().numpy().view(np.uint16)
                r = ref[k].view(np.uint16)
                # +0 / -0 may differ in sign only where the reference's own sum order is sign-ambiguous? no: demand exact bits
                if not np.array_equal(g, r):
                    bad += 1
                    msgs.append(" ".join(str(x) for x in ("MISMATCH case", case, "kind", kind, "C", C, "E", E, H, W, "level", k, int((g != r).sum()),)))
        if case % 6 == 5:                                                                    # fp32 sums + counts path on the last kind
            sums = (rng.standard_normal((E, cells, C)) * 5).astype(np.float32)
            counts = rng.integers(0, 5, (E, cells)).astype(np.float32)
            got = eod.ops.read_pool(torch.from_numpy(sums).to(dev), torch.from_numpy(counts).to(dev), d_idx)
            for e in range(E):
                ref = R.read_frame(torch.from_numpy(sums[e]), torch.from_numpy(counts[e]), torch.from_numpy(idx[e]).long())
                for k in range(3):
                    if not np.array_equal(got[k][e].contiguous().cpu().numpy().view(np.uint16), ref[k][0].numpy().view(np.uint16)):
                        bad += 1
                        msgs.append(" ".join(str(x) for x in ("MISMATCH fp32-table case", case, "level", k,)))
    return n_cases, bad, msgs


def stress_paste(dev, cases):
    """Adversarial sweep of the mask-pasting kernels against oracle/paste.c (which tests/ pin to torch-CPU grid_sample): masks that sit
exactly on the threshold (constant 0.5, {0, 0.5, 1} lattices, +-1e-5 around 0.5), NaN / +-Inf probabilities, boxes with integer and
half-integer corners, sub-pixel boxes, boxes larger than or outside the image, odd mask sizes, thresholds 0.5 and 0.25.  Checks
eod_paste_masks (masks + observed) and the pasted object write's touched-cell set against the two-step path."""
    msgs = []
    rng = np.random.default_rng(1)
    H, W = 120, 160
    bad = n = 0
    for case in range(cases):
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
            # below 0.5 the library follows detectron2's CUDA path (whole image sampled), which the reference takes
            ref = oracle.paste_masks(probs, boxes, H, W, thr, skip_empty=thr >= 0.5)
            masks, observed = eod.ops.paste_masks(d_p, d_b, (H, W), thr, want_observed=True)
            got = masks[0].cpu().numpy()
            n += 1
            d = int((got != ref).sum()) + int((observed[0].cpu().numpy().astype(bool) != ref.any(0).reshape(-1)).sum())
            if d:
                bad += 1
                msgs.append(" ".join(str(x) for x in ("MISMATCH case", case, "kind", kind, "S", S, "thr", thr, "diff", d,)))
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
            msgs.append(" ".join(str(x) for x in ("MISMATCH write case", case,)))
    return n, bad, msgs


def stress_geometry(dev, cases):
    """Adversarial sweep of the quantisation (build_memory_data.py:135-143 semantics) - world coordinates that land within an ulp of a
half-cell boundary, negative / huge / out-of-map coordinates, several cell sizes - GPU (eod_quantize_world) against the torch-CPU
restatement (true division, round-half-even, clip), and the back-projection kernels (vec4 and scalar) against the C oracle on
random poses incl. depth 0 and far depths."""
    msgs = []
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
                msgs.append(" ".join(str(x) for x in ("MISMATCH quantize cell", cell, mw, mh, d,)))
    intr_cache = {}
    for case in range(cases):
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
                msgs.append(" ".join(str(x) for x in ("MISMATCH backproject case", case, key,)))
        # the C oracle itself against the torch-CPU restatement of Projector.forward + the flat-index lines (non-finite depths included)
        vf = math.radians(67.5)
        q2_t, out_t, h_t = R.projector_forward(torch.from_numpy(depth[None, None]), T, vf, mh, mw, float(cell), torch.from_numpy(s0 + s1), 0.5)
        ref2 = oracle.backproject_quantize(depth, T[0].numpy(), intr, s0 + s1, np.zeros(3, np.float32), cell, mw, mh, 0, 0.5)
        fin = np.isfinite(ref2["world"]).all(-1)
        if not np.array_equal(out_t[0].numpy(), ref2["outlier"].astype(bool)):
            bad += 1
            msgs.append(" ".join(str(x) for x in ("MISMATCH oracle-vs-torch outlier case", case, int((out_t[0].numpy() != ref2["outlier"].astype(bool)).sum()),)))
        if not np.array_equal(q2_t[0].numpy()[fin & (np.abs(ref2["world"]).max(-1) < 1e8)], ref2["q2"][fin & (np.abs(ref2["world"]).max(-1) < 1e8)]):
            bad += 1
            msgs.append(" ".join(str(x) for x in ("MISMATCH oracle-vs-torch q2 case", case,)))
        world_t = torch.from_numpy(ref2["world"])[None]
        flat_t = R.quantize_flat_index(world_t, torch.zeros(3), float(cell), mw, mh).numpy().reshape(H, W)
        if not np.array_equal(flat_t, ref2["idx"]):
            bad += 1
            msgs.append(" ".join(str(x) for x in ("MISMATCH oracle-vs-torch flat idx case", case, int((flat_t != ref2["idx"]).sum()),)))
    return 12 + cases, bad, msgs


def stress_dense_write(dev, cases):
    """Adversarial sweep of the dense write (all kernels: TMA ring, LDG, channels-last fp32 / bf16, deterministic) against the C oracle:
run structures random data does not produce (one cell for the whole frame, a new cell at every pixel, cells that change exactly at
the 32-pixel tile boundaries, two interleaved cells), sample masks with isolated pixels / empty tiles / only tile edges, feature
magnitudes from 1e-18 to 1e18 in neighbouring channels.  Per-cell means within 1e-5 of the per-channel feature scale; touched sets, per-cell sample
counts (through the means) and visibility counts exact."""
    msgs = []
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
    combos = [(C, pname, sname) for C in (128, 256) for pname in patterns for sname in samps]
    combos = combos[:: max(1, len(combos) * len(configs) // max(cases, 1))]          # `cases` bounds the number of (kernel, input) cases
    for C, pname, sname in combos:
        scale = np.where(np.arange(C) % 3 == 0, 1e-18, np.where(np.arange(C) % 3 == 1, 1.0, 1e18)).astype(np.float32)
        pat, sm = patterns[pname], samps[sname]
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
                    msgs.append(" ".join(str(x) for x in ("MISMATCH", name, "C", C, pname, sname, "episode", e, "max rel err", float(err.max()),)))
    return n, bad, msgs


def stress_objects(dev, cases):
    """Adversarial sweep of the object-regime write (byte-mask and pasted variants) against the oracle chain paste -> box_to_image_features
-> project_image_features (sparse restatement) -> accumulate: every object on the same box (3, 5, 6, 7 ... objects per pixel: the
non-power-of-two divides), up to 140 objects (beyond the 128-object bitmask path), single-pixel and full-image masks, every sampled
pixel in one cell / in its own cell, frames without detections.  Touched sets and visibility counts exact, sums within 1e-5 of scale."""
    msgs = []
    rng = np.random.default_rng(4)
    H, W, C, mw, mh = 64, 96, 128, 13, 11
    cells = mw * mh
    bad = n = 0
    for case in range(cases):
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
    return n, bad, msgs


def stress_fuse(dev, cases):
    """Stress of the persistent projection+fusion kernel against the tile-per-CTA kernel (bit-identical tile arithmetic expected):
many tiles per CTA, all level combinations, bias on/off, sum / mem_only.  Found the res-ring release race of r3."""
    msgs = []
    rng = np.random.default_rng(99)
    E, K, N = 24, 512, 256
    n_bad_total = [0]
    def run(shapes, use_bias, mode):
        lv = [torch.from_numpy((rng.standard_normal((E, h, w, K)) * 2).astype(np.float16)).to(dev) for h, w in shapes]
        Ws = [(rng.uniform(-1, 1, (N, K)) / math.sqrt(K)).astype(np.float32) for _ in shapes]
        ws = [ops.project_split_weights(torch.from_numpy(W).to(dev)) for W in Ws]
        bs = [torch.from_numpy(rng.standard_normal(N).astype(np.float32)).to(dev) if use_bias else None for _ in shapes]
        rs = [torch.from_numpy(rng.standard_normal((E, N, h, w)).astype(np.float32)).to(dev) for h, w in shapes]
        ref = ops.project_fuse_levels(lv, ws, bs, rs if mode == 0 else None, 1.0, mode, variant=1)
        got = ops.project_fuse_levels(lv, ws, bs, rs if mode == 0 else None, 1.0, mode, variant=2)
        for k, (h, w) in enumerate(shapes):
            d = (got[k] - ref[k]).abs()
            bad = (d > 1e-3).nonzero()
            msgs.append(" ".join(str(x) for x in (shapes, "bias", use_bias, "mode", mode, "level", k, "max diff", d.max().item(), "n bad", bad.shape[0], "of", d.numel(),)))
            n_bad_total[0] += int(bad.shape[0])
            if bad.shape[0]:
                e = bad[:, 0]; n = bad[:, 1]; pix = bad[:, 2] * w + bad[:, 3]
                tpe = -(-h * w // 256)
                msgs.append("   episodes " + str(torch.unique(e).tolist()[:10]))
                print("   episodes", torch.unique(e).tolist()[:10], "nblocks", torch.unique(n // 128).tolist(), "ptiles", torch.unique(pix // 256).tolist()[:20], "of", tpe,
                      "chan%128", (n % 128).min().item(), (n % 128).max().item(), "pix%256", (pix % 256).min().item(), (pix % 256).max().item())
                b0 = bad[0]; print("   sample", b0.tolist(), got[k][tuple(b0)].item(), ref[k][tuple(b0)].item())
    for shapes in ([(60, 80)], [(30, 40)], [(15, 20)], [(60, 80), (30, 40)], [(60, 80), (30, 40), (15, 20)]):
        for use_bias in (False, True):
            for mode in (1, 0):
                run(shapes, use_bias, mode)
    return len(msgs), n_bad_total[0], msgs




def stress_linear(dev, cases):
    """Randomised sweep of the tcgen05 row GEMM (eod_linear_rows, 3xTF32) against fp64: shapes from one row to several tiles, K with and
without a 32-wide tail, N over one and two TMEM column blocks, every stride combination (contiguous / transposed / padded / unaligned
operands), gathered input rows, scattered output rows with a device-side row count, rows of very different scale.  |error| <= 1e-5 of
the result's scale, and rows that must not be written keep their fill."""
    msgs = []
    rng = np.random.default_rng(515)
    bad = 0
    for case in range(cases):
        M = int(rng.choice([1, 2, 31, 127, 128, 129, 300, 700]))
        K = int(rng.choice([4, 20, 32, 36, 64, 96, 256, 500, 512]))
        N = 16 * int(rng.integers(1, 33))
        g = torch.Generator(device=dev).manual_seed(case)
        a = torch.randn((M, K), device=dev, generator=g) * torch.rand((M, 1), device=dev, generator=g).mul(8).sub(4).exp()
        w = torch.randn((N, K), device=dev, generator=g) / K ** 0.5
        b = torch.randn((N,), device=dev, generator=g) if case % 3 else None
        scale_f = float(rng.choice([1.0, 5.0, 0.125]))
        ka, kw = case % 4, (case // 4) % 3
        aa = a if ka == 0 else (a.t().contiguous().t() if ka == 1 else torch.zeros((M, K + 7), device=dev)[:, 5:5 + K].copy_(a) if ka == 2 else a)
        ww = w if kw == 0 else (w.t().contiguous().t() if kw == 1 else torch.zeros((N, K + 12), device=dev)[:, 3:3 + K].copy_(w))
        ref = (a.double() @ w.double().t() + (b.double() if b is not None else 0.0)) * scale_f
        tol = 1e-5 * float(ref.abs().max())
        if ka == 3:                                                         # gather + scatter + device-side count
            perm = torch.randperm(M, device=dev, generator=g)
            n_live = int(rng.integers(1, M + 1))
            out = torch.full((M + 4, N + 4), 3.0, device=dev)
            ops.linear_rows(a, ww, b, scale_f, out=out, a_off=(perm * K).to(torch.int64), m_count=torch.tensor([n_live], dtype=torch.int32, device=dev),
                            n_rows=M, y_dst=(perm + 2).to(torch.int64))
            live = perm[:n_live]
            err = float((out[live + 2][:, :N].double() - ref[live]).abs().max())
            keep = torch.ones(M + 4, dtype=torch.bool, device=dev)
            keep[live + 2] = False
            clean = bool((out[keep] == 3.0).all()) and bool((out[:, N:] == 3.0).all())
        else:
            got = ops.linear_rows(aa, ww, b, scale_f)
            err = float((got.double() - ref).abs().max())
            clean = True
        ok = err <= tol and clean
        bad += int(not ok)
        msgs.append(f"{'ok' if ok else 'MISMATCH'} linear case {case}: M={M} K={K} N={N} a-kind {ka} w-kind {kw} bias {b is not None} err/scale {err / max(tol * 1e5, 1e-30):.2e} clean {clean}")
    return cases, bad, msgs
