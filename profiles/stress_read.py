"""Randomised sweep of the read kernels against the C oracle (bit-exact fp16 levels): image sizes, channel counts, batch sizes
and adversarial cell patterns (checkerboards = 16 runs per window, 1-px stripes both ways, noise, large uniform areas, int64
indices, fp32 table + counts).  One-off stress, complements tests/test_gpu_parity.py::test_read_pool_*."""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from oracle import reference_ops as R
eod = importlib.import_module("embodied-object-detection_b200")
dev = torch.device("cuda:0")
rng = np.random.default_rng(2026)
n_cases = int(os.environ.get("CASES", 36))
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
                print("MISMATCH case", case, "kind", kind, "C", C, "E", E, H, W, "level", k, int((g != r).sum()))
    if case % 6 == 5:                                                                    # fp32 sums + counts path on the last kind
        sums = (rng.standard_normal((E, cells, C)) * 5).astype(np.float32)
        counts = rng.integers(0, 5, (E, cells)).astype(np.float32)
        got = eod.ops.read_pool(torch.from_numpy(sums).to(dev), torch.from_numpy(counts).to(dev), d_idx)
        for e in range(E):
            ref = R.read_frame(torch.from_numpy(sums[e]), torch.from_numpy(counts[e]), torch.from_numpy(idx[e]).long())
            for k in range(3):
                if not np.array_equal(got[k][e].contiguous().cpu().numpy().view(np.uint16), ref[k][0].numpy().view(np.uint16)):
                    bad += 1
                    print("MISMATCH fp32-table case", case, "level", k)
print("read stress:", n_cases, "cases,", bad, "mismatches")
sys.exit(1 if bad else 0)
