"""ORACLE-side checker of the benchmarked path (test infrastructure, NOT product code).

``check_dense_steps`` drives an ``EpisodeBatch`` through a few frames of the dense regime - the exact configuration
``bench.py`` times (E episodes in lock step, ``pipeline=True``, 480x640, CHW fp32 features) - and compares a few of
its episodes against the CPU restatement of the reference, frame by frame, in the reference's order
(custom_rcnn.py:489-515: the read of frame t sees the state left by frame t-1):

  * cell indices                      bit-exact   (oracle/geometry.c, SMNet/projector/core.py + build_memory_data.py:135-143)
  * pooled fp16 levels                bit-exact   (timm.py:147-168 on the state DOWNLOADED before the frame)
  * visibility counts                 bit-exact   (custom_rcnn.py:699-701,743)
  * per-cell fp32 sums                max|a-b| <= 1e-5 * max|ref|   (custom_rcnn.py:917-934, 696-697; summation order unpinned)
  * rows of never-visible cells       all zero

The oracle works on a COMPACT grid: the cells an episode ever sees are renumbered 0..U-1 in ascending order (a bijection
that none of the restated ops can observe: they only gather, group and scatter by cell id), so a 1000x1000x256 grid
costs the CPU megabytes, not gigabytes.

Used by tests/test_gpu_bench_path.py and, outside the timed region, by bench.py's ``parity_check`` block.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

from . import backproject_quantize
from . import reference_ops as R

SUM_TOL = 1e-5


def _u16(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().cpu().numpy().view(np.uint16)


def check_dense_steps(batch, frames: Sequence[dict], shifts_dev, intr, cell: float, host: Dict[int, dict], *, synced: bool = True,
                      reset: bool = True) -> dict:
    """batch: an EpisodeBatch.  frames[t] = dict(depth=(E,H,W) dev, pose=(E,12) dev, feat=(E,C,H,W) dev, or a callable
    returning it - evaluated on the caller's stream right before the step, so that frames can share feature slabs).
    host[e] = dict(depth=(T,H,W) np f32, T=(T,4,4) np f32, shift=(3,) np f32) for every CHECKED episode e.
    synced=True: join + synchronise before every frame and download the checked episodes' state, so that the fp16 levels
    can be compared bit for bit; synced=False: the frames are enqueued back to back (what the bench times) and only the
    per-frame indices and the final state are compared (levels are returned for a cross-check against a synced pass).
    Returns a dict of mismatch counters (all zero == parity) plus the recorded levels."""
    eps = sorted(host)
    E, C, H, W, n_cells = batch.E, batch.C, batch.H, batch.W, batch.n_cells
    dev = batch.device
    n_frames = len(frames)
    if reset:
        batch.reset()
    # ---- oracle geometry first: per checked episode the index planes and the compact renumbering ----
    zero3 = np.zeros(3, np.float32)
    idx_ref = {e: [backproject_quantize(host[e]["depth"][t], host[e]["T"][t], intr, zero3, host[e]["shift"], np.float32(cell),
                                         batch.map_w, batch.map_h, 0, 0.5, want=("idx",))["idx"] for t in range(n_frames)] for e in eps}
    union = {e: np.unique(np.concatenate([i.reshape(-1) for i in idx_ref[e]])) for e in eps}
    compact = {e: [torch.from_numpy(np.searchsorted(union[e], i).astype(np.int64)) for i in idx_ref[e]] for e in eps}
    rows_dev = {e: torch.from_numpy(union[e].astype(np.int64)).to(dev) for e in eps}

    out = {"episodes": eps, "frames": n_frames, "synced": synced, "idx_mismatch": 0, "level_mismatch": 0, "count_mismatch": 0,
           "sum_max_err_over_scale": 0.0, "sum_out_of_tol": 0, "norm16_mismatch": 0, "stray_rows": 0}
    o_sums = {e: torch.zeros(len(union[e]), C) for e in eps}
    o_counts = {e: torch.zeros(len(union[e])) for e in eps}
    got_idx: Dict[int, List[torch.Tensor]] = {e: [] for e in eps}
    got_levels: Dict[int, List[List[torch.Tensor]]] = {e: [] for e in eps}
    feats: Dict[int, List[torch.Tensor]] = {e: [] for e in eps}
    observed = torch.ones(H, W, dtype=torch.bool)

    def compare_state(e, t):
        s = batch.sums[e][rows_dev[e]].cpu()
        c = batch.counts[e][rows_dev[e]].cpu()
        out["count_mismatch"] += int((c != o_counts[e]).sum())
        scale = float(o_sums[e].abs().max())
        err = float((s - o_sums[e]).abs().max()) / max(scale, 1e-30)
        out["sum_max_err_over_scale"] = max(out["sum_max_err_over_scale"], err)
        out["sum_out_of_tol"] += int(err > SUM_TOL)

    for t, fr in enumerate(frames):
        before = {}
        if synced:
            batch.join()
            torch.cuda.synchronize(dev)
            before = {e: (batch.sums[e][rows_dev[e]].cpu(), batch.counts[e][rows_dev[e]].cpu()) for e in eps}
        feat = fr["feat"]() if callable(fr["feat"]) else fr["feat"]
        levels = batch.step(fr["depth"], fr["pose"], shifts_dev, intr, float(cell), feat)
        for e in eps:                                                      # ordered on the caller's stream
            got_levels[e].append([lv[e].clone() for lv in levels])
            got_idx[e].append(batch.idx[e].clone())
            feats[e].append(feat[e].clone() if not synced else None)
        if not synced:
            continue
        batch.join()
        torch.cuda.synchronize(dev)
        for e in eps:
            out["idx_mismatch"] += int((got_idx[e][t].cpu().numpy() != idx_ref[e][t]).sum())
            ref_levels = R.read_frame(before[e][0], before[e][1], compact[e][t])
            for k in range(3):
                out["level_mismatch"] += int((_u16(got_levels[e][t][k]) != ref_levels[k][0].numpy().view(np.uint16)).sum())
            o_sums[e], o_counts[e] = R.write_mean_frame(o_sums[e], o_counts[e], feat[e].cpu().unsqueeze(0), observed, compact[e][t], stride=1)
            compare_state(e, t)
    batch.join()
    torch.cuda.synchronize(dev)
    for e in eps:
        if not synced:
            for t in range(n_frames):
                out["idx_mismatch"] += int((got_idx[e][t].cpu().numpy() != idx_ref[e][t]).sum())
                o_sums[e], o_counts[e] = R.write_mean_frame(o_sums[e], o_counts[e], feats[e][t].cpu().unsqueeze(0), observed, compact[e][t], stride=1)
            compare_state(e, n_frames - 1)
        # rows outside the union were never visible: they must still be zero, and so must their counts
        nz = (batch.sums[e] != 0).any(dim=1) | (batch.counts[e] != 0)
        nz[rows_dev[e]] = False
        out["stray_rows"] += int(nz.sum())
        # the incrementally maintained fp16 table == a full re-normalisation of the downloaded state (custom_rcnn.py:764-774,1036)
        table16 = R.create_implicit_memory(batch.sums[e][rows_dev[e]].cpu(), batch.counts[e][rows_dev[e]].cpu()).half()
        out["norm16_mismatch"] += int((_u16(batch.norm16[e][rows_dev[e]]) != table16.numpy().view(np.uint16)).sum())
    out["levels"] = got_levels
    out["ok"] = not any(out[k] for k in ("idx_mismatch", "level_mismatch", "count_mismatch", "sum_out_of_tol", "norm16_mismatch", "stray_rows"))
    return out


def summary(res: dict) -> dict:
    """JSON-serialisable part of a check_dense_steps result."""
    return {k: v for k, v in res.items() if k != "levels"}
