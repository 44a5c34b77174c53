"""Parity ON THE BENCHMARKED CONFIGURATION (run on the B200 box with `-m gpu`): EpisodeBatch(E=64, pipeline=True) at 480x640,
C=256, CHW fp32 features, on the 500x500 grid bench.py times (BASELINE configs[1]) and on the north-star 1000x1000 grid.
Several episodes of the batch are compared against the CPU oracle frame by frame (oracle/path_check.py): indices, counts
and fp16 levels bit-exact, fp32 sums within 1e-5 of scale, never-visible rows zero; then the same frames are enqueued back
to back without any synchronisation (exactly what the timed loop does) and must land in the same state.
"""
import math

import numpy as np
import pytest
import torch

from oracle import path_check

pytestmark = pytest.mark.gpu

H, W, C, E = 480, 640, 256, 64
CHECKED = (0, 21, 42, 63)          # first, last and two in between: episode strides inside every (E, ...) tensor are exercised


def _setup(eod, cuda, mw, mh, cell, n_frames):
    eps = {e: eod.episodes.make_episode(1234 + e, n_frames, H, W, mw, mh, cell) for e in CHECKED}
    filler = eod.episodes.make_episode(999, n_frames, H, W, mw, mh, cell)      # the unchecked episodes replay one more trajectory
    intr = eod.compute_intrinsics(W, H, math.radians(67.5))
    all_eps = [eps.get(e, filler) for e in range(E)]
    shifts = torch.from_numpy(np.stack([np.concatenate([np.zeros(3, np.float32), ep.map_world_shift]) for ep in all_eps])).to(cuda)
    T = {e: eod.transform3d(torch.from_numpy(eps[e].xyzhe)) for e in CHECKED}
    T_fill = eod.transform3d(torch.from_numpy(filler.xyzhe))
    host = {e: dict(depth=eps[e].depth, T=T[e].numpy(), shift=eps[e].map_world_shift) for e in CHECKED}
    slabs = [torch.empty((E, C, H, W), device=cuda) for _ in range(2)]
    gen = torch.Generator(device=cuda)

    def feat_fn(t):
        def make():
            gen.manual_seed(77 + t)
            return slabs[t & 1].normal_(generator=gen)
        return make

    frames = []
    for t in range(n_frames):
        depth = torch.from_numpy(np.stack([ep.depth[t] for ep in all_eps])).to(cuda)
        pose = torch.stack([(T[e] if e in T else T_fill)[t, :3].reshape(12) for e in range(E)]).to(cuda)
        frames.append(dict(depth=depth, pose=pose, feat=feat_fn(t)))
    return frames, shifts, intr, host


@pytest.mark.parametrize("mw,mh,cell", [(500, 500, 0.2), (1000, 1000, 0.2), (1000, 1000, 0.05)], ids=["500x500@0.2", "1000x1000@0.2", "1000x1000@0.05"])
def test_bench_configuration_matches_oracle(eod, cuda, mw, mh, cell):
    n_frames = 3
    frames, shifts, intr, host = _setup(eod, cuda, mw, mh, cell, n_frames)
    batch = eod.EpisodeBatch(E, mw, mh, C, H, W, cuda, pipeline=True)
    # pass 1: synchronised between frames, every frame of the checked episodes against the oracle
    res = path_check.check_dense_steps(batch, frames, shifts, intr, cell, host, synced=True)
    assert res["ok"], path_check.summary(res)
    assert res["sum_max_err_over_scale"] <= path_check.SUM_TOL
    state = (batch.sums[list(CHECKED)].clone(), batch.counts.clone())
    # pass 2: the same frames back to back, as the timed loop of bench.py enqueues them
    res2 = path_check.check_dense_steps(batch, frames, shifts, intr, cell, host, synced=False)
    assert res2["ok"], path_check.summary(res2)
    assert torch.equal(batch.counts, state[1])                                   # every one of the 64 episodes
    scale = state[0].abs().max().item()
    assert (batch.sums[list(CHECKED)] - state[0]).abs().max().item() <= 1e-6 * scale      # reduction-order noise only
    for e in CHECKED:
        for t in range(n_frames):
            for a, b in zip(res["levels"][e][t], res2["levels"][e][t]):
                assert torch.allclose(a.float(), b.float(), rtol=2e-3, atol=2e-3 * max(1.0, a.float().abs().max().item()))
