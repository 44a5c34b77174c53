"""EpisodeRunner (lock-step waves over resident grid slots, per-slot reset, ragged episodes, longterm snapshot, save at
frame 0) against a SERIAL oracle loop that restates the reference's eval loop (custom_rcnn.py:441-539) over the
reference's own episode order and reset flags (SMNet/loader.py:97-117,289-293 via formats.order_files /
memory_reset_flag), for all three MODEL.TEST_TYPEs.  Run on the B200 box with `-m gpu`.
"""
import math
import os
import zlib

import numpy as np
import pytest
import torch

import oracle
from oracle import reference_ops as R

pytestmark = pytest.mark.gpu

H, W, C, MW, MH, CELL = 96, 128, 128, 60, 45, 0.2
SUM_TOL = 1e-5
SCENES = {"aaaaaaaaaaa_0": [3, 2, 4], "bbbbbbbbbbb_1": [2, 3], "ccccccccccc_0": [1, 3]}       # scene -> frames per sequence (ragged)


def _seed(*parts) -> int:
    return zlib.crc32(repr(parts).encode()) % 100000          # stable across processes (str hashes are salted)


def _dataset(eod, test_type, regime, seed=0):
    """Episodes in the loader's order: list of frame lists (host), plus per-episode shifts."""
    rng = np.random.default_rng(seed)
    files = [f"{scene}_{k}.h5" for scene, lens in SCENES.items() for k in range(len(lens))]
    rng.shuffle(files)                                                   # os.listdir order is arbitrary: order_files sorts
    ordered = eod.formats.order_files(files, test_type)
    intr = eod.compute_intrinsics(W, H, math.radians(67.5))
    cache = {}
    episodes, shifts = [], []
    for f in ordered:
        if f not in cache:
            scene, k = f[:13], int(f.split("_")[-1].split(".")[0])
            n = SCENES[scene][k]
            ep = eod.episodes.make_episode(_seed(scene, k), n, H, W, MW, MH, CELL)
            T = eod.transform3d(torch.from_numpy(ep.xyzhe))
            frames = []
            for i in range(n):
                idx = oracle.backproject_quantize(ep.depth[i], T[i].numpy(), intr, np.zeros(3, np.float32), ep.map_world_shift, np.float32(CELL),
                                                  MW, MH, 0, 0.5, want=("idx",))["idx"]
                fr = {"sequence_name": f, "_idx": idx}
                if regime == "dense-idx":
                    fr["proj_indices"] = idx[..., None].astype(np.int32)              # (H,W,1) int32 as memory_data/*.h5 holds them
                else:
                    fr["depth"], fr["pose"] = ep.depth[i], T[i, :3].reshape(12).numpy()
                if regime == "detections":
                    frng = np.random.default_rng(_seed(f, i))
                    if i == 1:                                                        # a frame without kept detections writes nothing (:686)
                        fr["box_features"], fr["mask_probs"], fr["boxes"] = np.zeros((0, C), np.float32), np.zeros((0, 28, 28), np.float32), np.zeros((0, 4), np.float32)
                    else:
                        fr["box_features"], fr["mask_probs"], fr["boxes"] = eod.episodes.make_mask_head_detections(frng, H, W, C, (2, 6), 28)
                else:
                    fr["feat"] = np.random.default_rng(_seed(f, i)).standard_normal((C, H, W)).astype(np.float32)
                frames.append(fr)
            cache[f] = (frames, np.concatenate([np.zeros(3, np.float32), ep.map_world_shift]))
        frames, sh = cache[f]
        episodes.append([dict(fr, memory_reset=eod.formats.memory_reset_flag(test_type, f, i)) for i, fr in enumerate(frames)])
        shifts.append(sh)
    return episodes, shifts, intr


def _oracle_serial(episodes, test_type, regime):
    """custom_rcnn.py:443-539, one model instance, one sequence after the other."""
    cells = MW * MH
    sums = counts = upd = None
    levels, saved = {}, {}
    for e, ep in enumerate(episodes):
        for i, fr in enumerate(ep):
            if fr["memory_reset"]:                                                        # :470-477
                sums, counts = torch.zeros(cells, C), torch.zeros(cells)
            if i == 0 and test_type == "longterm":                                        # :482-486
                upd = (sums, counts)
            if test_type in ("default", "episodic"):                                      # :489-491
                upd = (sums, counts)
            proj = torch.from_numpy(fr["_idx"]).long()
            levels[(e, i)] = R.read_frame(upd[0], upd[1], proj)                           # :504 + timm.py:147-168
            if regime == "detections":
                K = fr["box_features"].shape[0]
                if K:                                                                     # :686: no kept detection -> no write at all
                    masks = torch.from_numpy(oracle.paste_masks(fr["mask_probs"], fr["boxes"], H, W, 0.5))
                    img, obs = R.box_to_image_features(torch.from_numpy(fr["box_features"]), masks)
                    sums, counts = R.write_mean_frame(sums, counts, img, obs, proj, stride=8)
            else:
                sums, counts = R.write_mean_frame(sums, counts, torch.from_numpy(fr["feat"])[None], torch.ones(H, W, dtype=torch.bool), proj, stride=1)
            if i == 0:                                                                    # :518-530
                saved[fr["sequence_name"]] = (sums.clone(), counts.clone())
    return levels, saved


@pytest.mark.parametrize("test_type,regime", [("default", "dense-idx"), ("episodic", "dense-depth"), ("longterm", "dense-idx"),
                                              ("default", "detections"), ("longterm", "detections")])
def test_runner_matches_serial_reference_loop(eod, cuda, tmp_path, test_type, regime):
    episodes, shifts, intr = _dataset(eod, test_type, regime)
    ref_levels, ref_saved = _oracle_serial(episodes, test_type, regime)
    provider = eod.HostEpisodeProvider(episodes, cuda, shifts, k_max=8)
    runner = eod.EpisodeRunner(provider, MW, MH, C, n_slots=2, height=H, width=W, device=cuda, test_type=test_type, intr=intr, cell=CELL,
                               save_dir=str(tmp_path))
    got = {}

    def on_levels(step, levels):
        for s, a in enumerate(step.assign):
            if a is not None:
                got[a] = [lv[s].clone() for lv in levels]

    stats = runner.run(on_levels)
    torch.cuda.synchronize()
    n_frames = sum(len(ep) for ep in episodes)
    assert stats["frames"] == n_frames and len(got) == n_frames
    assert stats["steps"] < n_frames                                    # slots really ran side by side
    for key, lv in got.items():
        for k in range(3):
            a, b = lv[k].float().cpu(), ref_levels[key][k][0].float()
            tol = 2e-3 * max(1.0, float(b.abs().max()))                 # fp16 rounding of reduction-order noise in the sums
            assert torch.allclose(a, b, rtol=2e-3, atol=tol), (test_type, key, k, float((a - b).abs().max()))
    assert set(os.path.splitext(f)[0] for f in os.listdir(tmp_path / "memory")) == set(os.path.splitext(n)[0] for n in ref_saved)
    for name, (s_ref, c_ref) in ref_saved.items():
        semmap1, mem, obs = eod.formats.load_memory(str(tmp_path / "memory" / name))
        assert np.array_equal(obs, c_ref.numpy()), name                 # visibility counts: exact
        assert np.abs(mem - s_ref.numpy()).max() <= SUM_TOL * max(float(s_ref.abs().max()), 1e-30), name
        assert np.array_equal((mem != 0).any(1), (s_ref.numpy() != 0).any(1)), name      # touched-cell set (rows): exact; single elements may cancel to 0
        assert (semmap1 == 0).all()                                     # no classifier given: semmap stays -1 (+1 on load, loader.py:221)


def test_runner_rejects_out_of_range_indices(eod, cuda):
    """an index plane that does not belong to the map must raise (the reference's gather does), not corrupt memory."""
    episodes, shifts, intr = _dataset(eod, "episodic", "dense-idx")
    episodes[0][0]["proj_indices"] = episodes[0][0]["proj_indices"].copy()
    episodes[0][0]["proj_indices"][5, 7, 0] = MW * MH                  # one past the last cell
    runner = eod.EpisodeRunner(eod.HostEpisodeProvider(episodes, cuda, shifts), MW, MH, C, n_slots=2, height=H, width=W, device=cuda,
                               test_type="episodic", intr=intr, cell=CELL)
    with pytest.raises(IndexError):
        runner.run()
    batch = eod.EpisodeBatch(1, MW, MH, C, H, W, cuda)
    bad = torch.zeros((1, H, W), dtype=torch.int64, device=cuda)
    bad[0, 0, 0] = -1
    with pytest.raises(IndexError):
        batch.set_indices(bad)
    mem = eod.SpatialFeatureMemory(C, cuda, height=H, width=W)
    mem.reset(MW * MH)
    with pytest.raises(IndexError):
        mem.read_levels(torch.full((H, W), MW * MH + 3, dtype=torch.int64))
    with pytest.raises(IndexError):
        mem.write_image_features(torch.zeros(1, C, H, W), None, torch.full((H, W, 1), -7, dtype=torch.int32))


def test_per_slot_reset_and_active_mask_leave_other_slots_alone(eod, cuda):
    """step(active=, reset_mask=): an idle slot's grid (sums, counts, fp16 table) is bit-identical before and after, even when
    its feature slab holds NaNs; a reset slot is all-zero before its read; the others advance as without masks."""
    E = 3
    eps = [eod.episodes.make_episode(40 + e, 3, H, W, MW, MH, CELL) for e in range(E)]
    intr = eod.compute_intrinsics(W, H, math.radians(67.5))
    shifts = torch.from_numpy(np.stack([np.concatenate([np.zeros(3, np.float32), ep.map_world_shift]) for ep in eps])).to(cuda)
    gen = torch.Generator(device=cuda).manual_seed(1)
    a, b = eod.EpisodeBatch(E, MW, MH, C, H, W, cuda, pipeline=True), eod.EpisodeBatch(E, MW, MH, C, H, W, cuda)
    for t in range(3):
        depth = torch.from_numpy(np.stack([ep.depth[t] for ep in eps])).to(cuda)
        pose = eod.transform3d(torch.from_numpy(np.stack([ep.xyzhe[t] for ep in eps])))[:, :3].reshape(E, 12).to(cuda)
        feat = torch.randn((E, C, H, W), device=cuda, generator=gen)
        if t < 2:
            a.step(depth, pose, shifts, intr, CELL, feat)
            b.step(depth, pose, shifts, intr, CELL, feat)
            continue
        a.join()
        torch.cuda.synchronize()
        snap = (a.sums[1].clone(), a.counts[1].clone(), a.norm16[1].clone())
        poisoned = feat.clone()
        poisoned[1] = float("nan")                                       # the idle slot's slab must not even be read
        active = torch.tensor([1, 0, 1], dtype=torch.int32, device=cuda)
        reset = torch.tensor([0, 0, 1], dtype=torch.int32, device=cuda)
        levels = a.step(depth, pose, shifts, intr, CELL, poisoned, active=active, reset_mask=reset, inputs_ready=False)
        lv_b = b.step(depth, pose, shifts, intr, CELL, feat)
        a.join()
        torch.cuda.synchronize()
        assert torch.equal(a.sums[1], snap[0]) and torch.equal(a.counts[1], snap[1]) and torch.equal(a.norm16[1], snap[2])
        assert all(float(lv[2].float().abs().max()) == 0 for lv in levels)          # slot 2 was cleared before its read
        for k in range(3):
            assert torch.allclose(levels[k][0].float(), lv_b[k][0].float(), rtol=2e-3, atol=2e-3)
        assert torch.equal(a.counts[0], b.counts[0])
        assert (a.sums[0] - b.sums[0]).abs().max().item() <= 1e-6 * b.sums[0].abs().max().item()
        # slot 2 now holds exactly one frame
        fresh = eod.EpisodeBatch(1, MW, MH, C, H, W, cuda)
        fresh.step(depth[2:3], pose[2:3], shifts[2:3], intr, CELL, feat[2:3].contiguous())
        torch.cuda.synchronize()
        assert torch.equal(a.counts[2], fresh.counts[0])
        assert (a.sums[2] - fresh.sums[0]).abs().max().item() <= 1e-6 * fresh.sums[0].abs().max().item()
