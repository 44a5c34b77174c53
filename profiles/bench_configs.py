"""Throughput of the BASELINE configurations that are NOT the bench line (parity-test cases in bench.py's contract), so
that DESIGN.md can quote them: CUDA events, 2 warm-up frame-steps, inputs resident, two alternating feature slabs.

  cfg3   SMNet height-max projection: C=256, 0.02 m cells, 1000x1000 map, E episodes in lock step (geometry + write_max)
  cfg5   dense write+read at C=512 on a 1000x1000 grid (EpisodeBatch.step), E episodes per GPU
  obj    the reference's live regime: K<=16 detections per frame, 28x28 mask probabilities + boxes -> write_detections
         (mask pasting folded in) + read, C=512, 500x500 grid
  det    configs[1] with the deterministic segmented-reduce write
Prints one JSON line per case."""
import importlib, json, math, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
eod = importlib.import_module("embodied-object-detection_b200")
ops = eod.ops
dev = torch.device("cuda:0")
H, W = 480, 640
CASES = os.environ.get("CASES", "cfg3,cfg5,obj,det").split(",")
T = int(os.environ.get("FRAMES", 12))


def episodes(E, n_frames, mw, mh, cell):
    eps = [eod.episodes.make_episode(1234 + e, n_frames, H, W, mw, mh, cell) for e in range(E)]
    Tm = eod.transform3d(torch.from_numpy(np.stack([ep.xyzhe for ep in eps]).reshape(-1, 5))).reshape(E, n_frames, 4, 4)
    pose = Tm[:, :, :3, :].reshape(E, n_frames, 12).permute(1, 0, 2).contiguous().to(dev)
    depth = torch.from_numpy(np.stack([ep.depth for ep in eps])).permute(1, 0, 2, 3).contiguous().to(dev)
    shifts = torch.from_numpy(np.stack([np.concatenate([np.zeros(3, np.float32), ep.map_world_shift]) for ep in eps])).to(dev)
    return depth, pose, shifts


def timed(step, n_frames, warm=2):
    for t in range(warm):
        step(t)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for t in range(warm, n_frames):
        step(t)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (n_frames - warm)


intr = eod.compute_intrinsics(W, H, math.radians(67.5))

if "cfg3" in CASES:
    E, C, mw, mh, cell = 16, 256, 1000, 1000, 0.02
    depth, pose, shifts = episodes(E, 4, mw, mh, cell)
    cells = mw * mh
    feats = [torch.randn((E, H, W, C), device=dev) for _ in range(2)]
    state = torch.zeros((E, cells, C), device=dev)
    hmap = torch.zeros((E, cells), device=dev)
    key = torch.zeros((E, cells), dtype=torch.int64, device=dev)
    arg = torch.zeros((E, cells), dtype=torch.int32, device=dev)
    obs = torch.zeros((E, cells), dtype=torch.uint8, device=dev)
    geo = {}

    def step3(t):
        g = ops.backproject_quantize(depth[t % 4], pose[t % 4], shifts, intr, cell, mw, mh, 0, 0.5, want_outlier=True, want_height=True, out=geo)
        ops.write_max(g["height"], g["idx"], g["outlier"], feats[t & 1], hmap, key, arg, obs, state, eod._lib.LAYOUT_HWC, 1)

    ms = timed(step3, T)
    raised = float(obs.sum().item()) / E
    print(json.dumps({"case": "cfg3 height-max", "E": E, "C": C, "grid": [mw, mh], "cell_m": cell, "ms_per_frame_step": ms,
                      "frames_per_s": E / ms * 1e3, "observed_cells_per_episode": raised,
                      "note": "geometry (idx, outlier, height) + packed-key atomicMax pass + winner-row copy; inliers only"}))
    del feats, state, hmap, key, arg, obs
    torch.cuda.empty_cache()

if "cfg5" in CASES:
    E, C, mw, mh, cell = 32, 512, 1000, 1000, 0.2
    depth, pose, shifts = episodes(E, 4, mw, mh, cell)
    batch = eod.EpisodeBatch(E, mw, mh, C, H, W, dev, pipeline=True)
    slabs = [torch.randn((E, C, H, W), device=dev) for _ in range(2)]
    batch.profile(True)

    def step5(t):
        batch.step(depth[t % 4], pose[t % 4], shifts, intr, cell, slabs[t & 1])

    ms = timed(lambda t: (step5(t), batch.join()), T)
    torch.cuda.synchronize()
    st = batch.stage_ms()
    wbytes = E * (H * W * C * 4 + H * W * 4)
    print(json.dumps({"case": "cfg5 dense C=512 1000x1000", "E": E, "C": C, "grid": [mw, mh], "ms_per_frame_step": ms, "frames_per_s": E / ms * 1e3,
                      "stage_ms": st, "write_GBps": wbytes / st["write"] / 1e6}))
    del batch, slabs
    torch.cuda.empty_cache()

if "obj" in CASES:
    E, C, mw, mh, cell, Kmax = 64, 512, 500, 500, 0.2, 16
    depth, pose, shifts = episodes(E, 4, mw, mh, cell)
    batch = eod.EpisodeBatch(E, mw, mh, C, H, W, dev)
    rng = np.random.default_rng(0)
    dets = []
    for s in range(2):
        bf = np.zeros((E, Kmax, C), np.float32); pr = np.zeros((E, Kmax, 28, 28), np.float32); bx = np.zeros((E, Kmax, 4), np.float32)
        n = np.zeros(E, np.int32)
        for e in range(E):
            f, p, b = eod.episodes.make_mask_head_detections(rng, H, W, C, (4, Kmax), 28)
            n[e] = f.shape[0]; bf[e, : n[e]], pr[e, : n[e]], bx[e, : n[e]] = f, p, b
        dets.append(tuple(torch.from_numpy(a).to(dev) for a in (bf, pr, bx, n)))

    def stepo(t):
        batch.project(depth[t % 4], pose[t % 4], shifts, intr, cell)
        batch.read()
        bf, pr, bx, n = dets[t & 1]
        batch.write_detections(bf, pr, bx, n)

    ms = timed(stepo, T)
    print(json.dumps({"case": "object regime (paste folded in) C=512 500x500", "E": E, "Kmax": Kmax, "ms_per_frame_step": ms, "frames_per_s": E / ms * 1e3,
                      "note": "project + read + paste/observed + sample + count + write_objects_pasted + flush + finalize, serial on one stream"}))
    del batch
    torch.cuda.empty_cache()

if "det" in CASES:
    E, C, mw, mh, cell = 64, 256, 500, 500, 0.2
    depth, pose, shifts = episodes(E, 4, mw, mh, cell)
    batch = eod.EpisodeBatch(E, mw, mh, C, H, W, dev, variant=eod._lib.WRITE_DET)
    slabs = [torch.randn((E, C, H, W), device=dev) for _ in range(2)]
    batch.profile(True)
    ms = timed(lambda t: batch.step(depth[t % 4], pose[t % 4], shifts, intr, cell, slabs[t & 1]), T)
    torch.cuda.synchronize()
    print(json.dumps({"case": "configs[1] deterministic write", "E": E, "ms_per_frame_step": ms, "frames_per_s": E / ms * 1e3, "stage_ms": batch.stage_ms(),
                      "overflowed": batch._det_ws.overflowed()}))
