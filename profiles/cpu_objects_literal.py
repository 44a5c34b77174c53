"""CPU timing of the reference's LITERAL object-regime memory ops (oracle restatement of custom_rcnn.py:884-936, 696-743 with the
dense one-hot matmul, detectron2-style mask pasting, timm.py:147-168 read) on one synthetic frame sequence: the figure the
object-regime GPU numbers of DESIGN.md sit next to.  Test infrastructure: runs the oracle, not the product.  No GPU needed."""
import importlib, json, math, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from oracle import reference_ops as R
episodes = importlib.import_module("embodied-object-detection_b200.episodes")
torch.set_num_threads(os.cpu_count() or 1)
H, W, C, mw, mh, cell, T = 480, 640, 512, 500, 500, 0.2, int(os.environ.get("FRAMES", 3))
ep = episodes.make_episode(1234, T, H, W, mw, mh, cell)
Tm = R.transform3d(torch.from_numpy(ep.xyzhe)).numpy()
intr = R.intrinsics(W, H, math.radians(67.5))
rng = np.random.default_rng(0)
sums, counts = torch.zeros(mw * mh, C), torch.zeros(mw * mh)
stage = {"geometry": 0.0, "read": 0.0, "paste": 0.0, "box_to_image": 0.0, "project_onehot": 0.0, "accumulate": 0.0}
t_all = time.perf_counter()
for t in range(T):
    bf, probs, boxes = episodes.make_mask_head_detections(rng, H, W, C, (4, 16), 28)
    t0 = time.perf_counter()
    idx = oracle.backproject_quantize(ep.depth[t], Tm[t], intr, np.zeros(3, np.float32), ep.map_world_shift, np.float32(cell), mw, mh, 0, 0.5, want=("idx",))["idx"]
    proj = torch.from_numpy(idx).long()
    t1 = time.perf_counter(); stage["geometry"] += t1 - t0
    R.read_frame(sums, counts, proj)
    t2 = time.perf_counter(); stage["read"] += t2 - t1
    masks = R.paste_masks_in_image(torch.from_numpy(probs), torch.from_numpy(boxes), (H, W), 0.5)
    t3 = time.perf_counter(); stage["paste"] += t3 - t2
    img, obs = R.box_to_image_features(torch.from_numpy(bf), masks)
    t4 = time.perf_counter(); stage["box_to_image"] += t4 - t3
    mean, observed_mem = R.project_image_features_dense(img, obs, proj, mw * mh, stride=8)
    t5 = time.perf_counter(); stage["project_onehot"] += t5 - t4
    sums, counts = R.accumulate(sums, counts, mean, observed_mem, proj)
    stage["accumulate"] += time.perf_counter() - t5
sec = time.perf_counter() - t_all
print(json.dumps({"frames": T, "cores": torch.get_num_threads(), "s_per_frame": sec / T, "frames_per_s": T / sec,
                  "stage_s_per_frame": {k: v / T for k, v in stage.items()}}))
