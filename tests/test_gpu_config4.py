"""BASELINE config 4 as a parity case: the reference's eval loop (custom_rcnn.py:443-515) over one synthetic MP3D-shaped
episode with MODEL.MEMORY_TYPE implicit_memory, MAP_FEAT_FUSION sum, MAP_FEATURE_WEIGHT 5 - the memory path (geometry,
read, projection+fusion, mask pasting, write) on the custom kernels through the reference-facing classes, everything
else stock torch: a random-init torchvision R50 + FPN in bf16 autocast stands in for the timm/detectron2 backbone, and a
seeded generator stands in for the detector heads (detectron2 is absent; its outputs are INPUTS of the path).
Every frame is checked against the CPU oracle: cell indices exact, fused p3-p5 within the fp32 tolerance (the conv is an
accumulation), grid sums within tolerance, touched / visible sets and counts exact."""
import math

import numpy as np
import pytest
import torch

import oracle
from oracle import reference_ops as R

pytestmark = pytest.mark.gpu
SUM_TOL = 1e-5


class _R50FPN(torch.nn.Module):
    """Stand-in dense backbone: torchvision resnet50 stages + FeaturePyramidNetwork -> [p3, p4, p5] (256 ch, strides 8/16/32)."""

    def __init__(self):
        super().__init__()
        import torchvision
        r = torchvision.models.resnet50(weights=None)
        self.stem = torch.nn.Sequential(r.conv1, r.bn1, r.relu, r.maxpool)
        self.layers = torch.nn.ModuleList([r.layer1, r.layer2, r.layer3, r.layer4])
        self.fpn = torchvision.ops.FeaturePyramidNetwork([512, 1024, 2048], 256)

    def forward(self, x):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            x = self.stem(x)
            feats = {}
            for i, layer in enumerate(self.layers):
                x = layer(x)
                if i >= 1:
                    feats[f"c{i + 2}"] = x
            out = self.fpn(feats)
        return [out["c3"].float(), out["c4"].float(), out["c5"].float()]


def test_config4_eval_loop_matches_oracle(eod, cuda):
    torch.manual_seed(0)
    rng = np.random.default_rng(4)
    H, W, C, mw, mh, cell, T = 480, 640, 512, 500, 500, 0.2, 3
    cells = mw * mh
    cfg = eod.config.merge_from_list(eod.config.add_detic_memory_config(),
                                     ["MODEL.MEMORY_TYPE", "implicit_memory", "MODEL.MAP_FEAT_FUSION", "sum", "MODEL.MAP_FEATURE_WEIGHT", "5"])
    fusion = eod.config.build_memory_fusion(cfg).to(cuda).eval()
    body = _R50FPN().to(cuda).eval()
    backbone = eod.CustomRecurrentFPN(body, None, fusion, out_features=("p3", "p4", "p5")).eval()
    mem = eod.config.build_spatial_memory(cfg, mem_feat_dim=C)
    ep = eod.episodes.make_episode(4321, T, H, W, mw, mh, cell)
    proj = eod.Projector(math.radians(67.5), 1, H, W, mh, mw, cell, np.zeros(3, np.float32), 0.5, device=cuda)
    Tm = eod.transform3d(torch.from_numpy(ep.xyzhe))
    intr = R.intrinsics(W, H, math.radians(67.5))
    weights = [c.weight.detach().cpu() for c in fusion.merge_map_projections]
    biases = [c.bias.detach().cpu() for c in fusion.merge_map_projections]
    sums, counts = torch.zeros(cells, C), torch.zeros(cells)
    launches0 = eod.ops.launch_count
    for t in range(T):
        frame = {"memory_reset": t == 0, "sequence_name": "synthetic_episode_0", "memory": np.zeros((cells, 256), np.float32)}
        if frame["memory_reset"]:                                                      # custom_rcnn.py:470-477
            mem.reset(cells)
        # geometry (offline in the reference: build_memory_data.py:131-144) -> frame['proj_indices'] (480,640,1) int32
        idx = proj.flat_indices(torch.from_numpy(ep.depth[t])[None, None], Tm[t:t + 1], ep.map_world_shift)[0]
        ref_idx = oracle.backproject_quantize(ep.depth[t], Tm[t].numpy(), intr, np.zeros(3, np.float32), ep.map_world_shift,
                                              np.float32(cell), mw, mh, 0, 0.5, want=("idx",))["idx"]
        assert np.array_equal(idx.cpu().numpy().reshape(H, W), ref_idx)
        frame["proj_indices"] = idx.reshape(H, W, 1)
        proj_indices = frame["proj_indices"].to(torch.long).squeeze(2)                 # :496-497
        frame["memory"], frame["observations"] = mem.implicit_memory, mem.observations  # :500-501 (default / episodic)
        frame["memory"], frame["proj_indices"] = mem.create_implicit_memory(frame)      # :504-505
        assert torch.equal(frame["memory"].cpu(), R.create_implicit_memory(sums_gpu_state(mem)[0], sums_gpu_state(mem)[1]))
        # inference: backbone with the memory block (timm.py:91-213)
        image = torch.from_numpy(rng.uniform(0, 255, (1, 3, H, W)).astype(np.float32)).to(cuda)
        map_memory, projection, observations = mem.preprocess_spatial_memory([frame])   # :1019-1042
        with torch.no_grad():
            feats, _ = backbone(image, map_memory, projection, observations)
            results = body(image)
        levels = R.read_pool(map_memory[0].cpu(), projection[0].cpu())
        ref = R.project_and_fuse(levels, [r.cpu() for r in results], weights, biases, 5.0, "sum")
        for k, name in enumerate(("p3", "p4", "p5")):
            assert feats[name].shape == ref[k].shape
            assert (feats[name].cpu() - ref[k]).abs().max().item() <= SUM_TOL * ref[k].abs().max().item(), (t, name)
        # detector heads (stock, out of scope): seeded stand-in for inference_with_proposals before pasting (:876-880)
        bf, probs, boxes = eod.episodes.make_mask_head_detections(rng, H, W, C, (5, 12), 28)
        mem.update_implicit_memory((torch.from_numpy(boxes), torch.from_numpy(bf), torch.from_numpy(probs), None), proj_indices,
                                   torch.zeros(cells, 1), frame)
        masks = torch.from_numpy(oracle.paste_masks(probs, boxes, H, W, 0.5))
        img, obs = R.box_to_image_features(torch.from_numpy(bf), masks)
        sums, counts = R.write_mean_frame(sums, counts, img, obs, torch.from_numpy(ref_idx).long(), stride=8)
        got = mem.implicit_memory.cpu()
        assert (got - sums).abs().max().item() <= SUM_TOL * sums.abs().max().item(), t
        assert torch.equal(got == 0, sums == 0)                                        # touched-cell set
        assert torch.equal(mem.observations.cpu(), counts)                             # visibility counts
    assert counts.max().item() >= 2 and eod.ops.launch_count - launches0 >= T * 9


def sums_gpu_state(mem):
    return mem.implicit_memory.cpu(), mem.observations.cpu()
