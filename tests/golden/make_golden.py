"""Generate the golden fixtures under tests/golden/ by EXECUTING THE REFERENCE'S OWN SOURCE on the CPU.

Needs /root/reference (dev container only; the GPU box never runs this).  Nothing from the reference is copied
into the repo: its code is imported (projector package) or its cited line ranges are read from disk and
exec'd in a scratch namespace at generation time.

  geometry_*.npz   SMNet/projector (imported as is): Projector.forward(return_heights=True), PointCloud.forward,
                   then the quantise lines SMNet/build_memory_data.py:135-143 exec'd verbatim
  write_*.npz      methods box_to_image_features / project_image_features / create_implicit_memory of
                   detic/modeling/meta_arch/custom_rcnn.py extracted with ast and exec'd (torch.cuda.* constructors
                   and .cuda() patched to CPU), plus lines 696-701 of update_implicit_memory
  read_*.npz       lines 142-192 of detic/modeling/backbone/timm.py exec'd verbatim

Run:  python tests/golden/make_golden.py
"""
import ast
import importlib
import math
import os
import sys
import textwrap

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference/Detic"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "SMNet"))

eod_episodes = importlib.import_module("embodied-object-detection_b200.episodes")


def patch_cpu():
    torch.cuda.FloatTensor = torch.FloatTensor
    torch.cuda.BoolTensor = torch.BoolTensor
    torch.Tensor.cuda = lambda self, *a, **k: self


def ref_methods():
    """FunctionDefs of CustomRCNNRecurrent compiled stand-alone."""
    path = os.path.join(REF, "detic/modeling/meta_arch/custom_rcnn.py")
    src = open(path).read()
    tree = ast.parse(src)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "CustomRCNNRecurrent"][0]
    want = {"box_to_image_features", "project_image_features", "create_implicit_memory"}
    ns = {"torch": torch, "np": np, "math": math, "autocast": torch.cuda.amp.autocast}
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in want:
            mod = ast.Module(body=[fn], type_ignores=[])
            exec(compile(mod, path, "exec"), ns)
    return ns, src.splitlines()


def exec_lines(lines, lo, hi, ns):
    """exec reference source lines lo..hi (1-based, inclusive), dedented."""
    exec(textwrap.dedent("\n".join(lines[lo - 1:hi])), ns)
    return ns


def gen_geometry():
    from projector.core import _transform3D
    from projector.point_cloud import PointCloud
    from projector.projector import Projector
    vfov = math.radians(67.5)
    bm_lines = open(os.path.join(REF, "SMNet/build_memory_data.py")).read().splitlines()
    for name, (H, W, n_frames, seed, mw, mh, cell, room) in {
        "geometry_small": (96, 128, 6, 7, 500, 500, 0.02 * 10, (12.0, 9.0)),
        "geometry_full": (480, 640, 2, 1234, 500, 500, 0.02 * 10, (12.0, 9.0)),
        "geometry_fine": (96, 128, 4, 11, 1000, 1000, 0.02, (24.0, 18.0)),     # out-of-map pixels: clip + mask
    }.items():
        ep = eod_episodes.make_episode(seed, n_frames, H, W, mw, mh, cell, room)
        xyzhe = torch.from_numpy(ep.xyzhe)
        T = _transform3D(xyzhe)
        shift = torch.from_numpy(ep.map_world_shift)
        depth = torch.from_numpy(ep.depth)
        # Projector (create_coco_mp3d.py:94-102,157-159): world_shift_origin = map_world_shift
        pr = Projector(vfov, 1, H, W, mh, mw, cell, shift, 0.5, device=torch.device("cpu"))
        q2, outl, hts = [], [], []
        for t in range(n_frames):
            a, b, c = pr.forward(depth[t][None, None], T[t:t + 1], return_heights=True)
            q2.append(a[0].numpy().astype(np.int32)); outl.append(b[0].numpy()); hts.append(c[0].numpy())
        # PointCloud (build_data.py:98-103,209) then build_memory_data.py:135-143 verbatim
        pc = PointCloud(vfov, 1, H, W, torch.zeros(3), 0.5, device=torch.device("cpu"))
        world = torch.cat([pc.forward(depth[t][None, None], T[t:t + 1])[0] for t in range(n_frames)])
        ns = {"torch": torch, "np": np, "projection_indices": world.clone(), "map_world_shift": shift.clone(),
              "resolution": 0.02, "res_downsample": 10 if cell > 0.02 else 1, "map_height": mh, "map_width": mw}
        exec_lines(bm_lines, 135, 144, ns)
        flat = ns["pixels_in_map"][..., 0].astype(np.int32)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), depth=ep.depth, xyzhe=ep.xyzhe, T=T.numpy(),
                            map_world_shift=ep.map_world_shift, cell=np.float32(ns["resolution"] * ns["res_downsample"]),
                            map_w=mw, map_h=mh, vfov=vfov, q2=np.stack(q2), outlier=np.stack(outl), height=np.stack(hts),
                            world=world.numpy(), flat=flat)
        print(name, "unique cells/frame", [len(np.unique(f)) for f in flat], "outliers", np.stack(outl).mean())


def gen_robot():
    """robot.npz: the online robot geometry (robot_demo.py:40-90 _transform3D with the axis swap, :92-225 ProjectorUtils with the
    hard-coded K, :514-534 millimetre depth -> world -> column-major flat index), executed from the reference's own source:
    the two definitions are extracted with ast (the module itself imports detectron2), lines 518-534 are exec'd verbatim."""
    path = os.path.join(REF, "robot_demo.py")
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "np": np, "math": math}
    for node in tree.body:
        if (isinstance(node, ast.FunctionDef) and node.name == "_transform3D") or (isinstance(node, ast.ClassDef) and node.name == "ProjectorUtils"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    lines = src.splitlines()
    device = torch.device("cpu")
    H, W = 480, 640
    res = 0.2
    map_w = map_h = math.ceil(40 / res)                                   # robot_demo.py:471-474
    rng = np.random.default_rng(31)
    projector = ns["ProjectorUtils"](vfov=math.radians(58), hfov=math.radians(87), batch_size=1, feature_map_height=H, feature_map_width=W,
                                     output_height=map_h, output_width=map_w, gridcellsize=res,
                                     world_shift_origin=torch.zeros(3), z_clip_threshold=3, device=device)
    out = {"K": np.array([380.3127746582031, 379.828857421875, 315.81829833984375, 250.9555206298828], np.float64),
           "map_world_shift": np.array([-13, 0, -13], np.float32), "res": res, "map_w": map_w, "map_h": map_h}
    depth_mm, poses, Ts, flats = [], [], [], []
    for t in range(3):
        ep = eod_episodes.make_episode(900 + t, 1, H, W, map_w, map_h, res)
        depth_image = np.round(ep.depth[0] * 1000.0).astype(np.uint16)   # RealSense-style millimetres (16-bit png in the demo)
        pose_val = np.array([rng.uniform(-8, 8), rng.uniform(-8, 8), rng.uniform(-3.1, 3.1)])
        u = {"np": np, "torch": torch, "device": device, "depth_image": depth_image, "pose_val": pose_val, "projector": projector,
             "_transform3D": ns["_transform3D"], "map_world_shift": torch.FloatTensor([-13, 0, -13]), "res": res, "map_h": map_h, "map_w": map_w}
        exec_lines(lines, 514, 534, u)                                    # depth_var .. proj_indices, verbatim
        depth_mm.append(depth_image); poses.append(pose_val); Ts.append(u["T"].numpy()); flats.append(u["proj_indices"][..., 0].astype(np.int32))
    out.update(depth_mm=np.stack(depth_mm), pose_val=np.stack(poses), T=np.concatenate(Ts), flat=np.stack(flats))
    np.savez_compressed(os.path.join(HERE, "robot.npz"), **out)
    print("robot: unique cells/frame", [len(np.unique(f)) for f in flats])


def gen_write():
    patch_cpu()
    ns, lines = ref_methods()
    rng = np.random.default_rng(5)
    H, W, C = 480, 640, 512                      # hard-coded in the reference (custom_rcnn.py:886-887)
    mw, mh = 60, 45                              # small grid: the literal one-hot is (P', cells) bool
    ep = eod_episodes.make_episode(21, 3, H, W, mw, mh, 0.02 * 10, (12.0, 9.0))
    from projector.core import _transform3D
    import oracle
    T = _transform3D(torch.from_numpy(ep.xyzhe)).numpy()
    intr = (320.0, 359.1853942871094, 320.0, 240.0)
    sums = torch.zeros(mw * mh, C)
    counts = torch.zeros(mw * mh)
    out = {}
    for t in range(3):
        idx = oracle.backproject_quantize(ep.depth[t], T[t], intr, np.zeros(3, np.float32), ep.map_world_shift,
                                          np.float32(0.2), mw, mh, 0, 0.5, want=("idx",))["idx"]
        bf, masks = eod_episodes.make_detections(rng, H, W, C, (3, 6))
        proj = torch.from_numpy(idx).long()
        image_features, observed = ns["box_to_image_features"](None, torch.from_numpy(bf), torch.from_numpy(masks))
        mean, observed_mem = ns["project_image_features"](None, image_features, observed, [proj], [sums])
        # update_implicit_memory lines 696-701 verbatim, then the += of 742-743 on the flat views
        u = {"torch": torch, "memory": sums, "observed_mem": observed_mem, "proj_features": mean, "proj_indices": proj}
        exec_lines(lines, 696, 701, u)
        sums = sums + u["semmap_update"]
        counts = counts + u["observed_update"][:, 0]
        norm, _ = ns["create_implicit_memory"](None, {"memory": sums, "observations": counts, "proj_indices": proj})
        out.update({f"idx{t}": idx, f"box_features{t}": bf, f"masks{t}": np.packbits(masks, axis=None),
                    f"K{t}": masks.shape[0], f"observed{t}": np.packbits(observed.numpy(), axis=None),
                    f"img_sample{t}": image_features[0, :, ::16, ::16].numpy(),        # subsample: 629 MB otherwise
                    f"img_checksum{t}": image_features.double().sum(dim=(0, 2, 3)).numpy(),
                    f"mean{t}": mean.numpy(), f"observed_mem{t}": observed_mem.numpy(),
                    f"sums{t}": sums.numpy().copy(), f"counts{t}": counts.numpy().copy(), f"norm{t}": norm.numpy()})
        print("write frame", t, "K", masks.shape[0], "P", int(observed.sum()), "M", int(observed_mem.sum()),
              "V", len(np.unique(idx)))
    np.savez_compressed(os.path.join(HERE, "write_mean.npz"), map_w=mw, map_h=mh, **out)


def gen_paste():
    """paste.npz: detectron2's paste_masks_in_image is absent here (unpinned git dependency), so its published CPU
    algorithm is restated in oracle/reference_ops.py; the arithmetic that decides every output bit - the grid
    construction ops and F.grid_sample - is executed by torch-CPU (AVX2 / AVX-512 ATen build) at generation time."""
    sys.path.insert(0, ROOT)
    from oracle import reference_ops as R
    rng = np.random.default_rng(77)
    H, W = 480, 640
    out = {}
    for name, edge in (("plain", False), ("edge", True)):
        _, probs, boxes = eod_episodes.make_mask_head_detections(rng, H, W, 8, (10, 10), 28, edge_cases=edge)
        masks, values = R.paste_masks_in_image(torch.from_numpy(probs), torch.from_numpy(boxes), (H, W), 0.5, want_values=True)
        out[name + "_probs"], out[name + "_boxes"] = probs, boxes
        out[name + "_masks_bits"] = np.packbits(masks.numpy().reshape(-1))
        # sampled values of the first two objects only (bit pattern pin of the sampler arithmetic), cropped to their region
        for k in range(2):
            ys, xs = np.nonzero(values[k].numpy() != 0)
            y0, y1, x0, x1 = (ys.min(), ys.max() + 1, xs.min(), xs.max() + 1) if ys.size else (0, 1, 0, 1)
            out[f"{name}_val{k}_yx"] = np.array([y0, y1, x0, x1], np.int32)
            out[f"{name}_val{k}"] = values[k, y0:y1, x0:x1].numpy()
        print(name, "pasted pixels per object:", masks.reshape(masks.shape[0], -1).sum(1).tolist())
    out["cpu_capability"] = np.array(torch.backends.cpu.get_cpu_capability())
    np.savez_compressed(os.path.join(HERE, "paste.npz"), H=H, W=W, thr=0.5, **out)


def gen_read():
    patch_cpu()
    lines = open(os.path.join(REF, "detic/modeling/backbone/timm.py")).read().splitlines()
    torch.manual_seed(3)
    H, W, C, CO = 480, 640, 128, 64               # channel counts are free in timm.py:142-192; kept small for size
    mw, mh = 60, 45
    ep = eod_episodes.make_episode(33, 2, H, W, mw, mh, 0.02 * 10, (12.0, 9.0))
    from projector.core import _transform3D
    import oracle
    T = _transform3D(torch.from_numpy(ep.xyzhe)).numpy()
    intr = (320.0, 359.1853942871094, 320.0, 240.0)
    sums = torch.randn(mw * mh, C) * 30
    counts = torch.randint(0, 6, (mw * mh,)).float()
    sums[counts == 0] = 0
    convs = [torch.nn.Conv2d(C, CO, 1, bias=True) for _ in range(3)]
    out = {"sums": sums.numpy(), "counts": counts.numpy(), "map_w": mw, "map_h": mh}
    for k, c in enumerate(convs):
        out[f"w{k}"], out[f"b{k}"] = c.weight.detach().numpy(), c.bias.detach().numpy()
    for t in range(2):
        idx = oracle.backproject_quantize(ep.depth[t], T[t], intr, np.zeros(3, np.float32), ep.map_world_shift,
                                          np.float32(0.2), mw, mh, 0, 0.5, want=("idx",))["idx"]
        proj = torch.from_numpy(idx).long()
        mem = sums.clone()
        sel = counts > 1
        mem[sel] = mem[sel] / counts.unsqueeze(1)[sel]
        results = [torch.randn(1, CO, H >> s, W >> s).half().float() for s in (3, 4, 5)]
        for fusion in (("sum", "mem_only", "image_only") if t == 0 else ("sum",)):
            captured = {}

            class _Conv:
                def __init__(self, conv, k):
                    self.conv, self.k = conv, k

                def __call__(self, x):
                    captured[self.k] = x.detach().clone()         # the pooled fp16-valued level fed to the conv
                    return self.conv(x)

            self_ns = type("S", (), {})()
            self_ns.memory_type, self_ns.feat_fusion, self_ns.map_feature_weight = "implicit_memory", fusion, 5
            self_ns.merge_map_projections = [_Conv(c, k) for k, c in enumerate(convs)]
            ns = {"torch": torch, "F": F, "self": self_ns, "map_memory": [mem.to(torch.half)], "proj_indices": [proj],
                  "observations": [counts.to(torch.half)], "results": [r.clone() for r in results]}
            with torch.no_grad():
                exec_lines(lines, 142, 192, ns)
            for k in range(3):
                out[f"fused_{fusion}_{t}_{k}"] = ns["results"][k].numpy()
            if fusion == "sum":
                for k in range(3):
                    out[f"level{t}_{k}"] = captured[k].to(torch.half).numpy()
                    out[f"res{t}_{k}"] = results[k].half().numpy()
        out[f"idx{t}"] = idx
    np.savez_compressed(os.path.join(HERE, "read_fuse.npz"), **out)
    print("read: levels", [out[f"level0_{k}"].shape for k in range(3)])


def gen_semmap():
    """Explicit semantic map: custom_rcnn.py:747-751 exec'd verbatim on tensors shaped as :731-743 leave them (permuted
    views), then the reference's own visualise_clip_image_features (:938-1017, cv2 headless, visualise=False)."""
    import cv2
    patch_cpu()
    path = os.path.join(REF, "detic/modeling/meta_arch/custom_rcnn.py")
    src = open(path).read()
    lines = src.splitlines()
    tree = ast.parse(src)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "CustomRCNNRecurrent"][0]
    fn = [f for f in cls.body if isinstance(f, ast.FunctionDef) and f.name == "visualise_clip_image_features"][0]
    ns = {"torch": torch, "np": np, "cv2": cv2, "palette": np.zeros((256, 3), np.uint8)}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    rng = np.random.default_rng(12)
    mw, mh, C, K = 40, 30, 512, 21
    cells = mw * mh
    counts = rng.integers(0, 5, cells).astype(np.float32)
    sums = (rng.standard_normal((cells, C)) * rng.uniform(0.2, 30, (cells, 1))).astype(np.float32)
    sums[counts == 0] = 0
    zs = rng.standard_normal((C, K)).astype(np.float32)
    zs /= np.linalg.norm(zs, axis=0, keepdims=True)
    self_ns = type("S", (), {})()
    self_ns.semmap_features = torch.from_numpy(sums).reshape(mh, mw, C).permute(2, 0, 1).unsqueeze(0)      # :731-732
    self_ns.observation_count = torch.from_numpy(counts).reshape(mh, mw, 1).permute(2, 0, 1)               # :734-735
    u = {"torch": torch, "self": self_ns}
    exec_lines(lines, 747, 751, u)
    inten = u["observation_intensity"]
    out = {"sums": sums, "counts": counts, "zs_weight": zs, "map_w": mw, "map_h": mh, "intensity": inten.reshape(-1).numpy()}
    for thresh in (0.4, 0.1):
        sem = ns["visualise_clip_image_features"](None, self_ns.semmap_features, torch.from_numpy(zs), "semmap", mask=inten,
                                                  thresh=thresh, visualise=False)
        out[f"semmap_{thresh}"] = np.asarray(sem).astype(np.int32)
        print("semmap thresh", thresh, "classes kept", int((out[f'semmap_{thresh}'] >= 0).sum()), "of", cells)
    np.savez_compressed(os.path.join(HERE, "semmap.npz"), **out)


def gen_loader_order():
    """Episode ordering / reset flags: SMNet/loader.py:97-117 and :289-293 exec'd verbatim on a synthetic file list."""
    import json
    lines = open(os.path.join(REF, "SMNet/loader.py")).read().splitlines()
    rng = np.random.default_rng(4)
    files = [f"{scene}_{lvl}_{i}.h5" for scene, lvl, n in (("17DRP5sb8fy", 0, 50), ("1LXtFkjw3qL", 1, 50), ("2azQ1b91cZZ", 0, 50)) for i in range(n)]
    files = [files[i] for i in rng.permutation(len(files))]
    out = {"files": files}
    for test_type in ("default", "episodic", "longterm"):
        self_ns = type("S", (), {})()
        self_ns.files, self_ns.test_type = list(files), test_type
        exec_lines(lines, 97, 117, {"self": self_ns})
        out[f"order_{test_type}"] = self_ns.files
        flags = []
        for file in self_ns.files[:120]:
            for i in (0, 1):
                u = {"self": self_ns, "file": file, "i": i}
                exec_lines(lines, 289, 293, u)
                flags.append(bool(u["mem_reset"]))
        out[f"reset_{test_type}"] = flags
    json.dump(out, open(os.path.join(HERE, "loader_order.json"), "w"))
    print("loader order: longterm length", len(out["order_longterm"]))


def gen_explicit_map():
    """MODEL.MEMORY_TYPE explicit_map / map_gt with a generated semantic map: SMNet/loader.py:222 (`semmap_real = semmap_real + 1`) and
    :232-246 (memory = [zero row; class embeddings], proj_indices = semmap_real[proj_indices]) exec'd verbatim; the pooled levels of
    the resulting (21,512) table through timm.py:147-168 exec'd verbatim."""
    lines = open(os.path.join(REF, "SMNet/loader.py")).read().splitlines()
    rng = np.random.default_rng(21)
    H, W, mw, mh, C, K = 96, 128, 70, 50, 512, 20
    clip = rng.standard_normal((K, C)).astype(np.float32)
    clip /= np.linalg.norm(clip, axis=1, keepdims=True)
    semmap = rng.integers(-1, K, (mw * mh,)).astype(np.int64)
    semmap[rng.random(mw * mh) < 0.5] = -1
    yy, xx = np.mgrid[0:H, 0:W]
    proj = ((yy // 5) * mw // 3 + xx // 3) % (mw * mh)
    proj[::9, ::7] = rng.integers(0, mw * mh, proj[::9, ::7].shape)
    self_ns = type("S", (), {})()
    self_ns.clip_path, self_ns.clip_embeddings, self_ns.memory_type, self_ns.smnet_class_mapping = "clip.npy", clip, "map_gt", None
    u = {"self": self_ns, "np": np, "semmap_real": semmap.copy(), "proj_indices": proj.copy(), "semmap_gt": None, "print": lambda *a, **k: None}
    exec_lines(lines, 222, 222, u)
    exec_lines(lines, 232, 246, u)
    memory, idx = u["memory"], u["proj_indices"]
    assert memory.shape == (K + 1, C) and idx.shape == (H, W)
    # the read of that table: timm.py:142-192 exec'd verbatim (as gen_read does); the pooled fp16-valued levels are captured where
    # they enter the 1x1 projections
    patch_cpu()
    tl = open(os.path.join(REF, "detic/modeling/backbone/timm.py")).read().splitlines()
    captured = {}

    class _Conv:
        def __init__(self, k):
            self.k = k

        def __call__(self, x):
            captured[self.k] = x.detach().clone()
            return torch.zeros((x.shape[0], 8, x.shape[2], x.shape[3]))

    fpn = type("S", (), {})()
    fpn.memory_type, fpn.feat_fusion, fpn.map_feature_weight = "implicit_memory", "sum", 5
    fpn.merge_map_projections = [_Conv(k) for k in range(3)]
    ns = {"torch": torch, "F": F, "self": fpn, "map_memory": [torch.from_numpy(memory.astype(np.float32)).half()],
          "proj_indices": [torch.from_numpy(idx).long()], "observations": [torch.zeros(memory.shape[0]).half()],
          "results": [torch.zeros(1, 8, H >> s_, W >> s_) for s_ in (3, 4, 5)]}
    with torch.no_grad():
        exec_lines(tl, 142, 192, ns)
    levels = [captured[k].to(torch.half).numpy() for k in range(3)]
    np.savez_compressed(os.path.join(HERE, "explicit_map.npz"), clip=clip, semmap=semmap, proj=proj.astype(np.int64), memory=memory.astype(np.float32),
                        idx=idx.astype(np.int64), level0=levels[0], level1=levels[1], level2=levels[2])
    print("explicit_map:", memory.shape, idx.shape, [l.shape for l in levels])


if __name__ == "__main__":
    if "--only-explicit-map" in sys.argv:
        gen_explicit_map()
        sys.exit(0)
    if "--only-loader" in sys.argv:
        gen_loader_order()
        sys.exit(0)
    if "--only-semmap" in sys.argv:
        gen_semmap()
        sys.exit(0)
    if "--only-robot" in sys.argv:
        gen_robot()
        sys.exit(0)
    if "--only-paste" in sys.argv:
        gen_paste()
        sys.exit(0)
    gen_geometry()
    gen_write()
    gen_read()
    gen_semmap()
    gen_paste()
    gen_robot()
    gen_loader_order()
    gen_explicit_map()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
