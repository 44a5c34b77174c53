"""CPU-side checks: the C-ABI library loads and exports every symbol include/eod_memory.h declares, the product
never touches the oracle, CPU tensors are refused (no fallback), host logic (config keys, map dims, sharding)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "embodied-object-detection_b200")


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "eod_memory.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(eod_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(eod):
    eod.build.build()
    handle = ctypes.CDLL(eod.build.SO_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/eod_memory.h but not exported"
    # the ctypes table binds exactly the declared set
    assert sorted(eod._lib.SIGNATURES) == declared
    assert eod._lib.lib().eod_version() == 100
    # every entry point that launches kernels is in the launch-count table ops._call consults (a missing key raised only on the GPU)
    no_launch = {"eod_version", "eod_last_error", "eod_write_mean_det_workspace_bytes", "eod_write_mean_det_status_offset"}
    assert set(eod.ops._LAUNCHES) == set(declared) - no_launch


def test_header_is_plain_c(tmp_path):
    """include/eod_memory.h is the drop-in boundary: it must compile as C99 and as C++17 without warnings (no torch / CUDA types)."""
    src = tmp_path / "hdr.c"
    src.write_text('#include "eod_memory.h"\nint probe(void) { return eod_version(); }\n')
    inc = os.path.join(ROOT, "include")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only"], ["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++"]):
        res = subprocess.run(cmd + ["-I", inc, str(src)], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr


def test_no_library_gemm_in_product_code():
    """Every dense contraction of the path runs on the in-tree tcgen05 kernels (csrc/project_fuse.cu, csrc/linear.cu): the package
    holds no call into the library GEMMs (torch.matmul in the training path and the A7'' forward projection)."""
    pat = re.compile(r"torch\.(matmul|mm|bmm|addmm|einsum|baddbmm)\s*\(|F\.(linear|conv2d|conv1d|bilinear)\s*\(|\.matmul\(")
    bad = []
    for f in os.listdir(PKG):
        if f.endswith(".py"):
            for n, line in enumerate(open(os.path.join(PKG, f)), 1):
                code = line.split("#", 1)[0]
                if pat.search(code):
                    bad.append((f, n, line.strip()))
    assert not bad, bad


def test_library_is_sm100a_only(eod):
    out = subprocess.run(["cuobjdump", "--list-elf", eod.build.SO_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "oracle/" in src:
                    bad.append(f)
    assert not bad, bad


def test_cpu_tensors_are_refused(eod):
    with pytest.raises(eod.EodError):
        eod.ops.fuse(torch.zeros(8), torch.zeros(8), 5.0, eod._lib.FUSE_SUM)
    with pytest.raises(eod.EodError):
        eod.EpisodeBatch(1, 10, 10, 128, device="cpu")
    with pytest.raises(eod.EodError):
        eod.SpatialFeatureMemory(device="cpu")
    # the round-2 entry points have no CPU path either
    with pytest.raises(eod.EodError):
        eod.ops.linear_rows(torch.zeros(4, 16), torch.zeros(16, 16))
    with pytest.raises(eod.EodError):
        eod.ops.remap_indices(torch.zeros(4, dtype=torch.int64), torch.zeros(9, dtype=torch.int64), 21)
    with pytest.raises(eod.EodError):
        eod.ops.backproject_count(torch.zeros(1, 32, 32), torch.zeros(1, 12), torch.zeros(1, 6), (1.0, 1.0, 16.0, 16.0), 0.2, 10, 10,
                                  torch.zeros(1, 32, 32, dtype=torch.int32), torch.zeros(1, 100, dtype=torch.int32))


def test_badarg_codes_without_gpu(eod):
    lib = eod._lib.lib()
    assert lib.eod_fuse(None, None, 1.0, 0, 0, None, None) == -1
    assert b"eod_fuse" in lib.eod_last_error()
    assert lib.eod_read_pool(None, 0, None, None, 0, 1, 256, 480, 640, 10, None, None, None, None) == -1
    assert lib.eod_backproject_quantize(None, None, None, 1, 480, 640, 1.0, 1.0, 0.0, 0.0, 0.2, 10, 10, 0, 0.5,
                                        None, None, None, None, None, None) == -1


def test_config_keys_and_defaults(eod):
    cfg = eod.config.add_detic_memory_config()
    m = cfg.MODEL
    assert (m.MEMORY_TYPE, m.MAP_FEAT_FUSION, m.MAP_FEATURE_WEIGHT, m.MEMORY_FEATURE_WEIGHT) == ("", "", 500, 100)
    assert (m.MEMORY_CLS_SCORE_THRESH, m.MEMORY_OBS_SCORE_THRESH, m.TEST_TYPE) == (0.3, 0.4, "default")
    eod.config.merge_from_list(cfg, ["MODEL.MEMORY_TYPE", "implicit_memory", "MODEL.MAP_FEAT_FUSION", "sum",
                                     "MODEL.MAP_FEATURE_WEIGHT", "5"])
    fusion = eod.config.build_memory_fusion(cfg)
    assert fusion.map_feature_weight == 5.0 and fusion.feat_fusion == "sum"
    names = sorted(k for k, _ in fusion.state_dict().items())
    assert names == [f"map_merge_projection{i}.{p}" for i in (1, 2, 3) for p in ("bias", "weight")]
    assert fusion.state_dict()["map_merge_projection1.weight"].shape == (256, 512, 1, 1)
    cfg.MODEL.MAP_FEAT_FUSION = "ave"
    with pytest.raises(ValueError):
        eod.config.build_memory_fusion(cfg)
    cfg.MODEL.MEMORY_TYPE = "image_only"
    f2 = eod.config.build_memory_fusion(cfg)
    assert len(f2.state_dict()) == 0
    res = [torch.zeros(1, 256, 4, 4)]
    assert f2(res, None, None, None)[0] is res[0]        # image_only: FPN results untouched (timm.py:194-196)


def test_backbone_state_dict_is_the_reference_key_set(eod):
    """the three 1x1 convs must appear ONCE, under the reference's names (timm.py:78-86), so that a reference
    checkpoint loads with strict=True and saved checkpoints carry no keys the reference does not know."""
    fusion = eod.MemoryFusion("implicit_memory", "sum", 5.0)
    body = torch.nn.Conv2d(3, 8, 1)
    bb = eod.CustomRecurrentFPN(lambda x: [body(x)] * 3, None, fusion)
    bb.body = body
    keys = sorted(bb.state_dict().keys())
    assert keys == sorted(["body.weight", "body.bias"] + [f"map_merge_projection{i}.{p}" for i in (1, 2, 3) for p in ("bias", "weight")])
    ref_ckpt = {k: torch.randn_like(v) for k, v in bb.state_dict().items()}
    bb.load_state_dict(ref_ckpt, strict=True)
    assert torch.equal(fusion.map_merge_projection2.weight, ref_ckpt["map_merge_projection2.weight"])       # shared parameters
    assert sum(p.numel() for p in bb.parameters()) == sum(v.numel() for v in ref_ckpt.values())
    assert any("map_merge" in n for n, _ in bb.named_parameters())                                        # the LR / un-freeze rule still matches


def test_map_dims_lookup(eod):
    info = {"17DRP5sb8fy_0": {"dim": [1094, 1, 569]}}
    rep = {"apartment_0": {"dim": [120, 1, 80]}}
    mem = eod.SpatialFeatureMemory.__new__(eod.SpatialFeatureMemory)
    mem.semmap_gt_info, mem.replica_map_info, mem.downsample = info, rep, 10
    assert mem.map_dims("17DRP5sb8fy_0_12") == (110, 57)            # ceil(dim / 10), custom_rcnn.py:705-707
    assert mem.map_dims("apartment_0_3") == (120, 80)               # replica table, :710-725
    assert mem.map_dims("robot") == (200, 200)                      # default, :727-729


def test_sharding_partitions(eod):
    sh = eod.sharding
    for world in (1, 2, 4, 8):
        parts = [sh.shard_episodes(512, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(512))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    names = [f"scene{i % 5:08d}_{i // 5}" for i in range(40)]
    parts = [sh.shard_scenes(names, r, 2) for r in range(2)]
    assert sorted(sum(parts, [])) == list(range(40))
    for p in parts:
        assert len({names[i][:13] for i in p} & {names[i][:13] for i in parts[1 - parts.index(p)]}) == 0


def test_episode_generator_is_deterministic_and_mp3d_shaped(eod):
    a = eod.episodes.make_episode(1234, n_frames=2, H=96, W=128)
    b = eod.episodes.make_episode(1234, n_frames=2, H=96, W=128)
    assert np.array_equal(a.depth, b.depth) and np.array_equal(a.xyzhe, b.xyzhe)
    assert a.depth.dtype == np.float32 and a.depth.max() <= 10.0 and (a.depth == 0).mean() > 0.005
    assert np.allclose(a.xyzhe[:, 1], 1.25) and np.allclose(a.xyzhe[:, 4], np.pi)
    step = np.linalg.norm(np.diff(a.xyzhe[:, [0, 2]], axis=0), axis=1)
    assert ((np.abs(step - 0.1) < 1e-4) | (step < 1e-6)).all()


def _gloo_worker(rank, world, port, q):
    import importlib
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    eod = importlib.import_module("embodied-object-detection_b200")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, _, w = eod.sharding.init_from_env(backend="gloo")
    mine = eod.sharding.shard_episodes(11, r, w)
    tot = eod.sharding.gather_counters({"frames": 20.0 * len(mine), "checksum": float(sum(mine))}, torch.device("cpu"))
    mx = eod.sharding.max_over_ranks(float(rank + 1), torch.device("cpu"))
    q.put((rank, tot, mx))
    dist.destroy_process_group()


def test_gloo_world2_counters():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(60) for p in procs]
    for _, tot, mx in res:
        assert tot == {"checksum": float(sum(range(11))), "frames": 220.0}
        assert mx == 2.0


def test_episode_order_and_reset_flags_match_reference(eod):
    """formats.order_files / memory_reset_flag vs SMNet/loader.py:97-117,289-293 executed from the reference source
    (tests/golden/loader_order.json)."""
    import json
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "loader_order.json")))
    F = eod.formats
    for test_type in ("default", "episodic", "longterm"):
        order = F.order_files(g["files"], test_type)
        assert order == g[f"order_{test_type}"], test_type
        flags = [F.memory_reset_flag(test_type, f, i) for f in order[:120] for i in (0, 1)]
        assert flags == g[f"reset_{test_type}"], test_type


def test_store_round_trip_and_episode_dataset(eod, tmp_path):
    """npz container with the reference's dataset names: memory_data, saved memory (impicit_memory [sic]) and the
    episode iteration contract of SMNetDetectionLoader."""
    F = eod.formats
    root = tmp_path / "mp3d_val"
    cells, T, H, W = 60, 23, 8, 12
    rng = np.random.default_rng(0)
    for seq in ("sceneA_0_1", "sceneA_0_0", "sceneB_1_0"):
        proj = rng.integers(0, cells, (T, H, W)).astype(np.int64)
        F.write_memory_data(str(root / "memory_data" / (seq + ".h5")), proj, cells)
        F.write_store(str(root / "sensor_data" / (seq + ".h5")), {"rgb": rng.integers(0, 255, (T, H, W, 3)).astype(np.uint8),
                                                                   "depth": rng.uniform(0, 10, (T, H, W)).astype(np.float32)})
    ds = F.EpisodeDataset(str(root), test_type="default")
    assert [f.split(".")[0] for f in ds.files] == ["sceneA_0_0", "sceneA_0_1", "sceneB_1_0"]
    ep = ds[0]
    assert len(ep) == 20                                                    # max_sequence_length (loader.py:71)
    assert ep[0]["memory_reset"] and not ep[1]["memory_reset"] and not ds[1][0]["memory_reset"] and ds[2][0]["memory_reset"]
    assert ep[0]["proj_indices"].shape == (H, W, 1) and ep[0]["proj_indices"].dtype == np.int32
    assert ep[0]["memory_features"].shape == (cells, 256) and ep[0]["observations"] is None
    assert ep[3]["image"].shape == (H, W, 3) and ep[3]["depth"].shape == (H, W)
    assert all(fr[0]["memory_reset"] for fr in F.EpisodeDataset(str(root), test_type="episodic"))
    # saved memory round trip (custom_rcnn.py:527-530 -> loader.py:216-223)
    sem = rng.integers(-1, 20, cells)
    mem = rng.standard_normal((cells, 512)).astype(np.float32)
    obs = rng.integers(0, 5, cells).astype(np.float32)
    for seq in ("sceneA_0_0", "sceneA_0_1", "sceneB_1_0"):
        F.save_memory(str(tmp_path / "memory" / (seq + ".h5")), sem, mem, obs)
    d = F.open_store(str(tmp_path / "memory" / "sceneA_0_0.h5"))
    assert set(d) == {"semmap", "impicit_memory", "observations"} and d["semmap"].dtype == np.int32
    sem1, mem1, obs1 = F.load_memory(str(tmp_path / "memory" / "sceneA_0_0.h5"))
    assert np.array_equal(sem1, sem + 1) and np.array_equal(mem1, mem) and np.array_equal(obs1, obs)
    ds2 = F.EpisodeDataset(str(root), semmap_path=str(tmp_path / "memory"))
    assert np.array_equal(ds2[1][5]["memory_features"], mem) and np.array_equal(ds2[1][5]["observations"], obs)
    with pytest.raises(F.StoreError):
        F.open_store(str(tmp_path / "nope.h5"))                             # no silent zero-memory fallback


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs on the host alone (the oracle port timed on the CPU) and prints ONE JSON line with the
    keys the driver reads; under torchrun every rank but 0 exits without work."""
    import json
    import subprocess
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config",
                "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    idle = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                          capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert idle.returncode == 0 and idle.stdout.strip() == ""
