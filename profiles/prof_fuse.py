"""Projection + fusion stage at the bench batch (E=64 episodes, three levels 60x80 / 30x40 / 15x20, K=512 -> N=256):
the tcgen05 kernel (eod_project_fuse) next to the library path it replaces (fp32 matmul + bias + permute + eod_fuse).
CUDA events, 3 warm-up + 10 timed rounds; the inputs of consecutive rounds alternate between two sets (> L2).
PROF_ONLY=tc skips the library path (for ncu)."""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
eod = importlib.import_module("embodied-object-detection_b200")
ops = eod.ops
E = int(os.environ.get("PROF_E", 64)); K = int(os.environ.get("PROF_K", 512)); N = int(os.environ.get("PROF_N", 256))
ROUNDS = int(os.environ.get("PROF_ROUNDS", 10)); ONLY = os.environ.get("PROF_ONLY", "")
dev = torch.device("cuda:0")
torch.manual_seed(0)
shapes = [(60, 80), (30, 40), (15, 20)]
sets = []
for _ in range(2):
    lv = [torch.randn((E, h, w, K), device=dev).half() for h, w in shapes]
    rs = [torch.randn((E, N, h, w), device=dev) for h, w in shapes]
    sets.append((lv, rs))
Wt = [torch.randn((N, K), device=dev) / K ** 0.5 for _ in shapes]
bias = [torch.randn((N,), device=dev) for _ in shapes]
wsplit = [ops.project_split_weights(w) for w in Wt]
outs = [torch.empty((E, N, h, w), device=dev) for h, w in shapes]


def run_tc(lv, rs):                      # one persistent launch for the three levels
    ops.project_fuse_levels(lv, wsplit, bias, rs, 5.0, 0, outs, variant=2)


def run_tc_v1(lv, rs):                   # tile-per-CTA kernel, one launch per level
    ops.project_fuse_levels(lv, wsplit, bias, rs, 5.0, 0, outs, variant=1)


def run_lib(lv, rs):
    for k in range(3):
        mem = torch.matmul(lv[k].to(torch.float32), Wt[k].t()) + bias[k]
        mem = mem.permute(0, 3, 1, 2).contiguous()
        ops.fuse(rs[k], mem, 5.0, 0, outs[k])


def timed(fn):
    for i in range(3):
        fn(*sets[i & 1])
    torch.cuda.synchronize()
    evs = []
    for i in range(ROUNDS):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(*sets[i & 1]); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    return {"median_ms": ms[len(ms) // 2], "min_ms": ms[0]}


M = sum(E * h * w for h, w in shapes)
algo_bytes = M * K * 2 + 2 * M * N * 4 + 3 * 2 * N * K * 2
res = {"E": E, "K": K, "N": N, "rows": M, "algorithmic_bytes": algo_bytes, "flops_fp16_split": 2 * M * K * 2 * N}
res["tcgen05"] = timed(run_tc)
res["tcgen05"]["GBps"] = algo_bytes / res["tcgen05"]["median_ms"] / 1e6
res["tcgen05"]["TFLOPs"] = res["flops_fp16_split"] / res["tcgen05"]["median_ms"] / 1e9
if ONLY != "tc":
    res["tcgen05_tile_per_cta"] = timed(run_tc_v1)
    res["library"] = timed(run_lib)
    res["speedup"] = res["library"]["median_ms"] / res["tcgen05"]["median_ms"]
    a = [o.clone() for o in outs]; run_tc(*sets[0]); b = [o.clone() for o in outs]; run_lib(*sets[0])
    res["max_rel_diff_vs_library"] = max(((x - y).abs().max() / y.abs().max()).item() for x, y in zip(b, outs))
print(json.dumps(res))
